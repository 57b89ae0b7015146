// (2b) Backward of the fused field kernel: the activation-gradient (dgrad) chain, per 128-point tile, on chip.
//   d raw[.,16] -> heads -> fc_feat -> trunk -> d PE -> d(warped point, ambient) (+ embedding-grid scatter) ->
//   tanh' / fc_final / fc_ambient -> warp | hyper-sheet layers.
// Mirrors field_fwd.cu: same CTA shape, same TMA/MMA warp loops, transposed fp16 weight images streamed in the
// order of the backward plan (sahs_build_bwd_plan), activations' sign masks recorded by the training forward.
// d_raw is multiplied by a caller-chosen scale so that fp16 (11-bit significand; bf16's 8 bits give ~8% noise on the
// gradient of the warped point after the 2^9 gain of the encoding) does not underflow.
// Every layer's activation gradient (fp16, scaled) is an operand chunk in shared memory anyway; those chunks go to the
// gradient tape with TMA bulk stores (tile-major chunk images) for the wgrad kernel: dW_l = dY_l^T X_{l-1}.
// ref (what autograd differentiates in the reference): nerf/modules.py:254-295, :371-390, :444-462,
//      nerf/models.py:301-365, nerf/nerf_helpers.py:305-349.
#include <stdlib.h>
#include "field_dev.cuh"

namespace {

struct BwdIO {
  const float* d_raw;       // [P,16]
  const uint4* masks;       // [layer][P][2]
  const float* saves;       // [P,8]: warped point (3), ambient (<=2)
  __half* tape_d;           // gradient tape: [tiles][td_total / 64][128 x 64] fp16 chunk images, scaled by *scale
  float* grid_grad;         // channel-last [32,32,32,32] fp32, atomically accumulated (may be null)
  const float* scale;       // device scalar: d_raw is multiplied by it (fp16 range management); outputs carry the factor
};

// sequential reader of fp32 accumulator columns (16 at a time)
struct TmemCols {
  uint32_t taddr;
  uint32_t v[16];
  int cur = -1;
  __device__ __forceinline__ explicit TmemCols(uint32_t t) : taddr(t) {}
  __device__ __forceinline__ float get(int col) {
    if ((col >> 4) != cur) {
      cur = col >> 4;
      tmem_ld16(taddr + cur * 16, v);
      tmem_ld_wait();
    }
    return __uint_as_float(v[col & 15]);
  }
};

// d x_d += sum_k 2^k (dE_sin[k,d] cos(2^k x_d) - dE_cos[k,d] sin(2^k x_d)) (+ dE_x[d] if the input is included);
// consumes the encoding gradient in the forward's column order.
template <int L, bool INC, int D>
__device__ __forceinline__ int pe_backward(TmemCols& rd, int col, const float (&x)[D], float (&dx)[D]) {
  if (INC) {
#pragma unroll
    for (int d = 0; d < D; ++d) dx[d] += rd.get(col++);
  }
  float s[D], c[D];
#pragma unroll
  for (int k = 0; k < L; ++k) {
#pragma unroll
    for (int d = 0; d < D; ++d) {
      if (k % 5 == 0) {
        sincosf(x[d] * (float)(1 << k), &s[d], &c[d]);
      } else {
        const float s2 = 2.f * s[d] * c[d];
        c[d] = 1.f - 2.f * s[d] * s[d];
        s[d] = s2;
      }
    }
    float gs[D];
#pragma unroll
    for (int d = 0; d < D; ++d) gs[d] = rd.get(col++);
#pragma unroll
    for (int d = 0; d < D; ++d) {
      const float gc = rd.get(col++);
      dx[d] += (float)(1 << k) * (gs[d] * c[d] - gc * s[d]);
    }
  }
  return col;
}

// dgrad epilogue: accumulator (no bias) -> [+ rank-1 sigma term] -> x activation derivative -> fp16 -> X (the operand
// chunks then go to the gradient tape with TMA bulk stores)
template <int ACT, int NBLK, bool ADD_SIGMA, bool FCS = false>
__device__ __forceinline__ void bwd_epilogue(uint32_t tmem_row, int cbeg, uint4 mask, uint8_t* X, int row,
                                             float dsig, const float* __restrict__ w_alpha) {
  uint8_t* rowp = X + (row >> 3) * 1024 + (row & 7) * 128;
  const uint32_t mw[4] = {mask.x, mask.y, mask.z, mask.w};
  uint32_t va[16], vb[16];
  tmem_ld16(tmem_row + cbeg, va);
#pragma unroll
  for (int blk = 0; blk < NBLK; ++blk) {
    const int c0 = cbeg + 16 * blk;
    tmem_ld_wait();
    if (blk + 1 < NBLK) {
      if (blk & 1) tmem_ld16_prefetch(tmem_row + c0 + 16, va, vb[0]);
      else tmem_ld16_prefetch(tmem_row + c0 + 16, vb, va[0]);
    }
    uint32_t (&v)[16] = (blk & 1) ? vb : va;
    const uint32_t bits = mw[blk >> 1] >> ((blk & 1) * 16);
    uint32_t pk[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float g0 = __uint_as_float(v[2 * j]), g1 = __uint_as_float(v[2 * j + 1]);
      if (ADD_SIGMA) {
        g0 += dsig * ldc1<FCS>(w_alpha + c0 + 2 * j);
        g1 += dsig * ldc1<FCS>(w_alpha + c0 + 2 * j + 1);
      }
      if (ACT != ACT_NONE) {
        const float neg = ACT == ACT_LEAKY ? 0.01f : 0.f;
        g0 *= ((bits >> (2 * j)) & 1u) ? 1.f : neg;
        g1 *= ((bits >> (2 * j + 1)) & 1u) ? 1.f : neg;
      }
      pk[j] = pack2<true>(g0, g1);
    }
    uint8_t* chunk = rowp + (c0 >> 6) * kChunkBytes;
    const int u0 = (c0 & 63) >> 3;
#pragma unroll
    for (int q = 0; q < 2; ++q)
      *reinterpret_cast<uint4*>(chunk + ((((u0 + q) ^ row) & 7) << 4)) =
          make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
  }
}

// backward of the trilinear embedding gather: scatter w_c * d_emb into the grid gradient and return the gradient
// w.r.t. the (raw, un-normalised) warped coordinate.  ref: nerf/models.py:346-365 (grid_sample, align_corners=True).
// The scatter is split between the two worker groups (`half` = 16 of the 32 channels each) and uses 16-byte vector
// reductions (REDG.ADD.F32x4): 32 of them per point instead of 256 scalar atomics by half of the threads -- the scalar
// version was 60 % of the dgrad kernel.
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.v4.f32.add [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void grid_backward(const float* __restrict__ g, float* __restrict__ gg, bool scatter, int half,
                                              float x, float y, float z, const float (&de)[32], float (&dxyz)[3]) {
  const float sc = 0.5f * (SAHS_GRID_RES - 1);
  const float ix = __fmul_rn(__fadd_rn(x, 1.f), sc), iy = __fmul_rn(__fadd_rn(y, 1.f), sc),
              iz = __fmul_rn(__fadd_rn(z, 1.f), sc);   // same roundings as grid_gather16
  const float fx = floorf(ix), fy = floorf(iy), fz = floorf(iz);
  float gx = 0.f, gy = 0.f, gz = 0.f;
  // Corner order rotated by lane: the lanes of a warp are consecutive samples of a ray, and where the fine samples
  // cluster at a surface (any trained field) most of them fall into the same voxel -- all issuing the same corner's
  // reduction at the same time serialises 32 atomics on one L2 address (measured: dgrad 0.74 ms per level on white-noise
  // weights, 1.00 ms on the trained-like fixture).  Rotated, a warp instruction touches all 8 corners at once.
  const int rot = (int)(threadIdx.x & 7);
#pragma unroll 1
  for (int ci = 0; ci < 8; ++ci) {
    const int corner = (ci + rot) & 7;
    const int bx = corner & 1, by = (corner >> 1) & 1, bz = corner >> 2;
    const float xi = fx + bx, yi = fy + by, zi = fz + bz;
    const float wx = 1.f - fabsf(ix - xi), wy = 1.f - fabsf(iy - yi), wz = 1.f - fabsf(iz - zi);
    const bool ok = xi >= 0.f && xi <= SAHS_GRID_RES - 1 && yi >= 0.f && yi <= SAHS_GRID_RES - 1 && zi >= 0.f &&
                    zi <= SAHS_GRID_RES - 1;
    if (!ok) continue;
    const size_t off = (((size_t)(int)zi * SAHS_GRID_RES + (int)yi) * SAHS_GRID_RES + (int)xi) * SAHS_GRID_CH;
    const float w = wx * wy * wz;
    float dot = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 v = ldg_stream(g + off + 4 * q);
      dot += v.x * de[4 * q] + v.y * de[4 * q + 1] + v.z * de[4 * q + 2] + v.w * de[4 * q + 3];
    }
    if (scatter) {
#pragma unroll
      for (int q = 0; q < 8; ++q)
        if ((q >> 2) == half)
          red_add_v4(gg + off + 4 * q, w * de[4 * q], w * de[4 * q + 1], w * de[4 * q + 2], w * de[4 * q + 3]);
    }
    gx += (bx ? 1.f : -1.f) * wy * wz * dot;
    gy += (by ? 1.f : -1.f) * wx * wz * dot;
    gz += (bz ? 1.f : -1.f) * wx * wy * dot;
  }
  dxyz[0] = sc * gx; dxyz[1] = sc * gy; dxyz[2] = sc * gz;
}

template <class C, bool PAIR>
__global__ void __launch_bounds__(kThreads, 2)
field_bwd_kernel(const __grid_constant__ FieldPlan plan, const __grid_constant__ NetDims dm,
                 const uint8_t* __restrict__ packed_t, const float* __restrict__ fc, const float* __restrict__ grid,
                 const float* __restrict__ ro, const float* __restrict__ rd, const float* __restrict__ zv, int S,
                 long long P, BwdIO io, int* __restrict__ status) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* X = smem;
  uint8_t* slots = smem + kSmemX;
  // same shared-memory layouts as the forward kernel (PAIR: 3 x 8 KB ring + frame constants in shared memory)
  constexpr int kSlotsBytes = PAIR ? kPairSlots * kPairSlotBytes : kSmemSlots;
  constexpr int kFcBytes = PAIR ? kPairFcFloats * 4 : 0;
  float* fc_s = reinterpret_cast<float*>(smem + kSmemX + kSlotsBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kSmemX + kSlotsBytes + kFcBytes);
  constexpr int NS = PAIR ? kPairSlots : kSlots;
  uint64_t* full = bars;
  uint64_t* empty = bars + NS;
  uint64_t* a_ready = bars + 2 * NS;
  uint64_t* acc_ready = bars + 2 * NS + 1;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 3 * NS + 2);
  uint32_t* peer_tmem = tmem_ptr + 1;

  // warp index through a shuffle: the compiler then knows it is warp-uniform, keeps the role dispatch convergent and
  // the MMA issue loop on the uniform datapath (see field_fwd.cu)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const long long ntiles = (P + kTileRows - 1) / kTileRows;
  const long long npairs = (ntiles + 1) / 2;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  if (threadIdx.x == 0) {
    if (smem_u32(smem) & 1023u) { status[0] = 2; __trap(); }
    for (int i = 0; i < NS; ++i) { mbar_init(&full[i], (PAIR && rank == 0) ? 2 : 1); mbar_init(&empty[i], 1); }
    mbar_init(a_ready, PAIR ? 2 * (kWorkerThreads / 32) : kWorkerThreads);
    mbar_init(acc_ready, 1);
    fence_mbar_init();
  }
  if (PAIR) {
    for (int i = threadIdx.x; i < dm.fc_total; i += kThreads) fc_s[i] = __ldg(fc + i);
  }
  if (warp == kMmaWarp) {
    if (PAIR) tmem_alloc_pair(tmem_ptr, kTmemCols);
    else tmem_alloc(tmem_ptr, kTmemCols);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  if (PAIR) {
    if (rank == 1 && threadIdx.x == 0) st_cluster_u32(mapa_u32(peer_tmem, 0), tmem_base);
    cluster_sync_all();
    if (rank == 0 && threadIdx.x == 0 && *peer_tmem != tmem_base) {
      status[0] = 3;
      __threadfence_system();
      __trap();
    }
    __syncthreads();   // reconverge after the single-thread check
  }

  if (warp == kTmaWarp) {
    if (PAIR) tma_warp_loop_pair(plan, packed_t, slots, full, empty, npairs, rank, status, lane);
    else tma_warp_loop(plan, packed_t, slots, full, empty, ntiles, status, lane);
  } else if (warp == kMmaWarp) {
    if (!PAIR) mma_warp_loop(plan, X, slots, full, empty, a_ready, acc_ready, tmem_base, ntiles, status, lane);
    else if (rank == 0) mma_warp_loop_pair(plan, X, slots, full, empty, a_ready, acc_ready, tmem_base, npairs, status, lane);
    else relay_warp_loop_pair(plan, full, npairs, status, lane);
  } else {
    const float* fcw = PAIR ? fc_s : fc;   // frame constants: shared-memory copy in the pair kernel
    SyncT<PAIR> sy{a_ready, acc_ready, 0u, status, PAIR ? mapa_u32(a_ready, 0) : 0u};
    const int row = threadIdx.x & (kTileRows - 1);
    const int grp = threadIdx.x >> 7;
    const uint32_t tmem_row = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    const int wl = C::USE_W ? dm.w_layers : 0;
    const long long it_end = PAIR ? npairs : ntiles;
    const long long it_step = PAIR ? (long long)cluster_num_x() : (long long)gridDim.x;
    const long long it0 = PAIR ? (long long)cluster_id_x() : (long long)blockIdx.x;
    for (long long it = it0; it < it_end; it += it_step) {
      const long long tile = PAIR ? 2 * it + rank : it;   // PAIR: possibly one past the end (all rows masked)
      const long long p = tile * kTileRows + row;
      const bool valid = p < P;
      const long long pc = valid ? p : P - 1;
      const long long ray = pc / S;
      // gradient tape of this tile: tile-major chunk images (sahs_make_dims), written from X with TMA bulk stores
      uint8_t* tape_tile = reinterpret_cast<uint8_t*>(io.tape_d) + (size_t)tile * (dm.td_total / 64) * kChunkBytes;
      auto tape_put = [&](int chunk0, int nch, int col) {   // after every worker fenced its writes (signal_a)
        group_sync();
        if (threadIdx.x == 0 && tile < ntiles) {
          for (int i = 0; i < nch; ++i)
            tma_bulk_s2g(tape_tile + (size_t)(col / 64 + i) * kChunkBytes, X + (chunk0 + i) * kChunkBytes, kChunkBytes);
          tma_store_commit();
        }
      };
      auto tape_drain = [&]() {   // before X is overwritten: earlier bulk stores have finished reading it
        if (threadIdx.x == 0) tma_store_wait_read();
        group_sync();
      };
      tape_drain();
      auto mask_of = [&](int layer) -> uint4 {
        return __ldg(io.masks + ((size_t)layer * P + pc) * 2 + grp);
      };
      // ---- d raw -> output-layer operand (chunk 0, K = 16) ----
      float dr[16];
      {
        const float4* q = reinterpret_cast<const float4*>(io.d_raw + pc * SAHS_RAW_CH);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float4 v = valid ? __ldg(q + k) : make_float4(0.f, 0.f, 0.f, 0.f);
          dr[4 * k] = v.x; dr[4 * k + 1] = v.y; dr[4 * k + 2] = v.z; dr[4 * k + 3] = v.w;
        }
      }
      {
        const float sc = __ldg(io.scale);
#pragma unroll
        for (int k = 0; k < 16; ++k) dr[k] *= sc;
      }
      const float dsig = dr[15];
      if (grp == 0) {
        uint32_t pk[8];
#pragma unroll
        // column 15 (d sigma) stays in the operand: the transposed output-layer image has a zero row there, and the
        // tape copy of this chunk feeds the wgrad of fc_alpha
        for (int j = 0; j < 8; ++j) pk[j] = pack2<true>(dr[2 * j], dr[2 * j + 1]);
        uint8_t* rowp = X + (row >> 3) * 1024 + (row & 7) * 128;
        *reinterpret_cast<uint4*>(rowp + (((0 ^ row) & 7) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        *reinterpret_cast<uint4*>(rowp + (((1 ^ row) & 7) << 4)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
      }
      signal_a(sy);
      tape_put(0, 1, dm.td_out);
      // ---- heads ----
      for (int i = 3; i >= 0; --i) {   // pass producing d(head hidden i): output layer for i == 3, layers i+1 otherwise
        const uint4 m = mask_of(wl + dm.t_layers + i);
        wait_acc(sy, 5000 + i);
        tape_drain();
        bwd_epilogue<ACT_LEAKY, 8, false, PAIR>(tmem_row, grp * 128, m, X, row, 0.f, nullptr);
        signal_a(sy);
        tape_put(0, 2 * dm.hd / 64, dm.td_hh + i * 2 * dm.hd);
      }
      // d [PE(dir) | embedding]: only the embedding part carries gradient
      const float* sv = io.saves + pc * 8;
      float mapped[3] = {sv[0], sv[1], sv[2]};
      float amb[C::AMB_DIM > 0 ? C::AMB_DIM : 1] = {};
#pragma unroll
      for (int k = 0; k < C::AMB_DIM; ++k) amb[k] = sv[3 + k];
      float dmap[3] = {0.f, 0.f, 0.f};
      float damb[C::AMB_DIM > 0 ? C::AMB_DIM : 1] = {};
      wait_acc(sy, 5100);
      {
        TmemCols rdc(tmem_row);
        float de[32];
#pragma unroll
        for (int q = 0; q < 32; ++q) de[q] = rdc.get(C::DIR_DIM + q);
        grid_backward(grid, io.grid_grad, valid && io.grid_grad != nullptr, grp, mapped[0], mapped[1], mapped[2], de, dmap);
      }
      signal_a(sy);
      // d feat = dir/seg contributions + d sigma * fc_alpha
      wait_acc(sy, 5200);
      tape_drain();
      bwd_epilogue<ACT_NONE, 8, true, PAIR>(tmem_row, grp * 128, make_uint4(0, 0, 0, 0), X, row, dsig, fcw + dm.off_alpha);
      signal_a(sy);
      tape_put(0, dm.th / 64, dm.td_feat);
      // fc_feat^T -> d(trunk hidden L-1)
      {
        const uint4 m = mask_of(wl + dm.t_layers - 1);
        wait_acc(sy, 5300);
        tape_drain();
        bwd_epilogue<ACT_LEAKY, 8, false, PAIR>(tmem_row, grp * 128, m, X, row, 0.f, nullptr);
        signal_a(sy);
        tape_put(0, dm.th / 64, dm.td_th + (dm.t_layers - 1) * dm.th);
      }
      // ---- trunk ----
      for (int i = dm.t_layers - 1; i >= 0; --i) {
        if (i == dm.t_skip || i == 0) {
          wait_acc(sy, 5400 + i);
          TmemCols rdc(tmem_row);
          int col = pe_backward<C::XYZ_L, true, 3>(rdc, 0, mapped, dmap);
          if (C::AMB_PE > 0) pe_backward<C::AMB_L, C::AMB_INC, (C::AMB_DIM > 0 ? C::AMB_DIM : 1)>(rdc, col, amb, damb);
          if (i > 0) signal_a(sy);   // i == 0: the next pass (deformation layers) is released after d(h5) is written
        }
        if (i > 0) {
          const uint4 m = mask_of(wl + i - 1);
          wait_acc(sy, 5500 + i);
          tape_drain();
          bwd_epilogue<ACT_LEAKY, 8, false, PAIR>(tmem_row, grp * 128, m, X, row, 0.f, nullptr);
          signal_a(sy);
          tape_put(0, dm.th / 64, dm.td_th + (i - 1) * dm.th);
        }
      }
      if (C::USE_W) {
        // ---- tanh', fc_final / fc_ambient (fp32) -> d(last deformation hidden layer) ----
        float pt[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) pt[k] = __fadd_rn(ro[ray * 3 + k], __fmul_rn(rd[ray * 3 + k], zv[pc]));
        float dpre[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const float t = mapped[k] - pt[k];
          dpre[k] = dmap[k] * (1.f - t * t);
        }
        tape_drain();   // the trunk's last gradient chunks have left X
        if (grp == 0) {
          // d(pre-tanh dx, ambient) -> columns 0-15 of chunk 3 (idle during the deformation layers) -> tape item td_final
          float f8[8] = {dpre[0], dpre[1], dpre[2], 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int k = 0; k < C::AMB_DIM; ++k) f8[3 + k] = damb[k];
          uint8_t* r3 = X + 3 * kChunkBytes + (row >> 3) * 1024 + (row & 7) * 128;
          *reinterpret_cast<uint4*>(r3 + (((0 ^ row) & 7) << 4)) =
              make_uint4(pack2<true>(f8[0], f8[1]), pack2<true>(f8[2], f8[3]), pack2<true>(f8[4], f8[5]),
                         pack2<true>(f8[6], f8[7]));
          *reinterpret_cast<uint4*>(r3 + (((1 ^ row) & 7) << 4)) = make_uint4(0u, 0u, 0u, 0u);
        }
        const float* wf = fcw + dm.off_wfinal;
        const float* wa = wf + 3 * dm.wh + 4;
        const uint4 m5 = mask_of(dm.w_layers - 1);
        const uint32_t mw[4] = {m5.x, m5.y, m5.z, m5.w};
        uint8_t* rowp = X + (row >> 3) * 1024 + (row & 7) * 128;
#pragma unroll
        for (int blk = 0; blk < 6; ++blk) {
          const int c0 = grp * 96 + 16 * blk;
          const uint32_t bits = mw[blk >> 1] >> ((blk & 1) * 16);
          uint32_t pk[8];
#pragma unroll
          for (int j = 0; j < 16; j += 2) {
            float g[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int c = c0 + j + e;
              float v = 0.f;
              if (c < dm.wh) {
#pragma unroll
                for (int k = 0; k < 3; ++k) v += ldc1<PAIR>(wf + k * dm.wh + c) * dpre[k];
              } else {
#pragma unroll
                for (int k = 0; k < C::AMB_DIM; ++k) v += ldc1<PAIR>(wa + k * dm.hh + (c - dm.wh)) * damb[k];
              }
              g[e] = ((bits >> (j + e)) & 1u) ? v : 0.f;
            }
            pk[j >> 1] = pack2<true>(g[0], g[1]);
          }
          uint8_t* chunk = rowp + (c0 >> 6) * kChunkBytes;
          const int u0 = (c0 & 63) >> 3;
#pragma unroll
          for (int q = 0; q < 2; ++q)
            *reinterpret_cast<uint4*>(chunk + ((((u0 + q) ^ row) & 7) << 4)) =
                make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
        }
        signal_a(sy);
        tape_put(0, dm.whh / 64, dm.td_wh + (dm.w_layers - 1) * dm.whh);
        tape_put(3, 1, dm.td_final);
        for (int i = dm.w_layers - 1; i >= 1; --i) {
          const uint4 m = mask_of(i - 1);
          wait_acc(sy, 5600 + i);
          tape_drain();
          bwd_epilogue<ACT_RELU, 6, false, PAIR>(tmem_row, grp * 96, m, X, row, 0.f, nullptr);
          if (i > 1) signal_a(sy);
          else fence_proxy_async_smem();   // the last gradient chunk is only stored, not multiplied
          tape_put(0, dm.whh / 64, dm.td_wh + (i - 1) * dm.whh);
        }
      }
      tc_fence_before();
      group_sync();
    }
    if (threadIdx.x == 0) tma_store_wait_all();   // tape stores complete before the CTA exits
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();   // the peer's smem / TMEM / barriers stay alive until both CTAs are done
  if (warp == kMmaWarp) {
    tc_fence_after();
    if (PAIR) tmem_dealloc_pair(tmem_base, kTmemCols);
    else tmem_dealloc(tmem_base, kTmemCols);
  }
}

// The CTA-pair variant of the dgrad kernel is opt-in (SAHS_BWD_PAIR=1): its passes are shorter than the forward's and
// the pair's lock step costs more than the halved weight traffic saves (measured 1.68 vs 1.56 ms per level).
static bool bwd_pair_enabled() {
  const char* e = getenv("SAHS_BWD_PAIR");
  return e && e[0] == '1';
}

template <class C>
int launch_bwd(const HostPlan& hp, const void* packed_t, const float* fc, const float* grid, const float* ro,
               const float* rd, const float* z, int R, int S, BwdIO io, cudaStream_t st) {
  int* status = sahs_status_words(1);
  SAHS_CHECK_ARG(status, "cannot allocate the diagnostic word");
  const long long P = (long long)R * S;
  const long long ntiles = (P + kTileRows - 1) / kTileRows;
  const uint8_t* pk = (const uint8_t*)packed_t;
  if (bwd_pair_enabled() && hp.dims.fc_total <= kPairFcFloats) {
    auto kfn = field_bwd_kernel<C, true>;
    SAHS_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, kPairSmemTotal));
    const long long npairs = (ntiles + 1) / 2;
    long long nclusters = sahs_num_sms();
    if (nclusters > npairs) nclusters = npairs;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(2 * nclusters));
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = kPairSmemTotal;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    SAHS_CUDA(cudaLaunchKernelEx(&cfg, kfn, hp.plan, hp.dims, pk, fc, grid, ro, rd, z, S, P, io, status));
    SAHS_LAUNCH_CHECK();
    return SAHS_OK;
  }
  auto kfn = field_bwd_kernel<C, false>;
  SAHS_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal));
  long long grid_dim = 2LL * sahs_num_sms();
  if (grid_dim > ntiles) grid_dim = ntiles;
  kfn<<<(unsigned)grid_dim, kThreads, kSmemTotal, st>>>(hp.plan, hp.dims, pk, fc, grid, ro, rd, z, S, P, io, status);
  SAHS_LAUNCH_CHECK();
  return SAHS_OK;
}

}  // namespace


extern "C" int sahs_field_bwd(const sahs_model_spec* spec, int level, const void* packed_t, const float* frame_const,
                              const float* grid_cl, const float* ro, const float* rd, const float* z, int num_rays,
                              int num_samples, const float* d_raw, const float* scale, const void* masks,
                              const float* saves, void* tape_d, float* grid_grad, void* stream) {
  SAHS_CHECK_ARG(spec, "null spec");
  SAHS_CHECK_ARG(level == 0 || level == 1, "level must be 0 (coarse) or 1 (fine)");
  SAHS_CHECK_ARG(num_rays >= 0 && num_samples > 0, "bad extents");
  if (num_rays == 0) return SAHS_OK;
  SAHS_CHECK_ARG(packed_t && frame_const && grid_cl && ro && rd && z && d_raw && scale && masks && saves && tape_d, "null pointer");
  static thread_local HostPlan hp;
  int rc = sahs_build_bwd_plan(*spec, nullptr, hp);
  if (rc) return rc;
  const sahs_model_spec& s = *spec;
  SAHS_CHECK_ARG(s.xyz_inc && s.dir_inc && s.dir_L == 4 && s.use_grid, "unsupported encoding options");
  SAHS_CHECK_ARG(!hp.dims.use_w || hp.dims.whh == 192, "warp 128 + hyper 64 hidden units expected");
  BwdIO io{d_raw, (const uint4*)masks, saves, (__half*)tape_d, grid_grad, scale};
  cudaStream_t st = (cudaStream_t)stream;
#define SAHS_TRY(XL, AD, AL, AI, UW)                                                                        \
  if (s.xyz_L == XL && (UW ? (s.amb_dim == AD && s.amb_L == AL && (s.amb_inc != 0) == AI) : true) &&        \
      ((s.use_warp != 0) == UW))                                                                            \
    return launch_bwd<FieldCfg<XL, AD, AL, AI, 4, UW>>(hp, packed_t, frame_const, grid_cl, ro, rd, z, num_rays, \
                                                       num_samples, io, st);
  SAHS_TRY(10, 2, 4, true, true)
  SAHS_TRY(15, 1, 15, false, true)
  SAHS_TRY(10, 0, 0, false, false)
#undef SAHS_TRY
  sahs_set_error("sahs_field_bwd: no kernel instantiated for this model spec");
  return SAHS_EUNSUPPORTED;
}

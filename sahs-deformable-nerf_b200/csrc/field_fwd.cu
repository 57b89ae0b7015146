// (2) Fused field kernel: per 128-point tile, entirely on chip:
//   ray point -> PE -> [warp | hyper-sheet] MLPs -> tanh / ambient -> trilinear embedding gather -> PE ->
//   radiance trunk -> fc_feat / sigma -> direction and semantic heads -> raw[.,16]
// ref: nerf/train_utils.py:9-50 (run_network), nerf/models.py:301-380, :514-528, nerf/modules.py:254-295,
//      :371-390, :444-462, nerf/nerf_helpers.py:305-349.
//
// Execution model (sm_100a): 192 threads per CTA, 2 CTAs per SM (TMEM 2 x 256 columns, smem 2 x 113 KB).
//   warps 0-3  workers : thread t owns tile row t == TMEM lane t; encodings, epilogues (TMEM -> bias/act ->
//                        bf16 -> swizzled smem), small heads in fp32, final store
//   warp  4    TMA     : streams the packed weight image stage by stage with cp.async.bulk (UBLKCP)
//   warp  5    MMA     : single-thread tcgen05.mma issue (UTCHMMA), accumulators in TMEM
// Within a CTA the tile is processed pass by pass (bulk synchronous through two mbarriers); the co-resident CTA
// on the same SM fills the tensor pipe while this one runs its epilogue.
#include "sahs_common.cuh"
#include "field_plan.cuh"

namespace {

constexpr int kWorkerThreads = 128;
constexpr int kThreads = 192;
constexpr int kSlots = 3;
constexpr int kTmemCols = 256;
constexpr int kSmemX = 4 * kChunkBytes;                       // 64 KB
constexpr int kSmemSlots = kSlots * kStageSlotBytes;          // 48 KB
constexpr int kSmemBar = kSmemX + kSmemSlots;                 // barriers after the tiles
constexpr int kSmemTotal = kSmemBar + 128;

__device__ int g_field_status[4];

enum { ACT_NONE = 0, ACT_RELU = 1, ACT_LEAKY = 2 };

template <int ACT>
__device__ __forceinline__ float act_fn(float v) {
  if (ACT == ACT_RELU) return fmaxf(v, 0.f);
  if (ACT == ACT_LEAKY) return fmaxf(v, 0.01f * v);
  return v;
}

struct Sync {
  uint64_t* full;
  uint64_t* empty;
  uint64_t* a_ready;
  uint64_t* acc_ready;
  uint32_t acc_par;
  int* status;
};

__device__ __forceinline__ void signal_a(Sync& sy) {
  fence_proxy_async_smem();
  tc_fence_before();
  mbar_arrive(sy.a_ready);
}
__device__ __forceinline__ void wait_acc(Sync& sy, int tag) {
  mbar_wait(sy.acc_ready, sy.acc_par, sy.status, tag);
  sy.acc_par ^= 1;
  tc_fence_after();
}

// store N (multiple of 8) fp32 values of this row as bf16 into X chunks starting at chunk0
template <int N>
__device__ __forceinline__ void store_row(uint8_t* X, int chunk0, int row, const float (&e)[N]) {
  static_assert(N % 8 == 0, "row width must be a multiple of 8");
  uint8_t* rowp = X + (row >> 3) * 1024 + (row & 7) * 128;
#pragma unroll
  for (int u = 0; u < N / 8; ++u) {
    uint4 q;
    q.x = pack_bf16x2(e[8 * u + 0], e[8 * u + 1]);
    q.y = pack_bf16x2(e[8 * u + 2], e[8 * u + 3]);
    q.z = pack_bf16x2(e[8 * u + 4], e[8 * u + 5]);
    q.w = pack_bf16x2(e[8 * u + 6], e[8 * u + 7]);
    uint8_t* p = rowp + (chunk0 + (u >> 3)) * kChunkBytes + ((((u & 7) ^ row) & 7) << 4);
    *reinterpret_cast<uint4*>(p) = q;
  }
}

// positional encoding of D values into e[base ...] in the reference's order (nerf_helpers.py:341-349).
// Accurate sincosf every 5th octave (2^k * x is exact), double-angle recurrence in between (error < 2e-6).
template <int L, bool INC, int D, int N>
__device__ __forceinline__ void pe_fill(float (&e)[N], int base, const float (&x)[D]) {
  if (INC) {
#pragma unroll
    for (int d = 0; d < D; ++d) e[base + d] = x[d];
  }
  constexpr int o = INC ? D : 0;
#pragma unroll
  for (int d = 0; d < D; ++d) {
    float s = 0.f, c = 1.f;
#pragma unroll
    for (int k = 0; k < L; ++k) {
      if (k % 5 == 0) {
        sincosf(x[d] * (float)(1 << k), &s, &c);
      } else {
        float s2 = 2.f * s * c;
        float c2 = 1.f - 2.f * s * s;
        s = s2; c = c2;
      }
      e[base + o + (2 * k) * D + d] = s;
      e[base + o + (2 * k + 1) * D + d] = c;
    }
  }
}

// epilogue of one pass: TMEM accumulator -> +bias -> activation -> bf16 -> X (in place, swizzled)
// optional extras: fp32 dot with a weight row (sigma head), debug dump
template <int ACT, bool DOT>
__device__ __forceinline__ float epilogue_store(uint32_t tmem_row, int ncols, const float* __restrict__ bias,
                                                uint8_t* X, int row, const float* __restrict__ dot_w,
                                                float* __restrict__ dbg_row) {
  float dot = 0.f;
  uint8_t* rowp = X + (row >> 3) * 1024 + (row & 7) * 128;
  for (int c0 = 0; c0 < ncols; c0 += 32) {
    uint32_t v[32];
    tmem_ld32(tmem_row + c0, v);
    tmem_ld_wait();
    uint32_t pk[16];
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(bias + c0 + j));
      float f0 = act_fn<ACT>(__uint_as_float(v[j + 0]) + b.x);
      float f1 = act_fn<ACT>(__uint_as_float(v[j + 1]) + b.y);
      float f2 = act_fn<ACT>(__uint_as_float(v[j + 2]) + b.z);
      float f3 = act_fn<ACT>(__uint_as_float(v[j + 3]) + b.w);
      if (DOT) {
        const float4 w = __ldg(reinterpret_cast<const float4*>(dot_w + c0 + j));
        dot += f0 * w.x + f1 * w.y + f2 * w.z + f3 * w.w;
      }
      if (dbg_row) {
        dbg_row[c0 + j + 0] = f0; dbg_row[c0 + j + 1] = f1; dbg_row[c0 + j + 2] = f2; dbg_row[c0 + j + 3] = f3;
      }
      pk[j / 2] = pack_bf16x2(f0, f1);
      pk[j / 2 + 1] = pack_bf16x2(f2, f3);
    }
    uint8_t* chunk = rowp + (c0 >> 6) * kChunkBytes;
    const int u0 = (c0 & 63) >> 3;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      uint4 val = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
      *reinterpret_cast<uint4*>(chunk + ((((u0 + q) ^ row) & 7) << 4)) = val;
    }
  }
  return dot;
}

// trilinear gather from the channel-last embedding grid, ref: nerf/models.py:346-365 (align_corners=True,
// zero padding, raw warped coordinates; x -> last grid dim, z -> first)
__device__ __forceinline__ void grid_gather(const float* __restrict__ g, float x, float y, float z, float (&out)[32]) {
  const float sc = 0.5f * (SAHS_GRID_RES - 1);
  const float ix = (x + 1.f) * sc, iy = (y + 1.f) * sc, iz = (z + 1.f) * sc;
  const float fx = floorf(ix), fy = floorf(iy), fz = floorf(iz);
#pragma unroll
  for (int c = 0; c < 32; ++c) out[c] = 0.f;
#pragma unroll
  for (int corner = 0; corner < 8; ++corner) {
    const float xi = fx + (corner & 1), yi = fy + ((corner >> 1) & 1), zi = fz + (corner >> 2);
    const float w = (1.f - fabsf(ix - xi)) * (1.f - fabsf(iy - yi)) * (1.f - fabsf(iz - zi));
    const bool ok = xi >= 0.f && xi <= SAHS_GRID_RES - 1 && yi >= 0.f && yi <= SAHS_GRID_RES - 1 && zi >= 0.f &&
                    zi <= SAHS_GRID_RES - 1;
    if (ok) {
      const float4* p = reinterpret_cast<const float4*>(
          g + ((((size_t)(int)zi * SAHS_GRID_RES + (int)yi) * SAHS_GRID_RES + (int)xi) * SAHS_GRID_CH));
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 v = __ldg(p + q);
        out[4 * q + 0] += w * v.x; out[4 * q + 1] += w * v.y; out[4 * q + 2] += w * v.z; out[4 * q + 3] += w * v.w;
      }
    }
  }
}

template <int XYZ_L_, int AMB_DIM_, int AMB_L_, bool AMB_INC_, int DIR_L_, bool USE_W_, bool USE_GRID_>
struct FieldCfg {
  static constexpr int XYZ_L = XYZ_L_;
  static constexpr int AMB_DIM = AMB_DIM_;
  static constexpr int AMB_L = AMB_L_;
  static constexpr bool AMB_INC = AMB_INC_;
  static constexpr int DIR_L = DIR_L_;
  static constexpr bool USE_W = USE_W_;
  static constexpr bool USE_GRID = USE_GRID_;
  static constexpr int E0_DIM = 3 + 6 * XYZ_L;                                      // include_input is always on
  static constexpr int E0_PAD = (E0_DIM + 15) / 16 * 16;
  static constexpr int AMB_PE = USE_W ? ((AMB_INC ? AMB_DIM : 0) + 2 * AMB_DIM * AMB_L) : 0;
  static constexpr int E1_DIM = E0_DIM + AMB_PE;
  static constexpr int E1_PAD = (E1_DIM + 15) / 16 * 16;
  static constexpr int DIR_DIM = 3 + 6 * DIR_L;
  static constexpr int XTRA_DIM = DIR_DIM + (USE_GRID ? 32 : 0);
};

template <class C>
__global__ void __launch_bounds__(kThreads, 2)
field_fwd_kernel(const __grid_constant__ FieldPlan plan, const __grid_constant__ NetDims dm,
                 const uint8_t* __restrict__ packed, const float* __restrict__ fc, const float* __restrict__ grid,
                 const float* __restrict__ ro, const float* __restrict__ rd, const float* __restrict__ zv,
                 int S, long long P, float* __restrict__ raw_out, float* __restrict__ dbg, int dbg_pass,
                 int* __restrict__ status) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* X = smem;
  uint8_t* slots = smem + kSmemX;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kSmemBar);
  uint64_t* full = bars;               // [kSlots]
  uint64_t* empty = bars + kSlots;     // [kSlots]
  uint64_t* a_ready = bars + 2 * kSlots;
  uint64_t* acc_ready = bars + 2 * kSlots + 1;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * kSlots + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long ntiles = (P + kTileRows - 1) / kTileRows;

  if (threadIdx.x == 0) {
    if (smem_u32(smem) & 1023u) {  // UMMA 128B swizzle needs 1024-byte aligned tiles
      status[0] = 2;
      __trap();
    }
    for (int i = 0; i < kSlots; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(a_ready, kWorkerThreads);
    mbar_init(acc_ready, 1);
    fence_mbar_init();
  }
  if (warp == 5) tmem_alloc(tmem_ptr, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 4) {
    // ================================ TMA producer ==============================================
    uint32_t slot = 0, phase = 0;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      uint32_t off = 0;
      for (int st = 0; st < plan.num_stages; ++st) {
        const uint32_t bytes = (uint32_t)plan.st[st].n8 * 1024u;
        mbar_wait(&empty[slot], phase ^ 1, status, 100);
        if (lane == 0) {
          mbar_arrive_expect_tx(&full[slot], bytes);
          tma_bulk_g2s(slots + slot * kStageSlotBytes, packed + off, bytes, &full[slot]);
        }
        __syncwarp();
        off += bytes;
        if (++slot == kSlots) { slot = 0; phase ^= 1; }
      }
    }
  } else if (warp == 5) {
    // ================================ MMA issuer ================================================
    uint32_t slot = 0, phase = 0, a_par = 0;
    const uint32_t x_addr = smem_u32(X), s_addr = smem_u32(slots);
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      for (int st = 0; st < plan.num_stages; ++st) {
        const StageRec r = plan.st[st];
        const uint32_t flags = r.kflags >> 3, ksteps = r.kflags & 7;
        if (flags & ST_WAIT_A) {
          mbar_wait(a_ready, a_par, status, 200 + st);
          a_par ^= 1;
        }
        mbar_wait(&full[slot], phase, status, 400 + st);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t idesc = umma_idesc_bf16_m128((uint32_t)r.n8 * 8u);
          const uint64_t a0 = umma_smem_desc_sw128(x_addr + r.a_chunk * kChunkBytes);
          const uint64_t b0 = umma_smem_desc_sw128(s_addr + slot * kStageSlotBytes);
          const uint32_t d = tmem_base + (uint32_t)r.d_col8 * 8u;
          for (uint32_t k = 0; k < ksteps; ++k) {
            // advancing K by 16 bf16 = 32 bytes = 2 descriptor address units inside the 128B swizzle atom
            tc_mma_bf16(d, a0 + 2 * k, b0 + 2 * k, idesc, (k > 0 || !(flags & ST_FRESH)) ? 1u : 0u);
          }
          tc_commit(&empty[slot]);
          if (flags & ST_COMMIT) tc_commit(acc_ready);
        }
        __syncwarp();
        if (++slot == kSlots) { slot = 0; phase ^= 1; }
      }
    }
  } else {
    // ================================ workers ===================================================
    Sync sy{full, empty, a_ready, acc_ready, 0u, status};
    const int row = threadIdx.x;
    const uint32_t tmem_row = tmem_base + ((uint32_t)(warp * 32) << 16);
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const long long p = tile * kTileRows + row;
      const bool valid = p < P;
      const long long pc = valid ? p : P - 1;
      const long long ray = pc / S;
      const float zz = zv[pc];
      const float dir[3] = {rd[ray * 3 + 0], rd[ray * 3 + 1], rd[ray * 3 + 2]};
      float pt[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) pt[k] = __fadd_rn(ro[ray * 3 + k], __fmul_rn(dir[k], zz));
      float* dbg_row = (dbg && tile == 0) ? dbg + row * 256 : nullptr;
      float mapped[3] = {pt[0], pt[1], pt[2]};
      float amb[C::AMB_DIM > 0 ? C::AMB_DIM : 1] = {};

      if (C::USE_W) {
        // -------- deformation phase: warp | hyper-sheet --------
        {
          float e[C::E0_PAD];
#pragma unroll
          for (int i = C::E0_DIM; i < C::E0_PAD; ++i) e[i] = 0.f;
          pe_fill<C::XYZ_L, true, 3>(e, 0, pt);
          store_row<C::E0_PAD>(X, dm.e0_chunk_base, row, e);
        }
        signal_a(sy);
        for (int i = 0; i < dm.w_layers; ++i) {
          wait_acc(sy, 1000 + i);
          if (i == dm.w_skip && !dm.e0_resident) {
            float e[C::E0_PAD];
#pragma unroll
            for (int k = C::E0_DIM; k < C::E0_PAD; ++k) e[k] = 0.f;
            pe_fill<C::XYZ_L, true, 3>(e, 0, pt);
            store_row<C::E0_PAD>(X, 0, row, e);
            signal_a(sy);
            wait_acc(sy, 1100 + i);
          }
          const float* bias = fc + dm.off_wbias + i * dm.whh;
          if (i < dm.w_layers - 1) {
            epilogue_store<ACT_RELU, false>(tmem_row, dm.whh, bias, X, row, nullptr,
                                            (dbg_row && dbg_pass == SAHS_DBG_WARP(i)) ? dbg_row : nullptr);
            signal_a(sy);
          } else {
            // last hidden layer stays in fp32: dx = tanh(fc_final h_w), ambient = fc_ambient h_h
            const float* wf = fc + dm.off_wfinal;
            const float* bf = wf + 3 * dm.wh;
            const float* wa = bf + 3;
            const float* ba = wa + C::AMB_DIM * dm.hh;
            float acc_dx[3] = {0.f, 0.f, 0.f};
            float acc_am[C::AMB_DIM > 0 ? C::AMB_DIM : 1] = {};
            for (int c0 = 0; c0 < dm.whh; c0 += 32) {
              uint32_t v[32];
              tmem_ld32(tmem_row + c0, v);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const float h = fmaxf(__uint_as_float(v[j]) + __ldg(bias + c0 + j), 0.f);
                if (dbg_row && dbg_pass == SAHS_DBG_WARP(i)) dbg_row[c0 + j] = h;
                const int c = c0 + j;
                if (c0 < dm.wh) {
#pragma unroll
                  for (int k = 0; k < 3; ++k) acc_dx[k] += h * __ldg(wf + k * dm.wh + c);
                } else {
#pragma unroll
                  for (int k = 0; k < C::AMB_DIM; ++k) acc_am[k] += h * __ldg(wa + k * dm.hh + (c - dm.wh));
                }
              }
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) mapped[k] = pt[k] + tanhf(acc_dx[k] + __ldg(bf + k));
#pragma unroll
            for (int k = 0; k < C::AMB_DIM; ++k) amb[k] = acc_am[k] + __ldg(ba + k);
          }
        }
      }
      // -------- spatial embedding gather (kept packed until layers_dir.0) --------
      uint32_t emb_pk[16];
      if (C::USE_GRID) {
        float emb[32];
        grid_gather(grid, mapped[0], mapped[1], mapped[2], emb);
#pragma unroll
        for (int q = 0; q < 16; ++q) emb_pk[q] = pack_bf16x2(emb[2 * q], emb[2 * q + 1]);
        if (dbg_row && dbg_pass == SAHS_DBG_MAPPED) {
#pragma unroll
          for (int q = 0; q < 32; ++q) dbg_row[8 + q] = emb[q];
        }
      }
      if (dbg_row && dbg_pass == SAHS_DBG_MAPPED) {
        dbg_row[0] = mapped[0]; dbg_row[1] = mapped[1]; dbg_row[2] = mapped[2];
#pragma unroll
        for (int k = 0; k < C::AMB_DIM; ++k) dbg_row[3 + k] = amb[k];
      }
      // -------- trunk --------
      auto write_e1 = [&]() {
        float e[C::E1_PAD];
#pragma unroll
        for (int i = C::E1_DIM; i < C::E1_PAD; ++i) e[i] = 0.f;
        pe_fill<C::XYZ_L, true, 3>(e, 0, mapped);
        if (C::AMB_PE > 0) pe_fill<C::AMB_L, C::AMB_INC, (C::AMB_DIM > 0 ? C::AMB_DIM : 1)>(e, C::E0_DIM, amb);
        store_row<C::E1_PAD>(X, 0, row, e);
      };
      write_e1();
      signal_a(sy);
      for (int i = 0; i < dm.t_layers; ++i) {
        wait_acc(sy, 2000 + i);
        if (i == dm.t_skip) {
          write_e1();
          signal_a(sy);
          wait_acc(sy, 2100 + i);
        }
        epilogue_store<ACT_LEAKY, false>(tmem_row, dm.th, fc + dm.off_tbias + i * dm.th, X, row, nullptr,
                                         (dbg_row && dbg_pass == SAHS_DBG_TRUNK(i)) ? dbg_row : nullptr);
        signal_a(sy);
      }
      // fc_feat (no activation) + sigma = fc_alpha(feat) in fp32
      wait_acc(sy, 2200);
      float sigma = epilogue_store<ACT_NONE, true>(tmem_row, dm.th, fc + dm.off_featb, X, row, fc + dm.off_alpha,
                                                   (dbg_row && dbg_pass == SAHS_DBG_TRUNK(dm.t_layers)) ? dbg_row : nullptr);
      sigma += __ldg(fc + dm.off_alpha + dm.th);
      signal_a(sy);
      // -------- heads: layers_dir.0 = [feat | PE(dir) | emb], layers_seg.0 = feat --------
      wait_acc(sy, 3000);
      {
        float e[64];
#pragma unroll
        for (int i = 0; i < 64; ++i) e[i] = 0.f;
        pe_fill<C::DIR_L, true, 3>(e, 0, dir);
        store_row<64>(X, 0, row, e);
        if (C::USE_GRID) {
          // overwrite the embedding columns DIR_DIM .. DIR_DIM+31 (bf16 pairs are not 4-byte aligned when
          // DIR_DIM is odd, so go through the scalar path)
          uint8_t* rowp = X + (row >> 3) * 1024 + (row & 7) * 128;
#pragma unroll
          for (int q = 0; q < 32; ++q) {
            const int col = C::DIR_DIM + q;
            const uint32_t pair = emb_pk[q >> 1];
            const uint16_t hv = (q & 1) ? (uint16_t)(pair >> 16) : (uint16_t)(pair & 0xffffu);
            *reinterpret_cast<uint16_t*>(rowp + ((((col >> 3) ^ row) & 7) << 4) + (col & 7) * 2) = hv;
          }
        }
      }
      signal_a(sy);
      for (int i = 0; i < 4; ++i) {
        wait_acc(sy, 3100 + i);
        epilogue_store<ACT_LEAKY, false>(tmem_row, 2 * dm.hd, fc + dm.off_hbias + i * 2 * dm.hd, X, row, nullptr,
                                         (dbg_row && dbg_pass == SAHS_DBG_HEAD(i)) ? dbg_row : nullptr);
        signal_a(sy);
      }
      // -------- output layer: cols 0-2 rgb, 3-14 seg (+ sigma) --------
      wait_acc(sy, 3200);
      {
        uint32_t v[16];
        tmem_ld16(tmem_row, v);
        tmem_ld_wait();
        float o[16];
#pragma unroll
        for (int k = 0; k < 15; ++k) o[k] = __uint_as_float(v[k]) + __ldg(fc + dm.off_outb + k);
        o[15] = sigma;
        if (valid) {
          float4* dst = reinterpret_cast<float4*>(raw_out + p * SAHS_RAW_CH);
#pragma unroll
          for (int q = 0; q < 4; ++q) dst[q] = make_float4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
        }
      }
      tc_fence_before();
    }
  }
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

template <class C>
int launch_field(const HostPlan& hp, const void* packed, const float* fc, const float* grid, const float* ro,
                 const float* rd, const float* z, int R, int S, float* raw, float* dbg, int dbg_pass,
                 cudaStream_t st) {
  auto kfn = field_fwd_kernel<C>;
  static thread_local bool attr_set = false;
  if (!attr_set) {
    SAHS_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal));
    attr_set = true;
  }
  int* status = nullptr;
  SAHS_CUDA(cudaGetSymbolAddress((void**)&status, g_field_status));
  const long long P = (long long)R * S;
  const long long ntiles = (P + kTileRows - 1) / kTileRows;
  long long grid_dim = 2LL * sahs_num_sms();
  if (grid_dim > ntiles) grid_dim = ntiles;
  kfn<<<(unsigned)grid_dim, kThreads, kSmemTotal, st>>>(hp.plan, hp.dims, (const uint8_t*)packed, fc, grid, ro, rd, z, S,
                                                       P, raw, dbg, dbg_pass, status);
  SAHS_LAUNCH_CHECK();
  return SAHS_OK;
}

}  // namespace

extern "C" int sahs_field_status(int* out4_host) {
  SAHS_CUDA(cudaMemcpyFromSymbol(out4_host, g_field_status, sizeof(int) * 4));
  return SAHS_OK;
}

extern "C" int sahs_field_fwd(const sahs_model_spec* spec, int level, const void* packed, const float* frame_const,
                              const float* grid_cl, const float* ro, const float* rd, const float* z, int num_rays,
                              int num_samples, float* raw_out, float* debug, int debug_pass, void* stream) {
  SAHS_CHECK_ARG(spec && packed && frame_const && ro && rd && z && raw_out, "null pointer");
  SAHS_CHECK_ARG(level == 0 || level == 1, "level must be 0 (coarse) or 1 (fine)");
  SAHS_CHECK_ARG(num_rays >= 0 && num_samples > 0, "bad extents");
  SAHS_CHECK_ARG(!spec->use_grid || grid_cl, "grid pointer required when use_spatial_embeddings");
  if (num_rays == 0) return SAHS_OK;
  static thread_local HostPlan hp;
  int rc = sahs_build_host_plan(*spec, nullptr, hp);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const sahs_model_spec& s = *spec;
  SAHS_CHECK_ARG(s.xyz_inc && s.dir_inc && s.dir_L == 4, "include_input_xyz/dir and num_encoding_fn_dir=4 expected");
#define SAHS_TRY(XL, AD, AL, AI, UW)                                                                             \
  if (s.xyz_L == XL && (UW ? (s.amb_dim == AD && s.amb_L == AL && (s.amb_inc != 0) == AI) : true) &&             \
      ((s.use_warp != 0) == UW) && s.use_grid)                                                                   \
    return launch_field<FieldCfg<XL, AD, AL, AI, 4, UW, true>>(hp, packed, frame_const, grid_cl, ro, rd, z,      \
                                                               num_rays, num_samples, raw_out, debug, debug_pass, st);
  SAHS_TRY(10, 2, 4, true, true)     // config/audio/*.yml
  SAHS_TRY(15, 1, 15, false, true)   // config/expression/person_{2,3}.yml
  SAHS_TRY(10, 0, 0, false, false)   // config/expression/person_1.yml (no deformation, no hyper space)
#undef SAHS_TRY
  sahs_set_error("sahs_field_fwd: no kernel instantiated for this model spec");
  return SAHS_EUNSUPPORTED;
}

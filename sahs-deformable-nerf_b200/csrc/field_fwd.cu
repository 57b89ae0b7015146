// (2) Fused field kernel: per 128-point tile, entirely on chip:
//   ray point -> PE -> [warp | hyper-sheet] MLPs -> tanh / ambient -> trilinear embedding gather -> PE ->
//   radiance trunk -> fc_feat / sigma -> direction and semantic heads -> raw[.,16]
// ref: nerf/train_utils.py:9-50 (run_network), nerf/models.py:301-380, :514-528, nerf/modules.py:254-295,
//      :371-390, :444-462, nerf/nerf_helpers.py:305-349.
//
// Execution model (sm_100a): 320 threads per CTA, 2 CTAs per SM (TMEM 2 x 256 columns, smem 2 x 113 KB).
//   warps 0-7  workers : thread t owns tile row (t & 127) == TMEM lane; the two warp groups (t >> 7) split the
//                        columns of every epilogue (TMEM -> +bias -> activation -> 16-bit -> swizzled smem),
//                        the encodings, the small fp32 heads and the final store
//   warp  8    TMA     : streams the packed weight image stage by stage with cp.async.bulk (UBLKCP)
//   warp  9    MMA     : single-thread tcgen05.mma issue (UTCHMMA), accumulators in TMEM
// Within a CTA the tile is processed pass by pass (bulk synchronous through two mbarriers); the co-resident CTA
// on the same SM fills the tensor pipe while this one runs its epilogue.
//
// Precision: all MMA operands are fp16 (11-bit significand, saturating conversion; same tensor rate as bf16 and 8x
// finer), accumulation is fp32 in TMEM; dx, ambient and sigma heads are evaluated in fp32.  With more than 10
// encoding octaves the deformation phase runs in split precision (fp16 hi + lo operands, ~fp32 accuracy) because
// the encoding of the warped point amplifies its error by 2^(L-1).
#include <stdlib.h>
#include <string.h>
#include "field_dev.cuh"


namespace {

// TRAIN: additionally records what the backward pass needs -- every layer's activated output (fp16 "activation tape",
// tile-major chunk images written with TMA bulk stores, layout in sahs_make_dims), the sign bits of the pre-activations ([layer][P][2] x 128 bit) and the warped
// point / ambient coordinates ([P,8] fp32).
struct TrainOut {
  __half* tape_x;
  uint4* masks;
  float* saves;
};

// SAHS_DBG_PROF_PROD (per-pass worker events from the production instantiation, true speed) costs the production kernel
// a null check per pass (~1 % measured), so it is compiled in only with -DSAHS_PROF_PROD=1
// (profiles/r1_pass_timeline_production_*.txt were taken with such a build).
#ifndef SAHS_PROF_PROD
#define SAHS_PROF_PROD 0
#endif
constexpr bool kProfProd = SAHS_PROF_PROD != 0;

// One 128-point tile, start to finish, as executed by the 256 worker threads of a tile set (`tid` = 0..255 inside the
// set, named barrier `bar_id` belongs to the set): encodings, epilogues of all passes, fp32 tails, embedding gather,
// final store.  Shared by the one-tile-per-CTA kernels below and by the two-tile ("duo") kernel.
template <class C, bool DBG, bool TRAIN, bool PAIR>
__device__ __forceinline__ void field_tile_program(const NetDims& dm, SyncT<PAIR>& sy, uint8_t* X, const float* fcw,
                                                   float* xchg, const float* __restrict__ grid,
                                                   const float* __restrict__ ro, const float* __restrict__ rd,
                                                   const float* __restrict__ zv, int S, long long P, long long ntiles,
                                                   float* __restrict__ raw_out, float* __restrict__ dbg, int dbg_pass,
                                                   TrainOut tr, long long tile, int tid, int bar_id, uint32_t tmem_row,
                                                   long long* prof_buf, int iter) {
  constexpr bool kTrunkF16 = TRAIN || kRenderTrunkF16;   // trunk / head operand format (field_plan.cuh)
  const int row = tid & (kTileRows - 1);
  const int grp = tid >> 7;
    if (DBG || (kProfProd && prof_buf)) {
      sy.prof = nullptr;
      if (prof_buf && iter == 2 && (tid == 0 || tid == 255)) {
        sy.prof = prof_buf + (tid == 0 ? 0 : 4096);
        prof_event(sy.prof, 2);   // tile start
      }
    }
    const long long p = tile * kTileRows + row;
    const bool valid = p < P;
    const long long pc = valid ? p : P - 1;
    const long long ray = pc / S;
    const float zz = zv[pc];
    const float dir[3] = {rd[ray * 3 + 0], rd[ray * 3 + 1], rd[ray * 3 + 2]};
    float pt[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) pt[k] = __fadd_rn(ro[ray * 3 + k], __fmul_rn(dir[k], zz));
    float* dbg_row = (DBG && dbg && tile == 0) ? dbg + row * 256 : nullptr;
    // training: activation tape of this tile, tile-major chunk images (sahs_make_dims); operand chunks go there
    // straight from shared memory with TMA bulk stores
    uint8_t* tape_tile = TRAIN ? reinterpret_cast<uint8_t*>(tr.tape_x) + (size_t)tile * (dm.tx_total / 64) * kChunkBytes
                               : nullptr;
    // X chunks [chunk0, chunk0 + nch) -> tape columns [col, col + 64 nch).  Every worker has fenced its writes
    // (signal_a or an explicit fence.proxy.async) before calling.
    auto tape_put = [&](int chunk0, int nch, int col) {
      if (!TRAIN) return;
      group_sync(bar_id);
      if (tid == 0 && tile < ntiles) {   // (a pair's second CTA may own a tile past the end: nothing to store)
        for (int i = 0; i < nch; ++i)
          tma_bulk_s2g(tape_tile + (size_t)(col / 64 + i) * kChunkBytes, X + (chunk0 + i) * kChunkBytes, kChunkBytes);
        tma_store_commit();
      }
    };
    // before X is overwritten: the bulk stores issued so far have finished reading shared memory
    auto tape_drain = [&]() {
      if (!TRAIN) return;
      if (tid == 0) tma_store_wait_read();
      group_sync(bar_id);
    };
    tape_drain();
    auto mask_slot = [&](int layer) -> uint4* {
      return (TRAIN && valid) ? tr.masks + ((size_t)layer * P + p) * 2 + grp : nullptr;
    };
    float mapped[3] = {pt[0], pt[1], pt[2]};
    float amb[C::AMB_DIM > 0 ? C::AMB_DIM : 1] = {};

    if (C::USE_W && dm.w_split) {
      // -------- deformation phase in split precision: warp net, then hyper-sheet net --------
      auto write_e0s = [&]() {
        RowStreamSplit<C::E0_PAD / 8> st(X, 0, 2, row, grp);
        int col = pe_stream<C::XYZ_L, true, 3>(st, 0, pt);
#pragma unroll
        for (int i = C::E0_DIM; i < C::E0_PAD; ++i) st.put(col++, 0.f);
      };
      const float* wf = fcw + dm.off_wfinal;
      const float* bf = wf + 3 * dm.wh;
      const float* wa = bf + 4;
      const float* ba = wa + C::AMB_DIM * dm.hh;
      float* scratch = reinterpret_cast<float*>(X);   // [2][128][8]
#pragma unroll 1
      for (int net = 0; net < 2; ++net) {
        const int boff = net == 0 ? 0 : dm.wh;
        write_e0s();
        signal_a(sy);
        for (int i = 0; i < dm.w_layers; ++i) {
          wait_acc(sy, 1000 + 16 * net + i);
          if (i == dm.w_skip) {
            write_e0s();
            signal_a(sy);
            wait_acc(sy, 1100 + 16 * net + i);
          }
          const float* bias = fcw + dm.off_wbias + i * dm.whh + boff;
          float* dr = (DBG && dbg_row && dbg_pass == SAHS_DBG_WARP(i)) ? dbg_row : nullptr;
          if (i < dm.w_layers - 1) {
            if (net == 0) epilogue_split<DBG, 4, PAIR>(tmem_row, grp * 64, bias, X, row, 2, dr, 0);
            else epilogue_split<DBG, 2, PAIR>(tmem_row, grp * 32, bias, X, row, 1, dr, dm.wh);
            signal_a(sy);
          } else if (net == 0) {
            float part[3] = {0.f, 0.f, 0.f};
            final_partial<DBG, 4, 3, PAIR>(tmem_row, grp * 64, bias, wf, dm.wh, part, dr, 0);
#pragma unroll
            for (int k = 0; k < 3; ++k) scratch[(grp * 128 + row) * 8 + k] = part[k];
            group_sync(bar_id);
#pragma unroll
            for (int k = 0; k < 3; ++k)
              mapped[k] = pt[k] + tanhf(scratch[row * 8 + k] + scratch[(128 + row) * 8 + k] + ldc1<PAIR>(bf + k));
            group_sync(bar_id);
          } else {
            float part[C::AMB_DIM > 0 ? C::AMB_DIM : 1] = {};
            final_partial<DBG, 2, (C::AMB_DIM > 0 ? C::AMB_DIM : 1), PAIR>(tmem_row, grp * 32, bias, wa, dm.hh, part, dr, dm.wh);
#pragma unroll
            for (int k = 0; k < C::AMB_DIM; ++k) scratch[(grp * 128 + row) * 8 + k] = part[k];
            group_sync(bar_id);
#pragma unroll
            for (int k = 0; k < C::AMB_DIM; ++k)
              amb[k] = scratch[row * 8 + k] + scratch[(128 + row) * 8 + k] + ldc1<PAIR>(ba + k);
            group_sync(bar_id);
          }
        }
      }
    } else if (C::USE_W) {
      // -------- deformation phase: warp | hyper-sheet merged (fp16 operands) --------
      auto write_e0 = [&](int chunk0) {
        RowStream<true, C::E0_PAD / 8> st(X, chunk0, row, grp);
        int col = pe_stream<C::XYZ_L, true, 3>(st, 0, pt);
#pragma unroll
        for (int i = C::E0_DIM; i < C::E0_PAD; ++i) st.put(col++, 0.f);
      };
      write_e0(dm.e0_chunk_base);
      signal_a(sy);
      tape_put(dm.e0_chunk_base, dm.e0_chunks, dm.tx_e0);
      for (int i = 0; i < dm.w_layers - 1; ++i) {
        const float* bias = fcw + dm.off_wbias + i * dm.whh;
        const bool two_pass = (i == dm.w_skip && !dm.e0_resident);
        float4 b[4];
        if (!two_pass) load_bias<PAIR>(b, bias + grp * 96);
        wait_acc(sy, 1000 + i);
        tape_drain();
        if (two_pass) {
          write_e0(0);
          signal_a(sy);
          load_bias<PAIR>(b, bias + grp * 96);
          wait_acc(sy, 1100 + i);
        }
        // whh = 192: three 32-column blocks per group
        epilogue<ACT_RELU, true, false, DBG, 6, TRAIN, PAIR>(tmem_row, grp * 96, bias, b, X, row, nullptr,
                                                       (DBG && dbg_row && dbg_pass == SAHS_DBG_WARP(i)) ? dbg_row : nullptr,
                                                       mask_slot(i));
        signal_a(sy);
        tape_put(0, dm.whh / 64, dm.tx_wh + i * dm.whh);
      }
      {
        // last hidden layer stays in fp32: dx = tanh(fc_final h_w), ambient = fc_ambient h_h.  Group g reduces its
        // 96 columns; the 5 partial sums per row are exchanged through the (idle) X buffer.
        const int i = dm.w_layers - 1;
        wait_acc(sy, 1000 + i);
        tape_drain();
        if (DBG) prof_event(sy.prof, 201);
        if (i == dm.w_skip && !dm.e0_resident) { write_e0(0); signal_a(sy); wait_acc(sy, 1100 + i); }
        uint8_t* rowp5 = X + (row >> 3) * 1024 + (row & 7) * 128;
        const float* bias = fcw + dm.off_wbias + i * dm.whh;
        const float* wf = fcw + dm.off_wfinal;
        const float* bf = wf + 3 * dm.wh;          // [wf 3*wh | bf 4 | wa amb*hh | ba 4], all 16-byte aligned
        const float* wa = bf + 4;
        const float* ba = wa + C::AMB_DIM * dm.hh;
        float part[3 + (C::AMB_DIM > 0 ? C::AMB_DIM : 1)] = {};
        uint32_t hmask[3] = {0u, 0u, 0u};
#pragma unroll
        for (int blk = 0; blk < 3; ++blk) {
          const int c0 = grp * 96 + 32 * blk;
          uint32_t v[32];
          tmem_ld32(tmem_row + c0, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 bb = ldc4<PAIR>(bias + c0 + 4 * j);
            float h[4] = {fmaxf(__uint_as_float(v[4 * j + 0]) + bb.x, 0.f), fmaxf(__uint_as_float(v[4 * j + 1]) + bb.y, 0.f),
                          fmaxf(__uint_as_float(v[4 * j + 2]) + bb.z, 0.f), fmaxf(__uint_as_float(v[4 * j + 3]) + bb.w, 0.f)};
            if (TRAIN) {
#pragma unroll
              for (int q = 0; q < 4; ++q) hmask[blk] |= (h[q] > 0.f ? 1u : 0u) << (4 * j + q);
              // the tape needs this layer's output as an operand chunk too (X chunks 0-2 are free here)
              const int col = c0 + 4 * j;
              *reinterpret_cast<uint2*>(rowp5 + (col >> 6) * kChunkBytes + (((((col & 63) >> 3) ^ row) & 7) << 4) +
                                        (col & 7) * 2) = make_uint2(pack2<true>(h[0], h[1]), pack2<true>(h[2], h[3]));
            }
            if (DBG && dbg_row && dbg_pass == SAHS_DBG_WARP(i)) {
#pragma unroll
              for (int q = 0; q < 4; ++q) dbg_row[c0 + 4 * j + q] = h[q];
            }
            if (c0 < dm.wh) {
#pragma unroll
              for (int k = 0; k < 3; ++k) {
                const float4 w = ldc4<PAIR>(wf + k * dm.wh + c0 + 4 * j);
                part[k] += h[0] * w.x + h[1] * w.y + h[2] * w.z + h[3] * w.w;
              }
            } else {
#pragma unroll
              for (int k = 0; k < C::AMB_DIM; ++k) {
                const float4 w = ldc4<PAIR>(wa + k * dm.hh + (c0 + 4 * j - dm.wh));
                part[3 + k] += h[0] * w.x + h[1] * w.y + h[2] * w.z + h[3] * w.w;
              }
            }
          }
        }
        if (TRAIN) {
          uint4* ms = mask_slot(i);
          if (ms) *ms = make_uint4(hmask[0], hmask[1], hmask[2], 0u);
        }
        if (DBG) prof_event(sy.prof, 202);   // fp32 last deformation layer done
        if (TRAIN) {
          fence_proxy_async_smem();
          tape_put(0, dm.whh / 64, dm.tx_wh + i * dm.whh);
        }
        // [2][128][8] floats; training keeps chunks 0-2 for the tape store above and uses the (now idle) chunk 3
        float* scratch = reinterpret_cast<float*>(X + (TRAIN ? 3 * kChunkBytes : 0));
#pragma unroll
        for (int k = 0; k < 3 + C::AMB_DIM; ++k) scratch[(grp * 128 + row) * 8 + k] = part[k];
        group_sync(bar_id);
#pragma unroll
        for (int k = 0; k < 3; ++k)
          mapped[k] = pt[k] + tanhf(scratch[row * 8 + k] + scratch[(128 + row) * 8 + k] + ldc1<PAIR>(bf + k));
#pragma unroll
        for (int k = 0; k < C::AMB_DIM; ++k)
          amb[k] = scratch[row * 8 + 3 + k] + scratch[(128 + row) * 8 + 3 + k] + ldc1<PAIR>(ba + k);
        group_sync(bar_id);   // scratch is dead before E1 overwrites it
        tape_drain();
        if (DBG) prof_event(sy.prof, 203);   // tanh / ambient exchange done
      }
    }
    if (TRAIN && valid && grp == 0) {
      float* sv = tr.saves + p * 8;
      sv[0] = mapped[0]; sv[1] = mapped[1]; sv[2] = mapped[2];
#pragma unroll
      for (int k = 0; k < C::AMB_DIM; ++k) sv[3 + k] = amb[k];
    }
    // -------- spatial embedding gather: each group keeps its 16 channels packed until layers_dir.0 --------
    uint32_t emb_pk[8];
    {
      float emb[16];
      grid_gather16(grid, grp * 16, mapped[0], mapped[1], mapped[2], emb);
#pragma unroll
      for (int q = 0; q < 8; ++q) emb_pk[q] = pack2<kTrunkF16>(emb[2 * q], emb[2 * q + 1]);
      if (DBG && dbg_row && dbg_pass == SAHS_DBG_MAPPED) {
#pragma unroll
        for (int q = 0; q < 16; ++q) dbg_row[8 + grp * 16 + q] = emb[q];
        if (grp == 0) {
          dbg_row[0] = mapped[0]; dbg_row[1] = mapped[1]; dbg_row[2] = mapped[2];
#pragma unroll
          for (int k = 0; k < C::AMB_DIM; ++k) dbg_row[3 + k] = amb[k];
        }
      }
    }
    if (DBG) prof_event(sy.prof, 204);   // embedding gather done
    // -------- trunk (fp16 operands; bf16 in a -DSAHS_RENDER_BF16 build: kTrunkF16) --------
    auto write_e1 = [&]() {
      RowStream<kTrunkF16, C::E1_PAD / 8> st(X, 0, row, grp);
      int col = pe_stream<C::XYZ_L, true, 3>(st, 0, mapped);
      if (C::AMB_PE > 0) col = pe_stream<C::AMB_L, C::AMB_INC, (C::AMB_DIM > 0 ? C::AMB_DIM : 1)>(st, col, amb);
#pragma unroll
      for (int i = C::E1_DIM; i < C::E1_PAD; ++i) st.put(col++, 0.f);
    };
    write_e1();
    if (DBG) prof_event(sy.prof, 205);   // E1 written
    signal_a(sy);
    tape_put(0, dm.e1_chunks, dm.tx_e1);
    for (int i = 0; i < dm.t_layers; ++i) {
      const float* bias = fcw + dm.off_tbias + i * dm.th;
      float4 b[4];
      if (i != dm.t_skip) load_bias<PAIR>(b, bias + grp * 128);
      wait_acc(sy, 2000 + i);
      tape_drain();
      if (i == dm.t_skip) {
        write_e1();
        signal_a(sy);
        load_bias<PAIR>(b, bias + grp * 128);
        wait_acc(sy, 2100 + i);
      }
      epilogue<ACT_LEAKY, kTrunkF16, false, DBG, 8, TRAIN, PAIR>(
          tmem_row, grp * 128, bias, b, X, row, nullptr,
          (DBG && dbg_row && dbg_pass == SAHS_DBG_TRUNK(i)) ? dbg_row : nullptr,
          mask_slot((C::USE_W ? dm.w_layers : 0) + i));
      signal_a(sy);
      tape_put(0, dm.th / 64, dm.tx_th + i * dm.th);
    }
    // fc_feat (no activation) + sigma = fc_alpha(feat) in fp32 (partial dot per group)
    float4 bfe[4];
    load_bias<PAIR>(bfe, fcw + dm.off_featb + grp * 128);
    wait_acc(sy, 2200);
    tape_drain();
    float sigma = epilogue<ACT_NONE, kTrunkF16, true, DBG, 8, TRAIN, PAIR>(
        tmem_row, grp * 128, fcw + dm.off_featb, bfe, X, row, fcw + dm.off_alpha,
        (DBG && dbg_row && dbg_pass == SAHS_DBG_TRUNK(dm.t_layers)) ? dbg_row : nullptr,
        nullptr);
    if (grp == 1) xchg[row] = sigma;
    signal_a(sy);
    tape_put(0, dm.th / 64, dm.tx_feat);
    // -------- heads: layers_dir.0 = [feat | PE(dir) | emb], layers_seg.0 = feat --------
    wait_acc(sy, 3000);
    tape_drain();
    {
      // extra K-chunk: cols [0,27) PE(dir), [27,59) embedding, zero padding up to 64
      RowStream<kTrunkF16, 8, false> st(X, 0, row, 0);
      if (grp == 0) {
        int col = pe_stream<C::DIR_L, true, 3>(st, 0, dir);
#pragma unroll
        for (int i = C::DIR_DIM; i < 64; ++i) st.put(col++, 0.f);
      }
      group_sync(bar_id);     // group 0's zero fill precedes the scalar embedding stores of both groups
      uint8_t* rowp = X + (row >> 3) * 1024 + (row & 7) * 128;
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        const int col = C::DIR_DIM + grp * 16 + q;
        const uint32_t pair = emb_pk[q >> 1];
        const uint16_t hv = (q & 1) ? (uint16_t)(pair >> 16) : (uint16_t)(pair & 0xffffu);
        *reinterpret_cast<uint16_t*>(rowp + ((((col >> 3) ^ row) & 7) << 4) + (col & 7) * 2) = hv;
      }
    }
    signal_a(sy);
    tape_put(0, 1, dm.tx_xtra);
    for (int i = 0; i < 4; ++i) {
      const float* bias = fcw + dm.off_hbias + i * 2 * dm.hd;
      float4 b[4];
      load_bias<PAIR>(b, bias + grp * 128);
      wait_acc(sy, 3100 + i);
      tape_drain();
      epilogue<ACT_LEAKY, kTrunkF16, false, DBG, 8, TRAIN, PAIR>(
          tmem_row, grp * 128, bias, b, X, row, nullptr,
          (DBG && dbg_row && dbg_pass == SAHS_DBG_HEAD(i)) ? dbg_row : nullptr,
          mask_slot((C::USE_W ? dm.w_layers : 0) + dm.t_layers + i));
      signal_a(sy);
      tape_put(0, 2 * dm.hd / 64, dm.tx_hh + i * 2 * dm.hd);
    }
    // -------- output layer: cols 0-2 rgb, 3-14 seg (+ sigma) --------
    wait_acc(sy, 3200);
    if (grp == 0) {
      uint32_t v[16];
      tmem_ld16(tmem_row, v);
      tmem_ld_wait();
      float o[16];
#pragma unroll
      for (int k = 0; k < 15; ++k) o[k] = __uint_as_float(v[k]) + ldc1<PAIR>(fcw + dm.off_outb + k);
      o[15] = sigma + xchg[row] + ldc1<PAIR>(fcw + dm.off_alpha + dm.th);
      if (valid) {
        float4* dst = reinterpret_cast<float4*>(raw_out + p * SAHS_RAW_CH);
#pragma unroll
        for (int q = 0; q < 4; ++q) dst[q] = make_float4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
      }
    }
    tc_fence_before();
    group_sync(bar_id);   // xchg[] consumed before the next tile's fc_feat epilogue rewrites it
    if (DBG || kProfProd) prof_event(sy.prof, 3);   // tile end
}

template <class C, bool DBG, bool TRAIN, bool PAIR>
__global__ void __launch_bounds__(kThreads, 2)
field_fwd_kernel(const __grid_constant__ FieldPlan plan, const __grid_constant__ NetDims dm,
                 const uint8_t* __restrict__ packed, const float* __restrict__ fc, const float* __restrict__ grid,
                 const float* __restrict__ ro, const float* __restrict__ rd, const float* __restrict__ zv,
                 int S, long long P, float* __restrict__ raw_out, float* __restrict__ dbg, int dbg_pass,
                 int* __restrict__ status, TrainOut tr) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* X = smem;
  uint8_t* slots = smem + kSmemX;
  // PAIR: [X 64 KB | 3 x 8 KB slots | frame constants 22 KB | barriers | xchg]; else [X | 3 x 16 KB slots | barriers | xchg]
  constexpr int kSlotsBytes = PAIR ? kPairSlots * kPairSlotBytes : kSmemSlots;
  constexpr int kFcBytes = PAIR ? kPairFcFloats * 4 : 0;
  float* fc_s = reinterpret_cast<float*>(smem + kSmemX + kSlotsBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kSmemX + kSlotsBytes + kFcBytes);
  constexpr int NS = PAIR ? kPairSlots : kSlots;
  uint64_t* full = bars;               // [NS]
  uint64_t* empty = bars + NS;         // [NS]
  uint64_t* a_ready = bars + 2 * NS;
  uint64_t* acc_ready = bars + 2 * NS + 1;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 3 * NS + 2);
  uint32_t* peer_tmem = tmem_ptr + 1;  // PAIR: the peer reports its TMEM base here (must equal the leader's)
  float* xchg = reinterpret_cast<float*>(smem + kSmemX + kSlotsBytes + kFcBytes + 256);

  // warp index through a shuffle: the compiler then knows it is warp-uniform, keeps the role dispatch convergent and
  // the MMA issue loop on the uniform datapath (with the plain threadIdx.x >> 5 every tcgen05 instruction below gets a
  // divergence guard + R2UR moves with scoreboard waits, ~4x slower issue)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const long long ntiles = (P + kTileRows - 1) / kTileRows;
  const long long npairs = (ntiles + 1) / 2;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;

  if (threadIdx.x == 0) {
    if (smem_u32(smem) & 1023u) {  // UMMA 128B swizzle needs 1024-byte aligned tiles
      status[0] = 2;
      __trap();
    }
    // PAIR leader: a stage is complete when its own half (expect_tx arrive) and the peer's half (relayed arrive) landed
    for (int i = 0; i < NS; ++i) { mbar_init(&full[i], (PAIR && rank == 0) ? 2 : 1); mbar_init(&empty[i], 1); }
    if (PAIR) {
      mbar_init(a_ready, 2 * (kWorkerThreads / 32));   // one arrival per worker warp of both CTAs
    } else {
      mbar_init(a_ready, kWorkerThreads);
    }
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
    // barriers of both CTAs are initialised before anything arrives remotely; the pair's accumulators must sit at
    // the same TMEM columns in both SMs
    if (rank == 1 && threadIdx.x == 0) st_cluster_u32(mapa_u32(peer_tmem, 0), tmem_base);
    cluster_sync_all();
    if (rank == 0 && threadIdx.x == 0 && *peer_tmem != tmem_base) {
      status[0] = 3;
      status[1] = (int)tmem_base;
      status[2] = (int)*peer_tmem;
      __threadfence_system();
      __trap();
    }
    __syncthreads();   // reconverge after the single-thread check (keeps the role loops below warp-uniform)
  }

  // debug builds, dbg_pass == SAHS_DBG_PROF: block 0 records (tag, clock64) event pairs of its third tile into dbg,
  // 4096 long longs per role: worker thread 0, worker thread 255, MMA issuer, TMA producer
  // (SAHS_DBG_PROF_PROD: the production instantiation records the workers' per-pass events -- signal and accumulator
  // wake-up, already compiled into signal_a / wait_acc behind a null check -- so a tile can be timed at true speed)
  const bool prof_on = dbg && blockIdx.x == 0 &&
                       ((DBG && (dbg_pass == SAHS_DBG_PROF || dbg_pass == SAHS_DBG_PROF_LIGHT)) ||
                        (kProfProd && !DBG && dbg_pass == SAHS_DBG_PROF_PROD));
  long long* prof_buf = prof_on ? reinterpret_cast<long long*>(dbg) : nullptr;
  // SAHS_DBG_PROF_LIGHT: bit 0 of the pointers handed to the TMA / MMA loops asks for per-pass events only
  const uintptr_t prof_light = (dbg_pass == SAHS_DBG_PROF_LIGHT) ? 1u : 0u;
  if (warp == kTmaWarp) {
    long long* pb = prof_buf ? reinterpret_cast<long long*>(reinterpret_cast<uintptr_t>(prof_buf + 3 * 4096) | prof_light) : nullptr;
    if (PAIR) tma_warp_loop_pair(plan, packed, slots, full, empty, npairs, rank, status, lane, pb);
    else tma_warp_loop(plan, packed, slots, full, empty, ntiles, status, lane, pb);
  } else if (warp == kMmaWarp) {
    long long* pb = prof_buf ? reinterpret_cast<long long*>(reinterpret_cast<uintptr_t>(prof_buf + 2 * 4096) | prof_light) : nullptr;
    if (!PAIR) mma_warp_loop<DBG>(plan, X, slots, full, empty, a_ready, acc_ready, tmem_base, ntiles, status, lane, pb);
    else if (rank == 0) mma_warp_loop_pair<DBG>(plan, X, slots, full, empty, a_ready, acc_ready, tmem_base, npairs, status, lane, pb);
    else relay_warp_loop_pair(plan, full, npairs, status, lane);
  } else {
    // ================================ workers ===================================================
    const float* fcw = PAIR ? fc_s : fc;   // frame constants: shared-memory copy in the pair kernel
    SyncT<PAIR> sy{a_ready, acc_ready, 0u, status, PAIR ? mapa_u32(a_ready, 0) : 0u};
    const uint32_t tmem_row = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    // PAIR: `it` walks tile pairs and this CTA owns tile 2*it + rank (possibly past the end: all rows masked)
    const long long it_end = PAIR ? npairs : ntiles;
    const long long it_step = PAIR ? (long long)cluster_num_x() : (long long)gridDim.x;
    int iter = 0;
    const long long it0 = PAIR ? (long long)cluster_id_x() : (long long)blockIdx.x;
    for (long long it = it0; it < it_end; it += it_step, ++iter) {
      const long long tile = PAIR ? 2 * it + rank : it;
      field_tile_program<C, DBG, TRAIN, PAIR>(dm, sy, X, fcw, xchg, grid, ro, rd, zv, S, P, ntiles, raw_out, dbg, dbg_pass,
                                              tr, tile, (int)threadIdx.x, 1, tmem_row, prof_buf, iter);
    }
    if (TRAIN && threadIdx.x == 0) tma_store_wait_all();   // tape stores complete before the CTA exits
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


// ---- two-tile ("duo") render kernel: one cluster per SM pair, two tile pairs in flight, one in-order issuer --------
// (field_dev.cuh "duo variants").  Render path only: no tapes, no debug outputs.
template <class C, bool PROF>
__global__ void __launch_bounds__(kDuoThreads, 1)
field_fwd_duo_kernel(const __grid_constant__ FieldPlan plan, const __grid_constant__ NetDims dm,
                     const uint8_t* __restrict__ packed, const float* __restrict__ fc, const float* __restrict__ grid,
                     const float* __restrict__ ro, const float* __restrict__ rd, const float* __restrict__ zv,
                     int S, long long P, float* __restrict__ raw_out, int* __restrict__ status, long long* prof, int one_set) {
  extern __shared__ __align__(1024) uint8_t smem[];
  float* fc_s = reinterpret_cast<float*>(smem + kDuoOffFc);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kDuoOffBars);      // [set][full NS | empty NS | a_ready | acc_ready]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * kDuoBarsPerSet);
  uint32_t* peer_tmem = tmem_ptr + 1;
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const long long ntiles = (P + kTileRows - 1) / kTileRows;
  const long long npairs = (ntiles + 1) / 2;
  const uint32_t rank = cluster_ctarank();

  if (threadIdx.x == 0) {
    if (smem_u32(smem) & 1023u) {
      status[0] = 2;
      __trap();
    }
    for (int k = 0; k < 2; ++k) {
      uint64_t* b = bars + k * kDuoBarsPerSet;
      for (int i = 0; i < kDuoSlots; ++i) { mbar_init(&b[i], rank == 0 ? 2 : 1); mbar_init(&b[kDuoSlots + i], 1); }
      mbar_init(&b[2 * kDuoSlots], 2 * (kWorkerThreads / 32));   // a_ready: one arrival per worker warp of both CTAs
      mbar_init(&b[2 * kDuoSlots + 1], 1);                        // acc_ready
    }
    fence_mbar_init();
  }
  for (int i = threadIdx.x; i < dm.fc_total; i += kDuoThreads) fc_s[i] = __ldg(fc + i);
  if (warp == kDuoMmaWarp) tmem_alloc_pair(tmem_ptr, kDuoTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  if (rank == 1 && threadIdx.x == 0) st_cluster_u32(mapa_u32(peer_tmem, 0), tmem_base);
  cluster_sync_all();
  // one_set (measurement switch SAHS_DUO_ONE_SET=1): set 1 gets no tiles, set 0 all of them -- a pass then runs without a
  // concurrent epilogue of the other set
  const long long cid = cluster_id_x(), ncl = cluster_num_x();
  const long long step = one_set ? ncl : 2 * ncl;
  const long long first[2] = {cid, one_set ? npairs : cid + ncl};
  if (warp < 16) {
    // ================================ workers of set `k` ======================================
    const int k = warp >> 3;
    uint64_t* b = bars + k * kDuoBarsPerSet;
    uint64_t* a_ready = &b[2 * kDuoSlots];
    SyncT<true> sy{a_ready, &b[2 * kDuoSlots + 1], 0u, status, mapa_u32(a_ready, 0)};
    uint8_t* X = smem + k * kSmemX;
    float* xchg = reinterpret_cast<float*>(smem + kDuoOffXchg + k * 512);
    const int tid = (int)threadIdx.x & 255;
    const uint32_t tmem_row = tmem_base + (uint32_t)(k * 256) + ((uint32_t)((warp & 3) * 32) << 16);
    int iter = 0;
    for (long long it = first[k]; it < npairs; it += step, ++iter) {
      const long long tile = 2 * it + rank;   // possibly one past the end: all rows masked
      // SAHS_DBG_PROF_DUO: thread 0 of each set of cluster 0's leader records its 4th tile (signal / accumulator events)
      if (PROF) {
        sy.prof = (prof && cid == 0 && rank == 0 && tid == 0 && iter == 3) ? prof + k * 4096 : nullptr;
        prof_event(sy.prof, 2);
      }
      field_tile_program<C, false, false, true>(dm, sy, X, fc_s, xchg, grid, ro, rd, zv, S, P, ntiles, raw_out, nullptr, -1,
                                                TrainOut{nullptr, nullptr, nullptr}, tile, tid, 1 + k, tmem_row, nullptr,
                                                iter);
      if (PROF) prof_event(sy.prof, 3);
    }
  } else if (warp < kDuoMmaWarp) {
    const int k = warp - kDuoTmaWarp0;
    uint64_t* b = bars + k * kDuoBarsPerSet;
    tma_warp_loop_duo(plan, packed, smem + kDuoOffSlots + k * kDuoRingBytes, b, b + kDuoSlots, first[k], npairs, step, rank,
                      status, lane);
  } else if (warp == kDuoMmaWarp) {
    if (rank == 0) {
      // the pair's accumulators must sit at the same TMEM columns in both SMs.  Checked here by the whole (converged)
      // issuer warp on a shuffled value: a single-thread check ahead of the role dispatch makes the compiler treat the
      // issue loop as possibly divergent (BRA.DIV + per-instruction election + R2UR: ~4x slower issue).
      const uint32_t peer_base = __shfl_sync(0xffffffffu, *peer_tmem, 0);
      if (peer_base != tmem_base) {
        status[0] = 3;
        status[1] = (int)tmem_base;
        status[2] = (int)peer_base;
        __threadfence_system();
        __trap();
      }
      mma_warp_loop_duo<PROF>(plan, smem, smem + kSmemX, smem + kDuoOffSlots, smem + kDuoOffSlots + kDuoRingBytes, bars,
                        bars + kDuoBarsPerSet, tmem_base, first[0], first[1], step, npairs, status,
                        (PROF && prof && cid == 0 && lane == 0)
                            ? reinterpret_cast<long long*>(reinterpret_cast<uintptr_t>(prof + 2 * 4096) | (uintptr_t)(P & 1))
                            : nullptr);   // (odd point count: per-stage issuer events as well)
    } else {
      relay_warp_loop_duo(plan, bars, first[0], npairs, step, status, lane);
    }
  } else if (rank == 1) {
    relay_warp_loop_duo(plan, bars + kDuoBarsPerSet, first[1], npairs, step, status, lane);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // the peer's smem / TMEM / barriers stay alive until both CTAs are done
  if (warp == kDuoMmaWarp) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, kDuoTmemCols);
  }
}

// The two-tile kernel is opt-in (SAHS_FIELD_DUO=1).  Measured on one box, same frame, fine-level launch (profiles/
// r2_field_variants_ab.txt): two CTA pairs per SM pair 63.9 ms, two-tile kernel 65.1 ms.  Serving the passes first come
// first served does remove the lock step, but the kernel is not bound by the issue order: with the mbarrier-wait
// latency hidden (early peek) a pass still runs at ~92 cycles per N = 128 MMA instead of 70, because shared-memory
// bandwidth is saturated -- MMA operand reads (~41 % of the data pipe at 45 % tensor-pipe activity) plus the epilogues'
// loads and stores (40 %, ncu l1tex__data_pipe_lsu_wavefronts_mem_shared) -- and the GPU sits at its power cap
// (DESIGN.md section 5, round 2).
static bool field_duo_enabled() {
  const char* e = getenv("SAHS_FIELD_DUO");
  return e && e[0] == '1';
}

// The CTA-pair (cta_group::2) kernel is the default render path; SAHS_FIELD_PAIR=0 selects the single-CTA kernel
// (A/B measurements; also used for the per-pass debug outputs).
static bool field_pair_enabled() {
  const char* e = getenv("SAHS_FIELD_PAIR");
  return !(e && e[0] == '0');
}

template <class C>
int launch_field(const HostPlan& hp, const void* packed, const float* fc, const float* grid, const float* ro,
                 const float* rd, const float* z, int R, int S, float* raw, float* dbg, int dbg_pass,
                 cudaStream_t st, TrainOut tr = TrainOut{nullptr, nullptr, nullptr}) {
  int* status = sahs_status_words(0);
  SAHS_CHECK_ARG(status, "cannot allocate the diagnostic word");
  const long long P = (long long)R * S;
  const long long ntiles = (P + kTileRows - 1) / kTileRows;
  const uint8_t* pk = (const uint8_t*)packed;
  const bool prof_prod = dbg != nullptr && dbg_pass == SAHS_DBG_PROF_PROD;   // production kernel + per-pass events
  const bool prof = dbg != nullptr && (dbg_pass == SAHS_DBG_PROF || dbg_pass == SAHS_DBG_PROF_LIGHT);
  const bool prof_duo = dbg != nullptr && dbg_pass == SAHS_DBG_PROF_DUO;
  if ((!dbg || prof_duo) && !tr.tape_x && field_pair_enabled() && field_duo_enabled() && hp.dims.fc_total <= kPairFcFloats) {
    // render path: one cluster per SM pair, two tile pairs in flight per cluster
    auto kfn = prof_duo ? field_fwd_duo_kernel<C, true> : field_fwd_duo_kernel<C, false>;
    SAHS_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, kDuoSmemTotal));
    const long long npairs = (ntiles + 1) / 2;
    long long nclusters = sahs_num_sms() / 2;
    if (nclusters > (npairs + 1) / 2) nclusters = (npairs + 1) / 2;
    if (nclusters < 1) nclusters = 1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(2 * nclusters));
    cfg.blockDim = dim3(kDuoThreads);
    cfg.dynamicSmemBytes = kDuoSmemTotal;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    long long* prof_ptr = prof_duo ? reinterpret_cast<long long*>(dbg) : nullptr;
    static const int one_set = [] { const char* e = getenv("SAHS_DUO_ONE_SET"); return (e && e[0] == '1') ? 1 : 0; }();
    SAHS_CUDA(cudaLaunchKernelEx(&cfg, kfn, hp.plan, hp.dims, pk, fc, grid, ro, rd, z, S, P, raw, status, prof_ptr, one_set));
    SAHS_LAUNCH_CHECK();
    return SAHS_OK;
  }
  if ((!dbg || prof || prof_prod) && field_pair_enabled() && hp.dims.fc_total <= kPairFcFloats) {
    // CTA pairs: clusters of 2, two clusters resident per SM pair
    auto kfn = tr.tape_x ? field_fwd_kernel<C, false, true, true>
                         : (prof ? field_fwd_kernel<C, true, false, true> : field_fwd_kernel<C, false, false, true>);
    // measurement switch: SAHS_FIELD_ONE_PER_SM=1 pads the shared-memory request so that only one CTA fits per SM
    // (one cluster per SM pair): shows a CTA's MMA phases without a co-resident CTA sharing the tensor pipe
    static const bool one_per_sm = [] { const char* e = getenv("SAHS_FIELD_ONE_PER_SM"); return e && e[0] == '1'; }();
    const int smem_bytes = one_per_sm ? 160 * 1024 : kPairSmemTotal;
    SAHS_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    const long long npairs = (ntiles + 1) / 2;
    long long nclusters = one_per_sm ? sahs_num_sms() / 2 : sahs_num_sms();   // 2 CTAs per SM = one cluster per SM
    if (nclusters > npairs) nclusters = npairs;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(2 * nclusters));
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    SAHS_CUDA(cudaLaunchKernelEx(&cfg, kfn, hp.plan, hp.dims, pk, fc, grid, ro, rd, z, S, P, raw, dbg, dbg_pass,
                                 status, tr));
    SAHS_LAUNCH_CHECK();
    return SAHS_OK;
  }
  auto kfn = tr.tape_x ? field_fwd_kernel<C, false, true, false>
                       : ((dbg != nullptr && !prof_prod) ? field_fwd_kernel<C, true, false, false>
                                           : field_fwd_kernel<C, false, false, false>);
  SAHS_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal));
  long long grid_dim = 2LL * sahs_num_sms();
  if (grid_dim > ntiles) grid_dim = ntiles;
  kfn<<<(unsigned)grid_dim, kThreads, kSmemTotal, st>>>(hp.plan, hp.dims, pk, fc, grid, ro, rd, z, S,
                                                       P, raw, dbg, dbg_pass, status, tr);
  SAHS_LAUNCH_CHECK();
  return SAHS_OK;
}

}  // namespace

static int field_fwd_impl(const sahs_model_spec* spec, int level, const void* packed, const float* frame_const,
                          const float* grid_cl, const float* ro, const float* rd, const float* z, int num_rays,
                          int num_samples, float* raw_out, float* debug, int debug_pass, void* stream, bool train,
                          TrainOut tr) {
  SAHS_CHECK_ARG(spec, "null spec");
  SAHS_CHECK_ARG(level == 0 || level == 1, "level must be 0 (coarse) or 1 (fine)");
  SAHS_CHECK_ARG(num_rays >= 0 && num_samples > 0, "bad extents");
  if (num_rays == 0) return SAHS_OK;
  SAHS_CHECK_ARG(packed && frame_const && ro && rd && z && raw_out, "null pointer");
  SAHS_CHECK_ARG(spec->use_grid && grid_cl, "use_spatial_embeddings and its grid are required (all shipped configs)");
  static thread_local HostPlan hp;
  int rc = sahs_build_host_plan(*spec, nullptr, hp, train);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const sahs_model_spec& s = *spec;
  SAHS_CHECK_ARG(s.xyz_inc && s.dir_inc && s.dir_L == 4, "include_input_xyz/dir and num_encoding_fn_dir=4 expected");
  SAHS_CHECK_ARG(!hp.dims.use_w || hp.dims.whh == 192, "warp 128 + hyper 64 hidden units expected");
#define SAHS_TRY(XL, AD, AL, AI, UW)                                                                             \
  if (s.xyz_L == XL && (UW ? (s.amb_dim == AD && s.amb_L == AL && (s.amb_inc != 0) == AI) : true) &&             \
      ((s.use_warp != 0) == UW))                                                                                 \
    return launch_field<FieldCfg<XL, AD, AL, AI, 4, UW>>(hp, packed, frame_const, grid_cl, ro, rd, z, num_rays,  \
                                                         num_samples, raw_out, debug, debug_pass, st, tr);
  SAHS_TRY(10, 2, 4, true, true)     // config/audio/*.yml
  SAHS_TRY(15, 1, 15, false, true)   // config/expression/person_{2,3}.yml
  SAHS_TRY(10, 0, 0, false, false)   // config/expression/person_1.yml (no deformation, no hyper space)
#undef SAHS_TRY
  sahs_set_error("sahs_field_fwd: no kernel instantiated for this model spec");
  return SAHS_EUNSUPPORTED;
}

extern "C" int sahs_field_fwd(const sahs_model_spec* spec, int level, const void* packed, const float* frame_const,
                              const float* grid_cl, const float* ro, const float* rd, const float* z, int num_rays,
                              int num_samples, float* raw_out, float* debug, int debug_pass, void* stream) {
  return field_fwd_impl(spec, level, packed, frame_const, grid_cl, ro, rd, z, num_rays, num_samples, raw_out, debug,
                        debug_pass, stream, false, TrainOut{nullptr, nullptr, nullptr});
}

extern "C" int sahs_field_fwd_train(const sahs_model_spec* spec, int level, const void* packed_train,
                                    const float* frame_const, const float* grid_cl, const float* ro, const float* rd,
                                    const float* z, int num_rays, int num_samples, float* raw_out, void* tape_x,
                                    void* masks, float* saves, void* stream) {
  SAHS_CHECK_ARG(num_rays == 0 || (tape_x && masks && saves), "training buffers required");
  return field_fwd_impl(spec, level, packed_train, frame_const, grid_cl, ro, rd, z, num_rays, num_samples, raw_out,
                        nullptr, -1, stream, true, TrainOut{(__half*)tape_x, (uint4*)masks, saves});
}

// Device-side building blocks shared by the fused field kernels (forward: field_fwd.cu, backward: field_bwd.cu).
#pragma once
#include <cuda_fp16.h>
#include "sahs_common.cuh"
#include "field_plan.cuh"

namespace {


constexpr int kWorkerThreads = 256;
constexpr int kThreads = 320;
constexpr int kTmaWarp = 8, kMmaWarp = 9;
constexpr int kSlots = 3;
constexpr int kTmemCols = 256;
constexpr int kSmemX = 4 * kChunkBytes;                       // 64 KB
constexpr int kSmemSlots = kSlots * kStageSlotBytes;          // 48 KB
constexpr int kSmemBar = kSmemX + kSmemSlots;                 // barriers after the tiles
constexpr int kSmemXchg = kSmemBar + 256;                     // 128 floats exchanged between the two groups
constexpr int kSmemTotal = kSmemXchg + 512;
// CTA-pair variant (cluster of 2, tcgen05 cta_group::2): every CTA holds half of each weight stage (N/2 rows), so the
// L2 -> SM weight traffic per point is halved and a 24 KB ring is as deep (3 stages) as the single-CTA 48 KB one ...
#ifndef SAHS_PAIR_SLOTS
#define SAHS_PAIR_SLOTS 3
#endif
constexpr int kPairSlots = SAHS_PAIR_SLOTS;
constexpr int kPairSlotBytes = kStageSlotBytes / 2;
// ... and the 24 KB this frees hold the per-frame constant block (folded biases, fp32 head weights) in shared memory:
// the epilogues' bias reads become LDS instead of global loads through the 28 KB L1 (measured: 18 % of the kernel).
constexpr int kPairFcFloats = 5632;
constexpr int kPairSmemTotal = kSmemX + kPairSlots * kPairSlotBytes + kPairFcFloats * 4 + 256 + 512;

enum { ACT_NONE = 0, ACT_RELU = 1, ACT_LEAKY = 2 };

template <bool PAIR>
struct SyncT {
  uint64_t* a_ready;
  uint64_t* acc_ready;
  uint32_t acc_par;
  int* status;
  uint32_t a_leader;   // PAIR: shared::cluster address of the leader CTA's a_ready barrier
  long long* prof = nullptr;   // debug builds: (tag, clock64) event pairs of one tile iteration
};
__device__ __forceinline__ void prof_event(long long*& prof, long long tag) {
  if (prof) {
    prof[0] = tag;
    prof[1] = clock64();
    prof += 2;
  }
}
using Sync = SyncT<false>;

// "my part of the A operand is written" -> MMA issuer.  Single CTA: every worker thread arrives.  Pair: every thread
// fences its own writes, one lane per warp arrives on the leader CTA's barrier (8 warps x 2 CTAs).
template <bool PAIR>
__device__ __forceinline__ void signal_a(SyncT<PAIR>& sy) {
  fence_proxy_async_smem();
  tc_fence_before();
  if (PAIR) {
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive_cluster(sy.a_leader);
  } else {
    mbar_arrive(sy.a_ready);
  }
  prof_event(sy.prof, 1);
}
template <bool PAIR>
__device__ __forceinline__ void wait_acc(SyncT<PAIR>& sy, int tag) {
  mbar_wait(sy.acc_ready, sy.acc_par, sy.status, tag);
  sy.acc_par ^= 1;
  tc_fence_after();
  prof_event(sy.prof, 100000 + tag);
}
__device__ __forceinline__ void group_sync(int bar_id = 1) {  // the 256 worker threads of one tile set only
  asm volatile("bar.sync %0, 256;" ::"r"(bar_id) : "memory");
}

// fp32 pair -> packed 16-bit pair.  fp16 conversions saturate to +-65504 instead of producing inf.
template <bool F16>
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  if (F16) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
  }
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

__device__ __forceinline__ void add2(float& a0, float& a1, float b0, float b1) {  // FADD2
  unsigned long long a, b, d;
  asm("mov.b64 %0, {%1,%2};" : "=l"(a) : "f"(a0), "f"(a1));
  asm("mov.b64 %0, {%1,%2};" : "=l"(b) : "f"(b0), "f"(b1));
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  asm("mov.b64 {%0,%1}, %2;" : "=f"(a0), "=f"(a1) : "l"(d));
}

// activation on a packed 16-bit pair (HMNMX2 / HFMA2): relu and leaky-relu(0.01)
template <int ACT, bool F16>
__device__ __forceinline__ uint32_t act2(uint32_t p) {
  if (ACT == ACT_NONE) return p;
  if (F16) {
    __half2 v = *reinterpret_cast<__half2*>(&p);
    __half2 r = (ACT == ACT_RELU) ? __hmax2(v, __float2half2_rn(0.f)) : __hmax2(v, __hmul2(v, __float2half2_rn(0.01f)));
    return *reinterpret_cast<uint32_t*>(&r);
  }
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&p);
  __nv_bfloat162 r = (ACT == ACT_RELU) ? __hmax2(v, __float2bfloat162_rn(0.f))
                                       : __hmax2(v, __hmul2(v, __float2bfloat162_rn(0.01f)));
  return *reinterpret_cast<uint32_t*>(&r);
}

// Streams consecutive feature columns of one tile row into X (16-bit, 128B-swizzled K-chunks), eight columns
// (one 16-byte unit) at a time.  Both worker groups generate every column; a group only stores the units it owns
// (first or second half), which keeps the generators free of cross-group exchanges.  All indices are compile-time
// after unrolling.
template <bool F16, int NUNITS, bool SPLIT = true>
struct RowStream {
  uint8_t* rowp;   // X + row offset inside a chunk
  int chunk0, row, grp;
  float buf[8];
  __device__ __forceinline__ RowStream(uint8_t* X, int chunk0_, int row_, int grp_)
      : rowp(X + (row_ >> 3) * 1024 + (row_ & 7) * 128), chunk0(chunk0_), row(row_), grp(grp_) {}
  __device__ __forceinline__ void put(int col, float v) {
    buf[col & 7] = v;
    if ((col & 7) == 7) {
      const int u = col >> 3;
      const int owner = (u < (NUNITS + 1) / 2) ? 0 : 1;
      if (!SPLIT || owner == grp) {
        uint4 q;
        q.x = pack2<F16>(buf[0], buf[1]);
        q.y = pack2<F16>(buf[2], buf[3]);
        q.z = pack2<F16>(buf[4], buf[5]);
        q.w = pack2<F16>(buf[6], buf[7]);
        *reinterpret_cast<uint4*>(rowp + (chunk0 + (u >> 3)) * kChunkBytes + ((((u & 7) ^ row) & 7) << 4)) = q;
      }
    }
  }
};

// positional encoding of D values streamed in the reference's column order (nerf_helpers.py:341-349):
// [x_0..x_{D-1}] (if INC), then per octave k: sin(2^k x_d) for all d, cos(2^k x_d) for all d.
// Accurate sincosf every 5th octave (2^k * x is exact in fp32), double-angle recurrence in between (err < 2e-6).
template <int L, bool INC, int D, class Stream>
__device__ __forceinline__ int pe_stream(Stream& st, int col, const float (&x)[D]) {
  if (INC) {
#pragma unroll
    for (int d = 0; d < D; ++d) st.put(col++, x[d]);
  }
  float s[D], c[D];
#pragma unroll
  for (int k = 0; k < L; ++k) {
#pragma unroll
    for (int d = 0; d < D; ++d) {
      if (k % 5 == 0) {
        sincosf(x[d] * (float)(1 << k), &s[d], &c[d]);
      } else {
        // explicit roundings (same reason as grid_gather16): sin 2a = (2 sin a) cos a, cos 2a = 1 - (2 sin a) sin a
        const float t = __fmul_rn(2.f, s[d]);
        const float s2 = __fmul_rn(t, c[d]);
        c[d] = __fmaf_rn(-t, s[d], 1.f);
        s[d] = s2;
      }
    }
#pragma unroll
    for (int d = 0; d < D; ++d) st.put(col++, s[d]);
#pragma unroll
    for (int d = 0; d < D; ++d) st.put(col++, c[d]);
  }
  return col;
}

// bias / small-weight loads keep their lines in the (28 KB) L1 ahead of the streaming traffic
__device__ __forceinline__ float4 ldg_keep(const float* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::evict_last.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ float ldg_keep1(const float* p) {
  float v;
  asm volatile("ld.global.nc.L1::evict_last.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ float4 ldg_stream(const float* p) {   // read-once data: do not allocate in L1
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
// frame-constant loads: FCS = the block lives in shared memory (CTA-pair kernel), else global through L1
template <bool FCS>
__device__ __forceinline__ float4 ldc4(const float* p) {
  if (FCS) return *reinterpret_cast<const float4*>(p);
  return ldg_keep(p);
}
template <bool FCS>
__device__ __forceinline__ float ldc1(const float* p) {
  if (FCS) return *p;
  return ldg_keep1(p);
}
template <bool FCS = false>
__device__ __forceinline__ void load_bias(float4 (&b)[4], const float* __restrict__ bias) {
#pragma unroll
  for (int j = 0; j < 4; ++j) b[j] = ldc4<FCS>(bias + 4 * j);
}

// one 16-column block of an epilogue: +bias (FADD2), optional fp32 dot, pack, activation, swizzled store.
// The bias registers are dead after the adds, so the next block's bias is fetched before the pack/store part.
template <int ACT, bool F16, bool DOT, bool DBG, bool TRAIN = false, bool FCS = false>
__device__ __forceinline__ void epi_block(uint32_t (&v)[16], float4 (&b)[4], const float* __restrict__ next_bias, int c0,
                                          uint8_t* rowp, int row, const float* __restrict__ dot_w, float& dot,
                                          float* dbg_row, uint32_t* mbits = nullptr) {
  float f[16];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    f[4 * j + 0] = __uint_as_float(v[4 * j + 0]); f[4 * j + 1] = __uint_as_float(v[4 * j + 1]);
    f[4 * j + 2] = __uint_as_float(v[4 * j + 2]); f[4 * j + 3] = __uint_as_float(v[4 * j + 3]);
    add2(f[4 * j + 0], f[4 * j + 1], b[j].x, b[j].y);
    add2(f[4 * j + 2], f[4 * j + 3], b[j].z, b[j].w);
  }
  if (next_bias) load_bias<FCS>(b, next_bias);
  if (TRAIN) {
    // training: sign bits of the pre-activations (the activated values reach the tape as whole operand chunks)
    uint32_t bits = 0;
#pragma unroll
    for (int j = 0; j < 16; ++j) bits |= (f[j] > 0.f ? 1u : 0u) << j;
    if (mbits) *mbits = bits;
  }
  uint32_t pk[8];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float f0 = f[4 * j], f1 = f[4 * j + 1], f2 = f[4 * j + 2], f3 = f[4 * j + 3];
    if (DOT) {
      const float4 w = ldc4<FCS>(dot_w + c0 + 4 * j);
      dot += f0 * w.x + f1 * w.y + f2 * w.z + f3 * w.w;
    }
    pk[2 * j] = act2<ACT, F16>(pack2<F16>(f0, f1));
    pk[2 * j + 1] = act2<ACT, F16>(pack2<F16>(f2, f3));
    if (DBG && dbg_row) {
      const float lk = ACT == ACT_LEAKY ? 0.01f : 0.f;
      const bool a = ACT != ACT_NONE;
      dbg_row[c0 + 4 * j + 0] = a ? fmaxf(f0, lk * f0) : f0;
      dbg_row[c0 + 4 * j + 1] = a ? fmaxf(f1, lk * f1) : f1;
      dbg_row[c0 + 4 * j + 2] = a ? fmaxf(f2, lk * f2) : f2;
      dbg_row[c0 + 4 * j + 3] = a ? fmaxf(f3, lk * f3) : f3;
    }
  }
  uint8_t* chunk = rowp + (c0 >> 6) * kChunkBytes;
  const int u0 = (c0 & 63) >> 3;
#pragma unroll
  for (int q = 0; q < 2; ++q)
    *reinterpret_cast<uint4*>(chunk + ((((u0 + q) ^ row) & 7) << 4)) =
        make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
}

// Epilogue of one pass for this group's NBLK 16-column blocks starting at column cbeg.  TMEM loads are double
// buffered against the math of the previous block; `b` arrives pre-loaded with the first block's bias (fetched by
// the caller before it waited for the accumulator).
template <int ACT, bool F16, bool DOT, bool DBG, int NBLK, bool TRAIN = false, bool FCS = false>
__device__ __forceinline__ float epilogue(uint32_t tmem_row, int cbeg, const float* __restrict__ bias, float4 (&b)[4],
                                          uint8_t* X, int row, const float* __restrict__ dot_w, float* dbg_row,
                                          uint4* mask_out = nullptr) {
  float dot = 0.f;
  uint32_t mw[4] = {0u, 0u, 0u, 0u};
  uint8_t* rowp = X + (row >> 3) * 1024 + (row & 7) * 128;
  uint32_t va[16], vb[16];
  tmem_ld16(tmem_row + cbeg, va);
#pragma unroll
  for (int blk = 0; blk < NBLK; ++blk) {
    const int c0 = cbeg + 16 * blk;
    const float* nb = (blk + 1 < NBLK) ? bias + c0 + 16 : nullptr;
    tmem_ld_wait();
    uint32_t bits = 0;
    if (blk & 1) {
      if (blk + 1 < NBLK) tmem_ld16_prefetch(tmem_row + c0 + 16, va, vb[0]);
      epi_block<ACT, F16, DOT, DBG, TRAIN, FCS>(vb, b, nb, c0, rowp, row, dot_w, dot, dbg_row, &bits);
    } else {
      if (blk + 1 < NBLK) tmem_ld16_prefetch(tmem_row + c0 + 16, vb, va[0]);
      epi_block<ACT, F16, DOT, DBG, TRAIN, FCS>(va, b, nb, c0, rowp, row, dot_w, dot, dbg_row, &bits);
    }
    if (TRAIN) mw[blk >> 1] |= bits << ((blk & 1) * 16);
  }
  if (TRAIN && mask_out) *mask_out = make_uint4(mw[0], mw[1], mw[2], mw[3]);
  return dot;
}

// ---- split-precision (fp16 hi + lo) variants used by the deformation phase when the encoding has > 10 octaves ----
// Streams a row as two fp16 planes: hi = fp16(v) into chunk0.., lo = fp16(v - hi) into chunk0 + lo_off..
template <int NUNITS>
struct RowStreamSplit {
  uint8_t* rowp;
  int chunk0, lo_off, row, grp;
  float buf[8];
  __device__ __forceinline__ RowStreamSplit(uint8_t* X, int chunk0_, int lo_off_, int row_, int grp_)
      : rowp(X + (row_ >> 3) * 1024 + (row_ & 7) * 128), chunk0(chunk0_), lo_off(lo_off_), row(row_), grp(grp_) {}
  __device__ __forceinline__ void put(int col, float v) {
    buf[col & 7] = v;
    if ((col & 7) == 7) {
      const int u = col >> 3;
      const int owner = (u < (NUNITS + 1) / 2) ? 0 : 1;
      if (owner == grp) {
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          __half2 h = __floats2half2_rn(buf[2 * q], buf[2 * q + 1]);
          const float2 hf = __half22float2(h);
          hi[q] = *reinterpret_cast<uint32_t*>(&h);
          lo[q] = pack2<true>(buf[2 * q] - hf.x, buf[2 * q + 1] - hf.y);
        }
        uint8_t* p = rowp + (chunk0 + (u >> 3)) * kChunkBytes + ((((u & 7) ^ row) & 7) << 4);
        *reinterpret_cast<uint4*>(p) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4*>(p + lo_off * kChunkBytes) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
      }
    }
  }
};

// relu epilogue writing hi/lo fp16 planes: columns [cbeg, cbeg + 16*NBLK) of the accumulator
template <bool DBG, int NBLK, bool FCS = false>
__device__ __forceinline__ void epilogue_split(uint32_t tmem_row, int cbeg, const float* __restrict__ bias, uint8_t* X,
                                               int row, int lo_off, float* dbg_row, int dbg_col0) {
  uint8_t* rowp = X + (row >> 3) * 1024 + (row & 7) * 128;
#pragma unroll
  for (int blk = 0; blk < NBLK; ++blk) {
    const int c0 = cbeg + 16 * blk;
    float4 b[4];
    load_bias<FCS>(b, bias + c0);
    uint32_t v[16];
    tmem_ld16(tmem_row + c0, v);
    tmem_ld_wait();
    uint32_t hi[8], lo[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float f0 = fmaxf(__uint_as_float(v[4 * j + 0]) + b[j].x, 0.f), f1 = fmaxf(__uint_as_float(v[4 * j + 1]) + b[j].y, 0.f);
      const float f2 = fmaxf(__uint_as_float(v[4 * j + 2]) + b[j].z, 0.f), f3 = fmaxf(__uint_as_float(v[4 * j + 3]) + b[j].w, 0.f);
      if (DBG && dbg_row) {
        dbg_row[dbg_col0 + c0 + 4 * j + 0] = f0; dbg_row[dbg_col0 + c0 + 4 * j + 1] = f1;
        dbg_row[dbg_col0 + c0 + 4 * j + 2] = f2; dbg_row[dbg_col0 + c0 + 4 * j + 3] = f3;
      }
      __half2 h0 = __floats2half2_rn(f0, f1), h1 = __floats2half2_rn(f2, f3);
      const float2 g0 = __half22float2(h0), g1 = __half22float2(h1);
      hi[2 * j] = *reinterpret_cast<uint32_t*>(&h0);
      hi[2 * j + 1] = *reinterpret_cast<uint32_t*>(&h1);
      lo[2 * j] = pack2<true>(f0 - g0.x, f1 - g0.y);
      lo[2 * j + 1] = pack2<true>(f2 - g1.x, f3 - g1.y);
    }
    uint8_t* chunk = rowp + (c0 >> 6) * kChunkBytes;
    const int u0 = (c0 & 63) >> 3;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      uint8_t* p = chunk + ((((u0 + q) ^ row) & 7) << 4);
      *reinterpret_cast<uint4*>(p) = make_uint4(hi[4 * q], hi[4 * q + 1], hi[4 * q + 2], hi[4 * q + 3]);
      *reinterpret_cast<uint4*>(p + lo_off * kChunkBytes) = make_uint4(lo[4 * q], lo[4 * q + 1], lo[4 * q + 2], lo[4 * q + 3]);
    }
  }
}

// fp32 reduction of the last hidden layer against NOUT small-head weight rows: columns [cbeg, cbeg+16*NBLK)
template <bool DBG, int NBLK, int NOUT, bool FCS = false>
__device__ __forceinline__ void final_partial(uint32_t tmem_row, int cbeg, const float* __restrict__ bias,
                                              const float* __restrict__ w, int ld, float (&part)[NOUT], float* dbg_row,
                                              int dbg_col0) {
#pragma unroll
  for (int blk = 0; blk < NBLK; ++blk) {
    const int c0 = cbeg + 16 * blk;
    uint32_t v[16];
    tmem_ld16(tmem_row + c0, v);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 bb = ldc4<FCS>(bias + c0 + 4 * j);
      const float h0 = fmaxf(__uint_as_float(v[4 * j + 0]) + bb.x, 0.f), h1 = fmaxf(__uint_as_float(v[4 * j + 1]) + bb.y, 0.f);
      const float h2 = fmaxf(__uint_as_float(v[4 * j + 2]) + bb.z, 0.f), h3 = fmaxf(__uint_as_float(v[4 * j + 3]) + bb.w, 0.f);
      if (DBG && dbg_row) {
        dbg_row[dbg_col0 + c0 + 4 * j + 0] = h0; dbg_row[dbg_col0 + c0 + 4 * j + 1] = h1;
        dbg_row[dbg_col0 + c0 + 4 * j + 2] = h2; dbg_row[dbg_col0 + c0 + 4 * j + 3] = h3;
      }
#pragma unroll
      for (int k = 0; k < NOUT; ++k) {
        const float4 ww = ldc4<FCS>(w + k * ld + c0 + 4 * j);
        part[k] += h0 * ww.x + h1 * ww.y + h2 * ww.z + h3 * ww.w;
      }
    }
  }
}

// trilinear gather of 16 of the 32 channels from the channel-last embedding grid, ref: nerf/models.py:346-365
// (align_corners=True, zero padding, raw warped coordinates; x -> last grid dim, z -> first)
// Explicit rounding intrinsics: the result must not depend on how a particular kernel instantiation contracts
// (x + 1) * sc or the weight products into FMAs (the pair and single-CTA kernels are compared bit for bit).
__device__ __forceinline__ void grid_gather16(const float* __restrict__ g, int ch0, float x, float y, float z,
                                              float (&out)[16]) {
  const float sc = 0.5f * (SAHS_GRID_RES - 1);
  const float ix = __fmul_rn(__fadd_rn(x, 1.f), sc), iy = __fmul_rn(__fadd_rn(y, 1.f), sc),
              iz = __fmul_rn(__fadd_rn(z, 1.f), sc);
  const float fx = floorf(ix), fy = floorf(iy), fz = floorf(iz);
#pragma unroll
  for (int c = 0; c < 16; ++c) out[c] = 0.f;
  // Two corners (8 x 16-byte loads) in flight at a time, through L1: the 32 consecutive samples of a warp fall into a
  // handful of voxels, so most of these loads hit lines a neighbouring lane just brought in.  (The first version
  // used volatile no-allocate loads that went to L2 one by one: 8.7 K cycles per tile for this gather.)
#pragma unroll
  for (int cp = 0; cp < 8; cp += 2) {
    float w[2];
    const float4* p[2];
    bool ok[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int corner = cp + e;
      const float xi = fx + (corner & 1), yi = fy + ((corner >> 1) & 1), zi = fz + (corner >> 2);
      w[e] = __fmul_rn(__fmul_rn(__fsub_rn(1.f, fabsf(__fsub_rn(ix, xi))), __fsub_rn(1.f, fabsf(__fsub_rn(iy, yi)))),
                       __fsub_rn(1.f, fabsf(__fsub_rn(iz, zi))));
      ok[e] = xi >= 0.f && xi <= SAHS_GRID_RES - 1 && yi >= 0.f && yi <= SAHS_GRID_RES - 1 && zi >= 0.f &&
              zi <= SAHS_GRID_RES - 1;
      const int xc = ok[e] ? (int)xi : 0, yc = ok[e] ? (int)yi : 0, zc = ok[e] ? (int)zi : 0;
      p[e] = reinterpret_cast<const float4*>(
          g + ((((size_t)zc * SAHS_GRID_RES + yc) * SAHS_GRID_RES + xc) * SAHS_GRID_CH) + ch0);
    }
    float4 v[2][4];
#pragma unroll
    for (int e = 0; e < 2; ++e)
#pragma unroll
      for (int q = 0; q < 4; ++q) v[e][q] = __ldg(p[e] + q);
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const float we = ok[e] ? w[e] : 0.f;     // zero padding outside the grid (adds exact zeros, same result)
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        out[4 * q + 0] = __fmaf_rn(we, v[e][q].x, out[4 * q + 0]); out[4 * q + 1] = __fmaf_rn(we, v[e][q].y, out[4 * q + 1]);
        out[4 * q + 2] = __fmaf_rn(we, v[e][q].z, out[4 * q + 2]); out[4 * q + 3] = __fmaf_rn(we, v[e][q].w, out[4 * q + 3]);
      }
    }
  }
}


template <int XYZ_L_, int AMB_DIM_, int AMB_L_, bool AMB_INC_, int DIR_L_, bool USE_W_>
struct FieldCfg {
  static constexpr int XYZ_L = XYZ_L_;
  static constexpr int AMB_DIM = AMB_DIM_;
  static constexpr int AMB_L = AMB_L_;
  static constexpr bool AMB_INC = AMB_INC_;
  static constexpr int DIR_L = DIR_L_;
  static constexpr bool USE_W = USE_W_;
  static constexpr int E0_DIM = 3 + 6 * XYZ_L;                                      // include_input is always on
  static constexpr int E0_PAD = (E0_DIM + 15) / 16 * 16;
  static constexpr int AMB_PE = USE_W ? ((AMB_INC ? AMB_DIM : 0) + 2 * AMB_DIM * AMB_L) : 0;
  static constexpr int E1_DIM = E0_DIM + AMB_PE;
  static constexpr int E1_PAD = (E1_DIM + 15) / 16 * 16;
  static constexpr int DIR_DIM = 3 + 6 * DIR_L;
};

// ---- warp-role loops shared by the forward and backward kernels ------------------------------------------------
// TMA producer: streams the packed stage images of `plan` once per tile through the slot ring.
__device__ __forceinline__ void tma_warp_loop(const FieldPlan& plan, const uint8_t* __restrict__ packed, uint8_t* slots,
                                              uint64_t* full, uint64_t* empty, long long ntiles, int* status, int lane,
                                              long long* prof_base = nullptr) {
  uint32_t slot = 0, phase = 0;
  int iter = 0;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++iter) {
    uint32_t off = 0;
    // (bit 0 of prof_base set: per-pass events only, so this per-stage producer records nothing)
    long long* prof = (prof_base && !(reinterpret_cast<uintptr_t>(prof_base) & 1) && iter == 2 && lane == 0) ? prof_base : nullptr;
    for (int st = 0; st < plan.num_stages; ++st) {
      const uint32_t bytes = sahs_stage_bytes(plan.st[st]);
      mbar_wait(&empty[slot], phase ^ 1, status, 100);
      prof_event(prof, 40000 + st);
      if (lane == 0) {
        mbar_arrive_expect_tx(&full[slot], bytes);
        tma_bulk_g2s(slots + slot * kStageSlotBytes, packed + off, bytes, &full[slot]);
      }
      __syncwarp();
      off += bytes;
      if (++slot == kSlots) { slot = 0; phase ^= 1; }
    }
  }
}

// The MMAs of one stage.  The issuer is a single warp on the uniform datapath and is itself ISSUE-BOUND: the generic
// form below costs ~12 uniform instructions per MMA (descriptor moves, k-step tests, branches), ~90 per stage, and one
// warp retires about one dependent instruction every 4 cycles -- 362 cycles per 4-MMA stage against 256-272 of tensor
// time (measured with one tile set alone, no contention: profiles/r2_duo_timeline_oneset.txt).  Nearly every stage is
// the full K = 64 without a split-precision partner, so that case is straight-line code: four MMAs, descriptor += 2.
template <bool PAIR>
__device__ __forceinline__ void issue_stage_mmas(const StageRec& r, uint32_t flags, uint32_t ksteps, uint64_t a_base,
                                                 uint64_t b0, uint32_t d, uint32_t idesc) {
  const uint64_t a0 = a_base + (uint64_t)r.a_off;
  const uint32_t acc0 = (flags & ST_FRESH) ? 0u : 1u;
  auto mma = [&](uint64_t a, uint64_t b, uint32_t acc) {
    if (elect_one()) {
      if (PAIR) tc_mma_pair(d, a, b, idesc, acc);
      else tc_mma_bf16(d, a, b, idesc, acc);
    }
  };
  if (ksteps == 4 && r.a_chunk2 == 0xFF) {
    mma(a0, b0, acc0);
    mma(a0 + 2, b0 + 2, 1u);
    mma(a0 + 4, b0 + 4, 1u);
    mma(a0 + 6, b0 + 6, 1u);
    return;
  }
  mma(a0, b0, acc0);
  if (ksteps > 1) mma(a0 + 2, b0 + 2, 1u);
  if (ksteps > 2) mma(a0 + 4, b0 + 4, 1u);
  if (ksteps > 3) mma(a0 + 6, b0 + 6, 1u);
  if (r.a_chunk2 != 0xFF) {   // split precision: the residual (lo) activations times the same weights
    const uint64_t a1 = a_base + (uint64_t)r.a2_off;
    mma(a1, b0, 1u);
    if (ksteps > 1) mma(a1 + 2, b0 + 2, 1u);
    if (ksteps > 2) mma(a1 + 4, b0 + 4, 1u);
    if (ksteps > 3) mma(a1 + 6, b0 + 6, 1u);
  }
}

// MMA issuer.  All 32 lanes of the warp run the loop converged on warp-uniform values; every tcgen05 instruction is
// guarded by elect.sync and the barrier waits spin inside asm (mbar_wait_uniform).  ptxas then keeps the whole loop on
// the uniform datapath (LDCU / UPRMT / UTCHMMA / UTCBAR back to back, no R2UR, no per-thread election loops).  Issued
// from an `if (lane == 0)` region instead, every tcgen05 instruction is wrapped in an election loop whose back-branch
// waits for the instruction's operand read: the issuer then runs in lock step with the tensor pipe (measured ~1000
// cycles per 4-MMA stage instead of 256).  Passes are delimited by the a_ready / acc_ready barriers.
template <bool PAIR, bool PROF>
__device__ __forceinline__ void mma_issue_loop(const FieldPlan& plan, uint8_t* X, uint8_t* slots, uint64_t* full,
                                               uint64_t* empty, uint64_t* a_ready, uint64_t* acc_ready,
                                               uint32_t tmem_base, long long first, long long count, long long step,
                                               int* status, long long* prof_base) {
  constexpr uint32_t NS = PAIR ? kPairSlots : kSlots;
  constexpr uint32_t SLOT_UNITS = (PAIR ? kPairSlotBytes : kStageSlotBytes) >> 4;   // descriptor address units (16 B)
  uint32_t slot = 0, phase = 0, a_par = 0;
  const uint64_t a_base = umma_smem_desc_sw128(smem_u32(X));
  const uint64_t b_base = umma_smem_desc_sw128(smem_u32(slots));
  const uint64_t b_base_wide = umma_smem_desc_sw64(smem_u32(slots));   // ST_WIDE stages: [256 x 32], 64-byte rows
  // software pipeline: the next stage's record is fetched and its `full` barrier peeked while this stage's MMAs run
  StageRec r = plan.st[0];
  uint32_t token = 0;
  int iter = 0;
  for (long long it = first; it < count; it += step, ++iter) {
    const bool light = PROF && (reinterpret_cast<uintptr_t>(prof_base) & 1);     // per-pass events only
    long long* prof = (PROF && prof_base && iter == 2 && (threadIdx.x & 31) == 0)
                          ? reinterpret_cast<long long*>(reinterpret_cast<uintptr_t>(prof_base) & ~(uintptr_t)1) : nullptr;
    for (int st = 0; st < plan.num_stages; ++st) {
      const uint32_t flags = r.kflags >> 3, ksteps = r.kflags & 7;
      if (flags & ST_WAIT_A) {
        mbar_wait_uniform<PAIR>(a_ready, a_par, status, 200 + st);
        a_par ^= 1;
        if (PROF) prof_event(prof, 10000 + st);
      }
      mbar_wait_token(&full[slot], phase, token, status, 400 + st);   // PAIR: both halves of the stage have landed
      if (PROF && !light) prof_event(prof, 20000 + st);
      tc_fence_after();
      // An mbarrier wait costs ~200 cycles even when the phase completed long ago, so the NEXT stage's `full` barrier
      // is peeked (non-blocking) BEFORE this stage's MMAs are issued: the peek's latency hides under the MMA issue and
      // the wait at the top of the next iteration is a register test whenever the weights were already there
      // (round 1 peeked after the MMAs, with nothing left to hide the latency under; profiles/r2_duo_timeline_*.txt).
      const uint32_t cur = slot;
      if (++slot == NS) { slot = 0; phase ^= 1; }
      token = mbar_peek(&full[slot], phase);
      const uint32_t idesc = PAIR ? (r.idesc ^ (((128u >> 4) ^ (256u >> 4)) << 24)) : r.idesc;   // M field 128 -> 256
      // descriptor address field counts 16-byte units; K advances 16 elements = 32 bytes = 2 units per MMA
      const uint64_t a0 = a_base + (uint64_t)r.a_off;
      const uint64_t b0 = ((flags & ST_WIDE) ? b_base_wide : b_base) + (uint64_t)(cur * SLOT_UNITS);
      const uint32_t d = tmem_base + (uint32_t)r.d_col8 * 8u;
      const uint32_t acc0 = (flags & ST_FRESH) ? 0u : 1u;
      auto mma = [&](uint64_t a, uint64_t b, uint32_t acc) {
        if (elect_one()) {
          if (PAIR) tc_mma_pair(d, a, b, idesc, acc);
          else tc_mma_bf16(d, a, b, idesc, acc);
        }
      };
      auto commit = [&](uint64_t* bar) {
        if (elect_one()) {
          if (PAIR) tc_commit_pair(bar);
          else tc_commit(bar);
        }
      };
      mma(a0, b0, acc0);
      if (ksteps > 1) mma(a0 + 2, b0 + 2, 1u);
      if (ksteps > 2) mma(a0 + 4, b0 + 4, 1u);
      if (ksteps > 3) mma(a0 + 6, b0 + 6, 1u);
      if (r.a_chunk2 != 0xFF) {   // split precision: the residual (lo) activations times the same weights
        const uint64_t a1 = a_base + (uint64_t)r.a2_off;
        mma(a1, b0, 1u);
        if (ksteps > 1) mma(a1 + 2, b0 + 2, 1u);
        if (ksteps > 2) mma(a1 + 4, b0 + 4, 1u);
        if (ksteps > 3) mma(a1 + 6, b0 + 6, 1u);
      }
      commit(&empty[cur]);
      if (flags & ST_COMMIT) commit(acc_ready);
      if (PROF && (!light || (flags & ST_COMMIT))) prof_event(prof, 30000 + st);
      r = plan.st[(st + 1 < plan.num_stages) ? st + 1 : 0];
    }
  }
}

template <bool PROF = false>
__device__ __forceinline__ void mma_warp_loop(const FieldPlan& plan, uint8_t* X, uint8_t* slots, uint64_t* full,
                                              uint64_t* empty, uint64_t* a_ready, uint64_t* acc_ready, uint32_t tmem_base,
                                              long long ntiles, int* status, int lane, long long* prof_base = nullptr) {
  mma_issue_loop<false, PROF>(plan, X, slots, full, empty, a_ready, acc_ready, tmem_base, blockIdx.x, ntiles,
                              gridDim.x, status, prof_base);
}

// ---- CTA-pair variants ------------------------------------------------------------------------------------
// Tile schedule: cluster c processes tile pairs c, c + nclusters, ...; CTA `rank` of the pair owns tile 2*pair + rank.
// TMA producer of one CTA: its half (rows [rank*n/2, (rank+1)*n/2)) of every stage.
__device__ __forceinline__ void tma_warp_loop_pair(const FieldPlan& plan, const uint8_t* __restrict__ packed,
                                                   uint8_t* slots, uint64_t* full, uint64_t* empty, long long npairs,
                                                   uint32_t rank, int* status, int lane, long long* prof_base = nullptr) {
  uint32_t slot = 0, phase = 0;
  int iter = 0;
  const long long pt0 = cluster_id_x(), pt_step = cluster_num_x();
  for (long long pt = pt0; pt < npairs; pt += pt_step, ++iter) {
    uint32_t off = 0;
    // (bit 0 of prof_base set: per-pass events only, so this per-stage producer records nothing)
    long long* prof = (prof_base && !(reinterpret_cast<uintptr_t>(prof_base) & 1) && iter == 2 && lane == 0) ? prof_base : nullptr;
    for (int st = 0; st < plan.num_stages; ++st) {
      const uint32_t half = sahs_stage_bytes(plan.st[st]) / 2;
      mbar_wait_cluster(&empty[slot], phase ^ 1, status, 100);
      prof_event(prof, 40000 + st);
      if (lane == 0) {
        mbar_arrive_expect_tx(&full[slot], half);
        tma_bulk_g2s(slots + slot * kPairSlotBytes, packed + off + rank * half, half, &full[slot]);
      }
      __syncwarp();
      off += 2 * half;
      if (++slot == kPairSlots) { slot = 0; phase ^= 1; }
    }
  }
}

// Peer CTA (rank 1): forwards "my half of the stage has landed" to the leader's full barrier of the same slot (which
// counts two arrivals per phase: the leader's own expect_tx arrive and this one).
__device__ __forceinline__ void relay_warp_loop_pair(const FieldPlan& plan, uint64_t* full, long long npairs, int* status,
                                                     int lane) {
  uint32_t slot = 0, phase = 0;
  uint32_t remote[kPairSlots];
#pragma unroll
  for (int i = 0; i < kPairSlots; ++i) remote[i] = mapa_u32(&full[i], 0);
  const long long pt0 = cluster_id_x(), pt_step = cluster_num_x();
  for (long long pt = pt0; pt < npairs; pt += pt_step) {
    for (int st = 0; st < plan.num_stages; ++st) {
      mbar_wait(&full[slot], phase, status, 500 + st);
      if (lane == 0) {
        uint32_t r = remote[0];
#pragma unroll
        for (int i = 1; i < kPairSlots; ++i) r = (slot == (uint32_t)i) ? remote[i] : r;
        mbar_arrive_cluster(r);
      }
      __syncwarp();
      if (++slot == kPairSlots) { slot = 0; phase ^= 1; }
    }
  }
}

// Leader CTA (rank 0): issues the M=256 MMAs of the pair.
template <bool PROF = false>
__device__ __forceinline__ void mma_warp_loop_pair(const FieldPlan& plan, uint8_t* X, uint8_t* slots, uint64_t* full,
                                                   uint64_t* empty, uint64_t* a_ready,
                                                   uint64_t* acc_ready, uint32_t tmem_base, long long npairs, int* status,
                                                   int lane, long long* prof_base = nullptr) {
  mma_issue_loop<true, PROF>(plan, X, slots, full, empty, a_ready, acc_ready, tmem_base, cluster_id_x(), npairs,
                             cluster_num_x(), status, prof_base);
}


// ---- two-tile ("duo") variants ---------------------------------------------------------------------------
// One cluster (CTA pair, cta_group::2) per SM pair owns TWO tile pairs at a time, "set 0" and "set 1": each set has its
// own 256 worker threads, activation buffer X, 256 TMEM columns, weight ring and a_ready / acc_ready barriers, and ONE
// issuer thread serves both sets pass by pass, first come first served.  A pass then runs at the full tensor rate and
// the two sets settle into anti-phase -- one set's epilogue under the other set's MMAs -- instead of the lock step two
// independent co-resident CTAs fall into (both in their MMA phase at half rate each, then both in their epilogue with the
// pipe idle).  Set k of cluster c processes tile pairs c + (2 j + k) * nclusters.
constexpr int kDuoThreads = 640;                 // 16 worker warps + 2 TMA + MMA / relay + relay
#ifndef SAHS_DUO_SLOTS
#define SAHS_DUO_SLOTS 4
#endif
constexpr int kDuoSlots = SAHS_DUO_SLOTS;
constexpr int kDuoTmemCols = 512;
constexpr int kDuoTmaWarp0 = 16, kDuoMmaWarp = 18, kDuoRelayWarp1 = 19;
constexpr int kDuoRingBytes = kDuoSlots * kPairSlotBytes;                                 // 32 KB per set
constexpr int kDuoOffSlots = 2 * kSmemX;                                                   // [X0 | X1 | ring0 | ring1 | fc | bars | xchg]
constexpr int kDuoOffFc = kDuoOffSlots + 2 * kDuoRingBytes;
constexpr int kDuoOffBars = kDuoOffFc + kPairFcFloats * 4;
constexpr int kDuoOffXchg = kDuoOffBars + 512;
constexpr int kDuoSmemTotal = kDuoOffXchg + 2 * 512;
constexpr int kDuoBarsPerSet = 2 * kDuoSlots + 2;   // full[NS], empty[NS], a_ready, acc_ready

// TMA producer of one set: this CTA's half of every stage, once per tile pair of the set
__device__ __forceinline__ void tma_warp_loop_duo(const FieldPlan& plan, const uint8_t* __restrict__ packed, uint8_t* slots,
                                                  uint64_t* full, uint64_t* empty, long long first, long long npairs,
                                                  long long step, uint32_t rank, int* status, int lane) {
  uint32_t slot = 0, phase = 0;
  for (long long pt = first; pt < npairs; pt += step) {
    uint32_t off = 0;
    for (int st = 0; st < plan.num_stages; ++st) {
      const uint32_t half = sahs_stage_bytes(plan.st[st]) / 2;
      mbar_wait_cluster(&empty[slot], phase ^ 1, status, 100);
      if (lane == 0) {
        mbar_arrive_expect_tx(&full[slot], half);
        tma_bulk_g2s(slots + slot * kPairSlotBytes, packed + off + rank * half, half, &full[slot]);
      }
      __syncwarp();
      off += 2 * half;
      if (++slot == kDuoSlots) { slot = 0; phase ^= 1; }
    }
  }
}

// peer CTA: forwards "my half of the stage has landed" to the leader's full barrier of the same slot
__device__ __forceinline__ void relay_warp_loop_duo(const FieldPlan& plan, uint64_t* full, long long first, long long npairs,
                                                    long long step, int* status, int lane) {
  uint32_t slot = 0, phase = 0;
  uint32_t remote[kDuoSlots];
#pragma unroll
  for (int i = 0; i < kDuoSlots; ++i) remote[i] = mapa_u32(&full[i], 0);
  for (long long pt = first; pt < npairs; pt += step) {
    for (int st = 0; st < plan.num_stages; ++st) {
      mbar_wait(&full[slot], phase, status, 500 + st);
      if (lane == 0) {
        uint32_t r = remote[0];
#pragma unroll
        for (int i = 1; i < kDuoSlots; ++i) r = (slot == (uint32_t)i) ? remote[i] : r;
        mbar_arrive_cluster(r);
      }
      __syncwarp();
      if (++slot == kDuoSlots) { slot = 0; phase ^= 1; }
    }
  }
}

// Which set may issue its next pass: polls the two a_ready barriers (starting with `prefer`) until one of the sets that
// still has work has its operand ready.  The spin lives in one asm block so that the surrounding issue loop stays
// warp-uniform straight-line code (see mbar_wait_uniform).  live0 / live1: the set still has passes to issue.
__device__ __forceinline__ uint32_t duo_pick(uint64_t* a_ready0, uint32_t par0, uint32_t live0, uint64_t* a_ready1,
                                             uint32_t par1, uint32_t live1, uint32_t prefer, int* status) {
  uint32_t k;
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t.reg .u32 c;\n\t"
      "mov.u32 c, 0;\n\t"
      "mov.u32 %0, %7;\n\t"
      "PICK_%=:\n\t"
      "setp.eq.u32 q, %0, 0;\n\t"
      "@!q bra TRY1_%=;\n\t"
      "setp.eq.u32 p, %3, 0;\n\t"
      "@p bra NEXT_%=;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra NEXT_%=;\n\t"
      "TRY1_%=:\n\t"
      "setp.eq.u32 p, %6, 0;\n\t"
      "@p bra NEXT_%=;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%4], %5;\n\t"
      "@p bra DONE_%=;\n\t"
      "NEXT_%=:\n\t"
      "xor.b32 %0, %0, 1;\n\t"
      "add.u32 c, c, 1;\n\t"
      "setp.lt.u32 p, c, 0x4000000;\n\t"
      "@p bra PICK_%=;\n\t"
      "st.global.u32 [%8], 1;\n\t"
      "st.global.u32 [%8+4], 777;\n\t"
      "fence.sc.sys;\n\t"
      "trap;\n\t"
      "DONE_%=:\n\t}"
      : "=&r"(k)
      : "r"(smem_u32(a_ready0)), "r"(par0), "r"(live0), "r"(smem_u32(a_ready1)), "r"(par1), "r"(live1), "r"(prefer),
        "l"(status)
      : "memory");
  return __shfl_sync(0xffffffffu, k, 0);
}

// A stage record decoded into what the issue path needs (descriptors, accumulator address, flags).  Decoding is a
// constant-bank load plus a dependent chain of ~25 uniform-datapath instructions (~100 cycles): done at the top of a
// stage it sits between the previous stage's last MMA and this stage's first one and drains the tensor pipe's queue --
// measured 362 cycles per 4-MMA stage with one tile set alone, against 270 of MMA time.  So the NEXT stage is decoded
// right after the current stage's MMAs and commit have been issued, while they execute.
struct StageDec {
  uint64_t a0, a1;          // A descriptors (a1: split-precision partner chunk)
  uint32_t d, idesc, flags, ksteps;
  uint32_t fast;            // 1: K = 64, no partner: four straight MMAs; 2: ST_WIDE K = 32: two straight MMAs
  uint32_t wide;
};
__device__ __forceinline__ StageDec duo_decode(const FieldPlan& plan, int st, uint64_t a_base, uint32_t tmem_set) {
  // two 8-byte constant loads (records sit at 8 + 16 i in the kernel parameter block: 8-byte aligned)
  const uint2 lo = reinterpret_cast<const uint2*>(&plan.st[st])[0], hi = reinterpret_cast<const uint2*>(&plan.st[st])[1];
  const uint4 q = make_uint4(lo.x, lo.y, hi.x, hi.y);
  // StageRec: n8 | kflags | a_chunk | d_col8 || a_chunk2 | pad[3] || idesc || a_off | a2_off
  StageDec s;
  const uint32_t kflags = (q.x >> 8) & 0xffu, a_chunk2 = q.y & 0xffu;
  s.flags = kflags >> 3;
  s.ksteps = kflags & 7u;
  s.d = tmem_set + ((q.x >> 24) & 0xffu) * 8u;
  s.idesc = q.z ^ (((128u >> 4) ^ (256u >> 4)) << 24);   // M field 128 -> 256
  s.a0 = a_base + (uint64_t)(q.w & 0xffffu);
  s.a1 = a_base + (uint64_t)(q.w >> 16);
  s.wide = (s.flags & ST_WIDE) ? 1u : 0u;
  s.fast = (a_chunk2 != 0xffu) ? 0u : ((s.ksteps == 4u && !s.wide) ? 1u : ((s.ksteps == 2u && s.wide) ? 2u : 0u));
  if (a_chunk2 == 0xffu) s.a1 = 0;
  return s;
}

// per-set issue state of the duo issuer
struct DuoSet {
  uint32_t slot = 0, phase = 0, a_par = 0;
  int st = 0;
  long long it;
  StageDec next;            // the set's next stage, already decoded
};

// One pass (the stages up to and including the one flagged ST_COMMIT) of set K; the set's a_ready phase has completed.
// An mbarrier wait costs ~200 cycles even when the phase completed long ago (measured: profiles/r2_duo_timeline_*.txt --
// 201 cycles per stage, every stage).  So the NEXT stage's `full` barrier is peeked (non-blocking test_wait) BEFORE this
// stage's MMAs are issued: the peek's latency hides under the MMA issue, and the wait at the top of the next iteration
// is a register test whenever the weights were already there (CUTLASS' consumer_try_wait token idiom, one stage
// earlier than round 1 had it).
template <int K, bool PROF>
__device__ __forceinline__ void duo_issue_pass(const FieldPlan& plan, DuoSet& s, uint64_t a_base, uint64_t b_base,
                                               uint64_t b_base_wide, uint64_t* full, uint64_t* empty, uint64_t* acc_ready, uint32_t tmem_set,
                                               long long step, int* status, long long*& prof, bool rec) {
  constexpr uint32_t SLOT_UNITS = kPairSlotBytes >> 4;
  s.a_par ^= 1;
  tc_fence_after();
  uint32_t flags;
  uint32_t token = mbar_peek(&full[s.slot], s.phase);
  do {
    const StageDec c = s.next;
    flags = c.flags;
    if (PROF && rec) prof_event(prof, 25000 + 1000 * K + s.st);
    mbar_wait_token(&full[s.slot], s.phase, token, status, 400 + s.st);   // both halves of the stage have landed
    if (PROF && rec) prof_event(prof, 20000 + 1000 * K + s.st);
    const uint32_t slot = s.slot;
    const uint64_t b0 = (c.wide ? b_base_wide : b_base) + (uint64_t)(slot * SLOT_UNITS);
    const uint32_t acc0 = (flags & ST_FRESH) ? 0u : 1u;
    auto mma = [&](uint64_t a, uint64_t b, uint32_t acc) {
      if (elect_one()) tc_mma_pair(c.d, a, b, c.idesc, acc);
    };
    if (c.fast == 2u) {
      mma(c.a0, b0, acc0);
      mma(c.a0 + 2, b0 + 2, 1u);
    } else if (c.fast == 1u) {
      mma(c.a0, b0, acc0);
      mma(c.a0 + 2, b0 + 2, 1u);
      mma(c.a0 + 4, b0 + 4, 1u);
      mma(c.a0 + 6, b0 + 6, 1u);
    } else {
      mma(c.a0, b0, acc0);
      if (c.ksteps > 1) mma(c.a0 + 2, b0 + 2, 1u);
      if (c.ksteps > 2) mma(c.a0 + 4, b0 + 4, 1u);
      if (c.ksteps > 3) mma(c.a0 + 6, b0 + 6, 1u);
      if (c.a1 != 0) {   // split precision: the residual (lo) activations times the same weights
        mma(c.a1, b0, 1u);
        if (c.ksteps > 1) mma(c.a1 + 2, b0 + 2, 1u);
        if (c.ksteps > 2) mma(c.a1 + 4, b0 + 4, 1u);
        if (c.ksteps > 3) mma(c.a1 + 6, b0 + 6, 1u);
      }
    }
    if (elect_one()) tc_commit_pair(&empty[slot]);
    if (flags & ST_COMMIT) {
      if (elect_one()) tc_commit_pair(acc_ready);
    }
    // ---- off the critical path (the MMAs above are executing): next stage's weights peeked, its record decoded ----
    if (++s.slot == kDuoSlots) { s.slot = 0; s.phase ^= 1; }
    token = mbar_peek(&full[s.slot], s.phase);
    if (++s.st == plan.num_stages) { s.st = 0; s.it += step; }
    s.next = duo_decode(plan, s.st, a_base, tmem_set);
  } while (!(flags & ST_COMMIT));
}

// Leader CTA: the single in-order issuer of both sets.
template <bool PROF>
__device__ __forceinline__ void mma_warp_loop_duo(const FieldPlan& plan, uint8_t* X0, uint8_t* X1, uint8_t* slots0,
                                                  uint8_t* slots1, uint64_t* bars0, uint64_t* bars1, uint32_t tmem_base,
                                                  long long first0, long long first1, long long step, long long npairs,
                                                  int* status, long long* prof_in = nullptr) {
  // bit 0 of prof_in: per-stage events too (before / after the wait for the weights)
  const bool prof_stages = (reinterpret_cast<uintptr_t>(prof_in) & 1) != 0;
  long long* prof = reinterpret_cast<long long*>(reinterpret_cast<uintptr_t>(prof_in) & ~(uintptr_t)1);
  DuoSet s0, s1;
  s0.it = first0;
  s1.it = first1;
  const uint64_t a0 = umma_smem_desc_sw128(smem_u32(X0)), a1 = umma_smem_desc_sw128(smem_u32(X1));
  const uint64_t b0 = umma_smem_desc_sw128(smem_u32(slots0)), b1 = umma_smem_desc_sw128(smem_u32(slots1));
  const uint64_t bw0 = umma_smem_desc_sw64(smem_u32(slots0)), bw1 = umma_smem_desc_sw64(smem_u32(slots1));
  uint64_t *full0 = bars0, *empty0 = bars0 + kDuoSlots, *a_rdy0 = bars0 + 2 * kDuoSlots, *acc0 = bars0 + 2 * kDuoSlots + 1;
  uint64_t *full1 = bars1, *empty1 = bars1 + kDuoSlots, *a_rdy1 = bars1 + 2 * kDuoSlots, *acc1 = bars1 + 2 * kDuoSlots + 1;
  s0.next = duo_decode(plan, 0, a0, tmem_base);
  s1.next = duo_decode(plan, 0, a1, tmem_base + 256u);
  uint32_t prefer = 0;
  for (;;) {
    const uint32_t live0 = s0.it < npairs ? 1u : 0u, live1 = s1.it < npairs ? 1u : 0u;
    if (!(live0 | live1)) break;
    const uint32_t k = duo_pick(a_rdy0, s0.a_par, live0, a_rdy1, s1.a_par, live1, prefer, status);
    // SAHS_DBG_PROF_DUO: pass start / issue end of both sets' 4th tile (tags 10000 / 30000 + 1000 set + first stage)
    const bool rec = PROF && prof && ((k == 0 && s0.it == first0 + 3 * step) || (k == 1 && s1.it == first1 + 3 * step));
    const long long tag = 1000 * k + (k == 0 ? s0.st : s1.st);
    if (PROF && rec) prof_event(prof, 10000 + tag);
    if (k == 0) duo_issue_pass<0, PROF>(plan, s0, a0, b0, bw0, full0, empty0, acc0, tmem_base, step, status, prof, rec && prof_stages);
    else duo_issue_pass<1, PROF>(plan, s1, a1, b1, bw1, full1, empty1, acc1, tmem_base + 256u, step, status, prof, rec && prof_stages);
    if (PROF && rec) prof_event(prof, 30000 + tag);
    prefer = k ^ 1u;
  }
}

}  // namespace

// Shared device/host helpers for libsahs_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>

#include "../../include/sahs_b200.h"

// ------------------------------------------------------------------------------------------------
// host side: error plumbing (no exceptions cross the C ABI)
// ------------------------------------------------------------------------------------------------
void sahs_set_error(const char* fmt, ...);
extern std::atomic<uint64_t> g_sahs_launches;

#define SAHS_CHECK_ARG(cond, msg)                                  \
  do {                                                             \
    if (!(cond)) {                                                 \
      sahs_set_error("%s: %s", __func__, msg);                     \
      return SAHS_EINVAL;                                          \
    }                                                              \
  } while (0)

#define SAHS_CUDA(call)                                                                   \
  do {                                                                                    \
    cudaError_t e__ = (call);                                                             \
    if (e__ != cudaSuccess) {                                                             \
      sahs_set_error("%s: %s failed: %s", __func__, #call, cudaGetErrorString(e__));      \
      return SAHS_ECUDA;                                                                  \
    }                                                                                     \
  } while (0)

#define SAHS_LAUNCH_CHECK()                                                               \
  do {                                                                                    \
    g_sahs_launches.fetch_add(1, std::memory_order_relaxed);                              \
    cudaError_t e__ = cudaGetLastError();                                                 \
    if (e__ != cudaSuccess) {                                                             \
      sahs_set_error("%s: kernel launch failed: %s", __func__, cudaGetErrorString(e__));  \
      return SAHS_ECUDA;                                                                  \
    }                                                                                     \
  } while (0)

int sahs_num_sms();
int* sahs_status_words(int which);   // device-visible diagnostic word of kernel 0 fwd / 1 dgrad / 2 wgrad / 3 Stage-II conv

// ------------------------------------------------------------------------------------------------
// device side
// ------------------------------------------------------------------------------------------------
#ifdef __CUDACC__

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- Philox-4x32-10 (Salmon et al. 2011): the library's one random source -------------------------------------
struct RngArg {   // device-side copy of sahs_rng (on == 0: no in-kernel draws)
  unsigned long long seed;
  const unsigned long long* counter;
  uint32_t stream;
  uint32_t on;
};
inline RngArg rng_arg(const sahs_rng* r) {
  RngArg a{0ull, nullptr, 0u, 0u};
  if (r) { a.seed = r->seed; a.counter = r->counter_dev; a.stream = r->stream; a.on = 1u; }
  return a;
}
__device__ __forceinline__ uint4 philox4x32_10(uint64_t ctr, uint32_t stream, uint64_t key) {
  uint32_t c0 = (uint32_t)ctr, c1 = (uint32_t)(ctr >> 32), c2 = stream, c3 = 0u;
  uint32_t k0 = (uint32_t)key, k1 = (uint32_t)(key >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
    const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
    const uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}
__device__ __forceinline__ uint64_t rng_key(const RngArg& r) { return r.seed + (r.counter ? *r.counter : 0ull); }
// U[0,1) on the 2^-24 grid (what torch.rand returns for fp32)
__device__ __forceinline__ float rng_uniform(const RngArg& r, uint64_t idx) {
  return (float)(philox4x32_10(idx, r.stream, rng_key(r)).x >> 8) * (1.0f / 16777216.0f);
}
// N(0,1): Box-Muller, u1 in (0,1]
__device__ __forceinline__ float rng_normal(const RngArg& r, uint64_t idx) {
  const uint4 q = philox4x32_10(idx, r.stream, rng_key(r));
  const float u1 = ((float)(q.x >> 8) + 1.0f) * (1.0f / 16777216.0f);
  const float u2 = (float)(q.y >> 8) * (1.0f / 16777216.0f);
  return sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
}

// ---- mbarrier ----------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, 0x4000;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// try_wait carries a suspend-time hint (0x4000 ns): the waiting thread sleeps in hardware until the phase completes or
// the hint expires instead of re-issuing the poll (the polls were 38 % of all issued instructions of the field kernel).
// Bounded wait: a protocol bug must surface as a reported failure, never as a hung GPU box.
// status[0] = 1 on timeout, status[1] = tag.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int* status, int tag) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 24)) {
      if (status) {
        status[0] = 1;
        status[1] = tag;
        status[2] = (int)blockIdx.x;
        status[3] = (int)threadIdx.x;
        __threadfence_system();
      }
      __trap();
    }
  }
}

// ---- async proxy / TMA bulk copy -----------------------------------------------------------------
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// global -> shared bulk copy executed by the TMA unit, completion counted on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// shared -> global bulk copy by the TMA unit (bulk async-group completion).  Used for the training tapes: an
// operand chunk the epilogue has just written to shared memory goes to HBM as one 16 KB burst instead of 128 x 8
// scattered 16-byte stores.
// The stores carry an L2 evict-first policy: the tapes are written once and read much later, and as ordinary
// write-allocate traffic the 3.6 GB per step kept pushing the (L2-resident) weight image and frame constants out of L2
// (training forward 0.88 -> 0.69 ms per level with the hint).
__device__ __forceinline__ void tma_bulk_s2g(void* gmem_dst, const void* smem_src, uint32_t bytes) {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(gmem_dst),
               "r"(smem_u32(smem_src)), "r"(bytes), "l"(pol)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ---- tcgen05 / TMEM --------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {  // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // same warp that allocated
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// tcgen05.commit: arrive on an mbarrier once all previously issued MMAs of this thread completed
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, bf16 x bf16 -> fp32, M=128, N from idesc, K=16
__device__ __forceinline__ void tc_mma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread t of the warp gets row (lane base + t)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
// Prefetching variant: `pin` is a register of the block that is processed while this load is in flight.  The fake
// in/out dependence keeps the compiler from sinking the load below that block's math (which it otherwise does to
// shorten the destination registers' live ranges, serialising TMEM latency with the epilogue math).
__device__ __forceinline__ void tmem_ld16_prefetch(uint32_t taddr, uint32_t (&v)[16], uint32_t& pin) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%17];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "+r"(pin)
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- UMMA descriptors (K-major, SWIZZLE_128B, bf16; layout verified against
//      cute/arch/mma_sm100_desc.hpp SmemDescriptor / InstrDescriptor bit fields) ---------------------
// operand tile in smem: rows of 64 bf16 (128 B), 8-row groups of 1024 B, 16-byte units XOR-swizzled by row%8
__device__ __forceinline__ uint64_t umma_smem_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);          // start address, bits [0,14)
  d |= (uint64_t)0 << 16;                               // LBO (unused for swizzled K-major), bits [16,30)
  d |= (uint64_t)(1024 >> 4) << 32;                     // SBO = 1024 B between 8-row groups, bits [32,46)
  d |= (uint64_t)1 << 46;                               // descriptor version 1 (sm_100)
  d |= (uint64_t)2 << 61;                               // layout type SWIZZLE_128B
  return d;
}
// kind::f16 instruction descriptor: D fp32, A/B both bf16 (format 1) or both fp16 (format 0), K-major, M=128
__device__ __forceinline__ uint32_t umma_idesc_m128(uint32_t n, bool f16) {
  const uint32_t fmt = f16 ? 0u : 1u;
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}
// byte offset of element (row, col) inside one [rows x 64] bf16 K-chunk with 128B swizzle
__host__ __device__ __forceinline__ uint32_t sw128_offset(uint32_t row, uint32_t col) {
  return (row >> 3) * 1024u + (row & 7u) * 128u + ((((col >> 3) ^ row) & 7u) << 4) + (col & 7u) * 2u;
}

// K-major SWIZZLE_64B operand tile (rows of 32 16-bit elements = 64 B, 8-row groups of 512 B, 16-byte units XOR-swizzled
// by (row >> 1) & 3): the B operand of ST_WIDE stages.  Layout type 4, SBO = 512 B.
__device__ __forceinline__ uint64_t umma_smem_desc_sw64(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;
  return d;
}
__host__ __device__ __forceinline__ uint32_t sw64_offset(uint32_t row, uint32_t col) {
  return (row >> 3) * 512u + (row & 7u) * 64u + ((((col >> 3) ^ (row >> 1)) & 3u) << 4) + (col & 7u) * 2u;
}

// ---- warp-converged issue -------------------------------------------------------------------------------
// tcgen05.mma / tcgen05.commit take uniform-register operands.  Issued from a divergent `if (lane == 0)` region, ptxas
// wraps each of them in a per-thread election loop whose back-branch waits until the instruction has consumed its
// operands -- for a commit that is after every earlier MMA has drained, i.e. the issuer runs in lock step with the
// tensor pipe (measured: ~1000 cycles per commit).  The variants below are called by all 32 lanes of a converged warp
// with warp-uniform arguments; elect.sync picks the lane that issues.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0, laneid = 0;
  asm volatile(
      "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
      "elect.sync rx|px, %2;\n\t"
      "@px mov.s32 %1, 1;\n\t"
      "mov.s32 %0, rx;\n\t}"
      : "+r"(laneid), "+r"(pred)
      : "r"(0xFFFFFFFF));
  return pred != 0;
}
// Bounded mbarrier wait whose spin loop lives inside one asm block: the compiler sees straight-line, warp-uniform
// code around it (a C++ spin loop makes everything after it "possibly divergent" and pushes the MMA issue loop off
// the uniform datapath).  On timeout: status[0] = 1, status[1] = tag, then trap.  (CLUSTER marks barriers the peer CTA arrives on; same wait.)
template <bool CLUSTER>
__device__ __forceinline__ void mbar_wait_uniform(uint64_t* bar, uint32_t parity, int* status, int tag) {
  if (CLUSTER) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .u32 c;\n\t"
        "mov.u32 c, 0;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, 0x4000;\n\t"
        "@p bra DONE_%=;\n\t"
        "add.u32 c, c, 1;\n\t"
        "setp.lt.u32 p, c, 0x1000000;\n\t"
        "@p bra WAIT_%=;\n\t"
        "st.global.u32 [%2], 1;\n\t"
        "st.global.u32 [%2+4], %3;\n\t"
        "fence.sc.sys;\n\t"
        "trap;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
        "r"(parity), "l"(status), "r"(tag)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .u32 c;\n\t"
        "mov.u32 c, 0;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, 0x4000;\n\t"
        "@p bra DONE_%=;\n\t"
        "add.u32 c, c, 1;\n\t"
        "setp.lt.u32 p, c, 0x1000000;\n\t"
        "@p bra WAIT_%=;\n\t"
        "st.global.u32 [%2], 1;\n\t"
        "st.global.u32 [%2+4], %3;\n\t"
        "fence.sc.sys;\n\t"
        "trap;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
        "r"(parity), "l"(status), "r"(tag)
        : "memory");
  }
}
__device__ __forceinline__ void tc_mma_f16_w(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
  if (elect_one()) tc_mma_bf16(d_tmem, a_desc, b_desc, idesc, accumulate);
}
__device__ __forceinline__ void tc_commit_w(uint64_t* bar) {
  if (elect_one()) tc_commit(bar);
}

// Non-blocking look at a barrier phase (1 = complete).  Issued right after a stage's MMAs for the NEXT stage's
// barrier, so that the wait at the top of the next iteration is normally a register test (CUTLASS' peek / token idiom).
__device__ __forceinline__ uint32_t mbar_peek(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok;
}
// mbar_wait_uniform that is skipped when the peeked token already saw the phase complete
__device__ __forceinline__ void mbar_wait_token(uint64_t* bar, uint32_t parity, uint32_t token, int* status, int tag) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .u32 c;\n\t"
      "setp.ne.u32 p, %4, 0;\n\t"
      "@p bra DONE_%=;\n\t"
      "mov.u32 c, 0;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, 0x4000;\n\t"
      "@p bra DONE_%=;\n\t"
      "add.u32 c, c, 1;\n\t"
      "setp.lt.u32 p, c, 0x1000000;\n\t"
      "@p bra WAIT_%=;\n\t"
      "st.global.u32 [%2], 1;\n\t"
      "st.global.u32 [%2+4], %3;\n\t"
      "fence.sc.sys;\n\t"
      "trap;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
      "r"(parity), "l"(status), "r"(tag), "r"(token)
      : "memory");
}

// ---- CTA pairs (cluster of 2, tcgen05 cta_group::2) ---------------------------------------------------
// The three getters below must be called by a converged warp: the value goes through a shuffle so that the compiler
// knows it is warp-uniform (see the note on the warp index in field_fwd.cu).
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return __shfl_sync(0xffffffffu, r, 0);
}
__device__ __forceinline__ uint32_t cluster_id_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
  return __shfl_sync(0xffffffffu, r, 0);
}
__device__ __forceinline__ uint32_t cluster_num_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r));
  return __shfl_sync(0xffffffffu, r, 0);
}
__device__ __forceinline__ void cluster_sync_all() {   // every thread of both CTAs
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `p` (a shared::cta pointer of this CTA) inside CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(const void* p, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(rank));
  return r;
}
// Remote arrive with the default semantics (what CUTLASS' ClusterBarrier::arrive(cta_id) emits).  The explicit
// .release.cluster form costs a MEMBAR + ERRBAR per arrive, and waiting with .acquire.cluster makes ptxas emit
// CCTL.IVALL (a full L1 invalidate) after every wait -- with one wait per weight stage that wiped the L1-resident
// biases ~140 times per tile.  The data guarded here lives in shared memory and is read by the async proxy only.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void st_cluster_u32(uint32_t cluster_addr, uint32_t v) {
  asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(cluster_addr), "r"(v) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, 0x4000;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// bounded wait on a barrier that the peer CTA arrives on
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity, int* status, int tag) {
  uint32_t spins = 0;
  while (!mbar_try_wait_cluster(bar, parity)) {
    if (++spins > (1u << 24)) {
      if (status) {
        status[0] = 1;
        status[1] = tag;
        status[2] = (int)blockIdx.x;
        status[3] = (int)threadIdx.x;
        __threadfence_system();
      }
      __trap();
    }
  }
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_dst, uint32_t ncols) {  // one full warp in each CTA
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// commit of the pair's MMAs: arrives on the barrier at this smem offset in both CTAs
__device__ __forceinline__ void tc_commit_pair(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"((uint16_t)3)
      : "memory");
}
// D[tmem, both CTAs] (+)= A[smem, 128 rows per CTA] * B[smem, N/2 rows per CTA]^T; issued by the leader CTA only
__device__ __forceinline__ void tc_mma_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_mma_pair_w(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                              uint32_t accumulate) {
  if (elect_one()) tc_mma_pair(d_tmem, a_desc, b_desc, idesc, accumulate);
}
__device__ __forceinline__ void tc_commit_pair_w(uint64_t* bar) {
  if (elect_one()) tc_commit_pair(bar);
}
// kind::f16 instruction descriptor for the pair: M = 256 (128 rows in each CTA)
__device__ __forceinline__ uint32_t umma_idesc_m256(uint32_t n, bool f16) {
  const uint32_t fmt = f16 ? 0u : 1u;
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((n >> 3) << 17) | ((256u >> 4) << 24);
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

#endif  // __CUDACC__

// Micro-benchmark: tcgen05.mma completion time for N = 128 vs N = 256 with the B operand in different shared-memory
// layouts, alone, with a concurrent st.shared stream (an epilogue's stores, stores=1) or a concurrent tcgen05.ld stream
// (an epilogue's accumulator reads, stores=2) from the other four warps, at 1 and 2 CTAs per SM.  Question (experiment (d)): does N = 256 (A tile read once instead of twice) pay, and in
// which layout?  layout 0: K-major SWIZZLE_128B rows of 64 k (N x 128 B);  1: K-major SWIZZLE_64B rows of 32 k;
// 2: K-major no swizzle (core matrices 128 B apart in k, 8-row groups 512 B apart);  3: MN-major SWIZZLE_128B panels of
// [32 k x 128 B].  Operand data is zero (timing only).  Diagnostic only; build:
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o scripts/build/mma_operand_microbench scripts/mma_operand_microbench.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../sahs-deformable-nerf_b200/csrc/sahs_common.cuh"

void sahs_set_error(const char*, ...) {}
std::atomic<uint64_t> g_sahs_launches{0};
int sahs_num_sms() { return 148; }

// mode 0: cta_group::1; mode 1: cta_group::2 (cluster of 2, leader issues)
// stores: when nonzero, the 4 other warps hammer st.shared (16 B per thread) into a scratch area while MMAs run
template <int MODE, int ELECT>
__global__ void __launch_bounds__(192, 2) bench(int n, int nmma, int per_commit, int stores, long long* out, int layout) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* A = smem;               // 16 KB
  uint8_t* B = smem + 16384;       // up to 32 KB
  uint8_t* scratch = smem + 49152; // 16 KB
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 65536);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar + 2);
  volatile int* stop = reinterpret_cast<volatile int*>(bar + 3);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_init(bar + 1, 1);
    *stop = 0;
    fence_mbar_init();
  }
  if (warp == 0) {
    if (MODE == 1) tmem_alloc_pair(tmem_ptr, 256);
    else tmem_alloc(tmem_ptr, 256);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr;
  const uint32_t rank = MODE == 1 ? cluster_ctarank() : 0;
  if (MODE == 1) cluster_sync_all();
  if (warp == 0) {
    if (rank == 0) {
      long long t0 = 0, t1 = 0, t2 = 0;
      for (int rep = 0; rep < 3; ++rep) {
        __syncwarp();
        t0 = clock64();
        if (ELECT == 2) {
          // converged warp, one election per group of 4 MMAs (+ commit), as a pipelined GEMM main loop would
          uint32_t idesc = MODE == 1 ? umma_idesc_m256(n, true) : umma_idesc_m128(n, true);
          if (layout == 3) idesc |= 1u << 16;                       // B MN-major
          const uint64_t a0 = umma_smem_desc_sw128(smem_u32(A));
          uint64_t b0 = umma_smem_desc_sw128(smem_u32(B));
          uint32_t kstep = 2, tstep = 0;                              // address units (16 B) per K16 step / per K32 tile
          const uint32_t rows = MODE == 1 ? n / 2 : n;                // B rows this CTA holds
          if (layout == 1) {                                          // SWIZZLE_64B: SBO 512 B, layout type 4
            b0 = (uint64_t)((smem_u32(B) & 0x3FFFF) >> 4) | ((uint64_t)(512 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)4 << 61);
            tstep = rows * 64 / 16;
          } else if (layout == 2) {                                   // no swizzle: LBO 128 B (k), SBO 512 B (rows)
            b0 = (uint64_t)((smem_u32(B) & 0x3FFFF) >> 4) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(512 >> 4) << 32) | ((uint64_t)1 << 46);
            kstep = 16; tstep = rows * 64 / 16;
          } else if (layout == 3) {                                   // MN-major SW128: LBO 4 KB (panels), SBO 1 KB
            b0 = (uint64_t)((smem_u32(B) & 0x3FFFF) >> 4) | ((uint64_t)(4096 >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
            kstep = 128; tstep = rows * 64 / 16;
          }
          for (int g = 0; g < nmma / 4; ++g) {
            const uint32_t d = tmem + (n == 128 ? (g & 1) * 128 : 0);
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                // layout 0: four K16 steps along one 128-byte row; others: two K32 tiles of two K16 steps each
                const uint64_t b = layout == 0 ? b0 + 2 * k : b0 + (k >> 1) * tstep + (k & 1) * kstep;
                if (MODE == 1) tc_mma_pair(d, a0 + 2 * k, b, idesc, 1u);
                else tc_mma_bf16(d, a0 + 2 * k, b, idesc, 1u);
              }
              if (per_commit) {
                if (MODE == 1) tc_commit_pair(bar + 1);
                else tc_commit(bar + 1);
              }
            }
            __syncwarp();
          }
          if (MODE == 1) tc_commit_pair_w(bar);
          else tc_commit_w(bar);
        } else if (ELECT == 1) {
          // all lanes converged, warp-uniform operands, elect.sync inside the wrappers
          const uint32_t idesc = MODE == 1 ? umma_idesc_m256(n, true) : umma_idesc_m128(n, true);
          const uint64_t a0 = umma_smem_desc_sw128(smem_u32(A));
          const uint64_t b0 = umma_smem_desc_sw128(smem_u32(B));
          int since = 0;
          for (int i = 0; i < nmma; ++i) {
            const int k = i & 3;
            if (MODE == 1) tc_mma_pair_w(tmem + (n == 128 ? (i & 4) * 32 : 0), a0 + 2 * k, b0 + 2 * k, idesc, 1u);
            else tc_mma_f16_w(tmem + (n == 128 ? (i & 4) * 32 : 0), a0 + 2 * k, b0 + 2 * k, idesc, 1u);
            if (per_commit && ++since == per_commit && i != nmma - 1) {
              since = 0;
              if (MODE == 1) tc_commit_pair_w(bar + 1);
              else tc_commit_w(bar + 1);
            }
          }
          if (MODE == 1) tc_commit_pair_w(bar);
          else tc_commit_w(bar);
        } else
        if (lane == 0) {
          const uint32_t idesc = MODE == 1 ? umma_idesc_m256(n, true) : umma_idesc_m128(n, true);
          const uint64_t a0 = umma_smem_desc_sw128(smem_u32(A));
          const uint64_t b0 = umma_smem_desc_sw128(smem_u32(B));
          for (int i = 0; i < nmma; ++i) {
            const int k = i & 3;
            if (MODE == 1) tc_mma_pair(tmem + (n == 128 ? (i & 4) * 32 : 0), a0 + 2 * k, b0 + 2 * k, idesc, 1u);
            else tc_mma_bf16(tmem + (n == 128 ? (i & 4) * 32 : 0), a0 + 2 * k, b0 + 2 * k, idesc, 1u);
            if (per_commit && (i % per_commit) == per_commit - 1 && i != nmma - 1) {
              // intermediate commits to a dummy barrier (as the ring's empty barriers)
              if (MODE == 1) tc_commit_pair(bar + 1);
              else tc_commit(bar + 1);
            }
          }
          if (MODE == 1) tc_commit_pair(bar);
          else tc_commit(bar);
        }
        t1 = clock64();
        __syncwarp();
        mbar_wait(bar, rep & 1, nullptr, 0);
        t2 = clock64();
      }
      if (lane == 0) {
        out[blockIdx.x * 2 + 0] = t1 - t0;
        out[blockIdx.x * 2 + 1] = t2 - t0;
      }
      *stop = 1;
    } else {
      // peer CTA of a pair: wait for the three commits
      for (int rep = 0; rep < 3; ++rep) mbar_wait(bar, rep & 1, nullptr, 0);
      *stop = 1;
    }
  } else if (stores == 2 && warp >= 2) {
    // an epilogue's other half: tcgen05.ld of the accumulator columns (each warp its own lane quarter), back to back
    const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t v[16], acc = 0;
    int it = 0;
    while (!*stop) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        tmem_ld16(trow + ((it * 8 + j) & 15) * 16, v);
        tmem_ld_wait();
        acc += v[0] ^ v[15];
      }
      ++it;
    }
    if (acc == 0x12345678u) *stop = 2;
  } else if (stores && warp >= 2) {
    uint4 v = make_uint4(threadIdx.x, 1, 2, 3);
    int it = 0;
    while (!*stop) {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        *reinterpret_cast<uint4*>(scratch + ((threadIdx.x * 16 + j * 2048 + it * 16) & 16383)) = v;
      ++it;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (MODE == 1) cluster_sync_all();
  if (warp == 0) {
    tc_fence_after();
    if (MODE == 1) tmem_dealloc_pair(tmem, 256);
    else tmem_dealloc(tmem, 256);
  }
}

template <int MODE, int ELECT = 0>
void run(int n, int nmma, int per_commit, int stores, int ctas_per_sm, int layout = 0) {
  const int smem = 65536 + 64 + (ctas_per_sm == 1 ? 60000 : 0);   // pad to force one CTA per SM
  auto k = bench<MODE, ELECT>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int grid = 148 * ctas_per_sm;
  long long* out;
  cudaMalloc(&out, grid * 2 * sizeof(long long));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(192);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = MODE == 1 ? 2 : 1;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, k, n, nmma, per_commit, stores, out, layout);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("mode %d n %d: CUDA error %s\n", MODE, n, cudaGetErrorString(e));
    exit(1);
  }
  std::vector<long long> h(grid * 2);
  cudaMemcpy(h.data(), out, grid * 2 * sizeof(long long), cudaMemcpyDeviceToHost);
  double issue = 0, done = 0;
  int cnt = 0;
  for (int b = 0; b < grid; ++b) {
    if (MODE == 1 && (b & 1)) continue;
    issue += h[2 * b];
    done += h[2 * b + 1];
    ++cnt;
  }
  issue /= cnt;
  done /= cnt;
  const double mac_per_sm = (double)128 * n * 16 * nmma;   // per SM (in pair mode each SM does 128 rows)
  const double ideal = mac_per_sm / 4096.0;
  static const char* names[4] = {"K-major SW128", "K-major SW64 ", "K-major none ", "MN-major SW128"};
  printf("cta_group::%d N=%3d B %s stores=%d ctas/SM=%d: done %7.0f cyc for %d MMAs (%.1f cyc/MMA), %.0f%% of 4096 MAC/cyc/SM "
         "per CTA, x%d CTAs\n",
         MODE + 1, n, names[layout], stores, ctas_per_sm, done, nmma, done / nmma, 100.0 * ideal / done, ctas_per_sm);
  (void)issue;
  cudaFree(out);
}

int main() {
  for (int cps = 1; cps <= 2; ++cps)
    for (int stores = 0; stores <= 2; ++stores) {   // 0: MMAs only, 1: + st.shared stream, 2: + tcgen05.ld stream
      run<1, 2>(128, 128, 1, stores, cps, 0);
      for (int layout = 0; layout < 4; ++layout) run<1, 2>(256, 128, 1, stores, cps, layout);
      run<0, 2>(128, 128, 1, stores, cps, 0);
      for (int layout = 0; layout < 4; ++layout) run<0, 2>(256, 128, 1, stores, cps, layout);
    }
  return 0;
}

// Micro-benchmark: what an N = 128, K = 16 tcgen05.mma (cta_group::2, M = 256) costs in the field kernel's issue PATTERN:
// four K-steps per 64-wide A chunk, chunks cycling through a 64 KB activation buffer, B slots cycling through a 4 x 8 KB
// ring, `dep` consecutive MMAs accumulating into the same TMEM columns before the other column half is used, a commit
// after every 4 MMAs -- with zero or random operand data, over a long run (power management included).
// Question: the render kernel's trunk pass takes 92 cycles per MMA with one tile set
// alone and nothing else running, where round 1's micro-benchmark (zero data, one operand tile, 128 MMAs) saw 66.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o scripts/build/mma_pattern_microbench scripts/mma_pattern_microbench.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_fp16.h>
#include "../sahs-deformable-nerf_b200/csrc/sahs_common.cuh"

void sahs_set_error(const char*, ...) {}
std::atomic<uint64_t> g_sahs_launches{0};
int sahs_num_sms() { return 148; }

__global__ void __launch_bounds__(192, 1) bench(int nmma, int dep, int data, int footprint, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* A = smem;                // 4 chunks x 16 KB
  uint8_t* B = smem + 65536;        // 4 slots x 8 KB
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 65536 + 32768);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t seed = 1234567u + blockIdx.x * 977u + threadIdx.x;
  for (int i = threadIdx.x; i < (65536 + 32768) / 4; i += blockDim.x) {
    uint32_t v = 0;
    if (data) {
      seed = seed * 1664525u + 1013904223u;
      const float f0 = ((seed >> 8) & 0xffff) / 65536.f - 0.5f;
      seed = seed * 1664525u + 1013904223u;
      const float f1 = ((seed >> 8) & 0xffff) / 65536.f - 0.5f;
      __half2 h = __floats2half2_rn(f0 * 0.25f, f1 * 0.25f);
      v = *reinterpret_cast<uint32_t*>(&h);
    }
    reinterpret_cast<uint32_t*>(smem)[i] = v;
  }
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_init(bar + 1, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc_pair(tmem_ptr, 256);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr;
  const uint32_t rank = cluster_ctarank();
  cluster_sync_all();
  if (warp == 0) {
    if (rank == 0) {
      __syncwarp();
      const long long t0 = clock64();
      const uint32_t idesc = umma_idesc_m256(128, true);
      const uint64_t a_base = umma_smem_desc_sw128(smem_u32(A));
      const uint64_t b_base = umma_smem_desc_sw128(smem_u32(B));
      for (int g = 0; g < nmma / 4; ++g) {
        const uint32_t d = tmem + (((g * 4) / dep) & 1) * 128;
        const uint64_t a0 = a_base + (footprint ? (uint64_t)((g & 3) * (16384 >> 4)) : 0);
        const uint64_t b0 = b_base + (footprint ? (uint64_t)((g & 3) * (8192 >> 4)) : 0);
        const uint32_t fresh = ((g * 4) % 64 == 0) ? 0u : 1u;   // a new accumulation every 64 MMAs (keeps values finite)
        if (elect_one()) {
          tc_mma_pair(d, a0, b0, idesc, fresh);
          tc_mma_pair(d, a0 + 2, b0 + 2, idesc, 1u);
          tc_mma_pair(d, a0 + 4, b0 + 4, idesc, 1u);
          tc_mma_pair(d, a0 + 6, b0 + 6, idesc, 1u);
          tc_commit_pair(bar + 1);
        }
        __syncwarp();
      }
      tc_commit_pair_w(bar);
      __syncwarp();
      mbar_wait(bar, 0, nullptr, 0);
      const long long t2 = clock64();
      if (lane == 0) out[blockIdx.x] = t2 - t0;
    } else {
      mbar_wait(bar, 0, nullptr, 0);
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc_pair(tmem, 256);
  }
}

void run(int nmma, int dep, int data, int footprint) {
  const int smem = 65536 + 32768 + 64 + 100000;   // pad to force one CTA per SM
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int grid = 148;
  long long* out;
  cudaMalloc(&out, grid * sizeof(long long));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(192);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  cudaError_t e = cudaLaunchKernelEx(&cfg, bench, nmma, dep, data, footprint, out);
  cudaEventRecord(e1);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); exit(1); }
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  std::vector<long long> h(grid);
  cudaMemcpy(h.data(), out, grid * sizeof(long long), cudaMemcpyDeviceToHost);
  double done = 0; int cnt = 0;
  for (int b = 0; b < grid; b += 2) { done += h[b]; ++cnt; }
  done /= cnt;
  printf("nmma %8d dep %3d data %s footprint %s: %.1f cyc/MMA (%.0f%% of 4096 MAC/cyc/SM), %.2f ms -> %.0f MHz, %.0f TFLOP/s\n", nmma, dep,
         data ? "random" : "zero  ", footprint ? "4 chunks / 4 slots" : "one tile          ", done / nmma, 100.0 * 64.0 * nmma / done,
         ms, done / ms / 1e3, 2.0 * 256 * 128 * 16 * (double)nmma * 74 / (ms * 1e-3) / 1e12);
  cudaFree(out);
}

int main() {
  for (int nmma : {4096, 400000})
    for (int data = 0; data < 2; ++data)
      for (int fp = 0; fp < 2; ++fp)
        for (int dep : {4, 16, 64}) run(nmma, dep, data, fp);
  return 0;
}

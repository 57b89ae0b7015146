// Micro-benchmark: TMEM -> register read throughput (tcgen05.ld 32x32b) per SM as a function of the number of warps,
// the load width and whether loads are waited one by one.  Diagnostic only; build:
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o scripts/build/tmem_microbench scripts/tmem_microbench.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../sahs-deformable-nerf_b200/csrc/sahs_common.cuh"

void sahs_set_error(const char*, ...) {}
std::atomic<uint64_t> g_sahs_launches{0};
int sahs_num_sms() { return 148; }
int* sahs_status_words(int) { return nullptr; }

// WIDTH: 16 or 32 columns per load; DEPTH: loads in flight before tcgen05.wait::ld
template <int WIDTH, int DEPTH>
__global__ void __launch_bounds__(512, 1) bench(int iters, long long* out, uint32_t* sink) {
  __shared__ uint32_t tmem_slot;
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  if (warp == 0) tmem_alloc(&tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = tmem_slot + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int c = 0; c < 256; c += WIDTH * DEPTH) {
      if (WIDTH == 16) {
        uint32_t v[DEPTH][16];
#pragma unroll
        for (int d = 0; d < DEPTH; ++d) tmem_ld16(base + c + d * 16, v[d]);
        tmem_ld_wait();
#pragma unroll
        for (int d = 0; d < DEPTH; ++d)
#pragma unroll
          for (int j = 0; j < 16; ++j) acc ^= v[d][j];
      } else {
        uint32_t v[DEPTH][32];
#pragma unroll
        for (int d = 0; d < DEPTH; ++d) tmem_ld32(base + c + d * 32, v[d]);
        tmem_ld_wait();
#pragma unroll
        for (int d = 0; d < DEPTH; ++d)
#pragma unroll
          for (int j = 0; j < 32; ++j) acc ^= v[d][j];
      }
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  if (acc == 0x12345678u) sink[0] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_slot, 256); }
}

template <int WIDTH, int DEPTH>
void run(int nwarps) {
  const int iters = 200;
  long long* out;
  uint32_t* sink;
  cudaMalloc(&out, 148 * sizeof(long long));
  cudaMalloc(&sink, 4);
  bench<WIDTH, DEPTH><<<148, nwarps * 32>>>(iters, out, sink);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
  std::vector<long long> h(148);
  cudaMemcpy(h.data(), out, 148 * sizeof(long long), cudaMemcpyDeviceToHost);
  double cyc = 0;
  for (auto v : h) cyc += v;
  cyc /= 148;
  const double bytes = (double)iters * nwarps * 32 * 256 * 4;   // every warp reads its 32 lanes x 256 columns per iteration
  printf("ld 32x32b.x%-2d depth %d warps %2d: %8.0f cycles, %6.1f B/cycle/SM, %5.1f cycles per warp-load\n", WIDTH, DEPTH, nwarps,
         cyc, bytes / cyc, cyc / (iters * (256.0 / WIDTH)));
  cudaFree(out);
  cudaFree(sink);
}

int main() {
  for (int nw : {1, 2, 4, 8, 16}) {
    run<16, 1>(nw);
    run<16, 2>(nw);
    run<16, 4>(nw);
    run<32, 1>(nw);
    run<32, 2>(nw);
  }
  return 0;
}

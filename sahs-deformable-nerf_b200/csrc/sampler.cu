// Semantic-weighted ray sampler on the device.
// ref: train_stage_rays_auto.py:390-420 -- probs = sum_c sample_prob[c] * mask[..., c], normalised, then
//      np.random.choice(H*W, num_random_rays, replace=False, p=probs) on the host (a GPU -> CPU round trip per step).
//
// Sampling n of N items without replacement with probabilities proportional to w_i, drawn sequentially (what
// np.random.choice does), has the same distribution as taking the n smallest keys k_i = E_i / w_i with E_i ~ Exp(1)
// i.i.d. (Efraimidis & Spirakis 2006, "exponential clocks").  So the whole draw is data parallel:
//   1. key kernel      : w_i from the one-hot mask and the 12 class probabilities, E_i from a counter-based RNG
//                        (Philox-4x32-10 keyed by the caller's seed, counter = pixel index), k_i as order-preserving bits;
//   2. radix select    : the n-th smallest key by four 8-bit histogram passes (one CTA-wide histogram per pass);
//   3. compaction      : indices with key < threshold, then ties, appended through one atomic counter.
// HBM-bound: 4*C + 4 bytes read + 4 written per pixel in the key pass, then 4 reads of the 4-byte key.
// The draw is reproducible for a given seed as a SET; the order of the indices in the output is not defined.
#include "sahs_common.cuh"

namespace {

struct SelState {           // lives in the caller's workspace
  uint32_t hist[256];
  uint32_t prefix;          // key bits fixed so far
  uint32_t mask_bits;       // which bits are fixed
  uint32_t remaining;       // rank of the wanted key among the keys that match the prefix
  uint32_t n_below;         // keys strictly below the threshold (filled after the last pass)
  uint32_t count_lt, count_eq;
  uint32_t positive;        // pixels with a positive weight
};

__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
  const uint32_t hi0 = __umulhi(M0, c[0]), lo0 = M0 * c[0];
  const uint32_t hi1 = __umulhi(M1, c[2]), lo1 = M1 * c[2];
  const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}
// Philox-4x32-10 (Salmon et al. 2011): counter (i, 0, 0, 0), key = seed
__device__ __forceinline__ uint32_t philox_u32(uint64_t i, uint64_t seed) {
  uint32_t c[4] = {(uint32_t)i, (uint32_t)(i >> 32), 0u, 0u};
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return c[0];
}

__global__ void sample_key_kernel(const int32_t* __restrict__ mask, const float* __restrict__ class_prob, int64_t n,
                                  int C, int n_select, uint64_t seed, const unsigned long long* __restrict__ seed_dev,
                                  uint32_t* __restrict__ keys, SelState* st) {
  if (seed_dev) seed += *seed_dev;   // capturable form: the per-step part of the seed lives in device memory
  __shared__ float prob[32];
  __shared__ uint32_t pos;
  if (threadIdx.x < 32) prob[threadIdx.x] = threadIdx.x < C ? class_prob[threadIdx.x] : 0.f;
  if (threadIdx.x == 0) pos = 0;
  if (blockIdx.x == 0 && threadIdx.x == 0) st->remaining = (uint32_t)(n_select - 1);   // 0-based rank of the wanted key
  __syncthreads();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    float w = 0.f;
    for (int c = 0; c < C; ++c) w += prob[c] * (float)mask[i * C + c];
    uint32_t key = 0x7F800000u;                                  // +inf: never selected
    if (w > 0.f) {
      const float u = ((float)(philox_u32((uint64_t)i, seed) >> 8) + 0.5f) * (1.0f / 16777216.0f);   // (0, 1)
      const float k = -logf(u) / w;                            // Exp(1) / w  (positive finite)
      key = __float_as_uint(fminf(k, 3.0e38f));                  // positive floats order like their bit patterns
      atomicAdd(&pos, 1u);
    }
    keys[i] = key;
  }
  __syncthreads();
  if (threadIdx.x == 0 && pos) atomicAdd(&st->positive, pos);
}

// histogram of byte `shift/8` of the keys that match the already fixed prefix
__global__ void sample_hist_kernel(const uint32_t* __restrict__ keys, int64_t n, int shift, SelState* st) {
  __shared__ uint32_t h[256];
  h[threadIdx.x] = 0;   // blockDim.x == 256
  __syncthreads();
  const uint32_t prefix = st->prefix, mbits = st->mask_bits;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t k = keys[i];
    if ((k & mbits) == prefix) atomicAdd(&h[(k >> shift) & 255u], 1u);
  }
  __syncthreads();
  if (h[threadIdx.x]) atomicAdd(&st->hist[threadIdx.x], h[threadIdx.x]);
}

// picks the bin that holds the wanted rank, fixes its byte in the prefix, clears the histogram
__global__ void sample_pick_kernel(int shift, SelState* st) {
  if (threadIdx.x != 0) return;
  uint32_t rem = st->remaining, acc = 0;
  int bin = 255;
  for (int b = 0; b < 256; ++b) {
    const uint32_t c = st->hist[b];
    if (acc + c > rem) { bin = b; break; }
    acc += c;
  }
  st->remaining = rem - acc;
  st->prefix |= (uint32_t)bin << shift;
  st->mask_bits |= 255u << shift;
  st->n_below += acc;                 // keys in lower bins of this pass are strictly below the final threshold
  for (int b = 0; b < 256; ++b) st->hist[b] = 0;
}

__global__ void sample_compact_kernel(const uint32_t* __restrict__ keys, int64_t n, int n_select, SelState* st,
                                      int64_t* __restrict__ out) {
  const uint32_t thr = st->prefix;            // the n_select-th smallest key
  const uint32_t n_lt = st->n_below;          // keys strictly below it (< n_select)
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t k = keys[i];
    if (k < thr) {
      out[atomicAdd(&st->count_lt, 1u)] = i;
    } else if (k == thr) {
      const uint32_t slot = n_lt + atomicAdd(&st->count_eq, 1u);   // ties fill what is left
      if (slot < (uint32_t)n_select) out[slot] = i;
    }
  }
}

}  // namespace

static int weighted_sample_impl(const int32_t* mask, const float* class_prob, int64_t num_pixels, int num_classes,
                                int num_select, uint64_t seed, const unsigned long long* seed_dev, int64_t* out_indices,
                                void* workspace, size_t workspace_bytes, void* stream) {
  SAHS_CHECK_ARG(num_pixels >= 0 && num_select >= 0 && num_select <= num_pixels, "bad extents");
  SAHS_CHECK_ARG(num_classes >= 1 && num_classes <= 32, "1..32 classes");
  if (num_select == 0) return SAHS_OK;
  SAHS_CHECK_ARG(mask && class_prob && out_indices && workspace, "null pointer");
  constexpr size_t kHeader = 2048;
  static_assert(sizeof(SelState) <= kHeader, "state must fit the reserved header");
  const size_t need = kHeader + (size_t)num_pixels * sizeof(uint32_t);
  SAHS_CHECK_ARG(workspace_bytes >= need, "workspace too small (2048 + 4 * num_pixels bytes)");
  cudaStream_t st = (cudaStream_t)stream;
  SelState* state = (SelState*)workspace;
  uint32_t* keys = (uint32_t*)((uint8_t*)workspace + kHeader);
  SAHS_CUDA(cudaMemsetAsync(state, 0, sizeof(SelState), st));
  const unsigned blocks = (unsigned)((num_pixels + 255) / 256);
  sample_key_kernel<<<blocks, 256, 0, st>>>(mask, class_prob, num_pixels, num_classes, num_select, seed, seed_dev, keys,
                                            state);
  SAHS_LAUNCH_CHECK();
  const unsigned hb = blocks < 592u ? blocks : 592u;
  for (int shift = 24; shift >= 0; shift -= 8) {
    sample_hist_kernel<<<hb, 256, 0, st>>>(keys, num_pixels, shift, state);
    SAHS_LAUNCH_CHECK();
    sample_pick_kernel<<<1, 32, 0, st>>>(shift, state);
    SAHS_LAUNCH_CHECK();
  }
  sample_compact_kernel<<<hb, 256, 0, st>>>(keys, num_pixels, num_select, state, out_indices);
  SAHS_LAUNCH_CHECK();
  return SAHS_OK;
}

extern "C" int sahs_weighted_sample(const int32_t* mask, const float* class_prob, int64_t num_pixels, int num_classes,
                                    int num_select, uint64_t seed, int64_t* out_indices, void* workspace,
                                    size_t workspace_bytes, void* stream) {
  return weighted_sample_impl(mask, class_prob, num_pixels, num_classes, num_select, seed, nullptr, out_indices, workspace,
                              workspace_bytes, stream);
}

extern "C" int sahs_weighted_sample_dev(const int32_t* mask, const float* class_prob, int64_t num_pixels, int num_classes,
                                        int num_select, uint64_t seed, const unsigned long long* seed_counter_dev,
                                        int64_t* out_indices, void* workspace, size_t workspace_bytes, void* stream) {
  SAHS_CHECK_ARG(seed_counter_dev, "null seed counter");
  return weighted_sample_impl(mask, class_prob, num_pixels, num_classes, num_select, seed, seed_counter_dev, out_indices,
                              workspace, workspace_bytes, stream);
}

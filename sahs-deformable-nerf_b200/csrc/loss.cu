// Stage-I training loss, forward and backward in two small launches.
// ref: nerf/nerf_helpers.py:14-62 (MaskCrossEntropyLoss, MaskMSELoss) and their assembly in
//      train_stage_rays_auto.py:455-468:
//   per level (coarse, fine):  diff_i = sum_c (rgb_ic - target_ic)^2,  ce_i = -sum_k mask_ik * log(p_ik + 1e-10)
//     l2 = mean_i diff_i,  c = mean_i ce_i,  count_k = max(1, #{i: mask_ik != 0})
//     m_l2[k] = sum_i diff_i * mask_ik / count_k,  m_c[k] = sum_i ce_i * mask_ik / count_k
//     level loss = l2 + ce_weight * c + mouth_weight * sum_{k in [mouth_lo, mouth_hi)} (m_l2[k] + m_c[k])
//   sample_prob = s / sum(s),  s = sum over levels of (m_l2 + m_c)        (the dynamic per-class sampling weight)
// The reference spends ~75 elementwise/reduction launches and as many autograd nodes on this per step; a 2048-ray
// batch is 0.6 MB, so the whole thing is launch bound.  Two launches: (1) every CTA reduces the 64 sums of its rays
// (registers -> warp shuffles -> fixed-order sum over the warps) into one row of a small workspace; (2) every CTA adds
// the rows in order (deterministic result), forms the per-class normalisers, and writes d loss / d map of its rays;
// CTA 0 also writes the loss, the statistics and sample_prob.  The cross entropy's target is the mask itself, as at
// the call site.
#include "sahs_common.cuh"

namespace {

constexpr int kC = 12;                 // semantic classes
constexpr int kMapCh = 3 + kC;         // rgb + class probabilities per ray
constexpr int kLossThreads = 256;
constexpr int kLossMaxBlocks = 64;     // rows of the workspace
constexpr int kAcc = kC + 2 * (2 + 2 * kC);   // counts, then per level: l2 sum, ce sum, m_l2[12], m_c[12]  (= 64)

__global__ void __launch_bounds__(kLossThreads)
stage1_loss_sums_kernel(const float* __restrict__ map_c, const float* __restrict__ map_f,
                        const float* __restrict__ target, const float* __restrict__ mask, int R,
                        float* __restrict__ partial) {
  __shared__ float part[kLossThreads / 32][kAcc];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int levels = map_f ? 2 : 1;
  float acc[kAcc];
#pragma unroll
  for (int a = 0; a < kAcc; ++a) acc[a] = 0.f;
  for (int i = blockIdx.x * kLossThreads + tid; i < R; i += gridDim.x * kLossThreads) {
    float m[kC];
#pragma unroll
    for (int k = 0; k < kC; ++k) {
      m[k] = mask[(size_t)i * kC + k];
      acc[k] += (m[k] != 0.f) ? 1.f : 0.f;
    }
    const float t0 = target[(size_t)i * 3], t1 = target[(size_t)i * 3 + 1], t2 = target[(size_t)i * 3 + 2];
#pragma unroll
    for (int lv = 0; lv < 2; ++lv) {
      if (lv < levels) {
        const float* row = (lv == 0 ? map_c : map_f) + (size_t)i * kMapCh;
        const float e0 = row[0] - t0, e1 = row[1] - t1, e2 = row[2] - t2;
        const float diff = e0 * e0 + e1 * e1 + e2 * e2;
        float ce = 0.f;
#pragma unroll
        for (int k = 0; k < kC; ++k)
          if (m[k] != 0.f) ce -= m[k] * logf(row[3 + k] + 1e-10f);
        float* a = acc + kC + lv * (2 + 2 * kC);
        a[0] += diff;
        a[1] += ce;
#pragma unroll
        for (int k = 0; k < kC; ++k) {
          a[2 + k] += diff * m[k];
          a[2 + kC + k] += ce * m[k];
        }
      }
    }
  }
#pragma unroll
  for (int a = 0; a < kAcc; ++a) {
    float v = acc[a];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) part[warp][a] = v;
  }
  __syncthreads();
  if (tid < kAcc) {
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < kLossThreads / 32; ++w) v += part[w][tid];
    partial[blockIdx.x * kAcc + tid] = v;
  }
}

__global__ void __launch_bounds__(kLossThreads)
stage1_loss_grad_kernel(const float* __restrict__ map_c, const float* __restrict__ map_f,
                        const float* __restrict__ target, const float* __restrict__ mask, int R, float ce_weight,
                        float mouth_weight, int mouth_lo, int mouth_hi, const float* __restrict__ partial,
                        float* __restrict__ stats, float* __restrict__ sample_prob, float* __restrict__ d_c,
                        float* __restrict__ d_f) {
  __shared__ float tot[kAcc];
  __shared__ float inv_count[kC];
  const int tid = threadIdx.x;
  const int levels = map_f ? 2 : 1;
  if (tid < kAcc) {
    float v = 0.f;
    for (int b = 0; b < (int)gridDim.x; ++b) v += partial[b * kAcc + tid];     // same order in every CTA
    tot[tid] = v;
  }
  __syncthreads();
  if (tid < kC) inv_count[tid] = 1.f / fmaxf(tot[tid], 1.f);
  __syncthreads();
  const float inv_r = 1.f / (float)R;
  if (blockIdx.x == 0 && tid == 0) {
    // stats: [0] total loss, per level l: [1+2l] l2, [2+2l] ce, then m_l2 / m_c per level at 5 + l*24
    float loss = 0.f, s[kC], ssum = 0.f;
    for (int k = 0; k < kC; ++k) s[k] = 0.f;
    for (int lv = 0; lv < levels; ++lv) {
      const float* a = tot + kC + lv * (2 + 2 * kC);
      const float l2 = a[0] * inv_r, c = a[1] * inv_r;
      float mouth = 0.f;
      for (int k = 0; k < kC; ++k) {
        const float ml2 = a[2 + k] * inv_count[k], mc = a[2 + kC + k] * inv_count[k];
        stats[5 + lv * 2 * kC + k] = ml2;
        stats[5 + lv * 2 * kC + kC + k] = mc;
        s[k] += ml2 + mc;
        if (k >= mouth_lo && k < mouth_hi) mouth += ml2 + mc;
      }
      stats[1 + 2 * lv] = l2;
      stats[2 + 2 * lv] = c;
      loss += l2 + ce_weight * c + mouth_weight * mouth;
    }
    stats[0] = loss;
    for (int k = 0; k < kC; ++k) ssum += s[k];
    for (int k = 0; k < kC; ++k) sample_prob[k] = s[k] / ssum;
  }
  for (int i = blockIdx.x * kLossThreads + tid; i < R; i += gridDim.x * kLossThreads) {
    float m[kC], mouth = 0.f;
#pragma unroll
    for (int k = 0; k < kC; ++k) {
      m[k] = mask[(size_t)i * kC + k];
      if (k >= mouth_lo && k < mouth_hi) mouth += m[k] * inv_count[k];
    }
    const float g_diff = inv_r + mouth_weight * mouth;              // d loss / d diff_i
    const float g_ce = ce_weight * inv_r + mouth_weight * mouth;    // d loss / d ce_i
    const float t[3] = {target[(size_t)i * 3], target[(size_t)i * 3 + 1], target[(size_t)i * 3 + 2]};
#pragma unroll
    for (int lv = 0; lv < 2; ++lv) {
      if (lv < levels) {
        const float* row = (lv == 0 ? map_c : map_f) + (size_t)i * kMapCh;
        float* drow = (lv == 0 ? d_c : d_f) + (size_t)i * kMapCh;
#pragma unroll
        for (int c = 0; c < 3; ++c) drow[c] = 2.f * g_diff * (row[c] - t[c]);
#pragma unroll
        for (int k = 0; k < kC; ++k) drow[3 + k] = (m[k] != 0.f) ? -g_ce * m[k] / (row[3 + k] + 1e-10f) : 0.f;
      }
    }
  }
}

}  // namespace

extern "C" int sahs_stage1_loss(const float* map_coarse, const float* map_fine, const float* target_rgb,
                                const float* mask, int num_rays, int num_classes, float ce_weight, float mouth_weight,
                                int mouth_lo, int mouth_hi, float* stats, float* sample_prob, float* d_map_coarse,
                                float* d_map_fine, float* workspace, void* stream) {
  SAHS_CHECK_ARG(num_rays >= 1, "at least one ray");
  SAHS_CHECK_ARG(num_classes == kC, "the maps carry 3 colour + 12 class channels");
  SAHS_CHECK_ARG(map_coarse && target_rgb && mask && stats && sample_prob && d_map_coarse && workspace, "null pointer");
  SAHS_CHECK_ARG(!map_fine || d_map_fine, "a fine map needs a fine gradient buffer");
  SAHS_CHECK_ARG(mouth_lo >= 0 && mouth_hi <= kC && mouth_lo <= mouth_hi, "bad mouth class range");
  int blocks = (num_rays + kLossThreads - 1) / kLossThreads;
  if (blocks > kLossMaxBlocks) blocks = kLossMaxBlocks;
  cudaStream_t st = (cudaStream_t)stream;
  stage1_loss_sums_kernel<<<blocks, kLossThreads, 0, st>>>(map_coarse, map_fine, target_rgb, mask, num_rays, workspace);
  SAHS_LAUNCH_CHECK();
  stage1_loss_grad_kernel<<<blocks, kLossThreads, 0, st>>>(map_coarse, map_fine, target_rgb, mask, num_rays, ce_weight,
                                                         mouth_weight, mouth_lo, mouth_hi, workspace, stats,
                                                         sample_prob, d_map_coarse, d_map_fine);
  SAHS_LAUNCH_CHECK();
  return SAHS_OK;
}

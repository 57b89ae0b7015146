// (3) alpha compositing: one warp per ray, lane = sample (mod 32), warp-level exclusive product scan.
// ref: nerf/volume_rendering_utils.py:7-78, nerf/nerf_helpers.py:99-120 (cumprod_exclusive),
//      nerf/train_utils.py:135-136 (background overwrite of the last sample).
// HBM-bound: 72*S + 144 algorithmic bytes per ray forward.
#include "sahs_common.cuh"

namespace {

constexpr int kWarpsPerBlock = 8;
constexpr int kMaxChunks = 8;  // S <= 256

// exp / reciprocal of the forward kernel: ex2.approx + rcp.approx (<= 2 ulp each; composited maps move by ~1e-7, far
// inside the 2e-5 parity bar) -- the accurate expf / IEEE division were a third of the kernel's instructions, and the
// kernel is issue-bound, not HBM-bound (ncu r1: issue active 65 %, DRAM 43 %).
__device__ __forceinline__ float fast_exp(float x) { return __expf(x); }
__device__ __forceinline__ float fast_rcp(float x) { return __fdividef(1.0f, x); }

// Sums of 16 per-lane values over the warp by recursive halving: 8 + 4 + 2 + 1 + 1 = 16 shuffles instead of 16 x 5.
// On return lane L holds the warp total of v[L >> 1] (in v[0]).
__device__ __forceinline__ void warp_sum16(float (&v)[16], int lane) {
#pragma unroll
  for (int half = 8, bit = 16; half >= 1; half >>= 1, bit >>= 1) {
    const bool up = (lane & bit) != 0;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      const float send = up ? v[i] : v[i + half];
      const float keep = up ? v[i + half] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
    }
  }
  v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

struct SampleColor {
  float c[SAHS_MAP_CH];
};

// activation of the 15 colour/semantic channels of one sample (last_raw: pass-through of raw values)
template <bool FAST = false>
__device__ __forceinline__ void activate(const float (&raw)[16], bool passthrough, bool seg_softmax, SampleColor& out) {
  auto ex = [](float x) { return FAST ? fast_exp(x) : expf(x); };
  auto rc = [](float x) { return FAST ? fast_rcp(x) : 1.0f / x; };
  if (passthrough) {
#pragma unroll
    for (int k = 0; k < SAHS_MAP_CH; ++k) out.c[k] = raw[k];
    return;
  }
  if (seg_softmax) {
#pragma unroll
    for (int k = 0; k < 3; ++k) out.c[k] = rc(1.0f + ex(-raw[k]));
    float m = raw[3];
#pragma unroll
    for (int k = 4; k < SAHS_MAP_CH; ++k) m = fmaxf(m, raw[k]);
    float den = 0.f;
#pragma unroll
    for (int k = 3; k < SAHS_MAP_CH; ++k) {
      out.c[k] = ex(raw[k] - m);
      den += out.c[k];
    }
    float inv = rc(den);
#pragma unroll
    for (int k = 3; k < SAHS_MAP_CH; ++k) out.c[k] *= inv;
  } else {
#pragma unroll
    for (int k = 0; k < SAHS_MAP_CH; ++k) out.c[k] = rc(1.0f + ex(-raw[k]));
  }
}

__device__ __forceinline__ void load_raw(const float* __restrict__ p, float (&raw)[16]) {
  const float4* q = reinterpret_cast<const float4*>(p);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float4 v = __ldg(q + k);
    raw[4 * k + 0] = v.x; raw[4 * k + 1] = v.y; raw[4 * k + 2] = v.z; raw[4 * k + 3] = v.w;
  }
}

// density noise of sample `idx`: the caller's tensor, or noise_std * N(0,1) drawn here (same value in fwd and bwd)
__device__ __forceinline__ float sample_noise(const float* __restrict__ noise, float noise_std, const RngArg& rng, size_t idx) {
  if (noise) return noise[idx];
  return rng.on ? noise_std * rng_normal(rng, (uint64_t)idx) : 0.f;
}

// per-sample density terms shared by forward and backward
struct Density {
  float sigma, delta, alpha, t;  // t = 1 - alpha + 1e-10
  bool on;                       // relu active
};

template <bool FAST = false>
__device__ __forceinline__ Density density(float raw_sigma, float noise, float z, float z_next, bool last, float rd_norm) {
  Density d;
  float pre = raw_sigma + noise;
  d.on = pre > 0.f;
  d.sigma = d.on ? pre : 0.f;
  if (last) d.sigma += 1e-6f;
  d.delta = (last ? 1e10f : (z_next - z)) * rd_norm;
  d.alpha = 1.0f - (FAST ? fast_exp(-d.sigma * d.delta) : expf(-d.sigma * d.delta));
  d.t = 1.0f - d.alpha + 1e-10f;
  return d;
}

template <int CHUNKS>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, 4)
composite_fwd_kernel(const float* __restrict__ raw, const float* __restrict__ z, const float* __restrict__ rd,
                     const float* __restrict__ noise, float noise_std, RngArg rng, const float* __restrict__ bg,
                     int apply_bg, int R, int S,
                     int white, float* __restrict__ rgb_map, float* __restrict__ disp, float* __restrict__ acc,
                     float* __restrict__ weights, float* __restrict__ depth) {
  const int lane = threadIdx.x & 31;
  const int warps_total = gridDim.x * kWarpsPerBlock;
  for (int r = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5); r < R; r += warps_total) {
    const float dx = rd[(size_t)r * 3], dy = rd[(size_t)r * 3 + 1], dz = rd[(size_t)r * 3 + 2];
    const float rd_norm = sqrtf(dx * dx + dy * dy + dz * dz);
    const float* zr = z + (size_t)r * S;
    // background of the last sample: fetched by 15 lanes at the start of the ray (one coalesced request) and handed to
    // the last lane by shuffles -- as 15 dependent-latency loads by the last lane alone it stalled the warp once per ray
    const float bgv = (bg && apply_bg && lane < SAHS_MAP_CH) ? __ldg(bg + (size_t)r * SAHS_MAP_CH + lane) : 0.f;
    const int last_chunk = (S - 1) >> 5;
    float carry = 1.0f;  // transmittance entering this 32-sample chunk
    float accum[SAHS_MAP_CH + 2];
#pragma unroll
    for (int k = 0; k < SAHS_MAP_CH + 2; ++k) accum[k] = 0.f;
#pragma unroll
    for (int m = 0; m < CHUNKS; ++m) {
      const int s = m * 32 + lane;
      const bool valid = s < S;
      const bool last = s == S - 1;
      float w = 0.f, zs = 0.f;
      SampleColor col;
      float t = 1.0f, alpha = 0.f;
      if (valid) {
        float rw[16];
        load_raw(raw + ((size_t)r * S + s) * SAHS_RAW_CH, rw);
        const float raw_sigma = rw[15];
        activate<true>(rw, bg != nullptr && last, bg != nullptr, col);
        zs = zr[s];
        float zn = last ? 0.f : zr[s + 1];
        Density d = density<true>(raw_sigma, sample_noise(noise, noise_std, rng, (size_t)r * S + s), zs, zn, last, rd_norm);
        t = d.t;
        alpha = d.alpha;
      }
      if (m == last_chunk && bg && apply_bg) {   // warp-uniform: the chunk that holds the ray's last sample
#pragma unroll
        for (int k = 0; k < SAHS_MAP_CH; ++k) {
          const float b = __shfl_sync(0xffffffffu, bgv, k);
          if (last) col.c[k] = b;                 // background overwrite of the last sample (passed through raw)
        }
      }
      // inclusive product scan of t over the warp, then shift to exclusive
      float incl = t;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        float up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl *= up;
      }
      float excl = __shfl_up_sync(0xffffffffu, incl, 1);
      if (lane == 0) excl = 1.0f;
      float T = carry * excl;
      carry *= __shfl_sync(0xffffffffu, incl, 31);
      if (valid) {
        w = alpha * T;
        weights[(size_t)r * S + s] = w;
#pragma unroll
        for (int k = 0; k < SAHS_MAP_CH; ++k) accum[k] += w * col.c[k];
        accum[SAHS_MAP_CH] += w * zs;
        accum[SAHS_MAP_CH + 1] += w;
      }
    }
    // 15 channel sums + the depth sum by recursive halving (lane L ends up with the total of value L >> 1), acc apart
    const float a = warp_sum(accum[SAHS_MAP_CH + 1]);
    float v16[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) v16[k] = accum[k];          // [0,15) channels, [15] = sum w z
    warp_sum16(v16, lane);
    const float tot = v16[0];
    const int ch = lane >> 1;
    if (!(lane & 1)) {
      if (ch < SAHS_MAP_CH) {
        rgb_map[(size_t)r * SAHS_MAP_CH + ch] = white ? tot + (1.0f - a) : tot;
      } else {
        depth[r] = tot;
        acc[r] = a;
        disp[r] = 1.0f / fmaxf(1e-10f, tot / a);
      }
    }
  }
}

// Backward w.r.t. raw.  Recomputes the forward per ray (cheaper than storing T and colours), then
//   g_w[s]   = d_w[s] + sum_c d_rgb[c]*col[s,c] + d_depth' * z[s] + d_acc'
//   dL/da[s] = g_w[s]*T[s] - (sum_{k>s} g_w[k]*w[k]) / t[s]          (t = 1 - alpha + 1e-10)
//   dL/dsig  = dL/da * delta * exp(-sigma*delta), gated by the relu.
template <int CHUNKS>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
composite_bwd_kernel(const float* __restrict__ raw, const float* __restrict__ z, const float* __restrict__ rd,
                     const float* __restrict__ noise, float noise_std, RngArg rng, const float* __restrict__ bg,
                     int apply_bg, int R, int S,
                     int white, const float* __restrict__ d_rgb_map, const float* __restrict__ d_disp,
                     const float* __restrict__ d_acc, const float* __restrict__ d_weights,
                     const float* __restrict__ d_depth, float* __restrict__ d_raw) {
  const int lane = threadIdx.x & 31;
  const int warps_total = gridDim.x * kWarpsPerBlock;
  for (int r = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5); r < R; r += warps_total) {
    const float dx = rd[(size_t)r * 3], dy = rd[(size_t)r * 3 + 1], dz = rd[(size_t)r * 3 + 2];
    const float rd_norm = sqrtf(dx * dx + dy * dy + dz * dz);
    const float* zr = z + (size_t)r * S;
    float g_rgb[SAHS_MAP_CH];
    float g_rgb_sum = 0.f;
#pragma unroll
    for (int k = 0; k < SAHS_MAP_CH; ++k) {
      g_rgb[k] = d_rgb_map ? d_rgb_map[(size_t)r * SAHS_MAP_CH + k] : 0.f;
      g_rgb_sum += g_rgb[k];
    }
    // pass 1: forward quantities kept in registers per chunk
    float w_[CHUNKS], T_[CHUNKS], t_[CHUNKS], zs_[CHUNKS], dadsig_[CHUNKS], gcol_[CHUNKS];
    float carry = 1.0f, dep = 0.f, a = 0.f;
#pragma unroll
    for (int m = 0; m < CHUNKS; ++m) {
      const int s = m * 32 + lane;
      const bool valid = s < S, last = s == S - 1;
      float t = 1.0f, alpha = 0.f;
      zs_[m] = 0.f; dadsig_[m] = 0.f; gcol_[m] = 0.f;
      if (valid) {
        float rw[16];
        load_raw(raw + ((size_t)r * S + s) * SAHS_RAW_CH, rw);
        const float raw_sigma = rw[15];
        if (last && bg && apply_bg) {
#pragma unroll
          for (int k = 0; k < SAHS_MAP_CH; ++k) rw[k] = bg[(size_t)r * SAHS_MAP_CH + k];
        }
        SampleColor col;
        activate(rw, bg != nullptr && last, bg != nullptr, col);
#pragma unroll
        for (int k = 0; k < SAHS_MAP_CH; ++k) gcol_[m] += g_rgb[k] * col.c[k];
        zs_[m] = zr[s];
        float zn = last ? 0.f : zr[s + 1];
        Density d = density(raw_sigma, sample_noise(noise, noise_std, rng, (size_t)r * S + s), zs_[m], zn, last, rd_norm);
        t = d.t; alpha = d.alpha;
        dadsig_[m] = d.on ? d.delta * expf(-d.sigma * d.delta) : 0.f;
      }
      float incl = t;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        float up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl *= up;
      }
      float excl = __shfl_up_sync(0xffffffffu, incl, 1);
      if (lane == 0) excl = 1.0f;
      T_[m] = carry * excl;
      carry *= __shfl_sync(0xffffffffu, incl, 31);
      t_[m] = t;
      w_[m] = valid ? alpha * T_[m] : 0.f;
      dep += w_[m] * zs_[m];
      a += w_[m];
    }
    dep = warp_sum(dep);
    a = warp_sum(a);
    float gd = d_depth ? d_depth[r] : 0.f;
    float ga = d_acc ? d_acc[r] : 0.f;
    if (d_disp) {
      float q = dep / a;
      if (q > 1e-10f) {
        float gq = -d_disp[r] / (q * q);
        gd += gq / a;
        ga += -gq * dep / (a * a);
      }
    }
    if (white) ga -= g_rgb_sum;
    // pass 2: suffix sums of g_w*w from the far end of the ray
    float suffix_carry = 0.f;
#pragma unroll
    for (int m = CHUNKS - 1; m >= 0; --m) {
      const int s = m * 32 + lane;
      const bool valid = s < S, last = s == S - 1;
      float gw = 0.f;
      if (valid) gw = (d_weights ? d_weights[(size_t)r * S + s] : 0.f) + gcol_[m] + gd * zs_[m] + ga;
      float x = gw * w_[m];
      float incl = x;  // inclusive suffix scan (towards higher lanes)
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        float dn = __shfl_down_sync(0xffffffffu, incl, o);
        if (lane + o < 32) incl += dn;
      }
      float excl = incl - x + suffix_carry;
      suffix_carry += __shfl_sync(0xffffffffu, incl, 0);
      if (valid) {
        float g_alpha = gw * T_[m] - excl / t_[m];
        float g_sigma = g_alpha * dadsig_[m];
        // colour channels: recompute activations for the local jacobian
        float rw[16];
        load_raw(raw + ((size_t)r * S + s) * SAHS_RAW_CH, rw);
        float out[16];
        const bool passthrough = bg != nullptr && last;
        if (passthrough) {
#pragma unroll
          for (int k = 0; k < SAHS_MAP_CH; ++k) out[k] = (apply_bg ? 0.f : w_[m] * g_rgb[k]);
        } else {
          SampleColor col;
          activate(rw, false, bg != nullptr, col);
          const int nsig = bg ? 3 : SAHS_MAP_CH;
#pragma unroll
          for (int k = 0; k < SAHS_MAP_CH; ++k)
            if (k < nsig) out[k] = w_[m] * g_rgb[k] * col.c[k] * (1.0f - col.c[k]);
          if (bg) {
            float dot = 0.f;
#pragma unroll
            for (int k = 3; k < SAHS_MAP_CH; ++k) dot += g_rgb[k] * col.c[k];
#pragma unroll
            for (int k = 3; k < SAHS_MAP_CH; ++k) out[k] = w_[m] * col.c[k] * (g_rgb[k] - dot);
          }
        }
        out[15] = g_sigma;
        float4* o4 = reinterpret_cast<float4*>(d_raw + ((size_t)r * S + s) * SAHS_RAW_CH);
#pragma unroll
        for (int k = 0; k < 4; ++k) o4[k] = make_float4(out[4 * k], out[4 * k + 1], out[4 * k + 2], out[4 * k + 3]);
      }
    }
  }
}

int pick_grid(int R) {
  int blocks = (R + kWarpsPerBlock - 1) / kWarpsPerBlock;
  int cap = sahs_num_sms() * 8;  // 8 blocks x 8 warps = 64 warps per SM
  return blocks < cap ? (blocks > 0 ? blocks : 1) : cap;
}

}  // namespace

#define DISPATCH_CHUNKS(S, BODY)                    \
  do {                                              \
    int chunks__ = ((S) + 31) / 32;                 \
    if (chunks__ <= 1) { constexpr int C = 1; BODY; }      \
    else if (chunks__ <= 2) { constexpr int C = 2; BODY; } \
    else if (chunks__ <= 4) { constexpr int C = 4; BODY; } \
    else { constexpr int C = 8; BODY; }                    \
  } while (0)

static int composite_fwd_impl(const float* raw, const float* z, const float* rd, const float* noise, float noise_std,
                              const sahs_rng* rng, const float* bg, int bg_ch, int apply_bg_overwrite, int num_rays,
                              int num_samples, int white_background, float* rgb_map, float* disp, float* acc,
                              float* weights, float* depth, void* stream) {
  if (num_rays == 0) return SAHS_OK;
  SAHS_CHECK_ARG(raw && z && rd && rgb_map && disp && acc && weights && depth, "null pointer");
  SAHS_CHECK_ARG(num_samples >= 1 && num_samples <= 32 * kMaxChunks, "num_samples must be in [1,256]");
  SAHS_CHECK_ARG(bg == nullptr || bg_ch == SAHS_MAP_CH, "background_prior must have 15 channels");
  if (num_rays == 0) return SAHS_OK;
  int grid = pick_grid(num_rays);
  DISPATCH_CHUNKS(num_samples, (composite_fwd_kernel<C><<<grid, kWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(
                                   raw, z, rd, noise, noise_std, rng_arg(rng), bg, apply_bg_overwrite, num_rays, num_samples,
                                   white_background, rgb_map, disp, acc, weights, depth)));
  SAHS_LAUNCH_CHECK();
  return SAHS_OK;
}

extern "C" int sahs_composite_fwd(const float* raw, const float* z, const float* rd, const float* noise,
                                  const float* bg, int bg_ch, int apply_bg_overwrite, int num_rays, int num_samples,
                                  int white_background, float* rgb_map, float* disp, float* acc, float* weights,
                                  float* depth, void* stream) {
  return composite_fwd_impl(raw, z, rd, noise, 0.f, nullptr, bg, bg_ch, apply_bg_overwrite, num_rays, num_samples,
                            white_background, rgb_map, disp, acc, weights, depth, stream);
}

extern "C" int sahs_composite_fwd_rng(const float* raw, const float* z, const float* rd, float noise_std,
                                      const sahs_rng* rng, const float* bg, int bg_ch, int apply_bg_overwrite,
                                      int num_rays, int num_samples, int white_background, float* rgb_map, float* disp,
                                      float* acc, float* weights, float* depth, void* stream) {
  return composite_fwd_impl(raw, z, rd, nullptr, noise_std, noise_std > 0.f ? rng : nullptr, bg, bg_ch, apply_bg_overwrite,
                            num_rays, num_samples, white_background, rgb_map, disp, acc, weights, depth, stream);
}

static int composite_bwd_impl(const float* raw, const float* z, const float* rd, const float* noise, float noise_std,
                              const sahs_rng* rng, const float* bg, int bg_ch, int apply_bg_overwrite, int num_rays,
                              int num_samples, int white_background, const float* d_rgb_map, const float* d_disp,
                              const float* d_acc, const float* d_weights, const float* d_depth, float* d_raw,
                              void* stream) {
  if (num_rays == 0) return SAHS_OK;
  SAHS_CHECK_ARG(raw && z && rd && d_raw, "null pointer");
  SAHS_CHECK_ARG(num_samples >= 1 && num_samples <= 32 * kMaxChunks, "num_samples must be in [1,256]");
  SAHS_CHECK_ARG(bg == nullptr || bg_ch == SAHS_MAP_CH, "background_prior must have 15 channels");
  int grid = pick_grid(num_rays);
  DISPATCH_CHUNKS(num_samples, (composite_bwd_kernel<C><<<grid, kWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(
                                   raw, z, rd, noise, noise_std, rng_arg(rng), bg, apply_bg_overwrite, num_rays, num_samples,
                                   white_background, d_rgb_map, d_disp, d_acc, d_weights, d_depth, d_raw)));
  SAHS_LAUNCH_CHECK();
  return SAHS_OK;
}

extern "C" int sahs_composite_bwd(const float* raw, const float* z, const float* rd, const float* noise,
                                  const float* bg, int bg_ch, int apply_bg_overwrite, int num_rays, int num_samples,
                                  int white_background, const float* d_rgb_map, const float* d_disp,
                                  const float* d_acc, const float* d_weights, const float* d_depth, float* d_raw,
                                  void* stream) {
  return composite_bwd_impl(raw, z, rd, noise, 0.f, nullptr, bg, bg_ch, apply_bg_overwrite, num_rays, num_samples,
                            white_background, d_rgb_map, d_disp, d_acc, d_weights, d_depth, d_raw, stream);
}

extern "C" int sahs_composite_bwd_rng(const float* raw, const float* z, const float* rd, float noise_std,
                                      const sahs_rng* rng, const float* bg, int bg_ch, int apply_bg_overwrite,
                                      int num_rays, int num_samples, int white_background, const float* d_rgb_map,
                                      const float* d_disp, const float* d_acc, const float* d_weights,
                                      const float* d_depth, float* d_raw, void* stream) {
  return composite_bwd_impl(raw, z, rd, nullptr, noise_std, noise_std > 0.f ? rng : nullptr, bg, bg_ch, apply_bg_overwrite,
                            num_rays, num_samples, white_background, d_rgb_map, d_disp, d_acc, d_weights, d_depth, d_raw,
                            stream);
}

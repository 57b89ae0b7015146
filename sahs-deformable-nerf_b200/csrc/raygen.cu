// (1) ray generation, stratified depths, positional encoding -- standalone kernels.
// The same arithmetic is used inline by the fused field kernel (field_fwd.cu); these entry points exist for
// the reference's public helpers (get_ray_bundle, positional_encoding) and for unit parity.
#include "sahs_common.cuh"

// ref: nerf/nerf_helpers.py:178-233.  One thread per pixel; fp32 ops in the reference's order with
// contraction disabled (__f*_rn) so the result is bit-identical to ATen's CPU kernels.
__global__ void ray_bundle_kernel(int H, int W, float fx, float fy, float wcx, float hcy,
                                  const float* __restrict__ c2w, float* __restrict__ ro, float* __restrict__ rd) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= H * W) return;
  int j = idx / W, i = idx - j * W;
  float d0 = __fdiv_rn(__fsub_rn((float)i, wcx), fx);
  float d1 = __fdiv_rn(-__fsub_rn((float)j, hcy), fy);
  float d2 = -1.0f;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    float r0 = c2w[k * 4 + 0], r1 = c2w[k * 4 + 1], r2 = c2w[k * 4 + 2];
    float v = __fadd_rn(__fadd_rn(__fmul_rn(d0, r0), __fmul_rn(d1, r1)), __fmul_rn(d2, r2));
    rd[(size_t)idx * 3 + k] = v;
    ro[(size_t)idx * 3 + k] = c2w[k * 4 + 3];
  }
}

extern "C" int sahs_get_ray_bundle(int height, int width, float fx, float fy, double cx, double cy,
                                   const float* c2w, float* ro, float* rd, void* stream) {
  SAHS_CHECK_ARG(height > 0 && width > 0 && c2w && ro && rd, "bad arguments");
  // width * cx is evaluated in double by the reference (python int * numpy float64) and then rounded to fp32
  float wcx = (float)((double)width * cx);
  float hcy = (float)((double)height * cy);
  int n = height * width;
  ray_bundle_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(height, width, fx, fy, wcx, hcy, c2w, ro, rd);
  SAHS_LAUNCH_CHECK();
  return SAHS_OK;
}

// ref: nerf/train_utils.py:93-113
__global__ void coarse_z_kernel(int R, int S, float near_, float far_, int lindisp, const float* __restrict__ t_vals,
                                const float* __restrict__ t_rand, RngArg rng, float* __restrict__ z_out) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)R * S) return;
  int s = (int)(idx % S);
  auto zat = [&](int k) -> float {
    float t = t_vals[k];
    float omt = __fsub_rn(1.0f, t);
    if (!lindisp) return __fadd_rn(__fmul_rn(near_, omt), __fmul_rn(far_, t));
    float a = __fmul_rn(__fdiv_rn(1.0f, near_), omt);
    float b = __fmul_rn(__fdiv_rn(1.0f, far_), t);
    return __fdiv_rn(1.0f, __fadd_rn(a, b));
  };
  float z = zat(s);
  if (t_rand || rng.on) {
    float lower = (s == 0) ? z : __fmul_rn(0.5f, __fadd_rn(z, zat(s - 1)));
    float upper = (s == S - 1) ? z : __fmul_rn(0.5f, __fadd_rn(zat(s + 1), z));
    const float tr = t_rand ? t_rand[idx] : rng_uniform(rng, (uint64_t)idx);
    z = __fadd_rn(lower, __fmul_rn(__fsub_rn(upper, lower), tr));
  }
  z_out[idx] = z;
}

extern "C" int sahs_coarse_z(int num_rays, int num_samples, float near_, float far_, int lindisp,
                             const float* t_vals, const float* t_rand, float* z_out, void* stream) {
  SAHS_CHECK_ARG(num_rays >= 0 && num_samples > 0, "bad arguments");
  int64_t n = (int64_t)num_rays * num_samples;
  if (n == 0) return SAHS_OK;
  SAHS_CHECK_ARG(t_vals && z_out, "null pointer");
  coarse_z_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(num_rays, num_samples, near_, far_,
                                                                               lindisp, t_vals, t_rand, rng_arg(nullptr),
                                                                               z_out);
  SAHS_LAUNCH_CHECK();
  return SAHS_OK;
}

extern "C" int sahs_coarse_z_rng(int num_rays, int num_samples, float near_, float far_, int lindisp,
                                 const float* t_vals, const sahs_rng* rng, float* z_out, void* stream) {
  SAHS_CHECK_ARG(num_rays >= 0 && num_samples > 0, "bad arguments");
  int64_t n = (int64_t)num_rays * num_samples;
  if (n == 0) return SAHS_OK;
  SAHS_CHECK_ARG(t_vals && z_out, "null pointer");
  coarse_z_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(num_rays, num_samples, near_, far_,
                                                                               lindisp, t_vals, nullptr, rng_arg(rng),
                                                                               z_out);
  SAHS_LAUNCH_CHECK();
  return SAHS_OK;
}

__global__ void rng_fill_kernel(float* __restrict__ out, int64_t n, RngArg rng, int normal, float scale) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = scale * (normal ? rng_normal(rng, (uint64_t)i) : rng_uniform(rng, (uint64_t)i));
}

extern "C" int sahs_rng_fill(float* out, int64_t n, const sahs_rng* rng, int normal, float scale, void* stream) {
  SAHS_CHECK_ARG(n >= 0 && rng, "bad arguments");
  if (n == 0) return SAHS_OK;
  SAHS_CHECK_ARG(out, "null pointer");
  rng_fill_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(out, n, rng_arg(rng), normal, scale);
  SAHS_LAUNCH_CHECK();
  return SAHS_OK;
}

// ref: nerf/nerf_helpers.py:305-349.  One thread per (row, input dim); accurate sincosf of the exactly
// scaled argument (2^k * x is exact in fp32), written in the reference's concatenation order.
__global__ void posenc_kernel(const float* __restrict__ x, int64_t n, int d, int L, int inc, float* __restrict__ out) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * d) return;
  int64_t row = idx / d;
  int c = (int)(idx - row * d);
  int width = d * (inc + 2 * L);
  float v = x[idx];
  float* o = out + row * width;
  if (inc) o[c] = v;
  int base = inc ? d : 0;
  float f = 1.0f;
  for (int k = 0; k < L; ++k) {
    float s, co;
    sincosf(v * f, &s, &co);
    o[base + (2 * k) * d + c] = s;
    o[base + (2 * k + 1) * d + c] = co;
    f *= 2.0f;
  }
}

extern "C" int sahs_positional_encoding(const float* x, int64_t n, int d, int num_freqs, int include_input,
                                        float* out, void* stream) {
  SAHS_CHECK_ARG(n >= 0 && d > 0 && num_freqs >= 0, "bad arguments");
  if (n == 0) return SAHS_OK;
  SAHS_CHECK_ARG(x && out, "null pointer");
  int64_t tot = n * d;
  posenc_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, n, d, num_freqs,
                                                                              include_input ? 1 : 0, out);
  SAHS_LAUNCH_CHECK();
  return SAHS_OK;
}

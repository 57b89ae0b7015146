// Adam on one flat fp32 parameter buffer.
// ref: train_stage_rays_auto.py:200-210 (torch.optim.Adam over the trainable parameters, default betas/eps, no weight
//      decay, no amsgrad) and :503-509 (exponential learning-rate decay, applied by the host as a scalar).
// The reference's optimizer walks 124 parameter tensors; with the parameters, gradients and both moments laid out as
// four flat buffers the step is one HBM-bound pass: 16 B read + 12 B written per parameter (2.8 M parameters -> 78 MB,
// ~15 us), and the data-parallel gradient average (1 / world size) is folded into the read of the gradient.
// Arithmetic follows torch.optim.Adam's single-tensor path: m = b1 m + (1 - b1) g;  v = b2 v + (1 - b2) g^2;
// p -= (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps).
#include "sahs_common.cuh"

namespace {

__global__ void __launch_bounds__(256)
adam_flat_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                 int64_t n, float step_size, float omb1, float beta2, float omb2, float sqrt_bc2, float eps,
                 float grad_scale) {
  const int64_t n4 = n >> 2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  // omb1 = 1 - beta1 and omb2 = 1 - beta2 are formed in double on the host, as torch does (1.f - 0.999f is off by 1e-5,
  // which is why the hyper-parameters cross the ABI as doubles)
  auto upd = [&](float& pp, float gg, float& mm, float& vv) {
    gg *= grad_scale;
    mm = mm + omb1 * (gg - mm);               // exp_avg.lerp_(grad, 1 - beta1)
    vv = beta2 * vv + omb2 * gg * gg;         // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
    const float denom = sqrtf(vv) / sqrt_bc2 + eps;
    pp -= step_size * (mm / denom);
  };
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 pp = reinterpret_cast<float4*>(p)[i], mm = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
    const float4 gg = reinterpret_cast<const float4*>(g)[i];
    upd(pp.x, gg.x, mm.x, vv.x);
    upd(pp.y, gg.y, mm.y, vv.y);
    upd(pp.z, gg.z, mm.z, vv.z);
    upd(pp.w, gg.w, mm.w, vv.w);
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
  }
  for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    upd(p[i], g[i], m[i], v[i]);
}

}  // namespace

extern "C" int sahs_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                              double lr, double beta1, double beta2, double eps, int step, float grad_scale,
                              void* stream) {
  SAHS_CHECK_ARG(n >= 0 && step >= 1, "bad extents / step counts from 1");
  if (n == 0) return SAHS_OK;
  SAHS_CHECK_ARG(params && grads && exp_avg && exp_avg_sq, "null pointer");
  SAHS_CHECK_ARG((((uintptr_t)params | (uintptr_t)grads | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) & 15) == 0,
                 "buffers must be 16-byte aligned");
  SAHS_CHECK_ARG(beta1 >= 0.0 && beta1 < 1.0 && beta2 >= 0.0 && beta2 < 1.0, "betas in [0, 1)");
  const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
  const float step_size = (float)(lr / bc1), sqrt_bc2 = (float)sqrt(bc2);
  const float omb1 = (float)(1.0 - beta1), omb2 = (float)(1.0 - beta2);
  int64_t blocks = (n / 4 + 255) / 256;
  const int64_t cap = (int64_t)sahs_num_sms() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  adam_flat_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(params, grads, exp_avg, exp_avg_sq, n, step_size,
                                                                      omb1, (float)beta2, omb2, sqrt_bc2, (float)eps, grad_scale);
  SAHS_LAUNCH_CHECK();
  return SAHS_OK;
}

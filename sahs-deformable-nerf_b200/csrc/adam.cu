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

// Capturable variant: the step count and the learning-rate schedule live in device memory, so a training step recorded
// in a CUDA graph replays with the right bias corrections and the decayed learning rate without host code.
// hyper (doubles): [0] lr0, [1] decay factor, [2] decay steps, [3] beta1, [4] beta2, [5] eps, [6] step t (>= 1, the
// step being applied), [7] schedule position i: lr = lr0 * factor^(i / steps) (ref: train_stage_rays_auto.py:503-509).
__global__ void __launch_bounds__(256)
adam_flat_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                     int64_t n, const double* __restrict__ hyper, float grad_scale) {
  __shared__ float sh[6];
  if (threadIdx.x == 0) {
    const double lr0 = hyper[0], factor = hyper[1], steps = hyper[2], b1 = hyper[3], b2 = hyper[4], t = hyper[6];
    const double lr = lr0 * pow(factor, hyper[7] / steps);
    const double bc1 = 1.0 - pow(b1, t), bc2 = 1.0 - pow(b2, t);
    sh[0] = (float)(lr / bc1);
    sh[1] = (float)(1.0 - b1);
    sh[2] = (float)b2;
    sh[3] = (float)(1.0 - b2);
    sh[4] = (float)sqrt(bc2);
    sh[5] = (float)hyper[5];
  }
  __syncthreads();
  const float step_size = sh[0], omb1 = sh[1], beta2 = sh[2], omb2 = sh[3], sqrt_bc2 = sh[4], eps = sh[5];
  const int64_t n4 = n >> 2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  auto upd = [&](float& pp, float gg, float& mm, float& vv) {
    gg *= grad_scale;
    mm = mm + omb1 * (gg - mm);
    vv = beta2 * vv + omb2 * gg * gg;
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

__global__ void adam_advance_kernel(double* hyper) {
  hyper[6] += 1.0;
  hyper[7] += 1.0;
}

__global__ void counter_add_kernel(unsigned long long* ctr, unsigned long long inc) { *ctr += inc; }

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

extern "C" int sahs_adam_step_dev(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                                  const double* hyper_dev, float grad_scale, void* stream) {
  SAHS_CHECK_ARG(n >= 0, "bad extents");
  if (n == 0) return SAHS_OK;
  SAHS_CHECK_ARG(params && grads && exp_avg && exp_avg_sq && hyper_dev, "null pointer");
  SAHS_CHECK_ARG((((uintptr_t)params | (uintptr_t)grads | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) & 15) == 0,
                 "buffers must be 16-byte aligned");
  int64_t blocks = (n / 4 + 255) / 256;
  const int64_t cap = (int64_t)sahs_num_sms() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  adam_flat_dev_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(params, grads, exp_avg, exp_avg_sq, n, hyper_dev,
                                                                          grad_scale);
  SAHS_LAUNCH_CHECK();
  return SAHS_OK;
}

extern "C" int sahs_adam_advance(double* hyper_dev, void* stream) {
  SAHS_CHECK_ARG(hyper_dev, "null pointer");
  adam_advance_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(hyper_dev);
  SAHS_LAUNCH_CHECK();
  return SAHS_OK;
}

extern "C" int sahs_counter_add(unsigned long long* counter_dev, unsigned long long inc, void* stream) {
  SAHS_CHECK_ARG(counter_dev, "null pointer");
  counter_add_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(counter_dev, inc);
  SAHS_LAUNCH_CHECK();
  return SAHS_OK;
}

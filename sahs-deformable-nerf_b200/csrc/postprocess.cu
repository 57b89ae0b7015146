// Frame post-processing: 15-channel composited map -> uint8 rgb + argmax semantic label + palette colours
// (SURVEY.md section 8f row 2).  ref: eval_stage_rays.py:221-227 (cast_to_image: clamp, *255, truncate -- what
// torchvision's ToPILImage does to a float tensor), nerf/utils.py:112-140 (label2color: argmax over the 12 classes,
// palette entries written in reversed channel order).  HBM-bound: 60 B in, 7 B out per ray; one thread per ray.
#include "sahs_common.cuh"

__constant__ uint8_t kSegPaletteBGR[12][3] = {
    {0, 0, 0}, {0, 0, 204}, {0, 153, 76}, {0, 204, 204}, {255, 51, 51}, {255, 255, 0},
    {0, 51, 102}, {0, 204, 102}, {0, 255, 255}, {204, 0, 0}, {51, 153, 255}, {0, 204, 0}};

__global__ void postprocess_kernel(const float* __restrict__ map, int64_t n, uint8_t* __restrict__ rgb,
                                   uint8_t* __restrict__ label, uint8_t* __restrict__ seg_color) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* m = map + i * SAHS_MAP_CH;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    float v = fminf(fmaxf(m[k], 0.f), 1.f);
    rgb[i * 3 + k] = (uint8_t)(v * 255.0f);          // truncation, like tensor.mul(255).byte()
  }
  int best = 0;
  float bv = m[3];
#pragma unroll
  for (int k = 1; k < 12; ++k) {
    float v = m[3 + k];
    if (v > bv) { bv = v; best = k; }                // first maximum wins, like torch.argmax on CPU
  }
  if (label) label[i] = (uint8_t)best;
  if (seg_color) {
#pragma unroll
    for (int k = 0; k < 3; ++k) seg_color[i * 3 + k] = kSegPaletteBGR[best][k];
  }
}

extern "C" int sahs_frame_postprocess(const float* map15, int64_t num_rays, uint8_t* rgb_u8, uint8_t* label_u8,
                                      uint8_t* seg_color_u8, void* stream) {
  SAHS_CHECK_ARG(num_rays >= 0, "bad extent");
  if (num_rays == 0) return SAHS_OK;
  SAHS_CHECK_ARG(map15 && rgb_u8, "null pointer");
  postprocess_kernel<<<(unsigned)((num_rays + 255) / 256), 256, 0, (cudaStream_t)stream>>>(map15, num_rays, rgb_u8,
                                                                                          label_u8, seg_color_u8);
  SAHS_LAUNCH_CHECK();
  return SAHS_OK;
}

// Depth map -> normal map (the eval script's `torch_normal_map`, ref: eval_stage_rays.py:116-151): back-project every
// pixel with the pinhole intrinsics, forward (or central) differences along rows and columns, cross product,
// normalise, n * 0.5 + 0.5, optional clean-up with the fine pass's background weight (normal -> white where the ray
// mostly hit the background), * 255.  One thread per output pixel; three depth reads (L1/L2 hits), 12 B written.
// The reference's meshgrid only broadcasts for square maps, which is all it is used on; so does this.
__global__ void normal_map_kernel(const float* __restrict__ depth, int n, int k, float fx, float fy, float cx, float cy,
                                  const float* __restrict__ weights, float* __restrict__ out) {
  const int m = n - k;                                   // output is [n-k, n-k, 3]
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= m * m) return;
  const int r = idx / m, c = idx - r * m;
  auto point = [&](int rr, int cc, float (&p)[3]) {
    const float d = depth[(size_t)rr * n + cc];
    p[0] = __fdiv_rn(__fmul_rn(__fsub_rn((float)cc, cx), d), fx);
    p[1] = -__fdiv_rn(__fmul_rn(__fsub_rn((float)rr, cy), d), fy);
    p[2] = d;
  };
  float p0[3], pr[3], pc[3];
  point(r, c, p0);
  point(r + k, c, pr);                                   // dx: difference along the first axis
  point(r, c + k, pc);                                   // dy: difference along the second axis
  float dx[3], dy[3];
#pragma unroll
  for (int q = 0; q < 3; ++q) { dx[q] = __fsub_rn(pr[q], p0[q]); dy[q] = __fsub_rn(pc[q], p0[q]); }
  // cross(dy, dx)
  float nv[3] = {__fsub_rn(__fmul_rn(dy[1], dx[2]), __fmul_rn(dy[2], dx[1])),
                 __fsub_rn(__fmul_rn(dy[2], dx[0]), __fmul_rn(dy[0], dx[2])),
                 __fsub_rn(__fmul_rn(dy[0], dx[1]), __fmul_rn(dy[1], dx[0]))};
  const float len = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(nv[0], nv[0]), __fmul_rn(nv[1], nv[1])), __fmul_rn(nv[2], nv[2])));
  float w = 0.f;
  if (weights) w = weights[(size_t)r * n + c];
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    float v = __fadd_rn(__fmul_rn(__fdiv_rn(nv[q], len), 0.5f), 0.5f);
    if (weights) {
      if (w > 0.22f) v = 1.0f;
      v = __fadd_rn(__fmul_rn(__fsub_rn(1.0f, w), v), w);   // (1 - mask) * normals + mask * 1
    }
    out[(size_t)idx * 3 + q] = __fmul_rn(v, 255.0f);
  }
}

extern "C" int sahs_normal_map(const float* depth, int size, float fx, float fy, float cx_rel, float cy_rel,
                               const float* weights, int central_difference, float* normals_out, void* stream) {
  const int k = central_difference ? 2 : 1;
  SAHS_CHECK_ARG(size > k, "depth map too small");
  SAHS_CHECK_ARG(depth && normals_out, "null pointer");
  const int m = size - k;
  normal_map_kernel<<<(unsigned)((m * m + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      depth, size, k, fx, fy, cx_rel * (float)size, cy_rel * (float)size, weights, normals_out);
  SAHS_LAUNCH_CHECK();
  return SAHS_OK;
}

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

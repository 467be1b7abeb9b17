// RoI crop of the observed image (reference: lib3d/cropping.py:74 ->
// torchvision.ops.roi_align(images, rois, (240,320), spatial_scale=1, sampling_ratio=4,
// aligned=False); torchvision 0.4.2 pinned by environment.yaml:10, CPU kernel
// csrc/ops/cpu/roi_align_kernel.cpp + roi_align_common.h).  Published rule restated:
//   roi_w = max(x2-x1, 1), bin_w = roi_w/320; sample x = x1 + pw*bin_w + (ix+.5)*bin_w/4, ix<4
//   a sample is dropped (weight 0) if x < -1 or x > W (same for y); else x = max(x, 0);
//   lo = (int)x; if lo >= W-1: lo = hi = W-1, x = lo; else hi = lo+1; bilinear; sum of the 16
//   samples / 16.
// Coordinates are computed with explicit round-to-nearest mul/add so that no FMA contraction
// moves a sample across the drop threshold relative to the CPU operator.
// The image is gathered through im_ids (the reference copies the full frame per hypothesis,
// integrated/pose_predictor.py:41).  Output: NCHW planes, the layout the stem kernel stages from.
#pragma once
#include "common.h"

namespace cosyb {

struct AxisSample {
  int lo, hi;
  float l, h;  // weights of hi / lo
  bool valid;
};

__device__ __forceinline__ AxisSample make_axis_sample(float start, float bin, int p, int i, int size) {
  // start + p*bin + ((i+.5)*bin)/4
  float c = __fadd_rn(__fadd_rn(start, __fmul_rn((float)p, bin)),
                      __fdiv_rn(__fmul_rn((float)i + 0.5f, bin), 4.0f));
  AxisSample s;
  s.valid = !(c < -1.0f || c > (float)size);
  if (c <= 0.f) c = 0.f;
  int lo = (int)c;
  int hi;
  if (lo >= size - 1) {
    lo = hi = size - 1;
    c = (float)lo;
  } else {
    hi = lo + 1;
  }
  if (!s.valid) lo = hi = 0;
  s.lo = lo;
  s.hi = hi;
  s.l = __fsub_rn(c, (float)lo);
  s.h = __fsub_rn(1.0f, s.l);
  return s;
}

constexpr int CROP_TX = 32, CROP_TY = 8;

__global__ void __launch_bounds__(CROP_TX* CROP_TY)
k_roi_crop(int B, const float* __restrict__ images, int n_images, int H, int W,
           const int32_t* __restrict__ im_ids, const float* __restrict__ boxes,
           float* __restrict__ crops) {
  const int b = blockIdx.z;
  const int pw = blockIdx.x * CROP_TX + threadIdx.x;
  const int ph = blockIdx.y * CROP_TY + threadIdx.y;
  if (pw >= RENDER_W || ph >= RENDER_H) return;
  const float x1 = boxes[b * 4 + 0], y1 = boxes[b * 4 + 1];
  const float x2 = boxes[b * 4 + 2], y2 = boxes[b * 4 + 3];
  const float roi_w = fmaxf(__fsub_rn(x2, x1), 1.0f);
  const float roi_h = fmaxf(__fsub_rn(y2, y1), 1.0f);
  const float bin_w = __fdiv_rn(roi_w, (float)RENDER_W);
  const float bin_h = __fdiv_rn(roi_h, (float)RENDER_H);
  AxisSample sx[4], sy[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    sx[i] = make_axis_sample(x1, bin_w, pw, i, W);
    sy[i] = make_axis_sample(y1, bin_h, ph, i, H);
  }
  const float* img = images + (size_t)im_ids[b] * 3 * H * W;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float* pl = img + (size_t)c * H * W;
    float acc = 0.f;
#pragma unroll
    for (int iy = 0; iy < 4; ++iy) {
      const float* rlo = pl + (size_t)sy[iy].lo * W;
      const float* rhi = pl + (size_t)sy[iy].hi * W;
#pragma unroll
      for (int ix = 0; ix < 4; ++ix) {
        float w1 = __fmul_rn(sy[iy].h, sx[ix].h), w2 = __fmul_rn(sy[iy].h, sx[ix].l);
        float w3 = __fmul_rn(sy[iy].l, sx[ix].h), w4 = __fmul_rn(sy[iy].l, sx[ix].l);
        float v = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(w1, __ldg(rlo + sx[ix].lo)),
                                                __fmul_rn(w2, __ldg(rlo + sx[ix].hi))),
                                      __fmul_rn(w3, __ldg(rhi + sx[ix].lo))),
                            __fmul_rn(w4, __ldg(rhi + sx[ix].hi)));
        if (sy[iy].valid && sx[ix].valid) acc = __fadd_rn(acc, v);
      }
    }
    crops[(((size_t)b * 3 + c) * RENDER_H + ph) * RENDER_W + pw] = __fdiv_rn(acc, 16.0f);
  }
}

}  // namespace cosyb

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
//
// The 16-sample sum is separable: out = 1/16 sum_iy vy (hy R[ylo] + ly R[yhi]) with the horizontal
// pass R[row][pw] = sum_ix vx (hx I[row][xlo] + lx I[row][xhi]).  A CTA owns CROP_PH output rows of
// one hypothesis, builds R for the few source rows they touch in shared memory (8 loads per entry
// instead of 64 per output pixel), then combines vertically.  Tiles that would need more than
// CROP_MAXR source rows (strong zoom-out) take the direct path.
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

constexpr int CROP_PH = 8;          // output rows per CTA
constexpr int CROP_MAXR = 16;       // source rows staged per CTA
constexpr int CROP_THREADS = 320;   // one thread per output column
constexpr int CROP_SMEM_FLOATS = CROP_MAXR * 3 * RENDER_W;

__global__ void __launch_bounds__(CROP_THREADS)
k_roi_crop(int B, const float* __restrict__ images, int n_images, int H, int W,
           const int32_t* __restrict__ im_ids, const float* __restrict__ boxes,
           float* __restrict__ crops) {
  extern __shared__ float s_R[];   // [rows][3][320]
  const int b = blockIdx.y;
  const int ph0 = blockIdx.x * CROP_PH;
  const int pw = threadIdx.x;
  const float x1 = boxes[b * 4 + 0], y1 = boxes[b * 4 + 1];
  const float x2 = boxes[b * 4 + 2], y2 = boxes[b * 4 + 3];
  const float roi_w = fmaxf(__fsub_rn(x2, x1), 1.0f);
  const float roi_h = fmaxf(__fsub_rn(y2, y1), 1.0f);
  const float bin_w = __fdiv_rn(roi_w, (float)RENDER_W);
  const float bin_h = __fdiv_rn(roi_h, (float)RENDER_H);
  const float* img = images + (size_t)min(max(im_ids[b], 0), n_images - 1) * 3 * H * W;   // ids outside the batch are clamped

  AxisSample sx[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) sx[i] = make_axis_sample(x1, bin_w, pw, i, W);

  // source rows touched by the valid y samples of this tile (sample coordinates are monotonic in ph, iy)
  int r0 = H, r1 = -1;
  for (int k = 0; k < CROP_PH * 4; ++k) {
    AxisSample s = make_axis_sample(y1, bin_h, ph0 + k / 4, k % 4, H);
    if (s.valid) {
      r0 = min(r0, s.lo);
      r1 = max(r1, s.hi);
    }
  }
  const int n_rows = r1 - r0 + 1;   // <= 0: every sample of the tile is outside the frame
  float* out = crops + ((size_t)b * 3 * RENDER_H + ph0) * RENDER_W + pw;

  if (n_rows <= 0) {
#pragma unroll
    for (int c = 0; c < 3; ++c)
      for (int p = 0; p < CROP_PH; ++p) out[((size_t)c * RENDER_H + p) * RENDER_W] = 0.f;
    return;
  }
  if (n_rows <= CROP_MAXR) {
    // horizontal pass: thread pw builds R[row][c][pw] for all staged rows
    for (int r = 0; r < n_rows; ++r) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float* row = img + ((size_t)c * H + (r0 + r)) * W;
        float acc = 0.f;
#pragma unroll
        for (int ix = 0; ix < 4; ++ix) {
          const float v = __fadd_rn(__fmul_rn(sx[ix].h, __ldg(row + sx[ix].lo)), __fmul_rn(sx[ix].l, __ldg(row + sx[ix].hi)));
          if (sx[ix].valid) acc = __fadd_rn(acc, v);
        }
        s_R[(r * 3 + c) * RENDER_W + pw] = acc;
      }
    }
    // vertical pass: only this thread's own column of R is read, no barrier needed
    for (int p = 0; p < CROP_PH; ++p) {
      float acc[3] = {0.f, 0.f, 0.f};
#pragma unroll
      for (int iy = 0; iy < 4; ++iy) {
        const AxisSample s = make_axis_sample(y1, bin_h, ph0 + p, iy, H);
        if (s.valid) {
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const float lo = s_R[((s.lo - r0) * 3 + c) * RENDER_W + pw], hi = s_R[((s.hi - r0) * 3 + c) * RENDER_W + pw];
            acc[c] = __fadd_rn(acc[c], __fadd_rn(__fmul_rn(s.h, lo), __fmul_rn(s.l, hi)));
          }
        }
      }
#pragma unroll
      for (int c = 0; c < 3; ++c) out[((size_t)c * RENDER_H + p) * RENDER_W] = __fdiv_rn(acc[c], 16.0f);
    }
    return;
  }
  // direct path (strong zoom-out): 16 bilinear samples per output pixel
  for (int p = 0; p < CROP_PH; ++p) {
    AxisSample sy[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) sy[i] = make_axis_sample(y1, bin_h, ph0 + p, i, H);
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
      out[((size_t)c * RENDER_H + p) * RENDER_W] = __fdiv_rn(acc, 16.0f);
    }
  }
}

}  // namespace cosyb

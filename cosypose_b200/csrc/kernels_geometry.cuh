// Per-hypothesis geometry: projection -> boxes -> crop box -> cropped intrinsics, the pose
// update head and the pose initialisers.  All fp32.  One CTA per hypothesis; the reductions over
// the 2000 sampled mesh points use warp shuffles.
#pragma once
#include "common.h"

namespace cosyb {

__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Phase A (reference: models/pose.py:45-67).
//   project_points_robust  lib3d/camera_geometry.py:18-31  (z clamped to >= 0.1)
//   boxes_from_uv          lib3d/camera_geometry.py:34-42
//   deepim_boxes           lib3d/cropping.py:7-47  (obs box == rend box on this path, lambda 1.4,
//                          aspect from the *input image* size, no clamping)
//   get_K_crop_resize      lib3d/camera_geometry.py:45-87
// The reference projects the same points twice (pose.py:52 and cropping.py:69); here once.
constexpr int GEO_THREADS = 128;

__global__ void __launch_bounds__(GEO_THREADS)
k_project_boxes(int B, const float* __restrict__ K, const float* __restrict__ TCO,
                const int32_t* __restrict__ label_ids, int n_labels, const float* __restrict__ pts_sampled,
                int n_sample, float img_aspect, float* __restrict__ boxes_rend,
                float* __restrict__ boxes_crop, float* __restrict__ K_crop) {
  const int b = blockIdx.x;
  if (b >= B) return;
  __shared__ float sP[12];
  __shared__ float sred[4][GEO_THREADS / 32];
  const float* Kb = K + b * 9;
  const float* Tb = TCO + b * 16;
  if (threadIdx.x < 12) {
    // P = K @ TCO[:3]  (3x4)
    int r = threadIdx.x / 4, c = threadIdx.x % 4;
    float acc = Kb[r * 3 + 0] * Tb[0 * 4 + c];
    acc = fmaf(Kb[r * 3 + 1], Tb[1 * 4 + c], acc);
    acc = fmaf(Kb[r * 3 + 2], Tb[2 * 4 + c], acc);
    sP[threadIdx.x] = acc;
  }
  __syncthreads();
  float P[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) P[i] = sP[i];
  const float* pts = pts_sampled + (size_t)min(max(label_ids[b], 0), n_labels - 1) * n_sample * 3;   // ids outside the table are clamped
  float umin = INFINITY, vmin = INFINITY, umax = -INFINITY, vmax = -INFINITY;
  for (int i = threadIdx.x; i < n_sample; i += GEO_THREADS) {
    float x = pts[i * 3 + 0], y = pts[i * 3 + 1], z = pts[i * 3 + 2];
    float su = fmaf(P[0], x, fmaf(P[1], y, fmaf(P[2], z, P[3])));
    float sv = fmaf(P[4], x, fmaf(P[5], y, fmaf(P[6], z, P[7])));
    float sz = fmaf(P[8], x, fmaf(P[9], y, fmaf(P[10], z, P[11])));
    sz = fmaxf(0.1f, sz);
    float u = su / sz, v = sv / sz;
    umin = fminf(umin, u);
    umax = fmaxf(umax, u);
    vmin = fminf(vmin, v);
    vmax = fmaxf(vmax, v);
  }
  umin = warp_min(umin);
  vmin = warp_min(vmin);
  umax = warp_max(umax);
  vmax = warp_max(vmax);
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  if (lane == 0) {
    sred[0][warp] = umin;
    sred[1][warp] = vmin;
    sred[2][warp] = umax;
    sred[3][warp] = vmax;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int w = 1; w < GEO_THREADS / 32; ++w) {
      umin = fminf(umin, sred[0][w]);
      vmin = fminf(vmin, sred[1][w]);
      umax = fmaxf(umax, sred[2][w]);
      vmax = fmaxf(vmax, sred[3][w]);
    }
    boxes_rend[b * 4 + 0] = umin;
    boxes_rend[b * 4 + 1] = vmin;
    boxes_rend[b * 4 + 2] = umax;
    boxes_rend[b * 4 + 3] = vmax;
    // projected object origin
    float cz = fmaxf(0.1f, P[11]);
    float xc = P[3] / cz, yc = P[7] / cz;
    float xdist = fmaxf(fabsf(umin - xc), fabsf(umax - xc));
    float ydist = fmaxf(fabsf(vmin - yc), fabsf(vmax - yc));
    float width = fmaxf(xdist, ydist * img_aspect) * 2.f * 1.4f;
    float height = fmaxf(xdist / img_aspect, ydist) * 2.f * 1.4f;
    float x1 = xc - width / 2.f, y1 = yc - height / 2.f;
    float x2 = xc + width / 2.f, y2 = yc + height / 2.f;
    boxes_crop[b * 4 + 0] = x1;
    boxes_crop[b * 4 + 1] = y1;
    boxes_crop[b * 4 + 2] = x2;
    boxes_crop[b * 4 + 3] = y2;
    // cropped + resized intrinsics
    const float fw = (float)RENDER_W, fh = (float)RENDER_H;
    float cw = x2 - x1, ch = y2 - y1;
    float cj = (x1 + x2) / 2.f, ci = (y1 + y2) / 2.f;
    float cx = Kb[2] + (cw - 1.f) / 2.f - cj;
    float cy = Kb[5] + (ch - 1.f) / 2.f - ci;
    float dx = cx - (cw - 1.f) / 2.f, dy = cy - (ch - 1.f) / 2.f;
    float sx = fw / cw, sy = fh / ch;
    float* Ko = K_crop + b * 9;
#pragma unroll
    for (int i = 0; i < 9; ++i) Ko[i] = Kb[i];
    Ko[0] = sx * Kb[0];
    Ko[4] = sy * Kb[4];
    Ko[2] = (fw - 1.f) / 2.f + sx * dx;
    Ko[5] = (fh - 1.f) / 2.f + sy * dy;
  }
}

// update_pose, pose_dim 9 (reference: models/pose.py:69-79):
//   compute_rotation_matrix_from_ortho6d  lib3d/rotations.py:6-21  (x, y, z stacked as COLUMNS)
//   apply_imagespace_predictions          lib3d/cosypose_ops.py:10-31
__device__ __forceinline__ void pose_update_one(const float* __restrict__ T, const float* __restrict__ Kc,
                                                const float* __restrict__ p, float* __restrict__ To) {
  float ax = p[0], ay = p[1], az = p[2], bx = p[3], by = p[4], bz = p[5];
  float n = sqrtf(ax * ax + ay * ay + az * az);
  float x0 = ax / n, x1 = ay / n, x2 = az / n;
  float z0 = x1 * bz - x2 * by, z1 = x2 * bx - x0 * bz, z2 = x0 * by - x1 * bx;
  float nz = sqrtf(z0 * z0 + z1 * z1 + z2 * z2);
  z0 /= nz; z1 /= nz; z2 /= nz;
  float y0 = z1 * x2 - z2 * x1, y1 = z2 * x0 - z0 * x2, y2 = z0 * x1 - z1 * x0;
  // dR columns are (x, y, z): dR[r][0]=x_r, dR[r][1]=y_r, dR[r][2]=z_r
  float dR[9] = {x0, y0, z0, x1, y1, z1, x2, y2, z2};
  float zsrc = T[11];
  float ztgt = p[8] * zsrc;
  float fx = Kc[0], fy = Kc[4];
  float xo = (p[6] / fx + T[3] / zsrc) * ztgt;
  float yo = (p[7] / fy + T[7] / zsrc) * ztgt;
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c)
      To[r * 4 + c] = dR[r * 3 + 0] * T[0 * 4 + c] + dR[r * 3 + 1] * T[1 * 4 + c] + dR[r * 3 + 2] * T[2 * 4 + c];
  To[3] = xo;
  To[7] = yo;
  To[11] = ztgt;
  To[12] = T[12];
  To[13] = T[13];
  To[14] = T[14];
  To[15] = T[15];
}

__global__ void k_update_pose(int B, const float* __restrict__ TCO, const float* __restrict__ K_crop,
                              const float* __restrict__ pose9, float* __restrict__ TCO_out) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float To[16];
  pose_update_one(TCO + b * 16, K_crop + b * 9, pose9 + b * 9, To);
#pragma unroll
  for (int i = 0; i < 16; ++i) TCO_out[b * 16 + i] = To[i];
}

// TCO_init_from_boxes(z_range=(1,1))            lib3d/cosypose_ops.py:121-135   (zup == 0)
// TCO_init_from_boxes_zup_autodepth             lib3d/cosypose_ops.py:138-173   (zup == 1)
__global__ void __launch_bounds__(GEO_THREADS)
k_tco_init(int B, int zup, const float* __restrict__ boxes, const float* __restrict__ K,
           const int32_t* __restrict__ label_ids, int n_labels, const float* __restrict__ pts_sampled, int n_sample,
           float* __restrict__ TCO) {
  const int b = blockIdx.x;
  if (b >= B) return;
  const float* Kb = K + b * 9;
  const float* bx = boxes + b * 4;
  const float fx = Kb[0], fy = Kb[4], cx = Kb[2], cy = Kb[5];
  const float uc = (bx[0] + bx[2]) / 2.f, vc = (bx[1] + bx[3]) / 2.f;
  float* T = TCO + b * 16;
  if (!zup) {
    if (threadIdx.x == 0) {
      const float z = 1.f;
      float t[16] = {1, 0, 0, ((uc - cx) * z) / fx, 0, 1, 0, ((vc - cy) * z) / fy, 0, 0, 1, z, 0, 0, 0, 1};
#pragma unroll
      for (int i = 0; i < 16; ++i) T[i] = t[i];
    }
    return;
  }
  // z-up + auto-depth: R = [[0,1,0],[0,0,-1],[-1,0,0]], z_guess = 1
  __shared__ float sred[4][GEO_THREADS / 32];
  const float tx = ((uc - cx) * 1.f) / fx, ty = ((vc - cy) * 1.f) / fy;
  const float* pts = pts_sampled + (size_t)min(max(label_ids[b], 0), n_labels - 1) * n_sample * 3;   // ids outside the table are clamped
  float xmin = INFINITY, xmax = -INFINITY, ymin = INFINITY, ymax = -INFINITY;
  for (int i = threadIdx.x; i < n_sample; i += GEO_THREADS) {
    float py = pts[i * 3 + 1], pz = pts[i * 3 + 2];
    float X = py + tx;    // row 0 of R = (0,1,0)
    float Y = -pz + ty;   // row 1 of R = (0,0,-1)
    xmin = fminf(xmin, X); xmax = fmaxf(xmax, X);
    ymin = fminf(ymin, Y); ymax = fmaxf(ymax, Y);
  }
  xmin = warp_min(xmin); ymin = warp_min(ymin); xmax = warp_max(xmax); ymax = warp_max(ymax);
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  if (lane == 0) { sred[0][warp] = xmin; sred[1][warp] = ymin; sred[2][warp] = xmax; sred[3][warp] = ymax; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < GEO_THREADS / 32; ++w) {
      xmin = fminf(xmin, sred[0][w]); ymin = fminf(ymin, sred[1][w]);
      xmax = fmaxf(xmax, sred[2][w]); ymax = fmaxf(ymax, sred[3][w]);
    }
    float dx3 = xmax - xmin, dy3 = ymax - ymin;
    float bdx = (bx[2] - bx[0]) + 1.f, bdy = (bx[3] - bx[1]) + 1.f;
    float zdx = fx * dx3 / bdx, zdy = fy * dy3 / bdy;
    float z = (zdy + zdx) / 2.f;
    float t[16] = {0, 1, 0, ((uc - cx) * z) / fx, 0, 0, -1, ((vc - cy) * z) / fy, -1, 0, 0, z, 0, 0, 0, 1};
#pragma unroll
    for (int i = 0; i < 16; ++i) T[i] = t[i];
  }
}

}  // namespace cosyb

// ADD / ADD-S pose errors (reference: lib3d/distances.py:5-21 `dists_add`,
// `dists_add_symmetric`, and the error statistics of evaluation/meters/pose_meters.py:84-89).
//   ADD    dists[b][j] = T_gt p_j - T_pred p_j
//   ADD-S  dists[b][i] = T_gt p_i - T_pred p_a(i),  a(i) = argmin_j |T_gt p_i - T_pred p_j|^2  (first minimum wins):
//          `dists[b, j, i] = gt_i - pred_j`, `argmin(dim=1)` runs over the PREDICTED points for every ground-truth point
// The reference materialises the [P, P, 3] difference tensor per pair (48 MB at P = 2000); here one CTA per pair keeps
// the transformed predicted points in shared memory and every thread scans them for its ground-truth points.
#pragma once
#include "common.h"
#include "kernels_ransac.cuh"

namespace cosyb {

constexpr int EVAL_THREADS = 256;

// dynamic smem: 3 * P floats (transformed predicted points, SoA)
__global__ void __launch_bounds__(EVAL_THREADS)
k_pose_errors(int n, int P, const float* __restrict__ T_pred, const float* __restrict__ T_gt,
              const float* __restrict__ points /*[n][P][3]*/, const int32_t* __restrict__ symmetric /*[n] or null*/,
              float* __restrict__ dists /*[n][P][3] or null*/, float* __restrict__ norm_avg, float* __restrict__ xyz_avg,
              float* __restrict__ tco_xyz, float* __restrict__ tco_norm) {
  extern __shared__ float s_gt[];
  __shared__ float s_red[4][EVAL_THREADS / 32];
  const int b = blockIdx.x, tid = threadIdx.x;
  if (b >= n) return;
  const Mat34 Tp = load34(T_pred + (size_t)b * 16), Tg = load34(T_gt + (size_t)b * 16);
  const float* pts = points + (size_t)b * P * 3;
  const bool sym = symmetric != nullptr && symmetric[b] != 0;
  float* px = s_gt; float* py = s_gt + P; float* pz = s_gt + 2 * P;
  for (int j = tid; j < P; j += EVAL_THREADS) {
    const float p[3] = {pts[j * 3], pts[j * 3 + 1], pts[j * 3 + 2]};
    float q[3];
    apply34(Tp, p, q);
    px[j] = q[0]; py[j] = q[1]; pz[j] = q[2];
  }
  __syncthreads();
  float a_n = 0.f, a_x = 0.f, a_y = 0.f, a_z = 0.f;
  for (int i = tid; i < P; i += EVAL_THREADS) {
    const float p[3] = {pts[i * 3], pts[i * 3 + 1], pts[i * 3 + 2]};
    float q[3];
    apply34(Tg, p, q);
    int a = i;
    if (sym) {
      float best = 3.402823466e+38f;
      a = 0;
      for (int j = 0; j < P; ++j) {
        // ((dx^2 + dy^2) + dz^2) with separate roundings, as `(dists ** 2).sum(-1)` evaluates it
        const float dx = q[0] - px[j], dy = q[1] - py[j], dz = q[2] - pz[j];
        const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
        if (d2 < best) { best = d2; a = j; }
      }
    }
    const float dx = q[0] - px[a], dy = q[1] - py[a], dz = q[2] - pz[a];
    if (dists) {
      float* o = dists + ((size_t)b * P + i) * 3;
      o[0] = dx; o[1] = dy; o[2] = dz;
    }
    a_n += sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
    a_x += fabsf(dx); a_y += fabsf(dy); a_z += fabsf(dz);
  }
  // fixed-order reduction: lanes by shuffle, warps in order
  float v[4] = {a_n, a_x, a_y, a_z};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
    if (tid % 32 == 0) s_red[k][tid / 32] = v[k];
  }
  __syncthreads();
  if (tid == 0) {
    float r[4] = {0.f, 0.f, 0.f, 0.f};
    for (int w = 0; w < EVAL_THREADS / 32; ++w)
      for (int k = 0; k < 4; ++k) r[k] += s_red[k][w];
    const float inv = 1.0f / (float)P;
    norm_avg[b] = r[0] * inv;
    xyz_avg[b * 3] = r[1] * inv; xyz_avg[b * 3 + 1] = r[2] * inv; xyz_avg[b * 3 + 2] = r[3] * inv;
    const float tx = Tp.m[3] - Tg.m[3], ty = Tp.m[7] - Tg.m[7], tz = Tp.m[11] - Tg.m[11];
    tco_xyz[b * 3] = fabsf(tx); tco_xyz[b * 3 + 1] = fabsf(ty); tco_xyz[b * 3 + 2] = fabsf(tz);
    tco_norm[b] = sqrtf(tx * tx + ty * ty + tz * tz);
  }
}

}  // namespace cosyb

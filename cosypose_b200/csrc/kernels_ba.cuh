// Object-level bundle adjustment kernels (reference: multiview/bundle_adjustment.py:159-222).
// The reference obtains the Jacobian of the reprojected points by autograd over
// [n_residuals, n_objects + n_views, 9] replicated parameters (:175-214) and builds J^T J on the
// device; here the derivatives are analytic (forward mode through the 6D -> rotation map) and
// J^T J, J^T e are produced directly.  The 9-D parameterisation is the reference's:
// [R[:,0], R[:,1], t] (extract_pose9d :159-162, compute_transform_from_pose9d
// lib3d/transform_ops.py:53-64, compute_rotation_matrix_from_ortho6d lib3d/rotations.py:6-21).
#pragma once
#include "common.h"
#include "kernels_ransac.cuh"

namespace cosyb {

// R (columns x,y,z) from 6 numbers a|b, plus dR/d(a|b): dR[k][r*3+c], k = 0..5
__device__ __forceinline__ void rot6d_with_jac(const float* p, float* R, float (*dR)[9]) {
  const float a[3] = {p[0], p[1], p[2]}, b[3] = {p[3], p[4], p[5]};
  const float na = sqrtf(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
  const float x[3] = {a[0] / na, a[1] / na, a[2] / na};
  float w[3] = {x[1] * b[2] - x[2] * b[1], x[2] * b[0] - x[0] * b[2], x[0] * b[1] - x[1] * b[0]};
  const float nw = sqrtf(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
  const float z[3] = {w[0] / nw, w[1] / nw, w[2] / nw};
  const float y[3] = {z[1] * x[2] - z[2] * x[1], z[2] * x[0] - z[0] * x[2], z[0] * x[1] - z[1] * x[0]};
#pragma unroll
  for (int r = 0; r < 3; ++r) { R[r * 3 + 0] = x[r]; R[r * 3 + 1] = y[r]; R[r * 3 + 2] = z[r]; }
  if (dR == nullptr) return;
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    float da[3] = {0.f, 0.f, 0.f}, db[3] = {0.f, 0.f, 0.f};
    if (k < 3) da[k] = 1.f; else db[k - 3] = 1.f;
    // dx = (da - x (x.da)) / |a|
    const float xda = x[0] * da[0] + x[1] * da[1] + x[2] * da[2];
    const float dx[3] = {(da[0] - x[0] * xda) / na, (da[1] - x[1] * xda) / na, (da[2] - x[2] * xda) / na};
    // dw = dx x b + x x db
    const float dw[3] = {dx[1] * b[2] - dx[2] * b[1] + x[1] * db[2] - x[2] * db[1],
                         dx[2] * b[0] - dx[0] * b[2] + x[2] * db[0] - x[0] * db[2],
                         dx[0] * b[1] - dx[1] * b[0] + x[0] * db[1] - x[1] * db[0]};
    const float zdw = z[0] * dw[0] + z[1] * dw[1] + z[2] * dw[2];
    const float dz[3] = {(dw[0] - z[0] * zdw) / nw, (dw[1] - z[1] * zdw) / nw, (dw[2] - z[2] * zdw) / nw};
    // dy = dz x x + z x dx
    const float dy[3] = {dz[1] * x[2] - dz[2] * x[1] + z[1] * dx[2] - z[2] * dx[1],
                         dz[2] * x[0] - dz[0] * x[2] + z[2] * dx[0] - z[0] * dx[2],
                         dz[0] * x[1] - dz[1] * x[0] + z[0] * dx[1] - z[1] * dx[0]};
#pragma unroll
    for (int r = 0; r < 3; ++r) { dR[k][r * 3 + 0] = dx[r]; dR[k][r * 3 + 1] = dy[r]; dR[k][r * 3 + 2] = dz[r]; }
  }
}

__device__ __forceinline__ Mat34 mat34_from_9d(const float* p) {
  float R[9];
  rot6d_with_jac(p, R, nullptr);
  Mat34 T;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    T.m[r * 4 + 0] = R[r * 3 + 0]; T.m[r * 4 + 1] = R[r * 3 + 1]; T.m[r * 4 + 2] = R[r * 3 + 2];
    T.m[r * 4 + 3] = p[6 + r];
  }
  return T;
}

// project_points (NOT the robust variant; lib3d/camera_geometry.py:4-15): K @ (T p), divide by z
__device__ __forceinline__ void project34(const float* K, const Mat34& T, const float* p, float* uv) {
  float P[3];
  apply34(T, p, P);
  const float s0 = K[0] * P[0] + K[1] * P[1] + K[2] * P[2];
  const float s1 = K[3] * P[0] + K[4] * P[1] + K[5] * P[2];
  const float s2 = K[6] * P[0] + K[7] * P[1] + K[8] * P[2];
  uv[0] = s0 / s2;
  uv[1] = s1 / s2;
}

// out[i] = inv(A[ia[i]]) @ B[ib[i]]   (invert_T(TWC) @ TWO: multiview_predictor.py:38, ransac.py:173)
__global__ void k_compose_inv(int64_t n, const float* __restrict__ A, const int32_t* __restrict__ ia,
                              const float* __restrict__ B, const int32_t* __restrict__ ib, float* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Mat34 T = mul34(inv34(load34(A + (size_t)(ia ? ia[i] : i) * 16)), load34(B + (size_t)(ib ? ib[i] : i) * 16));
  float* o = out + i * 16;
#pragma unroll
  for (int k = 0; k < 12; ++k) o[k] = T.m[k];
  o[12] = 0.f; o[13] = 0.f; o[14] = 0.f; o[15] = 1.f;
}

// align_TCO_cand (bundle_adjustment.py:164-173) = symmetric_distance_reprojected
// (lib3d/symmetric_distances.py:105-121) over the label's n_sym REAL symmetries (not the padded
// set): dist_k = mean_p || proj(K, cand_TCO S_k, p) - proj(K, TCO_est, p) ||, first minimum wins.
// One thread per candidate (n_cand is small).  Writes dists and cand_TCO @ S*.
__global__ void k_ba_align(int n_cand, int n_pts, const float* __restrict__ cand_TCO,
                           const int32_t* __restrict__ cand_obj, const int32_t* __restrict__ cand_view,
                           const int32_t* __restrict__ cand_label, const float* __restrict__ TWO_9d,
                           const float* __restrict__ TCW_9d, const float* __restrict__ K,
                           const float* __restrict__ points /*[L][n_pts][3]*/, const float* __restrict__ sym,
                           const int32_t* __restrict__ n_sym, int s_max, float* __restrict__ dists,
                           float* __restrict__ aligned /*[n_cand][16]*/) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_cand) return;
  const int l = cand_label[c], v = cand_view[c];
  const Mat34 Tc = load34(cand_TCO + (size_t)c * 16);
  const Mat34 Te = mul34(mat34_from_9d(TCW_9d + v * 9), mat34_from_9d(TWO_9d + cand_obj[c] * 9));
  const float* Kv = K + v * 9;
  const float* pts = points + (size_t)l * n_pts * 3;
  float best = 0.f;
  int best_k = -1;
  for (int k = 0; k < n_sym[l]; ++k) {
    Mat34 T1 = mul34(Tc, load34(sym + ((size_t)l * s_max + k) * 16));
    float d = 0.f;
    for (int p = 0; p < n_pts; ++p) {
      float pt[3] = {pts[p * 3], pts[p * 3 + 1], pts[p * 3 + 2]}, a[2], b[2];
      project34(Kv, T1, pt, a);
      project34(Kv, Te, pt, b);
      d += sqrtf((a[0] - b[0]) * (a[0] - b[0]) + (a[1] - b[1]) * (a[1] - b[1]));
    }
    d /= (float)n_pts;
    if (best_k < 0 || d < best) { best = d; best_k = k; }
  }
  if (best_k < 0) best_k = 0;
  dists[c] = best;
  Mat34 Ta = mul34(Tc, load34(sym + ((size_t)l * s_max + best_k) * 16));
  float* o = aligned + (size_t)c * 16;
#pragma unroll
  for (int k = 0; k < 12; ++k) o[k] = Ta.m[k];
  o[12] = 0.f; o[13] = 0.f; o[14] = 0.f; o[15] = 1.f;
}

// Residuals and compact Jacobian rows, one thread per (candidate, point):
//   rows r = (c * n_pts + p) * 2 + {0: x, 1: y}   (make_residuals_ids, bundle_adjustment.py:91-110)
//   errors[r] = y - yhat;  Jc[r][0:9] = d yhat / d TWO_9d[obj], Jc[r][9:18] = d yhat / d TCW_9d[view]
__global__ void k_ba_residuals(int n_cand, int n_pts, const float* __restrict__ aligned,
                               const int32_t* __restrict__ cand_obj, const int32_t* __restrict__ cand_view,
                               const int32_t* __restrict__ cand_label, const float* __restrict__ TWO_9d,
                               const float* __restrict__ TCW_9d, const float* __restrict__ K,
                               const float* __restrict__ points, float* __restrict__ errors,
                               float* __restrict__ Jc /*[n_res][18]*/) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_cand * n_pts) return;
  const int c = idx / n_pts, p = idx % n_pts;
  const int o = cand_obj[c], v = cand_view[c];
  const float* po = TWO_9d + o * 9;
  const float* pv = TCW_9d + v * 9;
  const float* Kv = K + v * 9;
  const float* ptp = points + ((size_t)cand_label[c] * n_pts + p) * 3;
  const float pt[3] = {ptp[0], ptp[1], ptp[2]};
  float Rwo[9], dRwo[6][9], Rcw[9], dRcw[6][9];
  rot6d_with_jac(po, Rwo, dRwo);
  rot6d_with_jac(pv, Rcw, dRcw);
  float q[3], P[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) q[r] = Rwo[r * 3] * pt[0] + Rwo[r * 3 + 1] * pt[1] + Rwo[r * 3 + 2] * pt[2] + po[6 + r];
#pragma unroll
  for (int r = 0; r < 3; ++r) P[r] = Rcw[r * 3] * q[0] + Rcw[r * 3 + 1] * q[1] + Rcw[r * 3 + 2] * q[2] + pv[6 + r];
  const float s0 = Kv[0] * P[0] + Kv[1] * P[1] + Kv[2] * P[2];
  const float s1 = Kv[3] * P[0] + Kv[4] * P[1] + Kv[5] * P[2];
  const float s2 = Kv[6] * P[0] + Kv[7] * P[1] + Kv[8] * P[2];
  const float yhat[2] = {s0 / s2, s1 / s2};
  // d(u,v)/dP: g[xy][j]
  float g[2][3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    g[0][j] = Kv[0 + j] / s2 - s0 * Kv[6 + j] / (s2 * s2);
    g[1][j] = Kv[3 + j] / s2 - s1 * Kv[6 + j] / (s2 * s2);
  }
  float ycand[2];
  project34(Kv, load34(aligned + (size_t)c * 16), pt, ycand);
  const size_t r0 = ((size_t)c * n_pts + p) * 2;
#pragma unroll
  for (int xy = 0; xy < 2; ++xy) {
    errors[r0 + xy] = ycand[xy] - yhat[xy];
    float* row = Jc + (r0 + xy) * 18;
    // gw = g . Rcw  (d/dq)
    float gw[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) gw[j] = g[xy][0] * Rcw[0 * 3 + j] + g[xy][1] * Rcw[1 * 3 + j] + g[xy][2] * Rcw[2 * 3 + j];
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      float dq[3], dP[3];
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        dq[r] = dRwo[k][r * 3] * pt[0] + dRwo[k][r * 3 + 1] * pt[1] + dRwo[k][r * 3 + 2] * pt[2];
        dP[r] = dRcw[k][r * 3] * q[0] + dRcw[k][r * 3 + 1] * q[1] + dRcw[k][r * 3 + 2] * q[2];
      }
      row[k] = gw[0] * dq[0] + gw[1] * dq[1] + gw[2] * dq[2];
      row[9 + k] = g[xy][0] * dP[0] + g[xy][1] * dP[1] + g[xy][2] * dP[2];
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      row[6 + j] = gw[j];
      row[15 + j] = g[xy][j];
    }
  }
}

// Normal equations from the compact rows.  Parameter a < 9*n_obj belongs to object a/9, otherwise to
// view (a - 9*n_obj)/9 (the order of torch.cat((J_TWO, J_TCW)), bundle_adjustment.py:251).
//   JtJ[a][b] = sum_r J[r][a] J[r][b];  Jte[a] = sum_r J[r][a] e[r]   (column n_params of the grid)
__global__ void k_ba_normal(int n_cand, int n_pts, int n_obj, int n_view, const int32_t* __restrict__ cand_obj,
                            const int32_t* __restrict__ cand_view, const float* __restrict__ Jc,
                            const float* __restrict__ errors, float* __restrict__ JtJ, float* __restrict__ Jte) {
  const int n_params = 9 * (n_obj + n_view);
  const int a = blockIdx.y * blockDim.y + threadIdx.y;
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= n_params || b > n_params) return;
  const bool a_obj = a < 9 * n_obj, b_obj = b < 9 * n_obj, b_err = b == n_params;
  const int a_id = a_obj ? a / 9 : (a - 9 * n_obj) / 9, a_k = a_obj ? a % 9 : 9 + (a - 9 * n_obj) % 9;
  const int b_id = b_obj ? b / 9 : (b - 9 * n_obj) / 9, b_k = b_obj ? b % 9 : 9 + (b - 9 * n_obj) % 9;
  float acc = 0.f;
  const int rows_per_cand = n_pts * 2;
  for (int c = 0; c < n_cand; ++c) {
    if ((a_obj ? cand_obj[c] : cand_view[c]) != a_id) continue;
    if (!b_err && (b_obj ? cand_obj[c] : cand_view[c]) != b_id) continue;
    const float* rows = Jc + (size_t)c * rows_per_cand * 18;
    if (b_err) {
      for (int r = 0; r < rows_per_cand; ++r) acc = fmaf(rows[r * 18 + a_k], errors[(size_t)c * rows_per_cand + r], acc);
    } else {
      for (int r = 0; r < rows_per_cand; ++r) acc = fmaf(rows[r * 18 + a_k], rows[r * 18 + b_k], acc);
    }
  }
  if (b_err) Jte[a] = acc; else JtJ[(size_t)a * n_params + b] = acc;
}

// loss = mean(min(e^2, thr))  (bundle_adjustment.py:205-208); single CTA, fixed order
__global__ void k_ba_loss(int n_res, const float* __restrict__ errors, float thr, float* __restrict__ loss) {
  __shared__ float s[256];
  float acc = 0.f;
  for (int i = threadIdx.x; i < n_res; i += 256) acc += fminf(errors[i] * errors[i], thr);
  s[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *loss = s[0] / (float)n_res;
}

// ---- float64 variants ------------------------------------------------------------------------------------
// The LM normal equations are ill conditioned in fp32 (pixel-scale Jacobians squared against lambda down to 1e-7):
// the reference's own fp32 result is ~1e-4 from the same code run in float64 (tests/golden/scene_state_*_fp64.npz).
// The problem is tiny (config 4: 2048 residuals x 216 parameters), so residuals, Jacobian, J^T J, J^T e, the loss and
// the damped solve are evaluated in float64 on the device; parameters and outputs stay fp32.
__device__ __forceinline__ void rot6d_with_jac_d(const double* p, double* R, double (*dR)[9]) {
  const double a[3] = {p[0], p[1], p[2]}, b[3] = {p[3], p[4], p[5]};
  const double na = sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
  const double x[3] = {a[0] / na, a[1] / na, a[2] / na};
  double w[3] = {x[1] * b[2] - x[2] * b[1], x[2] * b[0] - x[0] * b[2], x[0] * b[1] - x[1] * b[0]};
  const double nw = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
  const double z[3] = {w[0] / nw, w[1] / nw, w[2] / nw};
  const double y[3] = {z[1] * x[2] - z[2] * x[1], z[2] * x[0] - z[0] * x[2], z[0] * x[1] - z[1] * x[0]};
#pragma unroll
  for (int r = 0; r < 3; ++r) { R[r * 3 + 0] = x[r]; R[r * 3 + 1] = y[r]; R[r * 3 + 2] = z[r]; }
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    double da[3] = {0., 0., 0.}, db[3] = {0., 0., 0.};
    if (k < 3) da[k] = 1.; else db[k - 3] = 1.;
    const double xda = x[0] * da[0] + x[1] * da[1] + x[2] * da[2];
    const double dx[3] = {(da[0] - x[0] * xda) / na, (da[1] - x[1] * xda) / na, (da[2] - x[2] * xda) / na};
    const double dw[3] = {dx[1] * b[2] - dx[2] * b[1] + x[1] * db[2] - x[2] * db[1],
                          dx[2] * b[0] - dx[0] * b[2] + x[2] * db[0] - x[0] * db[2],
                          dx[0] * b[1] - dx[1] * b[0] + x[0] * db[1] - x[1] * db[0]};
    const double zdw = z[0] * dw[0] + z[1] * dw[1] + z[2] * dw[2];
    const double dz[3] = {(dw[0] - z[0] * zdw) / nw, (dw[1] - z[1] * zdw) / nw, (dw[2] - z[2] * zdw) / nw};
    const double dy[3] = {dz[1] * x[2] - dz[2] * x[1] + z[1] * dx[2] - z[2] * dx[1],
                          dz[2] * x[0] - dz[0] * x[2] + z[2] * dx[0] - z[0] * dx[2],
                          dz[0] * x[1] - dz[1] * x[0] + z[0] * dx[1] - z[1] * dx[0]};
#pragma unroll
    for (int r = 0; r < 3; ++r) { dR[k][r * 3 + 0] = dx[r]; dR[k][r * 3 + 1] = dy[r]; dR[k][r * 3 + 2] = dz[r]; }
  }
}

// as k_ba_residuals, float64 arithmetic and outputs (errors64 [n_res], Jc64 [n_res][18])
__global__ void k_ba_residuals_d(int n_cand, int n_pts, const float* __restrict__ aligned,
                                 const int32_t* __restrict__ cand_obj, const int32_t* __restrict__ cand_view,
                                 const int32_t* __restrict__ cand_label, const float* __restrict__ TWO_9d,
                                 const float* __restrict__ TCW_9d, const float* __restrict__ K,
                                 const float* __restrict__ points, double* __restrict__ errors, double* __restrict__ Jc) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_cand * n_pts) return;
  const int c = idx / n_pts, p = idx % n_pts;
  const int o = cand_obj[c], v = cand_view[c];
  double po[9], pv[9], Kv[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) { po[i] = TWO_9d[o * 9 + i]; pv[i] = TCW_9d[v * 9 + i]; Kv[i] = K[v * 9 + i]; }
  const float* ptp = points + ((size_t)cand_label[c] * n_pts + p) * 3;
  const double pt[3] = {ptp[0], ptp[1], ptp[2]};
  double Rwo[9], dRwo[6][9], Rcw[9], dRcw[6][9];
  rot6d_with_jac_d(po, Rwo, dRwo);
  rot6d_with_jac_d(pv, Rcw, dRcw);
  double q[3], P[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) q[r] = Rwo[r * 3] * pt[0] + Rwo[r * 3 + 1] * pt[1] + Rwo[r * 3 + 2] * pt[2] + po[6 + r];
#pragma unroll
  for (int r = 0; r < 3; ++r) P[r] = Rcw[r * 3] * q[0] + Rcw[r * 3 + 1] * q[1] + Rcw[r * 3 + 2] * q[2] + pv[6 + r];
  const double s0 = Kv[0] * P[0] + Kv[1] * P[1] + Kv[2] * P[2];
  const double s1 = Kv[3] * P[0] + Kv[4] * P[1] + Kv[5] * P[2];
  const double s2 = Kv[6] * P[0] + Kv[7] * P[1] + Kv[8] * P[2];
  const double yhat[2] = {s0 / s2, s1 / s2};
  double g[2][3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    g[0][j] = Kv[0 + j] / s2 - s0 * Kv[6 + j] / (s2 * s2);
    g[1][j] = Kv[3 + j] / s2 - s1 * Kv[6 + j] / (s2 * s2);
  }
  // the candidate's reprojection (aligned pose, fp32 data) in float64
  const float* Ta = aligned + (size_t)c * 16;
  double Pc[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) Pc[r] = (double)Ta[r * 4] * pt[0] + (double)Ta[r * 4 + 1] * pt[1] + (double)Ta[r * 4 + 2] * pt[2] + (double)Ta[r * 4 + 3];
  const double c0 = Kv[0] * Pc[0] + Kv[1] * Pc[1] + Kv[2] * Pc[2];
  const double c1 = Kv[3] * Pc[0] + Kv[4] * Pc[1] + Kv[5] * Pc[2];
  const double c2 = Kv[6] * Pc[0] + Kv[7] * Pc[1] + Kv[8] * Pc[2];
  const double ycand[2] = {c0 / c2, c1 / c2};
  const size_t r0 = ((size_t)c * n_pts + p) * 2;
#pragma unroll
  for (int xy = 0; xy < 2; ++xy) {
    errors[r0 + xy] = ycand[xy] - yhat[xy];
    double* row = Jc + (r0 + xy) * 18;
    double gw[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) gw[j] = g[xy][0] * Rcw[0 * 3 + j] + g[xy][1] * Rcw[1 * 3 + j] + g[xy][2] * Rcw[2 * 3 + j];
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      double dq[3], dP[3];
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        dq[r] = dRwo[k][r * 3] * pt[0] + dRwo[k][r * 3 + 1] * pt[1] + dRwo[k][r * 3 + 2] * pt[2];
        dP[r] = dRcw[k][r * 3] * q[0] + dRcw[k][r * 3 + 1] * q[1] + dRcw[k][r * 3 + 2] * q[2];
      }
      row[k] = gw[0] * dq[0] + gw[1] * dq[1] + gw[2] * dq[2];
      row[9 + k] = g[xy][0] * dP[0] + g[xy][1] * dP[1] + g[xy][2] * dP[2];
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      row[6 + j] = gw[j];
      row[15 + j] = g[xy][j];
    }
  }
}

__global__ void k_ba_normal_d(int n_cand, int n_pts, int n_obj, int n_view, const int32_t* __restrict__ cand_obj,
                              const int32_t* __restrict__ cand_view, const double* __restrict__ Jc,
                              const double* __restrict__ errors, double* __restrict__ JtJ, double* __restrict__ Jte) {
  const int n_params = 9 * (n_obj + n_view);
  const int a = blockIdx.y * blockDim.y + threadIdx.y;
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= n_params || b > n_params) return;
  const bool a_obj = a < 9 * n_obj, b_obj = b < 9 * n_obj, b_err = b == n_params;
  const int a_id = a_obj ? a / 9 : (a - 9 * n_obj) / 9, a_k = a_obj ? a % 9 : 9 + (a - 9 * n_obj) % 9;
  const int b_id = b_obj ? b / 9 : (b - 9 * n_obj) / 9, b_k = b_obj ? b % 9 : 9 + (b - 9 * n_obj) % 9;
  double acc = 0.;
  const int rows_per_cand = n_pts * 2;
  for (int c = 0; c < n_cand; ++c) {
    if ((a_obj ? cand_obj[c] : cand_view[c]) != a_id) continue;
    if (!b_err && (b_obj ? cand_obj[c] : cand_view[c]) != b_id) continue;
    const double* rows = Jc + (size_t)c * rows_per_cand * 18;
    if (b_err) {
      for (int r = 0; r < rows_per_cand; ++r) acc = fma(rows[r * 18 + a_k], errors[(size_t)c * rows_per_cand + r], acc);
    } else {
      for (int r = 0; r < rows_per_cand; ++r) acc = fma(rows[r * 18 + a_k], rows[r * 18 + b_k], acc);
    }
  }
  if (b_err) Jte[a] = acc; else JtJ[(size_t)a * n_params + b] = acc;
}

__global__ void k_ba_loss_d(int n_res, const double* __restrict__ errors, double thr, double* __restrict__ loss) {
  __shared__ double s[256];
  double acc = 0.;
  for (int i = threadIdx.x; i < n_res; i += 256) acc += fmin(errors[i] * errors[i], thr);
  s[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *loss = s[0] / (double)n_res;
}

// LM step h = (JtJ + lambda I)^-1 Jte (compute_lm_step, bundle_adjustment.py:216-222: pinverse of a positive
// definite matrix) by a Cholesky factorisation in float64: ONE CTA, right-looking; the packed lower triangle lives in
// shared memory when it fits (n <= 220: config 4 has n = 216 -> 187 KB), otherwise in `Ag` (global, L2 resident).  A pivot that is not positive (cannot happen
// for J^T J + lambda I in exact arithmetic) is replaced by lambda and counted in *n_bad.
constexpr int LM_THREADS = 1024;
constexpr int LM_MAX_N = 2048;
constexpr int LM_SMEM_MAX_N = 220;      // packed lower triangle in shared memory: n (n + 1) / 2 doubles <= 190 KB

// SMEM: the lower triangle lives in dynamic shared memory (packed rows), otherwise in `Ag` (global, n x n)
template <bool SMEM>
__global__ void __launch_bounds__(LM_THREADS)
k_lm_solve(int n, const double* __restrict__ JtJ, const double* __restrict__ Jte, double lambda,
           double* __restrict__ Ag, float* __restrict__ step, int* __restrict__ n_bad) {
  extern __shared__ double s_tri[];
  __shared__ double s_col[SMEM ? LM_SMEM_MAX_N : LM_MAX_N];
  __shared__ double s_y[SMEM ? LM_SMEM_MAX_N : LM_MAX_N];
  __shared__ int s_bad;
  const int tid = threadIdx.x;
  auto at = [&](int i, int j) -> double& {       // j <= i
    if (SMEM) return s_tri[i * (i + 1) / 2 + j];
    return Ag[(size_t)i * n + j];
  };
  if (tid == 0) s_bad = 0;
  for (int idx = tid; idx < n * n; idx += LM_THREADS) {
    const int i = idx / n, j = idx % n;
    if (j <= i) at(i, j) = JtJ[idx] + (i == j ? lambda : 0.0);
  }
  __syncthreads();
  for (int k = 0; k < n; ++k) {
    double akk = at(k, k);
    if (!(akk > 0.0)) { akk = lambda > 0.0 ? lambda : 1e-300; if (tid == 0) s_bad += 1; }
    const double d = sqrt(akk);
    for (int i = k + tid; i < n; i += LM_THREADS) s_col[i] = i == k ? d : at(i, k) / d;
    __syncthreads();
    for (int i = k + tid; i < n; i += LM_THREADS) at(i, k) = s_col[i];   // L[:, k]
    // trailing update of the lower triangle: A[i][j] -= L[i][k] L[j][k], k < j <= i
    const int m = n - k - 1;
    for (int idx = tid; idx < m * m; idx += LM_THREADS) {
      const int i = k + 1 + idx / m, j = k + 1 + idx % m;
      if (j <= i) at(i, j) -= s_col[i] * s_col[j];
    }
    __syncthreads();
  }
  // forward substitution L y = b, then back substitution L^T h = y (column sweeps, one sync per column)
  for (int i = tid; i < n; i += LM_THREADS) s_y[i] = Jte[i];
  __syncthreads();
  for (int k = 0; k < n; ++k) {
    const double yk = s_y[k] / at(k, k);
    __syncthreads();
    if (tid == 0) s_y[k] = yk;
    for (int i = k + 1 + tid; i < n; i += LM_THREADS) s_y[i] -= at(i, k) * yk;
    __syncthreads();
  }
  for (int k = n - 1; k >= 0; --k) {
    const double hk = s_y[k] / at(k, k);
    __syncthreads();
    if (tid == 0) s_y[k] = hk;
    for (int i = tid; i < k; i += LM_THREADS) s_y[i] -= at(k, i) * hk;
    __syncthreads();
  }
  for (int i = tid; i < n; i += LM_THREADS) step[i] = (float)s_y[i];
  if (tid == 0 && n_bad) *n_bad = s_bad;
}

}  // namespace cosyb

// Multiview RANSAC kernels: symmetric distance on the 8 AABB corners, camera-pose models per seed,
// scoring of every tentative match under every hypothesis.  fp32, tiny data: ALU/latency bound.
//   symmetric_distance_batched_fast  lib3d/symmetric_distances.py:38-57
//       d(T1,T2) = mean_p || (T1 S*) p - T2 p ||,  S* = argmin_S mean_p ||(T1 S) p - T2 p||^2
//       over ALL s_max (identity padded) symmetries of the label, first minimum wins
//   estimate_camera_poses            multiview/ransac.py:19-47
//   score_tmatches                   multiview/ransac.py:67-73
// A row is handled by GS cooperating lanes (GS = power of two <= 32 chosen from s_max).
#pragma once
#include "common.h"

namespace cosyb {

struct Mat34 { float m[12]; };  // rows of [R|t]

// lanes of one cooperating group inside the warp (groups may diverge from each other)
template <int GS>
__device__ __forceinline__ unsigned group_mask() {
  if constexpr (GS == 32) {
    return 0xffffffffu;
  } else {
    return ((1u << GS) - 1u) << (((threadIdx.x % 32) / GS) * GS);
  }
}

__device__ __forceinline__ Mat34 load34(const float* __restrict__ T) {
  Mat34 r;
#pragma unroll
  for (int i = 0; i < 12; ++i) r.m[i] = T[i];
  return r;
}
// C = A @ B for rigid [R|t] 3x4 blocks (bottom rows implied 0 0 0 1)
__device__ __forceinline__ Mat34 mul34(const Mat34& A, const Mat34& B) {
  Mat34 C;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float v = A.m[r * 4 + 0] * B.m[0 * 4 + c];
      v = fmaf(A.m[r * 4 + 1], B.m[1 * 4 + c], v);
      v = fmaf(A.m[r * 4 + 2], B.m[2 * 4 + c], v);
      if (c == 3) v += A.m[r * 4 + 3];
      C.m[r * 4 + c] = v;
    }
  }
  return C;
}
// invert_T (reference: lib3d/transform_ops.py:24-32): R^T, -R^T t
__device__ __forceinline__ Mat34 inv34(const Mat34& A) {
  Mat34 C;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
#pragma unroll
    for (int c = 0; c < 3; ++c) C.m[r * 4 + c] = A.m[c * 4 + r];
    float v = A.m[0 * 4 + r] * A.m[3];
    v = fmaf(A.m[1 * 4 + r], A.m[7], v);
    v = fmaf(A.m[2 * 4 + r], A.m[11], v);
    C.m[r * 4 + 3] = -v;
  }
  return C;
}
__device__ __forceinline__ void apply34(const Mat34& T, const float* p, float* o) {
#pragma unroll
  for (int r = 0; r < 3; ++r)
    o[r] = fmaf(T.m[r * 4 + 2], p[2], fmaf(T.m[r * 4 + 1], p[1], T.m[r * 4 + 0] * p[0])) + T.m[r * 4 + 3];
}

// Cooperative over GS lanes (lane = index inside the group).  Every lane returns the result.
template <int GS>
__device__ __forceinline__ float symdist_group(const Mat34& T1, const Mat34& T2,
                                               const float* __restrict__ aabb /*[8][3]*/,
                                               const float* __restrict__ sym /*[s_max][16]*/,
                                               int s_max, int lane, unsigned mask, int* best_sym) {
  float q[8][3], pts[8][3];
#pragma unroll
  for (int p = 0; p < 8; ++p) {
#pragma unroll
    for (int i = 0; i < 3; ++i) pts[p][i] = __ldg(aabb + p * 3 + i);
    apply34(T2, pts[p], q[p]);
  }
  float best_v = INFINITY;
  int best_s = 0x7fffffff;
  for (int s = lane; s < s_max; s += GS) {
    Mat34 S = load34(sym + (size_t)s * 16);
    Mat34 M = mul34(T1, S);
    float sum = 0.f;
#pragma unroll
    for (int p = 0; p < 8; ++p) {
      float o[3];
      apply34(M, pts[p], o);
      float dx = o[0] - q[p][0], dy = o[1] - q[p][1], dz = o[2] - q[p][2];
      sum += dx * dx + dy * dy + dz * dz;
    }
    float v = sum * 0.125f;
    if (v < best_v || (v == best_v && s < best_s)) { best_v = v; best_s = s; }
  }
#pragma unroll
  for (int o = GS / 2; o > 0; o >>= 1) {
    float ov = __shfl_xor_sync(mask, best_v, o, GS);
    int os = __shfl_xor_sync(mask, best_s, o, GS);
    if (ov < best_v || (ov == best_v && os < best_s)) { best_v = ov; best_s = os; }
  }
  if (best_s == 0x7fffffff) best_s = 0;  // all NaN: argmin of the first entry
  Mat34 S = load34(sym + (size_t)best_s * 16);
  Mat34 M = mul34(T1, S);
  float d = 0.f;
#pragma unroll
  for (int p = 0; p < 8; ++p) {
    float o[3];
    apply34(M, pts[p], o);
    float dx = o[0] - q[p][0], dy = o[1] - q[p][1], dz = o[2] - q[p][2];
    d += sqrtf(dx * dx + dy * dy + dz * dz);
  }
  if (best_sym) *best_sym = best_s;
  return d * 0.125f;
}

constexpr int RANSAC_THREADS = 128;

template <int GS>
__global__ void __launch_bounds__(RANSAC_THREADS)
k_symmetric_distance(int64_t n, const float* __restrict__ T1, const float* __restrict__ T2,
                     const int32_t* __restrict__ label_ids, const float* __restrict__ aabb,
                     const float* __restrict__ sym, int s_max, float* __restrict__ dists,
                     int32_t* __restrict__ best_sym) {
  int64_t row = ((int64_t)blockIdx.x * RANSAC_THREADS + threadIdx.x) / GS;
  int lane = threadIdx.x % GS;
  if (row >= n) row = n - 1;  // keep the whole group converged for the shuffles
  int l = label_ids[row];
  int bs;
  float d = symdist_group<GS>(load34(T1 + row * 16), load34(T2 + row * 16), aabb + (size_t)l * 24,
                              sym + (size_t)l * s_max * 16, s_max, lane, group_mask<GS>(), &bs);
  if (lane == 0 && ((int64_t)blockIdx.x * RANSAC_THREADS + threadIdx.x) / GS < n) {
    dists[row] = d;
    if (best_sym) best_sym[row] = bs;
  }
}

// seeds [6][n_seeds]: view1, view2, m1c1, m1c2, m2c1, m2c2
template <int GS>
__global__ void __launch_bounds__(RANSAC_THREADS)
k_ransac_models(int64_t n_seeds, const float* __restrict__ poses, const int32_t* __restrict__ cand_labels,
                const int32_t* __restrict__ seeds, const float* __restrict__ aabb,
                const float* __restrict__ sym, const int32_t* __restrict__ n_sym, int s_max,
                float* __restrict__ TC1C2) {
  int64_t gid = ((int64_t)blockIdx.x * RANSAC_THREADS + threadIdx.x) / GS;
  const bool live = gid < n_seeds;
  int64_t row = live ? gid : n_seeds - 1;
  int lane = threadIdx.x % GS;
  int a = seeds[2 * n_seeds + row], bb = seeds[3 * n_seeds + row];
  int g = seeds[4 * n_seeds + row], dd = seeds[5 * n_seeds + row];
  int lab_ab = cand_labels[a], lab_gd = cand_labels[g];
  Mat34 TC1Oa = load34(poses + (size_t)a * 16);
  Mat34 TObC2 = inv34(load34(poses + (size_t)bb * 16));
  Mat34 TC1Og = load34(poses + (size_t)g * 16);
  Mat34 TC2Od = load34(poses + (size_t)dd * 16);
  const float* sym_ab = sym + (size_t)lab_ab * s_max * 16;
  const float* sym_gd = sym + (size_t)lab_gd * s_max * 16;
  const float* aabb_gd = aabb + (size_t)lab_gd * 24;
  float best_v = 0.f;
  int best_k = -1;
  const int nk = n_sym[lab_ab];
  for (int k = 0; k < nk; ++k) {
    Mat34 TaS = mul34(TC1Oa, load34(sym_ab + (size_t)k * 16));
    Mat34 T2 = mul34(mul34(TaS, TObC2), TC2Od);
    float d = symdist_group<GS>(TC1Og, T2, aabb_gd, sym_gd, s_max, lane, group_mask<GS>(), nullptr);
    if (best_k < 0 || d < best_v) { best_v = d; best_k = k; }  // first minimum (cext scatter_argmin)
  }
  if (best_k < 0) best_k = 0;
  if (lane == 0 && live) {
    Mat34 T = mul34(mul34(TC1Oa, load34(sym_ab + (size_t)best_k * 16)), TObC2);
    float* o = TC1C2 + row * 16;
#pragma unroll
    for (int i = 0; i < 12; ++i) o[i] = T.m[i];
    o[12] = 0.f; o[13] = 0.f; o[14] = 0.f; o[15] = 1.f;
  }
}

// tmatches [3][n]: hypothesis_id, cand1, cand2
template <int GS>
__global__ void __launch_bounds__(RANSAC_THREADS)
k_ransac_score(int64_t n, const float* __restrict__ poses, const int32_t* __restrict__ cand_labels,
               const int32_t* __restrict__ tmatches, const float* __restrict__ TC1C2,
               const float* __restrict__ aabb, const float* __restrict__ sym, int s_max,
               float* __restrict__ dists) {
  int64_t gid = ((int64_t)blockIdx.x * RANSAC_THREADS + threadIdx.x) / GS;
  const bool live = gid < n;
  int64_t row = live ? gid : n - 1;
  int lane = threadIdx.x % GS;
  int hyp = tmatches[row], c1 = tmatches[n + row], c2 = tmatches[2 * n + row];
  int l = cand_labels[c1];
  Mat34 TWOa = load34(poses + (size_t)c1 * 16);
  Mat34 TWOb = mul34(load34(TC1C2 + (size_t)hyp * 16), load34(poses + (size_t)c2 * 16));
  float d = symdist_group<GS>(TWOa, TWOb, aabb + (size_t)l * 24, sym + (size_t)l * s_max * 16, s_max,
                              lane, group_mask<GS>(), nullptr);
  if (lane == 0 && live) dists[row] = d;
}


// ---- device-side inlier voting ------------------------------------------------------------------------------
// cosypose_cext.find_ransac_inliers (reference: csrc/cosypose_cext.cpp:107-216) without the device -> host copy of
// the distances: same integers out, bit for bit (tests/test_gpu_multiview.py compares with the host implementation,
// which is itself pinned to the reference's compiled extension).
//   k_vote_greedy   one warp per hypothesis: rows with dist <= thr, greedy one-to-one matching in ascending
//                   (dist, row) order = the reference's stable sort + first-fit scan, as "repeatedly take the smallest
//                   still-eligible row"; fp32 sum of the accepted distances in acceptance order.
//   k_vote_best     one warp per ordered view pair: most inliers (>= n_min), ties by smaller distance sum, then by
//                   lower hypothesis id; a pair whose best hypothesis is id 0 is dropped (cosypose_cext.cpp:203).
//   k_vote_emit     exclusive scan over the pairs (one block) and the ordered lists of inlier matches / best ids.
namespace vote {
constexpr int WARPS = 8;

__device__ __forceinline__ int lower_bound_i32(const int32_t* __restrict__ a, int n, int key) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (a[mid] < key) lo = mid + 1; else hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(WARPS * 32)
k_vote_greedy(int n_seeds, int n_mtc, const int32_t* __restrict__ hyp, const int32_t* __restrict__ c1,
              const int32_t* __restrict__ c2, const float* __restrict__ dists, float thr, uint8_t* __restrict__ flags,
              int32_t* __restrict__ acc_rows, int32_t* __restrict__ row_start, int32_t* __restrict__ n_inl,
              float* __restrict__ dsum) {
  const int h = blockIdx.x * WARPS + threadIdx.x / 32, lane = threadIdx.x % 32;
  if (h >= n_seeds) return;
  const int lo = lower_bound_i32(hyp, n_mtc, h), hi = lower_bound_i32(hyp, n_mtc, h + 1);
  for (int r = lo + lane; r < hi; r += 32) flags[r] = dists[r] <= thr ? 1 : 0;
  __syncwarp();
  int n = 0;
  float sum = 0.f;
  while (true) {
    float bd = 0.f;
    int br = 0x7fffffff;
    for (int r = lo + lane; r < hi; r += 32) {
      if (flags[r]) {
        const float d = dists[r];
        if (br == 0x7fffffff || d < bd) { bd = d; br = r; }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float od = __shfl_xor_sync(0xffffffffu, bd, o);
      const int orr = __shfl_xor_sync(0xffffffffu, br, o);
      if (orr != 0x7fffffff && (br == 0x7fffffff || od < bd || (od == bd && orr < br))) { bd = od; br = orr; }
    }
    if (br == 0x7fffffff) break;
    const int a1 = c1[br], a2 = c2[br];
    if (lane == 0) acc_rows[lo + n] = br;
    sum += bd;
    ++n;
    for (int r = lo + lane; r < hi; r += 32)
      if (c1[r] == a1 || c2[r] == a2) flags[r] = 0;
    __syncwarp();
  }
  if (lane == 0) { n_inl[h] = n; dsum[h] = sum; row_start[h] = lo; }
}

__global__ void __launch_bounds__(WARPS * 32)
k_vote_best(int n_pairs, const int32_t* __restrict__ pair_start, const int32_t* __restrict__ n_inl,
            const float* __restrict__ dsum, int n_min, int32_t* __restrict__ best_h, int32_t* __restrict__ cnt) {
  const int p = blockIdx.x * WARPS + threadIdx.x / 32, lane = threadIdx.x % 32;
  if (p >= n_pairs) return;
  int bh = -1, bn = 0;
  float bs = 3.402823466e+38f;
  for (int h = pair_start[p] + lane; h < pair_start[p + 1]; h += 32) {
    const int n = n_inl[h];
    const float s = dsum[h];
    if (n >= n_min && (n > bn || (n == bn && s < bs))) { bh = h; bn = n; bs = s; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const int oh = __shfl_xor_sync(0xffffffffu, bh, o), on = __shfl_xor_sync(0xffffffffu, bn, o);
    const float os = __shfl_xor_sync(0xffffffffu, bs, o);
    // the sequential scan keeps the FIRST hypothesis among full ties: lower id wins
    if (oh >= 0 && (bh < 0 || on > bn || (on == bn && (os < bs || (os == bs && oh < bh))))) { bh = oh; bn = on; bs = os; }
  }
  if (lane == 0) {
    const bool keep = bh > 0;                     // id 0 can never be selected (cosypose_cext.cpp:203)
    best_h[p] = keep ? bh : -1;
    cnt[p] = keep ? bn : 0;
  }
}

// single block: offsets of every kept pair, then the ordered outputs; counts[0] = inlier matches, counts[1] = pairs
__global__ void __launch_bounds__(1024)
k_vote_emit(int n_pairs, const int32_t* __restrict__ best_h, const int32_t* __restrict__ cnt,
            const int32_t* __restrict__ row_start, const int32_t* __restrict__ acc_rows, const int32_t* __restrict__ c1,
            const int32_t* __restrict__ c2, int32_t* __restrict__ out_c1, int32_t* __restrict__ out_c2,
            int32_t* __restrict__ best_out, int64_t* __restrict__ counts) {
  __shared__ int s_m[1024], s_b[1024];
  const int tid = threadIdx.x;
  const int per = (n_pairs + 1023) / 1024;
  const int p0 = min(n_pairs, tid * per), p1 = min(n_pairs, p0 + per);
  int m = 0, b = 0;
  for (int p = p0; p < p1; ++p) { m += cnt[p]; b += best_h[p] > 0 ? 1 : 0; }
  s_m[tid] = m;
  s_b[tid] = b;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {            // inclusive Hillis-Steele scan of both counters
    const int vm = tid >= o ? s_m[tid - o] : 0, vb = tid >= o ? s_b[tid - o] : 0;
    __syncthreads();
    s_m[tid] += vm;
    s_b[tid] += vb;
    __syncthreads();
  }
  int om = s_m[tid] - m, ob = s_b[tid] - b;
  for (int p = p0; p < p1; ++p) {
    const int h = best_h[p];
    if (h > 0) {
      best_out[ob++] = h;
      const int lo = row_start[h];
      for (int i = 0; i < cnt[p]; ++i) {
        const int r = acc_rows[lo + i];
        out_c1[om] = c1[r];
        out_c2[om] = c2[r];
        ++om;
      }
    }
  }
  if (tid == 1023) { counts[0] = s_m[1023]; counts[1] = s_b[1023]; }
}
}  // namespace vote

}  // namespace cosyb

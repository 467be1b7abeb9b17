// Host-side integer stages of multiview matching: seed enumeration, inlier voting, scatter-argmin,
// symmetry id expansion.  They must be bit-exact with the reference's pybind11 module
// (cosypose/csrc/cosypose_cext.cpp), so the same standard-library primitives are used where the
// result depends on them: std::shuffle + std::default_random_engine for the seed order
// (cosypose_cext.cpp:27-33), a stable ascending sort of the inlier distances (:19-25) and a float
// accumulator for the per-hypothesis distance sum (:180).  Data structures are flat arrays
// (sort + counting-sort buckets) instead of the reference's map / unordered_map of structs.
#include <algorithm>
#include <cfloat>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <numeric>
#include <random>
#include <vector>

#include "../../include/cosyb200.h"

namespace cosyb {
void set_error(const char* fmt, ...);
}

namespace {

struct TMatch {
  int32_t v1, v2, c1, c2;
};

std::vector<int32_t> shuffled_iota(int n, int seed) {
  std::vector<int32_t> v(n);
  std::iota(v.begin(), v.end(), 0);
  std::shuffle(v.begin(), v.end(), std::default_random_engine(seed));
  return v;
}

}  // namespace

extern "C" int cosyb200_ransac_infos(int n_cand, const int32_t* view_ids, const int32_t* label_ids,
                                     int n_ransac_iter, int seed, int64_t* n_seeds_out,
                                     int64_t* n_tmatches_out, int32_t* seeds, int32_t* tmatches) {
  if (n_cand < 0 || (n_cand > 0 && (!view_ids || !label_ids)) || !n_seeds_out || !n_tmatches_out) {
    cosyb::set_error("ransac_infos: bad arguments");
    return COSYB200_EINVAL;
  }
  // Tentative matches: ordered candidate pairs (n, m) in different views with equal labels,
  // grouped by ordered view pair (ascending (v1, v2)), (n, m) ascending inside a group.
  // Bucketing by label avoids the all-pairs scan; the sort restores the reference order.
  std::vector<int32_t> order(n_cand);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(),
                   [&](int32_t a, int32_t b) { return label_ids[a] < label_ids[b]; });
  std::vector<TMatch> tm;
  for (int i = 0; i < n_cand;) {
    int j = i;
    while (j < n_cand && label_ids[order[j]] == label_ids[order[i]]) ++j;
    for (int a = i; a < j; ++a)
      for (int b = i; b < j; ++b) {
        int32_t n = order[a], m = order[b];
        if (view_ids[n] != view_ids[m]) tm.push_back({view_ids[n], view_ids[m], n, m});
      }
    i = j;
  }
  std::sort(tm.begin(), tm.end(), [](const TMatch& x, const TMatch& y) {
    if (x.v1 != y.v1) return x.v1 < y.v1;
    if (x.v2 != y.v2) return x.v2 < y.v2;
    if (x.c1 != y.c1) return x.c1 < y.c1;
    return x.c2 < y.c2;
  });

  const bool fill = seeds != nullptr && tmatches != nullptr;
  // first pass: counts (needed for the [6][n] / [3][n] strides)
  int64_t n_seeds = 0, n_mtc = 0;
  for (size_t g0 = 0; g0 < tm.size();) {
    size_t g1 = g0;
    while (g1 < tm.size() && tm[g1].v1 == tm[g0].v1 && tm[g1].v2 == tm[g0].v2) ++g1;
    int64_t nt = (int64_t)(g1 - g0);
    int64_t n_pairs = std::min<int64_t>(std::max(n_ransac_iter, 0), nt * (nt - 1));
    n_seeds += n_pairs;
    n_mtc += n_pairs * nt;
    g0 = g1;
  }
  if (fill && (*n_seeds_out != n_seeds || *n_tmatches_out != n_mtc)) {
    cosyb::set_error("ransac_infos: caller sizes (%lld, %lld) do not match (%lld, %lld)",
                     (long long)*n_seeds_out, (long long)*n_tmatches_out, (long long)n_seeds,
                     (long long)n_mtc);
    return COSYB200_EINVAL;
  }
  *n_seeds_out = n_seeds;
  *n_tmatches_out = n_mtc;
  if (!fill) return COSYB200_OK;

  int64_t si = 0, mi = 0;
  for (size_t g0 = 0; g0 < tm.size();) {
    size_t g1 = g0;
    while (g1 < tm.size() && tm[g1].v1 == tm[g0].v1 && tm[g1].v2 == tm[g0].v2) ++g1;
    const int nt = (int)(g1 - g0);
    const TMatch* grp = tm.data() + g0;
    std::vector<int32_t> perm1 = shuffled_iota(nt, seed), perm2 = shuffled_iota(nt, seed + 1);
    int n_pairs = 0;
    for (int i1 = 0; i1 < nt && n_pairs < n_ransac_iter; ++i1) {
      for (int i2 = 0; i2 < nt && n_pairs < n_ransac_iter; ++i2) {
        const int m1 = perm1[i1], m2 = perm2[i2];
        if (m1 == m2) continue;
        seeds[0 * n_seeds + si] = grp[0].v1;
        seeds[1 * n_seeds + si] = grp[0].v2;
        seeds[2 * n_seeds + si] = grp[m1].c1;
        seeds[3 * n_seeds + si] = grp[m1].c2;
        seeds[4 * n_seeds + si] = grp[m2].c1;
        seeds[5 * n_seeds + si] = grp[m2].c2;
        for (int t = 0; t < nt; ++t, ++mi) {
          tmatches[0 * n_mtc + mi] = (int32_t)si;
          tmatches[1 * n_mtc + mi] = grp[t].c1;
          tmatches[2 * n_mtc + mi] = grp[t].c2;
        }
        ++si;
        ++n_pairs;
      }
    }
    g0 = g1;
  }
  return COSYB200_OK;
}

extern "C" int cosyb200_ransac_inliers(int64_t n_seeds, const int32_t* sv1, const int32_t* sv2,
                                       int64_t n_mtc, const int32_t* mtc_hyp, const int32_t* mtc_c1,
                                       const int32_t* mtc_c2, const float* dists, float thr,
                                       int n_min_inliers, int32_t* out_c1, int32_t* out_c2,
                                       int64_t* n_out, int32_t* best_out, int64_t* n_best) {
  if (n_seeds < 0 || n_mtc < 0 || !n_out || !n_best) {
    cosyb::set_error("ransac_inliers: bad arguments");
    return COSYB200_EINVAL;
  }
  // rows per hypothesis, in row order (stable counting sort on the hypothesis id)
  std::vector<int64_t> start(n_seeds + 1, 0);
  for (int64_t r = 0; r < n_mtc; ++r) {
    if (mtc_hyp[r] < 0 || mtc_hyp[r] >= n_seeds) {
      cosyb::set_error("ransac_inliers: hypothesis id %d out of range", mtc_hyp[r]);
      return COSYB200_EINVAL;
    }
    if (dists[r] <= thr) ++start[mtc_hyp[r] + 1];
  }
  for (int64_t h = 0; h < n_seeds; ++h) start[h + 1] += start[h];
  std::vector<int64_t> rows(start[n_seeds]);
  {
    std::vector<int64_t> fillp(start.begin(), start.end() - 1);
    for (int64_t r = 0; r < n_mtc; ++r)
      if (dists[r] <= thr) rows[fillp[mtc_hyp[r]]++] = r;
  }
  // greedy one-to-one matching per hypothesis in ascending distance (stable)
  std::vector<int32_t> n_inl(n_seeds, 0);
  std::vector<float> dsum(n_seeds, 0.f);
  std::vector<int64_t> ustart(n_seeds + 1, 0);
  std::vector<int64_t> urows;
  urows.reserve(rows.size());
  std::vector<int64_t> idx;
  std::vector<int32_t> used1, used2;
  for (int64_t h = 0; h < n_seeds; ++h) {
    idx.assign(rows.begin() + start[h], rows.begin() + start[h + 1]);
    std::stable_sort(idx.begin(), idx.end(), [&](int64_t a, int64_t b) { return dists[a] < dists[b]; });
    used1.clear();
    used2.clear();
    for (int64_t r : idx) {
      const int32_t c1 = mtc_c1[r], c2 = mtc_c2[r];
      if (std::find(used1.begin(), used1.end(), c1) != used1.end()) continue;
      if (std::find(used2.begin(), used2.end(), c2) != used2.end()) continue;
      used1.push_back(c1);
      used2.push_back(c2);
      urows.push_back(r);
      dsum[h] += dists[r];
      n_inl[h] += 1;
    }
    ustart[h + 1] = (int64_t)urows.size();
  }
  // best hypothesis per ordered view pair, pairs visited in ascending (v1, v2), hypotheses ascending
  std::vector<int64_t> horder(n_seeds);
  std::iota(horder.begin(), horder.end(), 0);
  std::stable_sort(horder.begin(), horder.end(), [&](int64_t a, int64_t b) {
    if (sv1[a] != sv1[b]) return sv1[a] < sv1[b];
    return sv2[a] < sv2[b];
  });
  int64_t no = 0, nb = 0;
  for (int64_t g0 = 0; g0 < n_seeds;) {
    int64_t g1 = g0;
    while (g1 < n_seeds && sv1[horder[g1]] == sv1[horder[g0]] && sv2[horder[g1]] == sv2[horder[g0]]) ++g1;
    int64_t best = -1;
    int32_t best_n = 0;
    float best_sum = FLT_MAX;
    for (int64_t i = g0; i < g1; ++i) {
      const int64_t h = horder[i];
      if (n_inl[h] >= n_min_inliers &&
          (n_inl[h] > best_n || (n_inl[h] == best_n && dsum[h] < best_sum))) {
        best = h;
        best_n = n_inl[h];
        best_sum = dsum[h];
      }
    }
    // The reference keeps a view pair only when its best hypothesis id is > 0
    // (cosypose_cext.cpp:203), i.e. hypothesis 0 can never be selected.  Preserved.
    if (best > 0) {
      if (best_out) best_out[nb] = (int32_t)best;
      ++nb;
      for (int64_t u = ustart[best]; u < ustart[best + 1]; ++u, ++no) {
        if (out_c1) out_c1[no] = mtc_c1[urows[u]];
        if (out_c2) out_c2[no] = mtc_c2[urows[u]];
      }
    }
    g0 = g1;
  }
  *n_out = no;
  *n_best = nb;
  return COSYB200_OK;
}

extern "C" int cosyb200_scatter_argmin(int64_t n, const float* values, const int32_t* group_ids,
                                       int64_t n_groups, int32_t* out) {
  if (n < 0 || n_groups < 0 || (n_groups > 0 && !out)) {
    cosyb::set_error("scatter_argmin: bad arguments");
    return COSYB200_EINVAL;
  }
  std::vector<char> seen(n_groups, 0);
  std::vector<float> lowest(n_groups, 0.f);
  for (int64_t g = 0; g < n_groups; ++g) out[g] = 0;
  for (int64_t i = 0; i < n; ++i) {
    const int32_t g = group_ids[i];
    if (g < 0 || g >= n_groups) {
      cosyb::set_error("scatter_argmin: group id %d out of range", g);
      return COSYB200_EINVAL;
    }
    if (!seen[g] || values[i] < lowest[g]) {
      seen[g] = 1;
      lowest[g] = values[i];
      out[g] = (int32_t)i;
    }
  }
  return COSYB200_OK;
}

extern "C" int cosyb200_expand_ids_for_symmetry(int64_t n, const int32_t* label_ids,
                                                const int32_t* n_sym_per_label, int64_t* n_out,
                                                int32_t* ids_expand, int32_t* sym_ids) {
  if (n < 0 || !n_out || (n > 0 && (!label_ids || !n_sym_per_label))) {
    cosyb::set_error("expand_ids_for_symmetry: bad arguments");
    return COSYB200_EINVAL;
  }
  int64_t o = 0;
  for (int64_t i = 0; i < n; ++i) {
    const int32_t ns = n_sym_per_label[label_ids[i]];
    for (int32_t k = 0; k < ns; ++k, ++o) {
      if (ids_expand) ids_expand[o] = (int32_t)i;
      if (sym_ids) sym_ids[o] = k;
    }
  }
  *n_out = o;
  return COSYB200_OK;
}

"""Pure-Python restatement of the reference's pybind11 module `cosypose_cext`
(cosypose/csrc/cosypose_cext.cpp).  TEST INFRASTRUCTURE (see oracle/__init__.py): small cases only.

Pinned against the reference's own extension compiled into oracle/_ref (tests/test_abi.py compares
the product with both).  The only third-party arithmetic is the seed order of make_ransac_infos,
which comes from libstdc++ (`std::shuffle` driven by `std::default_random_engine`,
cosypose_cext.cpp:27-33; the reference pins gxx_linux-64 7.3, this container has GCC 13.3 - the
algorithm below is the one both ship):
  * default_random_engine = minstd_rand0: x <- 16807 * x mod (2^31 - 1), seed 0 is replaced by 1;
  * uniform_int_distribution over a generator whose range is not a power of two: rejection
    sampling with `scaling = urng_range / n`, accept `r < n * scaling`, return `r / scaling`;
  * shuffle draws TWO swap positions from one variate (`x / b1`, `x % b1` with x uniform in
    [0, b0*b1)) while `urng_range / n >= n`, after one single draw when the length is even.
"""
import numpy as np

_M = 2147483647
_URNG_MIN, _URNG_MAX = 1, _M - 1


class _MinstdRand0:
    def __init__(self, seed):
        s = seed % _M
        self.x = 1 if s == 0 else s

    def __call__(self):
        self.x = (16807 * self.x) % _M
        return self.x


def _uniform_int(g, lo, hi):
    """libstdc++ uniform_int_distribution<size_t>{lo, hi}(g) for urng_range >= hi - lo."""
    urng_range = _URNG_MAX - _URNG_MIN
    urange = hi - lo
    assert urng_range > urange
    uerange = urange + 1
    scaling = urng_range // uerange
    past = uerange * scaling
    while True:
        r = g() - _URNG_MIN
        if r < past:
            return r // scaling + lo


def _shuffle(n, seed):
    """std::shuffle(iota(n), default_random_engine(seed))  (cosypose_cext.cpp:27-33)."""
    v = list(range(n))
    if n == 0:
        return v
    g = _MinstdRand0(seed)
    urng_range = _URNG_MAX - _URNG_MIN
    if urng_range // n >= n:
        i = 1
        if n % 2 == 0:
            j = _uniform_int(g, 0, 1)
            v[i], v[j] = v[j], v[i]
            i += 1
        while i < n:
            b0 = i + 1
            b1 = b0 + 1
            x = _uniform_int(g, 0, b0 * b1 - 1)
            p0, p1 = x // b1, x % b1
            v[i], v[p0] = v[p0], v[i]
            i += 1
            v[i], v[p1] = v[p1], v[i]
            i += 1
        return v
    for i in range(1, n):
        j = _uniform_int(g, 0, i)
        v[i], v[j] = v[j], v[i]
    return v


def make_ransac_infos(view_ids, labels, n_ransac_iter=100, seed=0):
    """cosypose_cext.cpp:36-105."""
    n = len(view_ids)
    groups = {}
    for a in range(n):
        for b in range(n):
            if view_ids[a] != view_ids[b] and labels[a] == labels[b]:
                groups.setdefault((int(view_ids[a]), int(view_ids[b])), []).append((a, b))
    seeds = {k: [] for k in ('view1', 'view2', 'match1_cand1', 'match1_cand2', 'match2_cand1', 'match2_cand2')}
    mtc = {k: [] for k in ('hypothesis_id', 'cand1', 'cand2')}
    n_seeds = 0
    for vp in sorted(groups):                       # std::map iteration order
        tm = groups[vp]
        perm1, perm2 = _shuffle(len(tm), seed), _shuffle(len(tm), seed + 1)
        n_pairs = 0
        for m1 in perm1:
            if n_pairs >= n_ransac_iter:
                break
            for m2 in perm2:
                if n_pairs >= n_ransac_iter:
                    break
                if m1 == m2:
                    continue
                seeds['view1'].append(vp[0])
                seeds['view2'].append(vp[1])
                seeds['match1_cand1'].append(tm[m1][0])
                seeds['match1_cand2'].append(tm[m1][1])
                seeds['match2_cand1'].append(tm[m2][0])
                seeds['match2_cand2'].append(tm[m2][1])
                for c1, c2 in tm:
                    mtc['hypothesis_id'].append(n_seeds)
                    mtc['cand1'].append(c1)
                    mtc['cand2'].append(c2)
                n_pairs += 1
                n_seeds += 1
    return ({k: np.asarray(v, dtype=np.int32) for k, v in seeds.items()},
            {k: np.asarray(v, dtype=np.int32) for k, v in mtc.items()})


def find_ransac_inliers(seeds_view1, seeds_view2, mtc_hypothesis_id, mtc_cand1, mtc_cand2, dists,
                        dist_threshold, n_min_inliers):
    """cosypose_cext.cpp:107-216 (fp32 distance sums; hypothesis 0 can never be selected, :203)."""
    dists = np.asarray(dists, dtype=np.float32)
    thr = np.float32(dist_threshold)
    n_hyp = len(seeds_view1)
    inl = [[] for _ in range(n_hyp)]
    for r in range(len(mtc_hypothesis_id)):
        if dists[r] <= thr:
            inl[int(mtc_hypothesis_id[r])].append(r)
    by_pair = {}
    for h in range(n_hyp):
        by_pair.setdefault((int(seeds_view1[h]), int(seeds_view2[h])), []).append(h)
    uniq, n_in, dsum = {}, np.zeros(n_hyp, np.int64), np.zeros(n_hyp, np.float32)
    for h in range(n_hyp):
        rows = sorted(inl[h], key=lambda r: dists[r])           # sorted() is stable
        used1, used2, keep = set(), set(), []
        for r in rows:
            c1, c2 = int(mtc_cand1[r]), int(mtc_cand2[r])
            if c1 not in used1 and c2 not in used2:
                used1.add(c1)
                used2.add(c2)
                keep.append((c1, c2))
                dsum[h] = np.float32(dsum[h] + dists[r])
                n_in[h] += 1
        uniq[h] = keep
    out1, out2, best_list = [], [], []
    for vp in sorted(by_pair):
        best, best_n, best_sum = -1, 0, np.finfo(np.float32).max
        for h in by_pair[vp]:
            if n_in[h] >= n_min_inliers and (n_in[h] > best_n or (n_in[h] == best_n and dsum[h] < best_sum)):
                best, best_n, best_sum = h, n_in[h], dsum[h]
        if best > 0:
            best_list.append(best)
            for c1, c2 in uniq[best]:
                out1.append(c1)
                out2.append(c2)
    return dict(inlier_matches_cand1=np.asarray(out1, dtype=np.int32),
                inlier_matches_cand2=np.asarray(out2, dtype=np.int32),
                best_hypotheses=np.asarray(best_list, dtype=np.int32))


def scatter_argmin(array, expand_ids):
    """cosypose_cext.cpp:218-245: index of the first minimum of each group."""
    best, low = {}, {}
    for n, (v, g) in enumerate(zip(np.asarray(array, dtype=np.float32), expand_ids)):
        g = int(g)
        if g not in best or v < low[g]:
            best[g], low[g] = n, v
    return np.asarray([best.get(g, 0) for g in range(len(best))], dtype=np.int32)


def expand_ids_for_symmetry(labels, n_symmetries):
    """cosypose_cext.cpp:247-259."""
    ids, sym = [], []
    for n, l in enumerate(labels):
        for k in range(n_symmetries[l]):
            ids.append(n)
            sym.append(k)
    return np.asarray(ids, dtype=np.int32), np.asarray(sym, dtype=np.int32)

"""CPU-side checks of the boundary: the C-ABI library loads and exports every symbol declared in
include/cosyb200.h, the compiled EfficientNet table equals the Python spec, argument errors map to
the reference's exception types, and the host-only integer stages are bit-exact against the
reference's own C++ extension (oracle/_ref, compiled from /root/reference by oracle/Makefile)
and against the pure-Python restatement (oracle/cext_oracle.py)."""
import ctypes
import re
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


def test_library_exports_every_declared_symbol():
    from cosypose_b200 import _lib
    header = (ROOT / 'include' / 'cosyb200.h').read_text()
    declared = sorted(set(re.findall(r'\b(cosyb200_[a-z0-9_]+)\s*\(', header)))
    assert len(declared) >= 20
    L = ctypes.CDLL(str(_lib.LIB_PATH))
    for name in declared:
        assert hasattr(L, name), f'{name} declared in include/cosyb200.h but not exported'
    assert sorted(_lib.EXPORTS) == declared
    assert _lib.lib().cosyb200_version() >= 100


def test_effnet_table_matches_spec():
    from cosypose_b200 import _lib, effnet_spec as spec
    L = _lib.lib()
    out = (ctypes.c_int32 * 11)()
    for b in spec.BLOCKS:
        assert L.cosyb200_effnet_block(b.idx, out) == 0
        assert list(out) == [b.k, b.s, b.e, b.cin, b.cexp, b.cse, b.cout, b.pad_lo, b.pad_hi, int(b.skip), 26]
    assert L.cosyb200_effnet_block(26, out) == _lib.EINVAL
    assert b'bad index' in L.cosyb200_last_error()
    # the shapes SURVEY.md section 8(d) lists
    shapes = dict((n, (h, w, c)) for n, h, w, c in spec.activation_shapes())
    assert shapes['stem'] == (120, 160, 40) and shapes['block18'] == (7, 10, 232) and shapes['head'] == (7, 10, 1536)
    # 574 state_dict entries minus one int64 num_batches_tracked per BatchNorm (78)
    assert len(spec.state_dict_layout()) == 574 - 78
    assert sum(int(np.prod(s)) for s in spec.state_dict_layout().values()) == 10798519 - 78


def test_error_mapping():
    from cosypose_b200 import _lib, engine as E
    with pytest.raises(AssertionError):
        E.ransac_infos(np.zeros(3, dtype=np.int32), np.zeros(4, dtype=np.int32), 5)
    with pytest.raises(AssertionError):
        E.scatter_argmin(np.zeros(3, np.float32), np.array([0, 5, 1], np.int32), n_groups=2)


def _ref_cext():
    sys.path.insert(0, str(ROOT / 'oracle' / '_ref'))
    try:
        import cosypose_cext
        return cosypose_cext
    except ImportError:
        return None


def _cases(n_trials=40):
    rs = np.random.RandomState(0)
    for t in range(n_trials):
        nv, n, nl = rs.randint(1, 6), rs.randint(0, 40), rs.randint(1, 6)
        yield t, rs, rs.randint(0, nv, n), rs.randint(0, nl, n), nl, int(rs.choice([0, 1, 5, 20, 200]))


@pytest.mark.parametrize('impl', ['reference_cext', 'python_oracle'])
def test_ransac_integer_stages_bit_exact(impl):
    from cosypose_b200 import engine as E
    from oracle import cext_oracle
    ref = _ref_cext() if impl == 'reference_cext' else cext_oracle
    if ref is None:
        pytest.skip('oracle/_ref not built (needs /root/reference)')
    keys = ['view1', 'view2', 'match1_cand1', 'match1_cand2', 'match2_cand1', 'match2_cand2']
    for t, rs, v, l, nl, it in _cases(40 if impl == 'reference_cext' else 12):
        labels = [f'obj_{i:06d}' for i in l]
        s, m = ref.make_ransac_infos(v.tolist(), labels, it, 0)
        seeds, tm = E.ransac_infos(v, l, it, 0)
        for i, k in enumerate(keys):
            assert np.array_equal(np.asarray(s[k]), seeds[i]), (t, k)
        for i, k in enumerate(['hypothesis_id', 'cand1', 'cand2']):
            assert np.array_equal(np.asarray(m[k]), tm[i]), (t, k)
        d = (rs.rand(tm.shape[1]) * 0.05).astype(np.float32)
        if t % 2:
            d = np.round(d, 2).astype(np.float32)   # ties exercise the stable sort / tie-breaks
        for thr, nmin in ((0.02, 3), (0.03, 1)):
            r = ref.find_ransac_inliers(np.asarray(s['view1'], np.int32), np.asarray(s['view2'], np.int32),
                                        np.asarray(m['hypothesis_id'], np.int32), np.asarray(m['cand1'], np.int32),
                                        np.asarray(m['cand2'], np.int32), d, thr, nmin)
            o = E.ransac_inliers(seeds[0], seeds[1], tm[0], tm[1], tm[2], d, thr, nmin)
            for k in ('inlier_matches_cand1', 'inlier_matches_cand2', 'best_hypotheses'):
                assert np.array_equal(np.asarray(r[k]), o[k]), (t, k)
        if len(v):
            g = np.unique(rs.randint(0, max(1, len(v) // 2), len(v)), return_inverse=True)[1].astype(np.int32)
            vals = np.round(rs.rand(len(v)), 1).astype(np.float32)
            assert np.array_equal(np.asarray(ref.scatter_argmin(vals, g)), E.scatter_argmin(vals, g))
        nsym = {f'obj_{i:06d}': int(rs.randint(1, 5)) for i in range(nl)}
        a, b = ref.expand_ids_for_symmetry(labels, nsym)
        a2, b2 = E.expand_ids_for_symmetry(l, [nsym[f'obj_{i:06d}'] for i in range(nl)])
        assert np.array_equal(np.asarray(a), a2) and np.array_equal(np.asarray(b), b2)


def test_ransac_infos_config4_counts():
    """BASELINE.json configs[3]: 8 views x 16 detections, unique labels per view, 2000 iterations:
    56 ordered view pairs x 240 seeds = 13 440 seeds, x 16 tentative matches = 215 040 rows
    (SURVEY.md section 8a, row a18)."""
    from cosypose_b200 import engine as E
    v = np.repeat(np.arange(8), 16)
    l = np.tile(np.arange(16), 8)
    seeds, tm = E.ransac_infos(v, l, 2000, 0)
    assert seeds.shape == (6, 13440) and tm.shape == (3, 215040)
    # hypothesis 0 can never win (reference quirk, cosypose_cext.cpp:203)
    d = np.full(tm.shape[1], 1.0, np.float32)
    d[tm[0] == 0] = 0.001
    o = E.ransac_inliers(seeds[0], seeds[1], tm[0], tm[1], tm[2], d, 0.02, 3)
    assert len(o['best_hypotheses']) == 0
    d[tm[0] == 1] = 0.0005     # strictly better than hypothesis 0 (a tie would keep 0 and drop the pair)
    o = E.ransac_inliers(seeds[0], seeds[1], tm[0], tm[1], tm[2], d, 0.02, 3)
    assert list(o['best_hypotheses']) == [1] and len(o['inlier_matches_cand1']) == 16

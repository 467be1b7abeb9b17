"""CPU tests of the host-side mirror of the reference interface: tensor collections, mesh tables,
pre-rendered view sequencing, sharding and the world_size-2 gather (gloo), product import rules."""
import os
import pickle
import subprocess
import sys
import textwrap
from pathlib import Path

import numpy as np
import pandas as pd
import pytest
import torch

ROOT = Path(__file__).resolve().parents[1]


def test_tensor_collection_semantics():
    from cosypose_b200.utils import tensor_collection as tc
    infos = pd.DataFrame(dict(label=['a', 'b', 'c'], view_id=[0, 0, 1]), index=[5, 6, 7])
    c = tc.PandasTensorCollection(infos, poses=torch.arange(48.).view(3, 4, 4), score=torch.tensor([1., 2, 3]))
    assert len(c) == 3 and list(c.infos.index) == [0, 1, 2]
    sub = c[[2, 0]]
    assert list(sub.infos['label']) == ['c', 'a'] and sub.poses.shape == (2, 4, 4) and sub.score.tolist() == [3., 1.]
    assert c[np.array([1])].poses[0, 0, 0] == 16
    c.poses = c.poses + 1                       # assignment to a registered name updates the tensor
    assert c.tensors['poses'][0, 0, 0] == 1
    c.extra = 'x'                               # other attributes are plain attributes
    assert 'extra' not in c.tensors
    with pytest.raises(AttributeError):
        c.missing
    m = c.merge_df(pd.DataFrame(dict(view_id=[0, 1], view_group=[7, 8])), on='view_id')
    assert list(m.infos['view_group']) == [7, 7, 8]
    cat = tc.concatenate([c, tc.PandasTensorCollection(pd.DataFrame()), sub])
    assert len(cat) == 5 and cat.poses.shape == (5, 4, 4)
    assert len(tc.concatenate([])) == 0
    c2 = pickle.loads(pickle.dumps(c))
    assert torch.equal(c2.poses, c.poses) and list(c2.infos['label']) == ['a', 'b', 'c']
    cl = c.clone()
    cl.poses[0, 0, 0] = -5
    assert c.poses[0, 0, 0] == 1
    assert c.float().poses.dtype == torch.float32 and c.device.type == 'cpu'
    assert 'poses' in repr(c)


def test_mesh_tables_padding_and_sampling():
    from cosypose_b200.engine import aabb_corners, sample_point_ids
    from cosypose_b200.lib3d.rigid_mesh_database import BatchedMeshes, pad_stack
    rs = np.random.RandomState(0)
    verts = [rs.rand(2100, 3), rs.rand(2500, 3), rs.rand(2001, 3)]
    syms = [np.eye(4)[None], np.tile(np.eye(4), (3, 1, 1))]
    db = BatchedMeshes.from_vertex_lists(['o1', 'o2', 'o3'], verts, syms + [np.eye(4)[None]])
    assert db.points.shape == (3, 2500, 3) and db.symmetries.shape == (3, 3, 4, 4)
    assert db.n_sym_mapping == {'o1': 1, 'o2': 3, 'o3': 1}
    # padding re-draws own points (reference rule), identity pads symmetries
    assert np.isin(db.points[0, 2100:].numpy().round(6), np.float32(verts[0]).round(6)).all()
    assert torch.equal(db.symmetries[0, 1], torch.eye(4))
    sel = db.select(np.array(['o3', 'o1', 'o3']))
    assert sel.points.shape == (3, 2500, 3)
    pts = sel.sample_points(2000, deterministic=True)
    ids = sample_point_ids(2500)
    assert torch.equal(pts, sel.points[:, torch.as_tensor(ids)])
    assert len(np.unique(ids)) == 2000
    with pytest.raises(KeyError):
        db.select(['nope'])
    box = aabb_corners(np.array([[[0., 0, 0], [1, 2, 3]]]))[0]
    assert box.tolist()[0] == [0, 2, 3] and box.tolist()[6] == [1, 0, 0] and box.tolist()[7] == [0, 0, 0]
    assert db.batched(aabb=True).points.shape == (3, 8, 3)
    assert pad_stack([np.zeros((1, 2)), np.ones((3, 2))], fill=np.array([7., 7.]))[0, 2, 0] == 7


def test_prerendered_view_sequencing():
    from cosypose_b200.rendering import PerCallRenderer, PreRenderedViews
    c = torch.arange(1 * 5, dtype=torch.float32).view(1, 5, 1, 1, 1).expand(1, 5, 3, 240, 320)
    r = (100 + torch.arange(2 * 5, dtype=torch.float32)).view(2, 5, 1, 1, 1).expand(2, 5, 3, 240, 320)
    v = PreRenderedViews([c, r], bsz_objects=4)
    # coarse: chunks of 4 then 1, one iteration each; refiner: 2 iterations per chunk
    assert v.prerendered(1, 4)[0, :, 0, 0, 0].tolist() == [0, 1, 2, 3]
    assert v.prerendered(1, 1)[0, :, 0, 0, 0].tolist() == [4]
    assert v.prerendered(2, 4)[1, :, 0, 0, 0].tolist() == [105, 106, 107, 108]
    assert v.render([0], None, None)[:, 0, 0, 0].tolist() == [104]
    assert v.render([0], None, None)[:, 0, 0, 0].tolist() == [109]
    assert v.render([0] * 4, None, None)[:, 0, 0, 0].tolist() == [0, 1, 2, 3]      # wrapped around
    assert not hasattr(PerCallRenderer(v), 'prerendered')
    with pytest.raises(AssertionError):
        v.prerendered(3, 4)


def test_shard_bounds():
    from cosypose_b200.sharding import shard_bounds
    for n, ws in ((512, 8), (10, 4), (3, 8), (0, 2)):
        spans = [shard_bounds(n, r, ws) for r in range(ws)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    assert shard_bounds(512, 3, 8) == (192, 256)


_GLOO_WORKER = textwrap.dedent('''
    import sys, numpy as np, pandas as pd, torch, torch.distributed as dist
    sys.path.insert(0, sys.argv[1])
    from cosypose_b200.sharding import gather_poses, shard_bounds, world
    from cosypose_b200.utils import tensor_collection as tc
    dist.init_process_group('gloo')
    rank, ws = world()
    n = 7
    full = torch.arange(n * 16, dtype=torch.float32).view(n, 4, 4)
    a, b = shard_bounds(n, rank, ws)
    out = gather_poses(full[a:b].clone(), n_total=n)
    assert torch.equal(out, full), (rank, out)
    # equal shards without n_total
    out2 = gather_poses(torch.full((3, 4, 4), float(rank)))
    assert out2.shape == (3 * ws, 4, 4) and out2[3 * rank, 0, 0] == rank
    c = tc.PandasTensorCollection(pd.DataFrame(dict(i=np.arange(a, b))), poses=full[a:b].clone())
    g = c.gather_distributed()
    assert list(g.infos['i']) == list(range(n)) and torch.equal(g.poses, full)
    dist.barrier()
    dist.destroy_process_group()
    sys.stdout.write('rank%dok ' % rank)
    sys.stdout.flush()
''')


def test_gather_world_size_2_gloo(tmp_path):
    """The N>1 exchange step on CPU: contiguous shards, one all-gather, every rank gets all poses."""
    script = tmp_path / 'worker.py'
    script.write_text(_GLOO_WORKER)
    env = dict(os.environ, MASTER_ADDR='127.0.0.1', OMP_NUM_THREADS='1')
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=2',
                        '--master-addr', '127.0.0.1', '--master-port', '29731', str(script), str(ROOT)],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert 'rank0ok' in r.stdout and 'rank1ok' in r.stdout


_SHARD_WORKER = textwrap.dedent('''
    import sys, types, numpy as np, pandas as pd, torch, torch.distributed as dist
    sys.path.insert(0, sys.argv[1])
    from cosypose_b200.integrated.pose_predictor import CoarseRefinePosePredictor
    from cosypose_b200.utils import tensor_collection as tc

    class FakeModel:
        """forward_indexed of models/pose.py with torch-only arithmetic: enough to exercise the predictor's shard /
        gather logic on CPU (every output depends on the hypothesis' own inputs only, like the real path)."""
        def __init__(self, scale):
            self.engine = types.SimpleNamespace(device=torch.device('cpu'))
            self.scale = scale
            self.cfg = types.SimpleNamespace(init_method='v0')
        def forward_indexed(self, images, im_ids, K, labels, TCO, n_iterations):
            out, T = {}, TCO
            for it in range(1, n_iterations + 1):
                Tn = T * self.scale + images[im_ids].mean(dim=(1, 2, 3))[:, None, None] + it
                out[f'iteration={it}'] = dict(TCO_output=Tn, TCO_input=T, K_crop=K * it, boxes_rend=Tn[:, 0, :4] + 1,
                                              boxes_crop=Tn[:, 1, :4] + 2)
                T = Tn
            return out

    dist.init_process_group('gloo')
    rank, ws = dist.get_rank(), dist.get_world_size()
    n = 11
    gen = torch.Generator().manual_seed(0)
    images = torch.rand((3, 3, 8, 8), generator=gen)
    K = torch.rand((3, 3, 3), generator=gen)
    infos = pd.DataFrame(dict(label=['obj_%06d' % (i % 4) for i in range(n)], batch_im_id=np.arange(n) % 3,
                              score=np.ones(n)))
    init = tc.PandasTensorCollection(infos=infos, poses=torch.rand((n, 4, 4), generator=gen))
    pred = CoarseRefinePosePredictor(FakeModel(0.5), FakeModel(0.25), bsz_objects=4)
    ref_final, ref_preds = pred.get_predictions(images, K, data_TCO_init=init, n_coarse_iterations=0,
                                                n_refiner_iterations=3)
    final, preds = pred.get_predictions(images, K, data_TCO_init=init, n_coarse_iterations=0,
                                        n_refiner_iterations=3, shard=True)
    assert list(preds.keys()) == list(ref_preds.keys())
    assert len(final) == n and torch.equal(final.poses, ref_final.poses)
    for k in ref_preds:
        if k == 'external_coarse':
            continue
        for f in ('poses', 'poses_input', 'K_crop', 'boxes_rend', 'boxes_crop'):
            assert torch.equal(getattr(preds[k], f), getattr(ref_preds[k], f)), (k, f)
        assert list(preds[k].infos['label']) == list(infos['label'])
    dist.barrier()
    dist.destroy_process_group()
    sys.stdout.write('rank%dok ' % rank)
    sys.stdout.flush()
''')


@pytest.mark.parametrize('nproc', [2, 3])
def test_sharded_predictor_gloo(tmp_path, nproc):
    """CoarseRefinePosePredictor.get_predictions(shard=True) with world_size 2 and 3 (uneven shards of 11
    hypotheses): every rank returns exactly the unsharded result, all iterations, after ONE all-gather."""
    script = tmp_path / 'worker.py'
    script.write_text(_SHARD_WORKER)
    env = dict(os.environ, MASTER_ADDR='127.0.0.1', OMP_NUM_THREADS='1')
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={nproc}',
                        '--master-addr', '127.0.0.1', '--master-port', str(29741 + nproc), str(script), str(ROOT)],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert all(f'rank{i}ok' in r.stdout for i in range(nproc))


def test_bop_csv_round_trip(tmp_path):
    from cosypose_b200.evaluation import bop_io
    from cosypose_b200.utils import tensor_collection as tc
    gen = torch.Generator().manual_seed(1)
    poses = torch.eye(4).repeat(3, 1, 1)
    poses[:, :3, 3] = torch.rand((3, 3), generator=gen)
    infos = pd.DataFrame(dict(scene_id=[1, 1, 2], view_id=[0, 3, 7], label=['obj_000004', 'obj_000011', 'obj_000004'],
                              score=[0.9, 0.5, 1.0]))
    p = tmp_path / 'r.csv'
    bop_io.tc_to_csv(tc.PandasTensorCollection(infos=infos, poses=poses), p)
    txt = p.read_text().splitlines()
    assert txt[0] == 'scene_id,im_id,obj_id,score,R,t,time' and len(txt) == 4
    assert txt[1].startswith('1,0,4,0.9,1.0 0.0 0.0 0.0 1.0 0.0 0.0 0.0 1.0,') and txt[1].endswith(',-1.0')
    back = bop_io.read_csv_candidates(p)
    assert list(back.infos['label']) == list(infos['label']) and list(back.infos['view_id']) == [0, 3, 7]
    assert (back.poses - poses).abs().max() < 1e-6


def test_product_never_imports_the_oracle():
    """The product package must not import, call or link anything under oracle/ (checker only)."""
    for p in (ROOT / 'cosypose_b200').rglob('*.py'):
        text = p.read_text()
        assert 'import oracle' not in text and 'from oracle' not in text, p
    for p in (ROOT / 'cosypose_b200' / 'csrc').iterdir():
        if p.suffix in ('.cu', '.cuh', '.h', '.cpp'):
            assert 'oracle' not in p.read_text(), p


def test_engine_fails_loudly_without_gpu():
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from cosypose_b200 import _lib
    from cosypose_b200.engine import Engine
    with pytest.raises(_lib.EngineError):
        Engine(0)


def test_bench_roofline_object():
    """bench.py's `roofline` object from a recorded engine profile (no GPU): the headline is the trunk against SURVEY
    8(d)'s algorithmic bytes (24.43 MB per forward and hypothesis) over the trunk's device time; per-kernel figures
    use each category's own one-kernel-per-stage bytes."""
    import importlib.util
    import json
    from pathlib import Path
    root = Path(__file__).resolve().parents[1]
    spec = importlib.util.spec_from_file_location('bench_mod', root / 'bench.py')
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    kb = bench.kernel_algorithmic_bytes()
    assert abs(sum(kb.values()) / 1e6 - 164.5) < 0.1          # one-kernel-per-stage bytes per forward
    assert bench.ALGO_BYTES_PER_FORWARD == 6107136 * 4
    # (launches, ms) over 5 profiled steps, the shape engine.profile_read() returns
    prof = {'geometry': (25, 0.5), 'roi_crop': (25, 3.15), 'stem': (25, 8.3), 'expand_1x1': (625, 38.6),
            'depthwise': (650, 43.4), 'squeeze_excite': (650, 7.1), 'project_1x1': (650, 50.9),
            'head_1x1': (25, 1.45), 'pool_fc_update': (25, 0.83), 'ransac': (0, 0.0)}
    r = bench.roofline_object(prof, 5, 320, 6537.0, 'measured (MEASURED_PEAKS.json)', step_ms=32.0)
    json.dumps(r)
    assert r['bound'] == 'hbm' and r['unit'] == 'GB/s' and 'trunk' in r['kernel']
    trunk_ms = sum(prof[c][1] for c in prof if c not in ('geometry', 'roi_crop', 'ransac')) / 5
    want = 320 * 6107136 * 4 / (trunk_ms * 1e-3) / 1e9
    assert abs(r['achieved'] - want) < 1e-6 * want and abs(r['frac'] - want / 6537.0) < 1e-9
    assert abs(r['frac_of_step'] - 320 * 6107136 * 4 / 32e-3 / 1e9 / 6537.0) < 1e-9
    assert r['algorithmic_bytes_per_launch'] == 64 * 6107136 * 4
    assert abs(r['launch_us'] - trunk_ms * 1e3 / 5) < 1e-6      # 5 forwards of 64 hypotheses per step
    assert set(r['by_kernel_gbs']) == {'stem', 'expand_1x1', 'depthwise', 'project_1x1', 'head_1x1'}
    for cfg in (1, 2, 3, 4):
        assert bench.CONFIGS[cfg]['name'] == f'configs[{cfg}]' and bench.metric_name(cfg)


def test_launch_plans_fit_the_sm():
    """Resource budgets of every launch configuration at the benchmark batch and at batch 1 (host code of the
    engine, no GPU): shared memory per SM (227 KB), TMEM columns, tile coverage, thread counts."""
    import ctypes
    from cosypose_b200 import _lib, effnet_spec as spec
    L = _lib.lib()
    out = (ctypes.c_int32 * 32)()
    SM_SMEM = 227 * 1024
    shapes = spec.activation_shapes()
    for batch in (1, 64, 4096):
        for b, (_, hi, wi, _), (_, ho, wo, _) in zip(spec.BLOCKS, shapes[1:-2], shapes[2:-1]):
            assert L.cosyb200_launch_plan(b.idx, batch, out) == 0
            v = list(out)
            tiled, R, n_strips, n_xt, wo_u, xu, n_chunks, smem = v[0:8]
            if tiled:
                assert R * n_strips >= ho and R * (n_strips - 1) < ho          # strips cover the rows exactly once
                assert wo_u * xu * n_xt == wo                                   # x tiles cover the columns
                assert n_chunks * 32 >= b.cexp > (n_chunks - 1) * 32
                halo = ((R - 1) * b.s + b.k) * ((wo_u * xu - 1) * b.s + b.k) * 32 * 4
                assert smem == halo and smem + 2048 <= SM_SMEM
                assert b.cse <= 128                                             # k_se_fc2's shared arrays
                assert ho <= 30                                                 # the larger blocks keep the rolling kernel
            else:
                assert ho >= 60 or (b.k == 5 and b.s == 2 and wo == 40)
            tiles, chunks, threads, th = v[8:12]
            assert 32 <= threads <= 256 and tiles >= 1 and th >= 1
            for off, N, K, M in ((12, b.cexp, b.cin, batch * hi * wi), (20, b.cout, b.cexp, batch * ho * wo)):
                if off == 12 and b.e == 1:
                    continue
                bn, n_tiles, nk, nb, resident, gsmem, ng, gtiles = v[off:off + 8]
                assert bn % 16 == 0 and 16 <= bn <= 64 and bn * n_tiles >= N    # UMMA N granularity at M = 128
                assert nk * 32 >= K > (nk - 1) * 32
                assert 2 <= nb <= 8 or (resident and nb == nk)
                assert bool(resident) == (nk <= nb)
                per_sm = 2 if ng == 1 else 1
                assert per_sm * (gsmem + 1024) <= SM_SMEM                       # CTAs per SM x (dynamic + static)
                assert 2 * 64 + 2 * 64 <= 256                                   # TMEM columns per CTA: 2 acc + 2 A slots
                assert gtiles == -(-M // 128) * n_tiles
                assert ng == (2 if gtiles <= 148 else 1)
    assert L.cosyb200_launch_plan(26, 64, out) == _lib.EINVAL


def test_gemm_work_split_invariants():
    """k_pw2's work split (host code, no GPU): every CTA is resident (grid <= SMs), shared memory fits, and with
    split_k the unit ranges tile the (m-tile, k-stage) space exactly with every m-tile shared by at most two
    CONSECUTIVE CTAs of its column, the first of which holds k-stage 0 (the owner) and never precedes its partner's
    hand-over in the partner's own order of work: the conditions the fixed-order partial-sum exchange relies on."""
    import ctypes
    from cosypose_b200 import _lib, effnet_spec as spec
    L = _lib.lib()
    out = (ctypes.c_int32 * 9)()
    shapes = spec.activation_shapes()
    n_sms = 148
    seen_split = set()
    for batch in (1, 4, 63, 64, 65, 256):
        layers = []
        for b, (_, hi, wi, _), (_, ho, wo, _) in zip(spec.BLOCKS, shapes[1:-2], shapes[2:-1]):
            if b.e != 1:
                layers.append((batch * hi * wi, b.cexp, b.cin))
            layers.append((batch * ho * wo, b.cout, b.cexp))
        layers.append((batch * 70, 1536, 384))
        for M, N, K in layers:
            assert L.cosyb200_pw2_plan(M, N, K, n_sms, out) == 0
            bn, small, n_tiles, nk, nb, resident, smem, grid, split_k = list(out)
            m_tiles = -(-M // 128)
            assert bn % 16 == 0 and 16 <= bn <= 192 and bn * n_tiles >= N and (small == 1) == (bn <= 96)
            assert nk * 32 >= K > (nk - 1) * 32
            assert 1 <= grid <= n_sms and grid % n_tiles == 0          # one CTA per SM, all co-resident
            assert smem + 2048 <= 227 * 1024
            P = grid // n_tiles
            if not split_k:
                assert P <= m_tiles                                     # whole m-tiles: part, part + P, ...
                continue
            seen_split.add((batch, M, N, K))
            assert nk >= 16
            U = m_tiles * nk
            bounds = [c * U // P for c in range(P + 1)]
            assert bounds[0] == 0 and bounds[-1] == U
            owners = {}
            for c in range(P):
                u0, u1 = bounds[c], bounds[c + 1]
                assert u1 > u0
                tiles = sorted({u // nk for u in range(u0, u1)})
                for t in tiles:
                    s_lo = max(u0, t * nk) - t * nk
                    s_hi = min(u1, (t + 1) * nk) - 1 - t * nk
                    owners.setdefault(t, []).append((c, s_lo, s_hi))
                    # a segment is a head (starts at stage 0) or a tail (ends at the last stage), never a middle piece
                    assert s_lo == 0 or s_hi == nk - 1
                    # a CTA's tail segment can only be the FIRST tile of its range: it is produced before anything else
                    if s_lo > 0:
                        assert t == tiles[0]
            assert sorted(owners) == list(range(m_tiles))
            for t, segs in owners.items():
                assert len(segs) <= 2
                if len(segs) == 2:
                    (c0, a0, b0), (c1, a1, b1) = segs
                    assert c1 == c0 + 1 and a0 == 0 and b0 + 1 == a1 and b1 == nk - 1
                else:
                    assert segs[0][1:] == (0, nk - 1)
    # the layers the split exists for, at the benchmark batch
    assert (64, 19200, 136, 816) in seen_split and (64, 4480, 232, 1392) in seen_split
    assert L.cosyb200_pw2_plan(19200, 136, 816, n_sms, out) == 0 and out[7] == 148
    assert L.cosyb200_pw2_plan(4480, 232, 1392, n_sms, out) == 0 and out[7] == 140
    assert L.cosyb200_pw2_plan(19200, 816, 136, n_sms, out) == 0 and out[8] == 0       # short K: whole tiles


def test_multiview_host_grouping_matches_pandas():
    """The numpy grouping of the multiview host path gives what the reference's pandas formulation gives
    (multiview/ransac.py:119-125: groupby('obj_id') size / sum / first)."""
    from cosypose_b200.multiview.ransac import make_obj_infos
    from cosypose_b200.utils import tensor_collection as tc
    rs = np.random.RandomState(0)
    n = 57
    infos = pd.DataFrame(dict(obj_id=rs.randint(0, 9, n), score=rs.rand(n), label=[f'obj_{i:06d}' for i in rs.randint(1, 5, n)],
                              view_id=rs.randint(0, 4, n)))
    got = make_obj_infos(tc.PandasTensorCollection(infos=infos, poses=torch.zeros(n, 4, 4)))
    ref = infos.loc[:, ['obj_id', 'score', 'label']].copy()
    gb = ref.groupby('obj_id')
    ref['n_cand'] = gb['score'].transform('size').astype(int)
    ref['score'] = gb['score'].transform('sum')
    ref = ref.groupby('obj_id').first().reset_index(drop=False)
    assert list(got.columns) == list(ref.columns) == ['obj_id', 'score', 'label', 'n_cand']
    assert got['obj_id'].tolist() == ref['obj_id'].tolist() and got['n_cand'].tolist() == ref['n_cand'].tolist()
    assert got['label'].tolist() == ref['label'].tolist()
    assert np.allclose(got['score'].to_numpy(), ref['score'].to_numpy(), rtol=1e-12)


def test_concatenate_single_collection_is_a_fresh_collection():
    from cosypose_b200.utils import tensor_collection as tc
    infos = pd.DataFrame(dict(a=[3, 1, 2]), index=[7, 8, 9])
    c = tc.PandasTensorCollection(infos=infos, poses=torch.arange(3.))
    out = tc.concatenate([c, tc.PandasTensorCollection(infos=pd.DataFrame(), poses=torch.zeros(0))])
    assert list(out.infos.index) == [0, 1, 2] and out.infos['a'].tolist() == [3, 1, 2]
    out.infos['b'] = 1                      # does not leak into the input
    assert 'b' not in c.infos.columns
    out.poses = out.poses + 1
    assert torch.equal(c.poses, torch.arange(3.))


def test_lazy_history_builds_scene_collections_on_first_access():
    from cosypose_b200.multiview.bundle_adjustment import LazyHistory

    class Problem:
        calls = 0

        def _scene_infos_many(self, states):
            Problem.calls += 1
            return [(f'objects{i}', f'cameras{i}') for i, _ in enumerate(states)]
    h = LazyHistory(dict(TWO_9d=[1, 2, 3], TCW_9d=[4, 5, 6], loss=[0.3, 0.2, 0.1]), Problem())
    assert h['loss'] == [0.3, 0.2, 0.1] and Problem.calls == 0           # plain keys do not trigger the conversion
    assert 'objects' in h and 'cameras' in h and Problem.calls == 0
    assert h['objects'] == ['objects0', 'objects1', 'objects2'] and Problem.calls == 1
    assert h['cameras'] == ['cameras0', 'cameras1', 'cameras2'] and Problem.calls == 1
    assert h.get('objects') == h['objects'] and h.get('missing', 5) == 5
    import pickle
    d = pickle.loads(pickle.dumps(h))
    assert type(d) is dict and d['objects'] == h['objects'] and set(d) >= {'TWO_9d', 'TCW_9d', 'loss', 'objects', 'cameras'}

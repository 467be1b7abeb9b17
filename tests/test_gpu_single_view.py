"""GPU parity of the single-view refinement path: CUDA engine (through the C-ABI) vs the CPU oracle
and vs the golden vectors produced by the unmodified reference (tests/golden/make_golden.py).

Tolerances (north star: pose matrices within 1e-4 of the reference; everything is fp32):
  geometry (boxes in pixels ~1e2..1e3)   1e-3 absolute  (~1e-6 relative)
  RoI crop (values in [0,1])             2e-6
  block-boundary activations             1e-4 * max|ref| per tap
  pose9 / TCO                            1e-4 (measured ~1e-6)
"""
import numpy as np
import pytest
import torch

from helpers import Workload, build_predictor, state_dict

pytestmark = pytest.mark.gpu

TOL_POSE = 1e-4


@pytest.fixture(scope='module')
def dev():
    return torch.device('cuda', 0)


@pytest.fixture(scope='module')
def small():
    return Workload(2, 3, 5, 1, 2)


@pytest.fixture(scope='module')
def engine(small, dev):
    from cosypose_b200.engine import Engine
    from cosypose_b200.lib3d.rigid_mesh_database import BatchedMeshes
    eng = Engine(0, max_batch=8)
    BatchedMeshes.from_tables(small.labels, small.points, small.sym, small.n_sym).install(eng)
    eng.load_pose_model(0, state_dict(0))
    eng.load_pose_model(1, state_dict(1))
    return eng


def _d(t, dev):
    return t.to(dev).contiguous()


def test_library_is_loaded(engine):
    """The product path runs in libcosyb200.so, not in a fallback."""
    import cosypose_b200._lib as L
    with open('/proc/self/maps') as f:
        assert 'libcosyb200.so' in f.read()
    assert L.lib().cosyb200_version() >= 100


def test_tco_init(engine, small, dev):
    from oracle import pose_oracle as po
    K_h = small.K[torch.as_tensor(small.im_ids)]
    lab = torch.as_tensor(small.label_ids).int()
    ref = po.TCO_init_from_boxes(small.boxes, K_h)
    out = engine.tco_init(_d(small.boxes, dev), _d(K_h, dev), _d(lab, dev))
    assert (out.cpu() - ref).abs().max() < 1e-6
    pts = po.select_points(small.points, small.label_ids)
    ref = po.TCO_init_from_boxes_zup_autodepth(small.boxes, pts, K_h)
    out = engine.tco_init(_d(small.boxes, dev), _d(K_h, dev), _d(lab, dev), zup=True)
    assert (out.cpu() - ref).abs().max() < 1e-5


def test_prepare_iter(engine, small, dev):
    from oracle import pose_oracle as po
    K_h = small.K[torch.as_tensor(small.im_ids)]
    lab = torch.as_tensor(small.label_ids).int()
    TCO = po.TCO_init_from_boxes(small.boxes, K_h)
    # perturb rotations so the projection is not axis aligned
    gen = torch.Generator().manual_seed(2)
    pose9 = torch.tensor([1., 0, 0, 0, 1, 0, 0, 0, 1]) + 0.3 * torch.randn((small.n, 9), generator=gen)
    pose9[:, 6:8] *= 0.1
    pose9[:, 8] = 1 + 0.1 * pose9[:, 8]
    TCO = po.update_pose(TCO, small.K[:1].repeat(small.n, 1, 1), pose9)
    pts = po.select_points(small.points, small.label_ids)
    _, Kc_o, br_o, bc_o = po.crop_inputs(small.images, small.im_ids, K_h, TCO, pts)
    br, bc, Kc = engine.prepare_iter(_d(K_h, dev), _d(TCO, dev), _d(lab, dev), (480, 640))
    assert (br.cpu() - br_o).abs().max() < 1e-3
    assert (bc.cpu() - bc_o).abs().max() < 1e-3
    rel = ((Kc.cpu() - Kc_o).abs() / Kc_o.abs().clamp_min(1.0)).max()
    assert rel < 1e-5


@pytest.mark.parametrize('case', ['inside', 'straddle', 'outside', 'tiny'])
def test_roi_crop(engine, small, dev, case):
    """Includes boxes that leave the frame: out-of-image samples are dropped, not edge-clamped
    (the compiled torchvision operator is the ground truth)."""
    import torchvision
    boxes = {
        'inside': [[100., 80., 420., 320.], [10., 10., 330., 250.]],
        'straddle': [[-40.5, -30.25, 200., 150.], [500., 300., 700.5, 520.]],
        'outside': [[-900., -700., -100., -100.], [700., 500., 1500., 1100.]],
        'tiny': [[300.2, 200.7, 300.6, 201.0], [0., 0., 640., 480.]],
    }[case]
    boxes = torch.tensor(boxes)
    im_ids = torch.tensor([0, 1], dtype=torch.int32)
    out = engine.roi_crop(_d(small.images, dev), _d(im_ids, dev), _d(boxes, dev)).cpu()
    rois = torch.cat((im_ids.float()[:, None], boxes), dim=1)
    ref = torchvision.ops.roi_align(small.images, rois, output_size=(240, 320), spatial_scale=1.0, sampling_ratio=4)
    assert (out - ref).abs().max() < 2e-6


def test_backbone_taps(engine, small, dev, golden_dir):
    """Every block-boundary activation against the oracle and the reference's golden taps."""
    from cosypose_b200 import effnet_spec as spec, synthetic as syn
    from oracle import pose_oracle as po
    x = syn.make_net_input(2, seed=21)
    taps_o = {}
    pose_o = po.net_forward(x, state_dict(0), taps_o)
    pose_e, taps_e = engine.net_forward(0, _d(x[:, :3], dev), _d(x[:, 3:], dev), taps=True)
    for name, *_ in spec.activation_shapes()[1:]:
        a = taps_e[name].cpu().permute(0, 3, 1, 2)
        b = taps_o[name]
        assert (a - b).abs().max() <= 1e-4 * b.abs().max(), name
    assert (pose_e.cpu() - pose_o).abs().max() < 1e-5
    g = np.load(golden_dir / 'backbone_b2.npz')
    assert np.abs(pose_e.cpu().numpy() - g['pose']).max() < 1e-5
    for i in (0, 1, 2, 5, 8, 13, 18, 25):
        a = taps_e[f'block{i}'][0].cpu().numpy()[::4, ::4]
        assert np.abs(a - g[f'block{i}']).max() <= 1e-4 * np.abs(g[f'block{i}']).max(), i


def test_update_pose_properties(engine, dev):
    from oracle import pose_oracle as po
    gen = torch.Generator().manual_seed(3)
    n = 257
    pose9 = torch.randn((n, 9), generator=gen)
    pose9[:, 8] = pose9[:, 8].abs() + 0.2
    TCO = torch.eye(4).repeat(n, 1, 1)
    TCO[:, :3, 3] = torch.randn((n, 3), generator=gen) * 0.2 + torch.tensor([0., 0., 1.])
    K = torch.tensor([[900., 0, 160], [0, 880., 120], [0, 0, 1]]).repeat(n, 1, 1)
    out = engine.update_pose(_d(TCO, dev), _d(K, dev), _d(pose9, dev)).cpu()
    ref = po.update_pose(TCO, K, pose9)
    assert (out - ref).abs().max() < 1e-5
    R = out[:, :3, :3]
    assert (R @ R.transpose(1, 2) - torch.eye(3)).abs().max() < 1e-5      # orthonormal
    assert (torch.linalg.det(R) - 1).abs().max() < 1e-5                    # proper rotation


@pytest.mark.parametrize('per_call', [False, True])
def test_get_predictions_small(small, dev, per_call, golden_dir):
    """CoarseRefinePosePredictor.get_predictions, 2 frames x 3 detections, chunks of 4, 1 coarse +
    2 refine: fused pre-rendered loop and the two-phase per-iteration loop give the reference's poses."""
    from cosypose_b200.utils import tensor_collection as tc
    pred, eng, views = build_predictor(small, 0, bsz_objects=4, per_call_renderer=per_call)
    det = tc.PandasTensorCollection(infos=small.infos(), bboxes=_d(small.boxes, dev))
    final, preds = pred.get_predictions(_d(small.images, dev), _d(small.K, dev), detections=det,
                                        n_coarse_iterations=1, n_refiner_iterations=2)
    g = np.load(golden_dir / 'single_view_small.npz')
    assert set(preds) == {'coarse/iteration=1', 'refiner/iteration=1', 'refiner/iteration=2'}
    assert np.abs(final.poses.cpu().numpy() - g['final_poses']).max() < TOL_POSE
    for k, v in preds.items():
        assert np.abs(v.poses.cpu().numpy() - g[f'{k}/poses']).max() < TOL_POSE, k
        assert np.abs(v.poses_input.cpu().numpy() - g[f'{k}/poses_input']).max() < TOL_POSE, k
        assert np.abs(v.boxes_rend.cpu().numpy() - g[f'{k}/boxes_rend']).max() < 2e-2, k
        assert np.abs(v.boxes_crop.cpu().numpy() - g[f'{k}/boxes_crop']).max() < 2e-2, k
        Kg = g[f'{k}/K_crop']
        assert (np.abs(v.K_crop.cpu().numpy() - Kg) / np.maximum(np.abs(Kg), 1)).max() < 1e-4, k
    assert list(final.infos['label']) == list(small.infos()['label'])


def test_get_predictions_cfg1(dev, golden_dir):
    """BASELINE.json configs[0]: single crop, 1 object, 1 coarse + 1 refine."""
    from cosypose_b200.utils import tensor_collection as tc
    w = Workload(1, 1, 3, 1, 1)
    pred, eng, views = build_predictor(w, 0)
    det = tc.PandasTensorCollection(infos=w.infos(), bboxes=_d(w.boxes, dev))
    final, _ = pred.get_predictions(_d(w.images, dev), _d(w.K, dev), detections=det)
    g = np.load(golden_dir / 'single_view_cfg1.npz')
    assert np.abs(final.poses.cpu().numpy() - g['final_poses']).max() < TOL_POSE


def test_get_predictions_zup_and_external_init(dev, golden_dir):
    from cosypose_b200.utils import tensor_collection as tc
    w = Workload(1, 2, 3, 1, 1)
    pred, eng, views = build_predictor(w, 0)
    pred.coarse_model.cfg.init_method = 'z-up+auto-depth'
    det = tc.PandasTensorCollection(infos=w.infos(), bboxes=_d(w.boxes, dev))
    final, preds = pred.get_predictions(_d(w.images, dev), _d(w.K, dev), detections=det)
    g = np.load(golden_dir / 'single_view_zup.npz')
    # The refiner iteration of THIS synthetic case crops the whole noisy frame at zoom ~1, where the random network
    # amplifies rounding ~20x per iteration: the reference run in float64 (single_view_zup_fp64.npz, the same
    # reference code with model.double()) differs from the reference run in float32 by 9.5e-5 on the final pose
    # (5e-6 after the coarse iteration), so the fp32 golden is itself only known to ~1e-4 here.  The engine is held to
    # 1e-4 against the float64 reference - the exact result of the reference's arithmetic - and, like the fp32
    # reference, to 2e-4 against the fp32 golden (two fp32 evaluations each within 1e-4 of the exact value).
    g64 = np.load(golden_dir / 'single_view_zup_fp64.npz')
    assert np.abs(preds['coarse/iteration=1'].poses_input.cpu().numpy() - g['coarse/iteration=1/poses_input']).max() < 1e-5
    assert np.abs(preds['coarse/iteration=1'].poses.cpu().numpy() - g['coarse/iteration=1/poses']).max() < TOL_POSE
    f = final.poses.cpu().numpy().astype(np.float64)
    assert float(g64['ref32_minus_fp64'][0]) > 5e-5          # the premise above
    assert np.abs(f - g64['final_poses']).max() < TOL_POSE
    assert np.abs(f - g['final_poses']).max() < 2 * TOL_POSE
    # external init: n_coarse_iterations must be 0, key 'external_coarse' is reported
    init = preds['coarse/iteration=1']
    views.reset()
    views._chunk = 1  # refiner stack
    final2, preds2 = pred.get_predictions(_d(w.images, dev), _d(w.K, dev), data_TCO_init=init,
                                          n_coarse_iterations=0, n_refiner_iterations=1)
    assert 'external_coarse' in preds2
    assert (final2.poses - final.poses).abs().max() < 1e-6
    with pytest.raises(AssertionError):
        pred.get_predictions(_d(w.images, dev), _d(w.K, dev), data_TCO_init=init, n_coarse_iterations=1)


def test_get_predictions_cfg2_all_hypotheses(dev, golden_dir):
    """BASELINE.json configs[1] at full size against the reference itself (tests/golden/make_golden_r2.py):
    every one of the 64 hypotheses, every iteration, within the north star's 1e-4."""
    from cosypose_b200.utils import tensor_collection as tc
    w = Workload(8, 8, 21, 1, 4)
    pred, eng, views = build_predictor(w, 0, bsz_objects=64)
    det = tc.PandasTensorCollection(infos=w.infos(), bboxes=_d(w.boxes, dev))
    final, preds = pred.get_predictions(_d(w.images, dev), _d(w.K, dev), detections=det, n_coarse_iterations=1,
                                        n_refiner_iterations=4)
    g = np.load(golden_dir / 'single_view_cfg2.npz')
    assert final.poses.shape == (64, 4, 4)
    worst = 0.0
    for k, v in preds.items():
        err = np.abs(v.poses.cpu().numpy() - g[f'{k}/poses']).max()
        worst = max(worst, err)
        assert err < TOL_POSE, (k, err)
    assert np.abs(final.poses.cpu().numpy() - g['final_poses']).max() < TOL_POSE
    print('cfg2: worst |dTCO| over 5 iterations x 64 hypotheses', worst)
    eng.close()


def test_get_predictions_symmetric_labels(dev, golden_dir):
    """Labels with 1 / 2 / 4 / 64 discrete symmetries through the single-view path (reference golden)."""
    from cosypose_b200.utils import tensor_collection as tc
    w = Workload(2, 4, 8, 1, 2, sym_counts=(1, 2, 4, 64))
    pred, eng, views = build_predictor(w, 0, bsz_objects=8)
    det = tc.PandasTensorCollection(infos=w.infos(), bboxes=_d(w.boxes, dev))
    final, _ = pred.get_predictions(_d(w.images, dev), _d(w.K, dev), detections=det, n_coarse_iterations=1,
                                    n_refiner_iterations=2)
    g = np.load(golden_dir / 'single_view_sym64.npz')
    assert np.abs(final.poses.cpu().numpy() - g['final_poses']).max() < TOL_POSE
    eng.close()


def test_uint8_views_match_float(small, dev):
    """uint8 NHWC views (the renderer's native output) give the same result as their float/255 form."""
    from cosypose_b200.rendering import PreRenderedViews
    from cosypose_b200.utils import tensor_collection as tc
    pred, eng, views = build_predictor(small, 0, bsz_objects=8)
    u8 = [(v * 255).round().to(torch.uint8) for v in (small.views_c, small.views_r)]
    f32 = [u.float() / 255 for u in u8]
    det = tc.PandasTensorCollection(infos=small.infos(), bboxes=_d(small.boxes, dev))
    outs = []
    for stages in (PreRenderedViews(f32, 8, device=dev),
                   PreRenderedViews.from_uint8([_d(u.permute(0, 1, 3, 4, 2), dev) for u in u8], 8)):
        pred.coarse_model.renderer = pred.refiner_model.renderer = stages
        final, _ = pred.get_predictions(_d(small.images, dev), _d(small.K, dev), detections=det,
                                        n_coarse_iterations=1, n_refiner_iterations=2)
        outs.append(final.poses.clone())
    assert (outs[0] - outs[1]).abs().max() < 1e-6


def test_batch_invariance_and_determinism(dev):
    """Full-size batch (64 hypotheses, 1+4 iterations): results do not depend on the chunking
    (hypotheses are independent), repeat bit-identically, and stay proper rigid transforms."""
    from cosypose_b200.utils import tensor_collection as tc
    w = Workload(8, 8, 21, 1, 4)
    det_infos = w.infos()
    res = []
    for bsz in (64, 64, 16):
        pred, eng, views = build_predictor(w, 0, bsz_objects=bsz)
        det = tc.PandasTensorCollection(infos=det_infos, bboxes=_d(w.boxes, dev))
        final, _ = pred.get_predictions(_d(w.images, dev), _d(w.K, dev), detections=det,
                                        n_coarse_iterations=1, n_refiner_iterations=4)
        res.append(final.poses.cpu())
        eng.close()
    assert torch.equal(res[0], res[1])                      # deterministic (no atomics)
    # chunking changes the GEMM work split only (which m-tiles share their K range between two CTAs depends on the
    # number of rows): fp32 summation order, amplified by 5 iterations; 1.2e-5 measured, the parity budget is 1e-4
    assert (res[0] - res[2]).abs().max() < 3e-5
    R = res[0][:, :3, :3]
    assert (R @ R.transpose(1, 2) - torch.eye(3)).abs().max() < 1e-4
    assert torch.isfinite(res[0]).all() and (res[0][:, 2, 3] > 0.1).all()
    # oracle on a sample of the batch (CPU finishes 8 hypotheses x 5 forwards in seconds)
    from oracle import pose_oracle as po
    sub = slice(0, 8)
    ref, _ = po.coarse_refine_predictions(
        w.images, w.K, w.boxes[sub], w.label_ids[sub], w.im_ids[sub], state_dict(0), state_dict(1), w.points,
        lambda stage, it, sl, T, Kc: (w.views_c if stage == 'coarse' else w.views_r)[it, sub][sl], 1, 4)
    assert (res[0][sub] - ref).abs().max() < TOL_POSE


def test_trunk_repeats_bit_identically(dev):
    """Eight trunk forwards at the benchmark batch: every block-boundary activation is bit-identical.  (The
    tensor-core kernels hand tiles between warps through mbarriers thousands of times per launch; a missing
    dependency or a split warp shows up here as a few differing rows, not in a tolerance test.)"""
    from cosypose_b200.engine import Engine
    B = 64
    gen = torch.Generator().manual_seed(3)
    crops = torch.rand((B, 3, 240, 320), generator=gen).to(dev)
    renders = torch.rand((B, 3, 240, 320), generator=gen).to(dev)
    for groups in (1, 0):
        eng = Engine(0, max_batch=B)
        eng.load_pose_model(0, state_dict(0))
        eng.set_option('tc_groups', groups)
        _, t0 = eng.net_forward(0, crops, renders, taps=True)
        t0 = {k: v.clone() for k, v in t0.items()}
        for _ in range(7):
            _, t1 = eng.net_forward(0, crops, renders, taps=True)
            for name in t0:
                assert torch.equal(t0[name], t1[name]), (groups, name)
        eng.set_option('tc_groups', 0)
        eng.close()


def test_errors(engine, small, dev):
    with pytest.raises(AssertionError):
        engine.net_forward(0, torch.zeros((9, 3, 240, 320), device=dev), torch.zeros((9, 3, 240, 320), device=dev))
    with pytest.raises(AssertionError):
        engine.prepare_iter(torch.zeros((2, 3, 3)), torch.zeros((2, 4, 4), device=dev),
                            torch.zeros(2, dtype=torch.int32, device=dev), (480, 640))
    from cosypose_b200.models.pose import PosePredictor
    with pytest.raises(ValueError):
        PosePredictor(engine, 0, None, None, pose_dim=7)

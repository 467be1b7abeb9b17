"""Pins the CPU oracle (oracle/pose_oracle.py) to the reference: it must reproduce the golden vectors
that tests/golden/make_golden.py produced by running the UNMODIFIED reference on CPU, and its RoI
crop must equal the compiled torchvision operator the reference calls.  Runs without a GPU."""
import numpy as np
import pytest
import torch

from helpers import Workload, state_dict

# Oracle and reference run the same torch CPU operators; thread counts / BLAS blocking differ
# between machines, so equality is to rounding, not bitwise.  The z-up case crops the whole noisy
# frame at zoom ~1, where this random network amplifies rounding differences ~100x: there the
# north-star tolerance itself (1e-4 on pose matrices) is used.
TOL = 2e-5
TOL_ZUP = 1e-4


def _checksum(t):
    return float(np.asarray(t, dtype=np.float64).sum())


def test_inputs_reproduce_from_seeds(golden_dir):
    g = np.load(golden_dir / 'single_view_small.npz')
    n_images, dets, n_labels, n_coarse, n_refine, bsz, zup = g['meta']
    w = Workload(n_images, dets, n_labels, n_coarse, n_refine)
    assert _checksum(w.images) == pytest.approx(float(g['chk_images']), rel=1e-9)
    assert _checksum(w.boxes) == pytest.approx(float(g['chk_boxes']), rel=1e-9)
    assert _checksum(w.views_r) == pytest.approx(float(g['chk_views']), rel=1e-9)
    assert _checksum(state_dict(1)['backbone._conv_head.weight']) == pytest.approx(float(g['chk_sd']), rel=1e-6)


@pytest.mark.parametrize('name', ['single_view_cfg1', 'single_view_small', 'single_view_zup'])
def test_oracle_matches_reference_predictor(golden_dir, name):
    from oracle import pose_oracle as po
    g = np.load(golden_dir / f'{name}.npz')
    n_images, dets, n_labels, n_coarse, n_refine, bsz, zup = [int(x) for x in g['meta']]
    w = Workload(n_images, dets, n_labels, n_coarse, n_refine)
    torch.set_num_threads(4)
    final, preds = po.coarse_refine_predictions(
        w.images, w.K, w.boxes, w.label_ids, w.im_ids, state_dict(0), state_dict(1), w.points,
        w.oracle_render_fn(), n_coarse, n_refine, bsz_objects=bsz,
        init_method='z-up+auto-depth' if zup else 'v0')
    tol = TOL_ZUP if zup else TOL
    assert np.abs(final.numpy() - g['final_poses']).max() < tol
    names = dict(poses='TCO_output', poses_input='TCO_input', K_crop='K_crop', boxes_rend='boxes_rend',
                 boxes_crop='boxes_crop')
    for k, v in preds.items():
        for gk, ok in names.items():
            ref = g[f'{k}/{gk}']
            err = np.abs(v[ok].numpy() - ref) / np.maximum(np.abs(ref), 1.0)
            assert err.max() < tol, (k, gk)


def test_oracle_backbone_matches_reference(golden_dir):
    from cosypose_b200 import synthetic as syn
    from oracle import pose_oracle as po
    g = np.load(golden_dir / 'backbone_b2.npz')
    x = syn.make_net_input(2, seed=21)
    assert _checksum(x) == pytest.approx(float(g['chk_x']), rel=1e-9)
    taps = {}
    torch.set_num_threads(4)
    pose = po.net_forward(x, state_dict(0), taps)
    assert np.abs(pose.numpy() - g['pose']).max() < TOL
    assert np.abs(taps['pooled'].numpy() - g['pooled']).max() < 1e-4
    for i in (0, 1, 2, 5, 8, 13, 18, 25):
        a = taps[f'block{i}'][0].permute(1, 2, 0).numpy()[::4, ::4]
        assert np.abs(a - g[f'block{i}']).max() <= 1e-4 * np.abs(g[f'block{i}']).max()
    sums = np.array([taps[f'block{i}'].double().sum().item() for i in range(26)])
    assert np.allclose(sums, g['block_sum'], rtol=1e-4)
    # activations stay O(1) through all 26 blocks: parity on this net is not vacuous
    assert g['block_mean_abs'].min() > 0.3 and g['block_mean_abs'].max() < 5


@pytest.mark.parametrize('box', [[100., 80., 420., 320.], [-40.5, -30.25, 200., 150.], [500., 300., 700.5, 520.],
                                 [-900., -700., -100., -100.], [300.2, 200.7, 300.6, 201.0], [0., 0., 64., 48.]])
def test_oracle_roi_align_equals_compiled_operator(box):
    """The oracle restates the compiled CPU operator's out-of-image rule (samples dropped, not
    edge-clamped); torchvision's pure-Python fallback differs and is not used."""
    import torchvision
    from oracle import pose_oracle as po
    gen = torch.Generator().manual_seed(0)
    images = torch.rand((2, 3, 48, 64), generator=gen)
    boxes = torch.tensor([box, box])
    im_ids = np.array([0, 1])
    out = po.roi_align_crop(images, im_ids, boxes)
    rois = torch.cat((torch.tensor([[0.], [1.]]), boxes), dim=1)
    ref = torchvision.ops.roi_align(images, rois, output_size=(240, 320), spatial_scale=1.0, sampling_ratio=4)
    assert (out - ref).abs().max() < 1e-6

"""Device rasteriser against its CPU restatement (oracle/raster_oracle.py), bit for bit, and the
in-engine render loop against the two-phase loop a host renderer needs."""
import numpy as np
import pytest
import torch

from helpers import Workload, state_dict

pytestmark = pytest.mark.gpu


def _table(n_labels=4, subdiv=2):
    from cosypose_b200 import synthetic
    from cosypose_b200.rendering import RenderMeshTable
    v, f, c = synthetic.make_render_meshes(n_labels, subdiv=subdiv)
    return RenderMeshTable(synthetic.make_labels(n_labels), v, f, c)


def _poses(B, seed=0):
    from scipy.spatial.transform import Rotation as R
    rs = np.random.RandomState(seed)
    T = np.tile(np.eye(4, dtype=np.float32), (B, 1, 1))
    T[:, :3, :3] = R.random(B, random_state=seed + 1).as_matrix().astype(np.float32)
    T[:, 2, 3] = rs.uniform(0.25, 0.6, B)
    T[:, :2, 3] = rs.uniform(-0.03, 0.03, (B, 2))
    K = np.tile(np.array([[900, 0, 160], [0, 900, 120], [0, 0, 1]], dtype=np.float32), (B, 1, 1))
    K[:, 0, 0] *= rs.uniform(0.7, 1.5, B).astype(np.float32)
    K[:, 1, 1] *= rs.uniform(0.7, 1.5, B).astype(np.float32)
    K[:, 0, 1] = rs.uniform(-2, 2, B)     # skew is honoured (simulator/camera.py:13)
    return T, K


@pytest.mark.parametrize('subdiv,zscale', [(2, 1.0), (4, 2.5)])
def test_frames_bit_exact_vs_oracle(subdiv, zscale):
    """subdiv 2: triangles of 10^2..10^5 pixels (the warp and the large-triangle tier); subdiv 4 seen from 2.5x the
    distance: pixel-sized triangles (the per-lane tier)."""
    from cosypose_b200.engine import Engine
    from oracle import raster_oracle as ro
    dev = torch.device('cuda', 0)
    tab = _table(subdiv=subdiv)
    eng = Engine(dev, max_batch=16)
    eng.set_render_meshes(tab.vertices, tab.colors, tab.faces, tab.face_offsets)
    B = 10
    T, K = _poses(B)
    T[:, 2, 3] *= zscale
    labels = np.array([0, 1, 2, 3, 0, 1, 2, 3, 0, 1], dtype=np.int32)
    T[4, 0, 3] = 0.08          # mostly outside the view on the right
    T[5, 2, 3] = 0.06          # box straddling the near plane: triangles in front of it are dropped
    T[6, 2, 3] = -0.4          # behind the camera: black frame
    T[7, 1, 1] = np.nan        # invalid pose: black frame (bullet_batch_renderer.py:27-38)
    K[8, 0, 0] = K[8, 1, 1] = 4000.0   # zoomed in: triangles spanning the whole view
    K[9, 0, 0] = K[9, 1, 1] = 5000.0
    ref, zref, ids = ro.render(tab.vertices, tab.colors, tab.faces, tab.face_offsets, labels, T, K)
    out, depth = eng.render(torch.from_numpy(labels).to(dev), torch.from_numpy(T).to(dev), torch.from_numpy(K).to(dev),
                            depth=True)
    torch.cuda.synchronize()
    got = out.cpu().numpy()
    assert np.array_equal(depth.cpu().numpy(), np.where(ids >= 0, zref, np.float32(0)))
    assert got.shape == (B, 240, 320, 3) and got.dtype == np.uint8
    assert np.array_equal(got, ref), f'{(got != ref).any(axis=-1).sum()} pixels differ'
    covered = (ids >= 0).reshape(B, -1).mean(axis=1)
    assert covered[0] > 0.05 / zscale ** 2 and covered[8] > 0.5 / zscale ** 2       # the cases are not vacuous
    assert covered[6] == 0 and covered[7] == 0 and not got[6].any() and not got[7].any()
    # float output: the same frames / 255, NCHW
    outf = eng.render(torch.from_numpy(labels).to(dev), torch.from_numpy(T).to(dev), torch.from_numpy(K).to(dev),
                      uint8=False).cpu().numpy()
    assert np.array_equal(outf, (ref.astype(np.float32) / np.float32(255)).transpose(0, 3, 1, 2))


def test_render_is_run_to_run_identical():
    from cosypose_b200.engine import Engine
    dev = torch.device('cuda', 0)
    tab = _table(subdiv=4)
    eng = Engine(dev, max_batch=16)
    eng.set_render_meshes(tab.vertices, tab.colors, tab.faces, tab.face_offsets)
    T, K = _poses(16, seed=3)
    labels = torch.arange(16, dtype=torch.int32, device=dev) % 4
    a = eng.render(labels, torch.from_numpy(T).to(dev), torch.from_numpy(K).to(dev)).clone()
    for _ in range(3):
        assert torch.equal(a, eng.render(labels, torch.from_numpy(T).to(dev), torch.from_numpy(K).to(dev)))
    assert a.any()


def test_render_requires_meshes():
    from cosypose_b200.engine import Engine
    dev = torch.device('cuda', 0)
    eng = Engine(dev, max_batch=4)
    T, K = _poses(2)
    with pytest.raises(RuntimeError, match='render meshes not set'):
        eng.render(torch.zeros(2, dtype=torch.int32, device=dev), torch.from_numpy(T).to(dev), torch.from_numpy(K).to(dev))


@pytest.mark.parametrize('graph', [0, 1])
def test_in_engine_loop_equals_two_phase_loop(graph):
    """All iterations inside the engine (refine_n without views, captured as one graph) == prepare_iter -> render ->
    refine_iter per iteration, the sequence a host renderer needs (models/pose.py:99-108), bit for bit."""
    from cosypose_b200.engine import Engine
    from cosypose_b200.integrated.pose_predictor import CoarseRefinePosePredictor
    from cosypose_b200.lib3d.rigid_mesh_database import BatchedMeshes
    from cosypose_b200.models.pose import PosePredictor
    from cosypose_b200.rendering import CudaRasterizer
    from cosypose_b200.utils import tensor_collection as tc
    dev = torch.device('cuda', 0)
    w = Workload(2, 4, 4, 1, 2)
    tab = _table(4)
    assert tab.labels == list(w.labels)

    class HostStyle:       # same frames, but through render(): forces the two-phase path
        def __init__(self, r):
            self.r = r

        def render(self, **kw):
            return self.r.render(**kw)

    results = []
    for in_engine in (True, False):
        eng = Engine(dev, max_batch=8)
        eng.set_option('graph', graph)
        mesh_db = BatchedMeshes.from_tables(w.labels, w.points, w.sym, w.n_sym)
        mesh_db.install(eng)
        rast = CudaRasterizer(eng, tab)
        renderer = rast if in_engine else HostStyle(rast)
        coarse = PosePredictor(eng, 0, renderer, mesh_db).load_state_dict(state_dict(0))
        refiner = PosePredictor(eng, 1, renderer, mesh_db).load_state_dict(state_dict(1))
        pred = CoarseRefinePosePredictor(coarse, refiner, bsz_objects=8)
        det = tc.PandasTensorCollection(infos=w.infos(), bboxes=w.boxes.to(dev))
        for _ in range(2):   # the second call replays the graph
            final, preds = pred.get_predictions(w.images.to(dev), w.K.to(dev), detections=det,
                                                n_coarse_iterations=1, n_refiner_iterations=2)
        torch.cuda.synchronize()
        results.append({k: v.poses.clone() for k, v in preds.items()})
        n_render = eng.profile_read()['render'][0]
        assert n_render == 2 * 2 * 3, n_render      # 2 launches per iteration, 3 iterations, 2 calls
    for k in results[0]:
        assert torch.equal(results[0][k], results[1][k]), k
    assert torch.isfinite(results[0]['refiner/iteration=2']).all()

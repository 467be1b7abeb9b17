"""Developer script (not a test): prints stage-by-stage max errors engine vs oracle on the GPU."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch
from cosypose_b200 import synthetic as syn, effnet_spec as spec
from cosypose_b200.engine import Engine
from oracle import pose_oracle as po
sys.path.insert(0, str(Path(__file__).resolve().parent))
from helpers import Workload, build_predictor, state_dict

torch.manual_seed(0)
dev = torch.device('cuda', 0)
w = Workload(2, 3, 5, 1, 2)
eng = Engine(0, max_batch=8)
from cosypose_b200.lib3d.rigid_mesh_database import BatchedMeshes
mesh_db = BatchedMeshes.from_tables(w.labels, w.points, w.sym, w.n_sym); mesh_db.install(eng)
eng.load_pose_model(0, state_dict(0))
im_ids = torch.as_tensor(w.im_ids); lab = torch.as_tensor(w.label_ids)
K_h = w.K[im_ids]
TCO = po.TCO_init_from_boxes(w.boxes, K_h)
d = lambda t: t.to(dev).contiguous()
T_e = eng.tco_init(d(w.boxes), d(K_h), d(lab.int()))
print('tco_init', (T_e.cpu() - TCO).abs().max().item())
pts = po.select_points(w.points, w.label_ids)
crop_o, Kc_o, br_o, bc_o = po.crop_inputs(w.images, w.im_ids, K_h, TCO, pts)
br, bc, Kc = eng.prepare_iter(d(K_h), d(TCO), d(lab.int()), (480, 640))
print('boxes_rend', (br.cpu() - br_o).abs().max().item(), 'boxes_crop', (bc.cpu() - bc_o).abs().max().item(),
      'K_crop', (Kc.cpu() - Kc_o).abs().max().item())
crop = eng.roi_crop(d(w.images), d(im_ids.int()), d(bc_o))
print('roi_crop', (crop.cpu() - crop_o).abs().max().item())
renders = w.views_c[0]
taps_o = {}
x = torch.cat((crop_o, renders), 1)
t = time.time(); pose_o = po.net_forward(x, state_dict(0), taps_o); print('oracle fwd s', time.time() - t)
pose_e, taps_e = eng.net_forward(0, d(crop_o), d(renders), taps=True)
torch.cuda.synchronize()
for name, h_, w_, c_ in spec.activation_shapes()[1:]:
    a = taps_e[name].cpu().permute(0, 3, 1, 2)
    b = taps_o[name]
    print(f'{name:8s} max|d| {(a - b).abs().max().item():.3e}  max|ref| {b.abs().max().item():.3e}')
print('pose9', (pose_e.cpu() - pose_o).abs().max().item())
To = po.update_pose(TCO, Kc_o, pose_o)
Te = eng.update_pose(d(TCO), d(Kc_o), d(pose_o))
print('update_pose', (Te.cpu() - To).abs().max().item())
# end to end
pred, eng2, views = build_predictor(w, 0, bsz_objects=4)
from cosypose_b200.utils import tensor_collection as tc
det = tc.PandasTensorCollection(infos=w.infos(), bboxes=d(w.boxes))
final, preds = pred.get_predictions(d(w.images), d(w.K), detections=det, n_coarse_iterations=1, n_refiner_iterations=2)
To, preds_o = po.coarse_refine_predictions(w.images, w.K, w.boxes, w.label_ids, w.im_ids, state_dict(0), state_dict(1),
                                           w.points, w.oracle_render_fn(), 1, 2, bsz_objects=4)
print('e2e final pose', (final.poses.cpu() - To).abs().max().item())
g = np.load(Path(__file__).parent / 'golden' / 'single_view_small.npz')
print('e2e vs golden', np.abs(final.poses.cpu().numpy() - g['final_poses']).max(), 'oracle vs golden', np.abs(To.numpy() - g['final_poses']).max())
# quick timing B=64
w2 = Workload(8, 8, 21, 1, 4)
pred, eng3, views = build_predictor(w2, 0, bsz_objects=64)
det = tc.PandasTensorCollection(infos=w2.infos(), bboxes=d(w2.boxes))
imgs, K = d(w2.images), d(w2.K)
for i in range(3):
    views.reset(); torch.cuda.synchronize(); t = time.time()
    final, preds = pred.get_predictions(imgs, K, detections=det, n_coarse_iterations=1, n_refiner_iterations=4)
    torch.cuda.synchronize(); dt = time.time() - t
    print(f'B=64 1+4 iters: {dt*1e3:.1f} ms -> {64/dt:.0f} hyp/s')
print('final z range', final.poses[:, 2, 3].min().item(), final.poses[:, 2, 3].max().item())

"""z-up golden case: distance of every GEMM implementation to the fp32 reference and to the float64 reference."""
import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, torch
from helpers import Workload, build_predictor
from cosypose_b200.utils import tensor_collection as tc
dev=torch.device('cuda',0)
g=np.load('/root/repo/tests/golden/single_view_zup.npz')
g64=np.load('/root/repo/tests/golden/single_view_zup_fp64.npz')
print('reference fp32 vs reference fp64: final', np.abs(g['final_poses']-g64['final_poses']).max(), 'coarse', np.abs(g['coarse/iteration=1/poses']-g64['coarse/iteration=1/poses']).max())
for impl, xdw in ((0,0),(0,1),(1,0),(2,0),(2,1)):
    w=Workload(1,2,3,1,1)
    pred,eng,views=build_predictor(w,0)
    eng.set_option('gemm_impl',impl)
    eng.set_option('xdw',xdw)
    pred.coarse_model.cfg.init_method='z-up+auto-depth'
    det=tc.PandasTensorCollection(infos=w.infos(),bboxes=w.boxes.to(dev))
    final,preds=pred.get_predictions(w.images.to(dev),w.K.to(dev),detections=det)
    f=final.poses.cpu().numpy().astype(np.float64); c=preds['coarse/iteration=1'].poses.cpu().numpy().astype(np.float64)
    print(f'impl {impl} xdw {xdw}: final vs ref32 {np.abs(f-g["final_poses"]).max():.3e} vs ref64 {np.abs(f-g64["final_poses"]).max():.3e} | coarse vs ref32 {np.abs(c-g["coarse/iteration=1/poses"]).max():.3e} vs ref64 {np.abs(c-g64["coarse/iteration=1/poses"]).max():.3e}')
    eng.close()

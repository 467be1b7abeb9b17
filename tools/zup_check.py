import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, torch
from helpers import Workload, build_predictor
from cosypose_b200.utils import tensor_collection as tc
dev=torch.device('cuda',0)
g=np.load('/root/repo/tests/golden/single_view_zup.npz')
for impl in (0,1,2):
    w=Workload(1,2,3,1,1)
    pred,eng,views=build_predictor(w,0)
    eng.set_option('gemm_impl',impl)
    pred.coarse_model.cfg.init_method='z-up+auto-depth'
    det=tc.PandasTensorCollection(infos=w.infos(),bboxes=w.boxes.to(dev))
    final,preds=pred.get_predictions(w.images.to(dev),w.K.to(dev),detections=det)
    print('impl',impl,'final err',np.abs(final.poses.cpu().numpy()-g['final_poses']).max(), 'coarse err', np.abs(preds['coarse/iteration=1'].poses.cpu().numpy()-g['coarse/iteration=1/poses']).max())

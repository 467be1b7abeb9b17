"""One trunk forward of 64 hypotheses (and optionally one multiview scene) inside a cudaProfilerStart/Stop range, for
ncu (`--profile-from-start off`).  COSYB200_GRAPH=0 is set so that every kernel is a plain launch.

    ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv \\
        --log-file gpurun_out/r02_ncu_dram_per_forward.csv python tools/prof_forward.py
    ncu --profile-from-start off --set full --import-source on -o gpurun_out/r02_full python tools/prof_forward.py --multiview
"""
import os
import sys
from pathlib import Path
os.environ['COSYB200_GRAPH'] = '0'
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / 'tests'))
import torch
from helpers import Scene, Workload, build_predictor
from cosypose_b200.utils import tensor_collection as tc
dev = torch.device('cuda', 0)
w = Workload(8, 8, 21, 1, 1)
pred, eng, views = build_predictor(w, 0, bsz_objects=64)
det = tc.PandasTensorCollection(infos=w.infos(), bboxes=w.boxes.to(dev))
images, K = w.images.to(dev), w.K.to(dev)
mv = None
if '--multiview' in sys.argv:
    from cosypose_b200.integrated.multiview_predictor import MultiviewScenePredictor
    sc = Scene(8, 16, 21, (1,), True, 0)
    mv = MultiviewScenePredictor(sc.mesh_db(), device=dev)
    cands, cams = sc.candidates(dev), sc.cameras(dev)


def run():
    views.reset()
    pred.get_predictions(images, K, detections=det, n_coarse_iterations=1, n_refiner_iterations=0)
    if mv is not None:
        mv.predict_scene_state(cands, cams, ransac_n_iter=2000, ransac_dist_threshold=0.02, ba_n_iter=2)


run()
torch.cuda.synchronize()
torch.cuda.profiler.start()
run()
torch.cuda.synchronize()
torch.cuda.profiler.stop()

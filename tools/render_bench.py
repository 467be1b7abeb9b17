"""Device time of the rasteriser at the benchmark batch (64 hypotheses), dense and coarse meshes.

    python tools/render_bench.py [--subdiv 5]
"""
import argparse
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / 'tests'))

import numpy as np  # noqa: E402
import torch  # noqa: E402
from cosypose_b200 import synthetic  # noqa: E402
from cosypose_b200.engine import Engine  # noqa: E402
from cosypose_b200.rendering import RenderMeshTable  # noqa: E402
from test_gpu_render import _poses  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--subdiv', type=int, default=5)
ap.add_argument('--batch', type=int, default=64)
args = ap.parse_args()
dev = torch.device('cuda', 0)
v, f, c = synthetic.make_render_meshes(4, subdiv=args.subdiv)
tab = RenderMeshTable(synthetic.make_labels(4), v, f, c)
eng = Engine(dev, max_batch=args.batch)
eng.set_render_meshes(tab.vertices, tab.colors, tab.faces, tab.face_offsets)
B = args.batch
T, K = _poses(B, seed=5)
T, K = torch.from_numpy(T).to(dev), torch.from_numpy(K).to(dev)
for name, labels in (('icosphere', [0, 2]), ('box', [1, 3]), ('mixed', [0, 1, 2, 3])):
    lab = torch.tensor([labels[i % len(labels)] for i in range(B)], dtype=torch.int32, device=dev)
    out = eng.render(lab, T, K)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    n = 20
    ev[0].record()
    for _ in range(n):
        eng.render(lab, T, K, out=out)
    ev[1].record()
    torch.cuda.synchronize()
    us = ev[0].elapsed_time(ev[1]) * 1e3 / n
    cov = float((out.sum(dim=-1) > 0).float().mean())
    faces = int(np.mean([tab.face_offsets[l + 1] - tab.face_offsets[l] for l in labels]))
    print(f'{name:10s} {faces:7d} faces/object  coverage {cov:.2f}  {us:8.1f} us per batch of {B}  ({us / B:.2f} us per view)')

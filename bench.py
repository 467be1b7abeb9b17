#!/usr/bin/env python
"""Benchmark of the refinement hot path (contract: task prompt section 4; metric: BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 1|2|3|4]

--config selects the BASELINE.json workload (index into `configs`; configs[0] is the CPU plumbing case and is a
parity test, not a bench line):
  1 (default)  64 synthetic "YCB-V" crops per GPU (8 frames of 640x480 x 8 detections, 21 labels), 1 coarse + 4
               refiner iterations.  With N GPUs the GLOBAL batch is 64 N hypotheses of rank-distinct data (weak
               scaling): every rank calls CoarseRefinePosePredictor.get_predictions(shard=True) with the global
               detection table, refines its contiguous shard and the per-iteration records are exchanged with one
               NCCL all-gather (libcosyb200's cosyb200_allgather_candidates), so each rank ends with all poses.
  2            512 hypotheses over 21 "T-LESS" classes with 1/2/4/64 symmetries (64 frames x 8 detections = 8 scenes
               of 8 views), 1 + 4 iterations, sharded over the N GPUs (strong scaling), all-gather, then the
               multiview candidate matching of every scene replicated on every rank.
  3            8-view scene, 16 detections per view: RANSAC candidate matching (13 440 seeds, 215 040 scored rows) +
               2 bundle-adjustment iterations through MultiviewScenePredictor.predict_scene_state; replicas only.
  4            BOP-style mix (132 labels), 2048 hypotheses (256 frames x 8), 1 coarse + 2 refiner iterations, sharded
               over the N GPUs (strong scaling).
A step = one pass of that workload.  `value` = hypotheses/s (scenes/s for config 3) with inputs resident in HBM,
CUDA-event timed over K steps after W warm-up steps, max over ranks.  `e2e` = the same through the public API with
HOST inputs: per step the frames, intrinsics, detections and uint8 views are copied from pinned host memory into
device buffers and the final poses are read back.  `roofline` = the EfficientNet-B3 trunk against SURVEY.md section
8(d)'s algorithmic bytes (24.43 MB per forward and hypothesis); per-kernel figures are in its sub-objects.
`--impl reference` / `cpu_baseline`: the UNMODIFIED reference (baseline/_ref, installed by __graft_entry__.build())
through its own CoarseRefinePosePredictor on the host cores, on a bounded sample of the same workload; if
baseline/_ref is absent the oracle port (oracle/pose_oracle.py) is timed instead and `kind` says so.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / 'tests'))

import numpy as np  # noqa: E402
import pandas as pd  # noqa: E402
import torch  # noqa: E402

UNIT = 'hypotheses/s'
ALGO_BYTES_PER_FORWARD = 6107136 * 4      # block-boundary activations, fp32
ALGO_FLOP_PER_FORWARD = 2 * 1456.6e6
BSZ = 64

# name, frames per GPU-or-total, detections per frame, labels, symmetry counts, coarse, refine, scaling
CONFIGS = {
    1: dict(name='configs[1]', frames=8, dets=8, labels=21, sym=(1,), n_coarse=1, n_refine=4, scaling='weak',
            text='64 synthetic YCB-V crops per GPU (8 frames 640x480 x 8 detections, 21 labels), 1 coarse + 4 refine '
                 'iters, random-init BN-calibrated EfficientNet-B3, pre-rendered views'),
    2: dict(name='configs[2]', frames=64, dets=8, labels=21, sym=(1, 2, 4, 64), n_coarse=1, n_refine=4, scaling='strong',
            text='512 hypotheses over 21 T-LESS-like classes with 1/2/4/64 symmetries (8 scenes x 8 views x 8 '
                 'detections), 1 coarse + 4 refine iters sharded over the GPUs, one all-gather, then multiview '
                 'candidate matching of every scene on the gathered set', matching=True),
    3: dict(name='configs[3]', text='8-view synthetic scene, 16 detections per view, multiview RANSAC matching '
                                    '(2000 iterations per view pair: 13440 seeds, 215040 scored rows) + 2 bundle-'
                                    'adjustment iterations; replicas only', scaling='weak'),
    4: dict(name='configs[4]', frames=256, dets=8, labels=132, sym=(1,), n_coarse=1, n_refine=2, scaling='strong',
            text='BOP-style mixed batch (132 labels: lm 15, tless 30, tudl 3, icbin 2, itodd 28, hb 33, ycbv 21), 2048 '
                 'hypotheses (256 frames x 8 detections), 1 coarse + 2 refine iters sharded over the GPUs'),
}


def metric_name(cfg):
    if cfg == 3:
        return 'multiview scenes/sec (8 views x 16 detections, RANSAC matching + 2 BA iters)'
    c = CONFIGS[cfg]
    return f'refined object-hypotheses/sec ({c["n_coarse"]} coarse + {c["n_refine"]} refine iters, 640x480)'


def kernel_algorithmic_bytes():
    """Bytes each kernel category must move per forward in the one-kernel-per-stage design (its input read once,
    its output written once, fp32; weights are L2-resident): the per-kernel roofline numerators.  The whole-trunk
    figure of SURVEY.md section 8(d), ALGO_BYTES_PER_FORWARD, counts only the block-boundary activations."""
    from cosypose_b200 import effnet_spec as spec
    shapes = spec.activation_shapes()
    out = dict(stem=(shapes[0][1] * shapes[0][2] * 6 + shapes[1][1] * shapes[1][2] * shapes[1][3]) * 4,
               expand_1x1=0, depthwise=0, project_1x1=0)
    for b, (_, hi, wi, _), (_, ho, wo, _) in zip(spec.BLOCKS, shapes[1:-2], shapes[2:-1]):
        if b.e != 1:
            out['expand_1x1'] += hi * wi * (b.cin + b.cexp) * 4
        out['depthwise'] += (hi * wi + ho * wo) * b.cexp * 4
        out['project_1x1'] += ho * wo * (b.cexp + b.cout * (2 if b.skip else 1)) * 4
    _, h, w, c = shapes[-2]
    out['head_1x1'] = h * w * (c + spec.N_FEATURES) * 4
    return out


# DRAM traffic of one trunk forward at the benchmark batch, from the committed ncu capture
# (profiles/r02_ncu_dram_per_forward.csv: dram__bytes_read.sum + dram__bytes_write.sum summed over every kernel of one
# forward of 64 hypotheses, crop and geometry included).
NCU_TRUNK_DRAM_BYTES_PER_FORWARD_BATCH = 5.171e9   # = 80.7 MB per hypothesis, 3.3x the algorithmic 24.43 MB (round 1: 148 MB)
TRUNK_CATS = ('stem', 'expand_1x1', 'depthwise', 'squeeze_excite', 'project_1x1', 'head_1x1', 'pool_fc_update')


def roofline_object(prof, prof_steps, fwd_per_step, peak, peak_src, step_ms=None):
    """`roofline` of the JSON line: the trunk forward (every launch between the crop and the pose update) against the
    HBM roofline at SURVEY.md section 8(d)'s algorithmic bytes: each stage reads its input once and writes its output
    once, nothing else touches HBM = 24.43 MB per forward and hypothesis.  achieved = those bytes / the trunk's device
    time, measured live with CUDA events around every launch on the launching stream (engine profiling mode);
    `frac_of_step` divides by the whole timed step instead (launch gaps, crop and geometry included).
    by_kernel_*: every kernel category against the bytes of a one-kernel-per-stage design (its own input + output)."""
    kbytes = kernel_algorithmic_bytes()
    ms = {c: prof[c][1] / prof_steps for c in prof}                 # device ms per step and category
    n_launch = {c: prof[c][0] / prof_steps for c in prof}
    by_kernel_gbs = {c: round(fwd_per_step * kbytes[c] / (ms[c] * 1e-3) / 1e9, 1) for c in kbytes if ms.get(c, 0) > 0}
    bb_ms = sum(ms[c] for c in TRUNK_CATS)
    bb_launches = sum(n_launch[c] for c in TRUNK_CATS)
    algo = fwd_per_step * ALGO_BYTES_PER_FORWARD
    achieved = algo / (bb_ms * 1e-3) / 1e9 if bb_ms > 0 else None
    tot = sum(ms.values())
    n_fwd_launches = fwd_per_step / BSZ                              # trunk forwards (of <= 64 hypotheses) per step
    out = {
        'bound': 'hbm',
        'kernel': 'EfficientNet-B3 trunk forward: stem, 26 MBConv blocks (fused expand+depthwise+pool kernel, '
                  'squeeze-excite, tcgen05 project), head, pool+FC+pose update',
        'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak if achieved else None,
        'peak_source': peak_src,
        'algorithmic_bytes_per_launch': ALGO_BYTES_PER_FORWARD * BSZ,
        'launch': 'one trunk forward of 64 hypotheses (%.0f kernel launches)' % (bb_launches / max(n_fwd_launches, 1e-9)),
        'launch_us': bb_ms * 1e3 / n_fwd_launches if n_fwd_launches else None,
        'traffic': NCU_TRUNK_DRAM_BYTES_PER_FORWARD_BATCH,
        'traffic_note': 'dram__bytes_read.sum + dram__bytes_write.sum over every kernel of one forward of 64 hypotheses '
                        '(profiles/r02_ncu_dram_per_forward.csv)',
        'share_of_step': bb_ms / tot if tot > 0 else None,
        'tflops': fwd_per_step * ALGO_FLOP_PER_FORWARD / (bb_ms * 1e-3) / 1e12 if bb_ms > 0 else None,
        'by_kernel_gbs': by_kernel_gbs,
        'by_kernel_frac': {c: round(v / peak, 4) for c, v in by_kernel_gbs.items()},
        'by_kernel_ms_per_step': {c: round(v, 4) for c, v in ms.items() if v > 0},
        'by_kernel_share': {c: round(v / tot, 4) for c, v in ms.items() if v > 0},
    }
    if step_ms:
        out['frac_of_step'] = algo / (step_ms * 1e-3) / 1e9 / peak
    return out


def measured_peaks():
    p = ROOT / 'MEASURED_PEAKS.json'
    if p.exists():
        d = json.loads(p.read_text())
        return float(d['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                          '-lms', '100', '-i', str(index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace('.', '').isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace('.', '').isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 6 and r[2 + i].lower() == 'active' for r in self.rows)]
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': reasons, 'samples': len(sm)}


# ---- workloads ------------------------------------------------------------------------------------------------
class SingleViewJob:
    """Global inputs of one get_predictions call over `world` ranks (every rank builds the same tables from seeds),
    plus this rank's pre-rendered views (seeded per rank: rank-distinct data)."""

    def __init__(self, cfg, rank, world):
        from cosypose_b200 import synthetic as syn
        c = CONFIGS[cfg]
        self.cfg, self.c = cfg, c
        frames = c['frames'] * (world if c['scaling'] == 'weak' else 1)
        self.labels = syn.make_labels(c['labels'])
        self.points, self.sym, self.n_sym = syn.make_mesh_tables(c['labels'], sym_counts=c['sym'])
        self.boxes, self.label_ids, self.im_ids = syn.make_detections(frames, c['dets'], c['labels'])
        self.n = len(self.label_ids)
        self.n_frames = frames
        self.K = syn.make_camera_K(frames)
        from cosypose_b200.sharding import shard_bounds
        self.start, self.stop = shard_bounds(self.n, rank, world)
        self.n_local = self.stop - self.start
        # frames this rank's hypotheses reference (contiguous: detections are frame-major)
        loc = self.im_ids[self.start:self.stop]
        self.f0, self.f1 = (int(loc.min()), int(loc.max()) + 1) if self.n_local else (0, 0)
        self.images_local = syn.make_images(self.f1 - self.f0, seed=5 + self.f0)     # [f1-f0,3,480,640]
        self.views_c = syn.make_renders(c['n_coarse'], self.n_local, seed=11 + 1000 * rank)
        self.views_r = syn.make_renders(c['n_refine'], self.n_local, seed=12 + 1000 * rank)
        self.scene_ids = self.im_ids // 8 if c.get('matching') else None

    def infos(self):
        d = dict(label=[self.labels[i] for i in self.label_ids], batch_im_id=self.im_ids, score=np.ones(self.n))
        if self.scene_ids is not None:
            d.update(scene_id=self.scene_ids, view_id=self.im_ids % 8, group_id=0)
        return pd.DataFrame(d)


def build_predictor(job, device_index, views_mode='prerendered'):
    from helpers import state_dict
    from cosypose_b200.engine import Engine
    from cosypose_b200.integrated.pose_predictor import CoarseRefinePosePredictor
    from cosypose_b200.lib3d.rigid_mesh_database import BatchedMeshes
    from cosypose_b200.models.pose import PosePredictor
    from cosypose_b200.rendering import PreRenderedViews
    eng = Engine(device_index, max_batch=BSZ)
    mesh_db = BatchedMeshes.from_tables(job.labels, job.points, job.sym, job.n_sym)
    mesh_db.install(eng)
    if views_mode == 'engine':
        # the engine's own rasteriser draws every iteration's views: even labels a 20 480-triangle
        # closed surface, odd labels a 12-triangle box
        from cosypose_b200 import synthetic as syn
        from cosypose_b200.rendering import CudaRasterizer, RenderMeshTable
        v, f, col = syn.make_render_meshes(len(job.labels), subdiv=5)
        views = CudaRasterizer(eng, RenderMeshTable(job.labels, v, f, col))
    else:
        views = PreRenderedViews([job.views_c, job.views_r], BSZ, device=eng.device)
    coarse = PosePredictor(eng, 0, views, mesh_db).load_state_dict(state_dict(0))
    refiner = PosePredictor(eng, 1, views, mesh_db).load_state_dict(state_dict(1))
    return CoarseRefinePosePredictor(coarse, refiner, bsz_objects=BSZ), eng, views, mesh_db


def run_single_view(args, cfg, rank, local_rank, world):
    import torch.distributed as dist
    from cosypose_b200.rendering import PreRenderedViews
    from cosypose_b200.utils import tensor_collection as tc
    c = CONFIGS[cfg]
    dev = torch.device('cuda', local_rank)
    job = SingleViewJob(cfg, rank, world)
    engine_views = args.views == 'engine'
    pred, eng, views, mesh_db = build_predictor(job, local_rank, args.views)
    if world > 1:
        eng.nccl_init()
    infos = job.infos()
    # device-resident global frame buffer: only this rank's frames are ever read (its hypotheses' batch_im_id)
    d_images = torch.zeros((job.n_frames, 3, 480, 640), dtype=torch.float32, device=dev)
    d_images[job.f0:job.f1] = job.images_local.to(dev)
    d_K, d_boxes = job.K.to(dev), job.boxes.to(dev)
    det = tc.PandasTensorCollection(infos=infos, bboxes=d_boxes)
    mv = None
    if c.get('matching'):
        from cosypose_b200.integrated.multiview_predictor import MultiviewScenePredictor
        from cosypose_b200.multiview.ransac import multiview_candidate_matching
        mv = MultiviewScenePredictor(mesh_db, engine=None, device=dev)

    def finish(final):
        """config 2: multiview candidate matching of every scene on the gathered candidates (replicated)."""
        if mv is None:
            return final.poses
        out = []
        for sid in range(int(job.scene_ids.max()) + 1):
            ids = np.where(job.scene_ids == sid)[0]
            cand = final[ids]
            cand.infos = cand.infos.reset_index(drop=True)
            out.append(multiview_candidate_matching(candidates=cand, mesh_db=mv.mesh_db_ransac, n_ransac_iter=2000,
                                                    dist_threshold=0.02))
        return final.poses

    def step_resident():
        if not engine_views:
            views.reset()
        final, _ = pred.get_predictions(d_images, d_K, detections=det, n_coarse_iterations=c['n_coarse'],
                                        n_refiner_iterations=c['n_refine'], shard=world > 1)
        return finish(final)

    # e2e leg: pinned host copies of this rank's inputs; the device landing buffers are allocated once, as a serving
    # loop would.  Views travel as uint8 NHWC, the reference renderer's native output (bullet_batch_renderer.py:70-83).
    h_images = job.images_local.pin_memory()
    h_K, h_boxes = job.K.pin_memory(), job.boxes.pin_memory()
    h_views = [] if engine_views else [
        (v[:, s0:s0 + BSZ] * 255).round().to(torch.uint8).permute(0, 1, 3, 4, 2).contiguous().pin_memory()
        for v in (job.views_c, job.views_r) for s0 in range(0, job.n_local, BSZ)]   # one stack per chunk, call order
    h_out = torch.empty((job.n, 4, 4), dtype=torch.float32).pin_memory()
    h2d = h_images.numel() * 4 + h_K.numel() * 4 + h_boxes.numel() * 4 + sum(v.numel() for v in h_views)
    d2h = h_out.numel() * 4
    # Two sets of landing buffers and a copy stream: the inputs of step i+1 are uploaded while step i computes (what a
    # serving loop does); every timed step still contains one full host->device copy of a step's inputs and one
    # device->host read of its result.
    class Landing:
        def __init__(self):
            self.images = torch.zeros_like(d_images)
            self.K, self.boxes = torch.empty_like(d_K), torch.empty_like(d_boxes)
            self.views = [torch.empty(v.shape, dtype=torch.uint8, device=dev) for v in h_views]
            self.rv = views if engine_views else PreRenderedViews.from_chunks(self.views, BSZ)
            self.det = tc.PandasTensorCollection(infos=infos, bboxes=self.boxes)
            self.ready, self.done = torch.cuda.Event(), torch.cuda.Event()
            self.done.record()

    sets = [Landing(), Landing()]
    copy_stream = torch.cuda.Stream(device=dev)
    e2e_state = {'i': 0}

    def upload(s):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(s.done)                 # the step that last read these buffers has finished
            s.images[job.f0:job.f1].copy_(h_images, non_blocking=True)
            s.K.copy_(h_K, non_blocking=True)
            s.boxes.copy_(h_boxes, non_blocking=True)
            for dst, src in zip(s.views, h_views):
                dst.copy_(src, non_blocking=True)
            s.ready.record(copy_stream)

    upload(sets[0])

    def step_e2e():
        i = e2e_state['i']
        e2e_state['i'] = i + 1
        s = sets[i % 2]
        cur = torch.cuda.current_stream()
        cur.wait_event(s.ready)
        if not engine_views:
            s.rv.reset()
        pred.coarse_model.renderer = s.rv
        pred.refiner_model.renderer = s.rv
        final, _ = pred.get_predictions(s.images, s.K, detections=s.det, n_coarse_iterations=c['n_coarse'],
                                        n_refiner_iterations=c['n_refine'], shard=world > 1)
        poses = finish(final)
        s.done.record(cur)
        upload(sets[(i + 1) % 2])                           # next step's inputs, under this step's compute
        h_out.copy_(poses, non_blocking=True)
        cur.synchronize()                                   # the step's result is on the host
        return poses

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            out = fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms, out

    for _ in range(args.warmup):
        step_resident()
    eng.profile_read(reset=True)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ms_total, out = timed(step_resident, args.steps)
    clocks = sampler.stop() if sampler else None
    counts = eng.profile_read(reset=True)
    launches = sum(n for n, _ in counts.values())

    # per-category device time (separate pass: event bracketing perturbs the pipeline and disables graph replay)
    eng.profile_enable(True)
    barrier()
    prof_steps = max(1, min(args.steps, 5))
    for _ in range(prof_steps):
        step_resident()
    torch.cuda.synchronize()
    prof = eng.profile_read(reset=True)
    eng.profile_enable(False)

    for _ in range(2):
        step_e2e()
    ms_e2e, _ = timed(step_e2e, args.steps)

    if rank == 0:
        hyps = job.n * args.steps
        fwd_per_step_rank = job.n_local * (c['n_coarse'] + c['n_refine'])
        peak, peak_src = measured_peaks()
        line = dict(
            metric=metric_name(cfg), value=hyps / (ms_total * 1e-3), unit=UNIT, n_gpus=world, steps=args.steps,
            warmup=args.warmup, ms_per_step=ms_total / args.steps, higher_is_better=True, scaling=c['scaling'],
            vs_baseline=None, dtype='f32', data='synthetic',
            config={'workload': f"{c['name']}: {c['text']}", 'hypotheses_total': job.n, 'hypotheses_per_gpu': job.n_local,
                    'forwards_per_hypothesis': c['n_coarse'] + c['n_refine'],
                    'l2': 'inputs larger than L2: %.0f MB of views + %.1f GB of activations stream per step and GPU'
                          % (job.n_local * (c['n_coarse'] + c['n_refine']) * 0.92, fwd_per_step_rank * 0.0244),
                    'collective': ('one ncclAllGather of %d floats per hypothesis and step (cosyb200_allgather_candidates)'
                                   % ((c['n_coarse'] + c['n_refine']) * 49)) if world > 1 else 'none',
                    'rank_data': 'rank-distinct frames, detections and views (shards of one global table)',
                    'views': ('rasterised by the engine inside every iteration (20 480-triangle surfaces and 12-triangle boxes); '
                              'the CPU baseline consumes pre-generated views') if engine_views else 'pre-rendered (uint8 NHWC on the e2e leg)'},
            clocks=clocks,
            e2e={'value': hyps / (ms_e2e * 1e-3), 'unit': UNIT, 'h2d_bytes_per_step': int(h2d),
                 'd2h_bytes_per_step': int(d2h), 'ms_per_step': ms_e2e / args.steps},
            gpu_launches=int(launches),
            roofline=roofline_object(prof, prof_steps, fwd_per_step_rank, peak, peak_src, ms_total / args.steps),
        )
        if world == 1 and not args.no_cpu_baseline:
            line['cpu_baseline'] = cpu_baseline(cfg, args.cpu_sample)
        print(json.dumps(line))


def run_multiview(args, rank, local_rank, world):
    """configs[3]: one scene per rank (replicas only: a scene is one matching problem and one view group)."""
    import torch.distributed as dist
    from helpers import Scene
    from cosypose_b200.integrated.multiview_predictor import MultiviewScenePredictor
    dev = torch.device('cuda', local_rank)
    sc = Scene(8, 16, 21, (1,), True, seed=rank)
    mv = MultiviewScenePredictor(sc.mesh_db(), device=dev)
    cands_dev, cams_dev = sc.candidates(dev), sc.cameras(dev)
    h_poses, h_K = sc.poses.pin_memory(), sc.K.pin_memory()

    def step_resident():
        return mv.predict_scene_state(cands_dev, cams_dev, ransac_n_iter=2000, ransac_dist_threshold=0.02, ba_n_iter=2)

    def step_e2e():
        cands_dev.poses.copy_(h_poses, non_blocking=True)
        cams_dev.K.copy_(h_K, non_blocking=True)
        out = mv.predict_scene_state(cands_dev, cams_dev, ransac_n_iter=2000, ransac_dist_threshold=0.02, ba_n_iter=2)
        out['scene/objects'].TWO.cpu()
        return out

    def timed(fn, steps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            out = fn()
        e1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms, out

    for _ in range(args.warmup):
        step_resident()
    eng = mv.engine
    eng.profile_read(reset=True)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ms_total, out = timed(step_resident, args.steps)
    clocks = sampler.stop() if sampler else None
    launches = sum(n for n, _ in eng.profile_read(reset=True).values())
    eng.profile_enable(True)
    for _ in range(3):
        step_resident()
    torch.cuda.synchronize()
    prof = eng.profile_read(reset=True)
    eng.profile_enable(False)
    ms_e2e, _ = timed(step_e2e, args.steps)
    if rank == 0:
        n_rows, n_seeds = 215040, 13440
        ransac_ms = prof['ransac'][1] / 3
        peak, peak_src = measured_peaks()
        line = dict(
            metric=metric_name(3), value=world * args.steps / (ms_total * 1e-3), unit='scenes/s', n_gpus=world,
            steps=args.steps, warmup=args.warmup, ms_per_step=ms_total / args.steps, higher_is_better=True,
            scaling='weak', vs_baseline=None, dtype='f32', data='synthetic',
            config={'workload': f"configs[3]: {CONFIGS[3]['text']}", 'seeds': n_seeds, 'scored_rows': n_rows,
                    'matched_candidates': int(len(out['ba_input'])), 'collective': 'none (replicas only)',
                    'l2': 'working set 42 MB of gathered poses per scoring launch, L2 resident by design'},
            clocks=clocks,
            e2e={'value': world * args.steps / (ms_e2e * 1e-3), 'unit': 'scenes/s',
                 'h2d_bytes_per_step': int(h_poses.numel() * 4 + h_K.numel() * 4), 'd2h_bytes_per_step': 16 * 16 * 4,
                 'ms_per_step': ms_e2e / args.steps},
            gpu_launches=int(launches),
            roofline={'bound': 'hbm', 'kernel': 'k_ransac_score + k_ransac_models + k_ba_* (device time of every launch '
                                                 'of the scene, CUDA events on the launching stream)',
                      'achieved': (n_rows * (3 * 64 + 4) + n_seeds * 5 * 64) / (ransac_ms * 1e-3) / 1e9 if ransac_ms > 0 else None,
                      'peak': peak, 'unit': 'GB/s', 'peak_source': peak_src,
                      'frac': ((n_rows * (3 * 64 + 4) + n_seeds * 5 * 64) / (ransac_ms * 1e-3) / 1e9 / peak) if ransac_ms > 0 else None,
                      'traffic': None, 'device_ms_per_scene': ransac_ms, 'rows_per_s': n_rows / (ransac_ms * 1e-3) if ransac_ms > 0 else None,
                      'note': 'latency / ALU bound by design (SURVEY.md 8d: 46 MB of algorithmic traffic per scene); the '
                              'step is dominated by host-side integer stages (seed enumeration, component analysis)'},
        )
        if world == 1 and not args.no_cpu_baseline:
            line['cpu_baseline'] = cpu_baseline(3, 0)
        print(json.dumps(line))


# ---- CPU arm: the unmodified reference on the host cores ----------------------------------------------------------
def _reference_predictor(cfg, n_hyp):
    """The reference's own CoarseRefinePosePredictor (baseline/_ref) on the first `n_hyp` hypotheses of the workload."""
    sys.path.insert(0, str(ROOT / 'baseline'))
    sys.path.insert(0, str(ROOT / 'tests' / 'golden'))
    import ref_harness
    cosypose = ref_harness.import_reference(force_cpu=True, use_installed=not ref_harness.available())
    import cosypose.utils.tensor_collection as rtc
    from cosypose.integrated.pose_predictor import CoarseRefinePosePredictor
    from make_golden import _RefRenderer, ref_mesh_db, ref_pose_model
    from cosypose_b200 import synthetic as syn
    c = CONFIGS[cfg]
    labels = syn.make_labels(c['labels'])
    points, sym, n_sym = syn.make_mesh_tables(c['labels'], sym_counts=c['sym'])
    mesh_db = ref_mesh_db(cosypose, labels, points, sym, n_sym)
    sd_c, sd_r = syn.make_pose_state_dict(0), syn.make_pose_state_dict(1)
    frames = max(1, -(-n_hyp // c['dets']))
    boxes, label_ids, im_ids = syn.make_detections(frames, c['dets'], c['labels'])
    boxes, label_ids, im_ids = boxes[:n_hyp], label_ids[:n_hyp], im_ids[:n_hyp]
    images, K = syn.make_images(frames), syn.make_camera_K(frames)
    views = [syn.make_renders(c['n_coarse'], n_hyp, seed=11), syn.make_renders(c['n_refine'], n_hyp, seed=12)]
    renderer = _RefRenderer(views, BSZ)
    pred = CoarseRefinePosePredictor(ref_pose_model(cosypose, sd_c, renderer, mesh_db),
                                     ref_pose_model(cosypose, sd_r, renderer, mesh_db), bsz_objects=BSZ)
    infos = pd.DataFrame(dict(label=[labels[i] for i in label_ids], batch_im_id=im_ids, score=np.ones(n_hyp)))
    det = rtc.PandasTensorCollection(infos=infos, bboxes=boxes)

    def step():
        renderer.i = 0
        with torch.no_grad():
            return pred.get_predictions(images, K, detections=det, n_coarse_iterations=c['n_coarse'],
                                        n_refiner_iterations=c['n_refine'])
    return step


def _port_step(cfg, n_hyp):
    from helpers import state_dict
    from oracle import pose_oracle as po
    from cosypose_b200 import synthetic as syn
    c = CONFIGS[cfg]
    points, sym, n_sym = syn.make_mesh_tables(c['labels'], sym_counts=c['sym'])
    frames = max(1, -(-n_hyp // c['dets']))
    boxes, label_ids, im_ids = syn.make_detections(frames, c['dets'], c['labels'])
    boxes, label_ids, im_ids = boxes[:n_hyp], label_ids[:n_hyp], im_ids[:n_hyp]
    images, K = syn.make_images(frames), syn.make_camera_K(frames)
    vc, vr = syn.make_renders(c['n_coarse'], n_hyp, seed=11), syn.make_renders(c['n_refine'], n_hyp, seed=12)
    sd_c, sd_r = state_dict(0), state_dict(1)

    def step():
        return po.coarse_refine_predictions(images, K, boxes, label_ids, im_ids, sd_c, sd_r, points,
                                            lambda stage, it, sl, T, Kc: (vc if stage == 'coarse' else vr)[it, sl],
                                            c['n_coarse'], c['n_refine'], bsz_objects=BSZ)
    return step


def _reference_available():
    return (ROOT / 'baseline' / '_ref' / 'cosypose').exists() or Path('/root/reference/cosypose').exists()


def _time_cpu(step, n_hyp, steps, warmup, threads):
    torch.set_num_threads(threads)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return n_hyp * steps / dt, dt / steps


def _reference_multiview_step():
    """The reference's MultiviewScenePredictor.predict_scene_state on the configs[3] scene, on CPU."""
    sys.path.insert(0, str(ROOT / 'baseline'))
    sys.path.insert(0, str(ROOT / 'tests' / 'golden'))
    import ref_harness
    cosypose = ref_harness.import_reference(force_cpu=True, use_installed=not ref_harness.available())
    import cosypose.utils.tensor_collection as rtc
    from cosypose.integrated.multiview_predictor import MultiviewScenePredictor
    from cosypose.lib3d.mesh_ops import get_meshes_bounding_boxes
    from make_golden import ref_mesh_db
    from cosypose_b200 import synthetic as syn
    labels = syn.make_labels(21)
    points, sym, n_sym = syn.make_mesh_tables(21, n_points=64, sym_counts=(1,))
    mesh_db = ref_mesh_db(cosypose, labels, get_meshes_bounding_boxes(points), sym, n_sym)
    scene = syn.make_multiview_scene(8, 16, 21, seed=0, unique_labels=True)
    infos = pd.DataFrame(dict(view_id=scene['view_ids'], label=[labels[i] for i in scene['label_ids']],
                              score=scene['scores'], scene_id=0, group_id=0, batch_im_id=scene['view_ids']))
    cands = rtc.PandasTensorCollection(infos=infos, poses=scene['poses'])
    cams = rtc.PandasTensorCollection(infos=pd.DataFrame(dict(view_id=np.arange(8), scene_id=0, batch_im_id=np.arange(8))),
                                      K=scene['K'], TWC=scene['TWC'])
    pred = MultiviewScenePredictor.__new__(MultiviewScenePredictor)
    pred.mesh_db_ransac = mesh_db
    pred.mesh_db_ba = mesh_db
    return lambda: pred.predict_scene_state(cands, cams, ransac_n_iter=2000, ransac_dist_threshold=0.02, ba_n_iter=2)


def cpu_baseline(cfg, sample_hyps, steps=1, warmup=1):
    """Bounded CPU sample of the same workload: all host cores, plus a smaller 1-thread sample (the reference's own
    default, cosypose/__init__.py:2-3).  The reference prints on import: stdout carries the JSON line only."""
    import contextlib
    with contextlib.redirect_stdout(sys.stderr):
        return _cpu_baseline(cfg, sample_hyps, steps, warmup)


def _cpu_baseline(cfg, sample_hyps, steps=1, warmup=1):
    cores = os.cpu_count()
    if cfg == 3:
        if not _reference_available():
            return {'value': None, 'unit': 'scenes/s', 'cores': cores, 'kind': 'reference', 'sample': 'baseline/_ref absent'}
        step = _reference_multiview_step()
        v, dt = _time_cpu(step, 1, steps, warmup, cores)
        return {'value': v, 'unit': 'scenes/s', 'cores': cores, 'kind': 'reference',
                'sample': f'the same 8 x 16 scene, {steps} step(s) after {warmup} warm-up, reference '
                          f'MultiviewScenePredictor.predict_scene_state on CPU ({cores} threads), {dt:.2f} s per scene'}
    sample_hyps = sample_hyps or 16
    use_ref = _reference_available()
    make = _reference_predictor if use_ref else _port_step
    v, dt = _time_cpu(make(cfg, sample_hyps), sample_hyps, steps, warmup, cores)
    n1 = min(4, sample_hyps)
    v1, dt1 = _time_cpu(make(cfg, n1), n1, 1, 0, 1)
    torch.set_num_threads(cores)
    c = CONFIGS[cfg]
    what = ('unmodified reference (baseline/_ref) CoarseRefinePosePredictor.get_predictions' if use_ref
            else 'oracle/pose_oracle.py (port; baseline/_ref absent)')
    return {'value': v, 'unit': UNIT, 'cores': cores, 'kind': 'reference' if use_ref else 'port',
            'value_1thread': v1,
            'sample': f'{sample_hyps} hypotheses x ({c["n_coarse"]}+{c["n_refine"]}) iterations x {steps} step(s) of the '
                      f'same workload, {what}, torch CPU fp32: {cores} threads {dt:.1f} s per step; 1 thread on '
                      f'{n1} hypotheses {dt1:.1f} s'}


def run_reference(args, cfg, rank):
    """`--impl reference`: the reference's CPU implementation of the path on all host cores.  Each step is a bounded
    sample of the workload, sized from a probe so that the K + W steps end within a few minutes."""
    if rank != 0:
        return
    import contextlib
    with contextlib.redirect_stdout(sys.stderr):
        line = _run_reference(args, cfg)
    print(json.dumps(line))


def _run_reference(args, cfg):
    cores = os.cpu_count()
    if cfg == 3:
        cb = _cpu_baseline(3, 0, steps=max(1, args.steps), warmup=min(args.warmup, 1))
        value, unit, ms = cb['value'], 'scenes/s', 1e3 / cb['value'] if cb['value'] else None
    else:
        use_ref = _reference_available()
        make = _reference_predictor if use_ref else _port_step
        probe_rate, _ = _time_cpu(make(cfg, 8), 8, 1, 0, cores)
        budget_s = 240.0
        n_steps = max(1, args.steps) + min(args.warmup, 1)
        sample = 8
        for cand in (64, 32, 16):
            if cand / probe_rate * n_steps <= budget_s:
                sample = cand
                break
        value, dt = _time_cpu(make(cfg, sample), sample, max(1, args.steps), min(args.warmup, 1), cores)
        n1 = 4
        v1, dt1 = _time_cpu(make(cfg, n1), n1, 1, 0, 1)
        c = CONFIGS[cfg]
        cb = {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'reference' if use_ref else 'port', 'value_1thread': v1,
              'sample': f'{sample} hypotheses x ({c["n_coarse"]}+{c["n_refine"]}) iterations per step, {args.steps} steps, '
                        + ('unmodified reference (baseline/_ref) CoarseRefinePosePredictor.get_predictions'
                           if use_ref else 'oracle/pose_oracle.py (port; baseline/_ref absent)')
                        + f', torch CPU fp32, {cores} threads, {dt:.1f} s per step; 1 thread on {n1} hypotheses: {dt1:.1f} s'}
        unit, ms = UNIT, dt * 1e3
    c = CONFIGS[cfg]
    line = dict(metric=metric_name(cfg), value=value, unit=unit, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=ms, higher_is_better=True, scaling=c['scaling'], vs_baseline=None, dtype='f32',
                data='synthetic', impl='reference', config={'workload': f"{c['name']}: {c['text']}"},
                cpu_baseline=cb,
                e2e={'value': value, 'unit': unit, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0})
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--config', type=int, default=1, choices=[1, 2, 3, 4])
    ap.add_argument('--cpu-sample', type=int, default=16, help='hypotheses in the CPU baseline sample')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--views', default='prerendered', choices=['prerendered', 'engine'],
                    help="'prerendered' (BASELINE.json: renders pre-generated, the default) or 'engine': every iteration's views "
                         'drawn by the engine rasteriser inside the loop')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'ours' else args.warmup

    rank = int(os.environ.get('RANK', 0))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    if args.impl == 'reference':
        return run_reference(args, args.config, rank)

    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    if args.config == 3:
        run_multiview(args, rank, local_rank, world)
    else:
        run_single_view(args, args.config, rank, local_rank, world)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()

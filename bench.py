#!/usr/bin/env python
"""Benchmark of the refinement hot path (contract: task prompt section 4; metric: BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (`config.workload`): BASELINE.json configs[1] - a batch of 64 synthetic "YCB-V" crops
(8 frames of 640x480 x 8 detections, 21 labels), 1 coarse + 4 refiner iterations, random-init
BN-calibrated EfficientNet-B3 weights, views pre-rendered (the renderer is a side input).
A step = one `CoarseRefinePosePredictor.get_predictions` over the 64 hypotheses of one GPU
(5 backbone forwards per hypothesis).  With N GPUs every rank refines its own 64 hypotheses
(weak scaling, config 3 = 512 hypotheses on 8 GPUs) and the refined poses are collected with one
NCCL all-gather per step.

`value`  : hypotheses/s, inputs resident in HBM, CUDA-event timed over K steps, max over ranks.
`e2e`    : same metric through the public API with HOST inputs: per step the frames, detections
           and uint8 views are copied from pinned memory, the final poses are read back.
`--impl reference` : the oracle port of the reference's CPU PyTorch path (oracle/pose_oracle.py)
           on the host cores, on a bounded sample of the same workload.
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
import torch  # noqa: E402

METRIC = 'refined object-hypotheses/sec (1 coarse + 4 refine iters, 640x480)'
UNIT = 'hypotheses/s'
N_IMAGES, DETS, N_LABELS, N_COARSE, N_REFINE = 8, 8, 21, 1, 4
BSZ = N_IMAGES * DETS
ALGO_BYTES_PER_FORWARD = 6107136 * 4      # block-boundary activations, fp32
ALGO_FLOP_PER_FORWARD = 2 * 1456.6e6


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


# DRAM traffic of the tensor-core 1x1 kernel from the committed `ncu --set full` capture
# (profiles/r01b_ncu_full_blocks0to4.csv: the 7 k_pw_gemm_tc launches of blocks 0-4, 64 hypotheses):
# sum of dram__bytes_read.sum + dram__bytes_write.sum against the algorithmic bytes of the same launches.
NCU_GEMM_CAPTURE = dict(launches=7, dram_bytes=2313.1e6, algorithmic_bytes=2575.8e6,
                        source='profiles/r01b_ncu_full_blocks0to4.csv')
GEMM_CATS = ('expand_1x1', 'project_1x1', 'head_1x1')


def roofline_object(prof, prof_steps, fwd_per_step, peak, peak_src):
    """`roofline` of the JSON line.  Dominant kernel = k_pw_gemm_tc (every 1x1 convolution: expand, project,
    head; ~58 % of the step).  achieved = algorithmic bytes of its launches / their CUDA-event device time
    (engine profiling mode brackets every launch with events on the launching stream); `trunk` is the
    whole-network figure of SURVEY.md section 8(d) (block-boundary activations only, i.e. what a fully fused
    trunk would move); by_kernel_* give every category against its own one-kernel-per-stage byte count."""
    kbytes = kernel_algorithmic_bytes()
    ms = {c: prof[c][1] / prof_steps for c in prof}                 # device ms per step and category
    n_launch = {c: prof[c][0] / prof_steps for c in prof}
    by_kernel_gbs = {c: round(fwd_per_step * kbytes[c] / (ms[c] * 1e-3) / 1e9, 1) for c in kbytes if ms[c] > 0}
    gemm_ms = sum(ms[c] for c in GEMM_CATS)
    gemm_bytes = fwd_per_step * sum(kbytes[c] for c in GEMM_CATS)
    gemm_launches = sum(n_launch[c] for c in GEMM_CATS)
    achieved = gemm_bytes / (gemm_ms * 1e-3) / 1e9 if gemm_ms > 0 else None
    backbone_cats = ('stem', 'expand_1x1', 'depthwise', 'squeeze_excite', 'project_1x1', 'head_1x1', 'pool_fc_update')
    bb_ms = sum(ms[c] for c in backbone_cats)
    trunk = fwd_per_step * ALGO_BYTES_PER_FORWARD / (bb_ms * 1e-3) / 1e9 if bb_ms > 0 else None
    tot = sum(ms.values())
    per_launch = gemm_bytes / gemm_launches if gemm_launches else None
    ratio = NCU_GEMM_CAPTURE['dram_bytes'] / NCU_GEMM_CAPTURE['algorithmic_bytes']
    return {
        'bound': 'hbm', 'kernel': 'k_pw_gemm_tc (tcgen05 3xTF32 1x1 convolutions: expand + project + head)',
        'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak if achieved else None,
        'peak_source': peak_src,
        'traffic': per_launch * ratio if per_launch else None,
        'traffic_note': f"mean algorithmic bytes per launch x {ratio:.3f}, the DRAM/algorithmic ratio ncu measured on "
                        f"{NCU_GEMM_CAPTURE['launches']} launches ({NCU_GEMM_CAPTURE['source']})",
        'algorithmic_bytes_per_launch': per_launch, 'launches_per_step': gemm_launches,
        'launch_us': gemm_ms * 1e3 / gemm_launches if gemm_launches else None,
        'share_of_step': gemm_ms / tot if tot > 0 else None,
        'trunk': {'kernel': 'EfficientNet-B3 trunk forward (stem + 26 MBConv + head), all launches',
                  'achieved': trunk, 'frac': trunk / peak if trunk else None, 'ms_per_step': bb_ms,
                  'algorithmic_bytes_per_step': fwd_per_step * ALGO_BYTES_PER_FORWARD,
                  'tflops': fwd_per_step * ALGO_FLOP_PER_FORWARD / (bb_ms * 1e-3) / 1e12 if bb_ms > 0 else None},
        'by_kernel_gbs': by_kernel_gbs,
        'by_kernel_frac': {c: round(v / peak, 4) for c, v in by_kernel_gbs.items()},
        'by_kernel_ms_per_step': {c: round(v, 4) for c, v in ms.items() if v > 0},
        'by_kernel_share': {c: round(v / tot, 4) for c, v in ms.items() if v > 0},
    }


def workload_config(world):
    return {'workload': 'configs[1]: 64 synthetic YCB-V crops per GPU (8 frames 640x480 x 8 detections, '
                        '21 labels), 1 coarse + 4 refine iters, random-init BN-calibrated EfficientNet-B3, '
                        'pre-rendered views', 'hypotheses_per_gpu': BSZ, 'forwards_per_hypothesis': 5,
            'l2': 'inputs larger than L2: 295 MB of views + 1.3 GB of activations stream per step',
            'collective': 'one NCCL all-gather of [64,4,4] poses per step' if world > 1 else 'none'}


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


def cpu_reference(steps, warmup, sample_hyps, threads=None):
    """The oracle port of the reference's CPU path, timed on the host cores."""
    from helpers import Workload, state_dict
    from oracle import pose_oracle as po
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    n_img = max(1, sample_hyps // DETS)
    w = Workload(n_img, min(DETS, sample_hyps), N_LABELS, N_COARSE, N_REFINE)
    sd_c, sd_r = state_dict(0), state_dict(1)

    def step():
        return po.coarse_refine_predictions(w.images, w.K, w.boxes, w.label_ids, w.im_ids, sd_c, sd_r, w.points,
                                            w.oracle_render_fn(), N_COARSE, N_REFINE, bsz_objects=BSZ)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return dict(value=w.n * steps / dt, unit=UNIT, cores=threads, kind='port',
                sample=f'{w.n} hypotheses x (1+4) iterations x {steps} steps of the same workload, '
                       f'oracle/pose_oracle.py (torch CPU fp32, {threads} threads), {dt:.1f} s'), dt / steps


def run_reference(args, rank):
    if rank != 0:
        return
    cb, ms = cpu_reference(max(1, args.steps), min(args.warmup, 1), sample_hyps=16)
    line = dict(metric=METRIC, value=cb['value'], unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=ms * 1e3, higher_is_better=True, scaling='weak', vs_baseline=None, dtype='f32',
                data='synthetic', impl='reference',
                config=workload_config(args.gpus),
                cpu_baseline=cb,
                e2e={'value': cb['value'], 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0})
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--cpu-sample', type=int, default=16, help='hypotheses in the CPU baseline sample')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'ours' else args.warmup

    rank = int(os.environ.get('RANK', 0))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    if args.impl == 'reference':
        return run_reference(args, rank)

    import torch.distributed as dist
    from helpers import Workload, build_predictor
    from cosypose_b200.sharding import gather_poses
    from cosypose_b200.utils import tensor_collection as tc

    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    # every rank owns 64 hypotheses of the global batch (different seeds per rank would change
    # nothing in cost; the same shard keeps the parity reference identical)
    w = Workload(N_IMAGES, DETS, N_LABELS, N_COARSE, N_REFINE)
    pred, eng, views = build_predictor(w, local_rank, bsz_objects=BSZ)
    infos = w.infos()
    d_images, d_K, d_boxes = w.images.to(dev), w.K.to(dev), w.boxes.to(dev)
    det = tc.PandasTensorCollection(infos=infos, bboxes=d_boxes)

    def step_resident():
        views.reset()
        final, _ = pred.get_predictions(d_images, d_K, detections=det, n_coarse_iterations=N_COARSE,
                                        n_refiner_iterations=N_REFINE)
        return gather_poses(final.poses) if world > 1 else final.poses

    # host-side copies for the e2e leg: frames fp32, views uint8 NHWC as the reference's renderer
    # delivers them (bullet_batch_renderer.py:70-83), detections, intrinsics
    h_images = w.images.pin_memory()
    h_K, h_boxes = w.K.pin_memory(), w.boxes.pin_memory()
    h_views = [(v * 255).round().to(torch.uint8).permute(0, 1, 3, 4, 2).contiguous().pin_memory()
               for v in (w.views_c, w.views_r)]
    h_out = torch.empty((w.n * world, 4, 4), dtype=torch.float32).pin_memory()
    h2d = h_images.numel() * 4 + h_K.numel() * 4 + h_boxes.numel() * 4 + sum(v.numel() for v in h_views)
    d2h = h_out.numel() * 4 // world

    from cosypose_b200.rendering import PreRenderedViews

    # device-side landing buffers of the e2e leg are allocated once, as a serving loop would: every step copies
    # that step's frames, intrinsics, detections and views from pinned host memory into them
    e_images, e_K, e_boxes = torch.empty_like(d_images), torch.empty_like(d_K), torch.empty_like(d_boxes)
    e_views = [torch.empty(v.shape, dtype=torch.uint8, device=dev) for v in h_views]
    e_rv = PreRenderedViews.from_uint8(e_views, BSZ)
    e_det = tc.PandasTensorCollection(infos=infos, bboxes=e_boxes)

    def step_e2e():
        e_images.copy_(h_images, non_blocking=True)
        e_K.copy_(h_K, non_blocking=True)
        e_boxes.copy_(h_boxes, non_blocking=True)
        for dst, src in zip(e_views, h_views):
            dst.copy_(src, non_blocking=True)
        e_rv.reset()
        pred.coarse_model.renderer = e_rv
        pred.refiner_model.renderer = e_rv
        final, _ = pred.get_predictions(e_images, e_K, detections=e_det, n_coarse_iterations=N_COARSE,
                                        n_refiner_iterations=N_REFINE)
        poses = gather_poses(final.poses) if world > 1 else final.poses
        h_out[:poses.shape[0]].copy_(poses, non_blocking=True)
        torch.cuda.current_stream().synchronize()
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

    # per-category device time (separate pass: event bracketing perturbs the pipeline)
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
    pred.coarse_model.renderer = views
    pred.refiner_model.renderer = views

    if rank == 0:
        hyps = w.n * world * args.steps
        value = hyps / (ms_total * 1e-3)
        fwd_per_step = w.n * (N_COARSE + N_REFINE)
        peak, peak_src = measured_peaks()
        line = dict(
            metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
            ms_per_step=ms_total / args.steps, higher_is_better=True, scaling='weak', vs_baseline=None,
            dtype='f32', data='synthetic',
            config=workload_config(world),
            clocks=clocks,
            e2e={'value': hyps / (ms_e2e * 1e-3), 'unit': UNIT, 'h2d_bytes_per_step': int(h2d),
                 'd2h_bytes_per_step': int(d2h), 'ms_per_step': ms_e2e / args.steps},
            gpu_launches=int(launches),
            roofline=roofline_object(prof, prof_steps, fwd_per_step, peak, peak_src),
        )
        if world == 1 and not args.no_cpu_baseline:
            cb, _ = cpu_reference(1, 1, sample_hyps=args.cpu_sample)
            line['cpu_baseline'] = cb
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()

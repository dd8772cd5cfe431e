#!/usr/bin/env python
"""Benchmark of the extract hot path (BASELINE.json metric: extract frames/s + kernel HBM GB/s vs peak).

Workload (config.workload): BASELINE.json configs[1] -- preprocess + crop/rotate kernels only (no R-CNN) on a
30-minute synthetic session (54,000 Kinect-v2-shaped 512x424 int16 frames, ROI box 240x240, 80x80 crops,
1000-frame chunks, use_tracking=False; instance masks + keypoints given).  One "step" = one pass over the whole
session.  `value` = frames/s with the session resident in HBM; `e2e` = the same metric with HOST (pinned) frames,
masks and keypoints copied to the GPU and the results (crops, scalars, keypoint table, flips) copied back inside
the timed region.  Under torchrun every rank extracts its own session (weak scaling, no collective).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, 'oracle')):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = 'extract_frames_per_s'
UNIT = 'frames/s'


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--frames', type=int, default=54000, help='frames per session (per GPU)')
    ap.add_argument('--launch-chunks', type=int, default=6, help='1000-frame chunks processed per kernel launch')
    ap.add_argument('--pool-frames', type=int, default=1000, help='distinct synthetic frames generated on the host')
    ap.add_argument('--geometry', default='kinect_v2', choices=['kinect_v2', 'azure'])
    ap.add_argument('--cpu-frames-per-worker', type=int, default=250)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--rcnn-frames', type=int, default=2000, help='frames for the secondary full-extract (R-CNN) figure; 0 = skip')
    ap.add_argument('--rcnn-batch', type=int, default=500)
    return ap.parse_args()


def workload_config(args, geom, n_gpus):
    return {
        'workload': f'configs[1]: prep + clean + features + angles/flips/filter + scalars/keypoints + crop/rotate, no R-CNN, '
                    f'{args.frames}-frame synthetic session per GPU',
        'frame': f'{geom.width}x{geom.height} int16', 'roi_box': None, 'crop': list(geom.crop_size), 'chunk_size': 1000,
        'frames_per_session': args.frames, 'sessions': n_gpus, 'frames_per_launch': args.launch_chunks * 1000,
        'use_tracking': False, 'parallelism': f'chunk-sharded x{n_gpus}, no collective',
        'l2_policy': 'inputs (23 GB session) far larger than the 126 MB L2; every frame is read from HBM once per step',
    }


# ------------------------------------------------------------------------------------------------
# clocks sampler (B200_PROFILING.md "clocks line")
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
              'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
              'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.samples = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--id={self.gpu_index}', f'--query-gpu={self.FIELDS}',
                                          '--format=csv,noheader,nounits', '-lms', '20'], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.time(), line.strip()))

    def wait_first(self, timeout_s):
        t_end = time.time() + timeout_s
        while self.proc is not None and not self.samples and time.time() < t_end:
            time.sleep(0.01)

    def stop(self, t0, t1):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ts, line in self.samples:
            parts = [p.strip() for p in line.split(',')]
            if len(parts) < 8:
                continue
            try:
                mx.append(float(parts[2]))
                if t0 <= ts <= t1 + 0.1:
                    sm.append(float(parts[1]))
                    for name, val in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), parts[4:8]):
                        if val.lower().startswith('active'):
                            reasons.add(name)
            except ValueError:
                continue
        if not sm:   # timed region shorter than the sampling period: fall back to the samples since the warm-up began
            for ts, line in self.samples[1:] or self.samples:
                parts = [p.strip() for p in line.split(',')]
                try:
                    sm.append(float(parts[1]))
                except (ValueError, IndexError):
                    pass
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port of the reference path on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_reference(args, workers=None):
    import cpu_baseline
    workers = workers or max(1, min(os.cpu_count() or 1, 64))
    frames, wall = cpu_baseline.run(workers, args.cpu_frames_per_worker, args.geometry)
    return {'value': frames / wall, 'unit': UNIT, 'cores': workers, 'kind': 'port',
            'sample': f'{workers} worker processes x {args.cpu_frames_per_worker} frames of the same synthetic workload '
                      f'through oracle/extract_oracle.py (numpy + OpenCV, 1 OpenCV thread per worker)',
            'frames': frames, 'wall_s': wall}


def run_reference_arm(args):
    from moseq2_detectron_extract_b200 import synthetic
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    geom = getattr(synthetic.SessionGeometry, args.geometry)()
    times = []
    base = None
    for i in range(args.warmup + args.steps):
        base = cpu_reference(args)
        if i >= args.warmup:
            times.append((base['frames'], base['wall_s']))
    frames = sum(t[0] for t in times)
    wall = sum(t[1] for t in times)
    value = frames / wall
    cfg = workload_config(args, geom, args.gpus)
    cfg['roi_box'] = [int(v) for v in synthetic.roi_bbox(synthetic.make_roi(geom))]
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': 1e3 * wall / max(args.steps, 1), 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'u8/int64/f64 (numpy + OpenCV on the host)', 'data': 'synthetic',
        'config': cfg,
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': base['cores'], 'kind': 'port', 'sample': base['sample']},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))

    # CPU baseline first (rank 0, N=1 only): before CUDA is initialised in this process, workers are spawned
    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_base = cpu_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    from moseq2_detectron_extract_b200 import _dev, _lib, synthetic
    from moseq2_detectron_extract_b200.engine import ChunkEngine
    from moseq2_detectron_extract_b200.proc import proc as P

    torch.cuda.set_device(local_rank)
    _dev.require_cuda()
    from moseq2_detectron_extract_b200.shard import bind_to_gpu_numa_node
    numa = bind_to_gpu_numa_node(local_rank)      # before any pinned allocation: host buffers land on the GPU's NUMA node
    if world > 1:
        # NCCL carries only the barrier and the max-over-ranks of two timing scalars (no data-path collective);
        # keep its banner off stdout so that the single JSON line is the only output
        os.environ['NCCL_DEBUG'] = 'WARN'
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    geom = getattr(synthetic.SessionGeometry, args.geometry)()
    cfg = synthetic.default_config(geom)
    roi, bg = synthetic.make_roi(geom), synthetic.make_background(geom)
    y0, x0, y1, x1 = synthetic.roi_bbox(roi)
    h, w = y1 - y0, x1 - x0
    H, W = geom.height, geom.width
    chunk = 1000
    n_frames = args.frames
    launch = args.launch_chunks * chunk

    # ---- synthetic session: `pool` distinct frames generated on the host, cycled (with a per-chunk roll) on the GPU
    pool = synthetic.generate_chunk(args.pool_frames, seed=rank, geom=geom, t0=0)
    pool_frames = torch.from_numpy(pool.frames).pin_memory()
    pool_masks = torch.from_numpy(pool.masks).pin_memory()
    pool_kpts = torch.from_numpy(pool.keypoints).pin_memory()
    d_pool_f, d_pool_m, d_pool_k = pool_frames.cuda(), pool_masks.cuda(), pool_kpts.cuda()
    idx = torch.cat([(torch.arange(min(chunk, n_frames - c), device='cuda') + 37 * (c // chunk)) % args.pool_frames
                     for c in range(0, n_frames, chunk)])
    frames = torch.empty((n_frames, H, W), dtype=torch.int16, device='cuda')
    for s in range(0, n_frames, 2000):
        frames[s:s + 2000] = d_pool_f[idx[s:s + 2000]]
    masks = d_pool_m[idx].contiguous()
    kpts = d_pool_k[idx].contiguous()
    del d_pool_f, d_pool_m, d_pool_k
    bg_d, roi_d = _dev.as_device(bg), _dev.as_device(roi.astype(np.uint8))
    prep_buf = _dev.empty((launch, h, w), torch.uint8)
    invalid = _dev.empty((launch,), torch.int32)
    engine = ChunkEngine()
    kw = dict(chunk_size=chunk, min_height=cfg['min_height'], max_height=cfg['max_height'], true_depth=cfg['true_depth'],
              crop_size=cfg['crop_size'])
    flags = _lib.MSQ_PREP_HAS_VMIN | _lib.MSQ_PREP_HAS_VMAX

    def resident_step():
        st = _dev.stream()
        for s in range(0, n_frames, launch):
            n = min(launch, n_frames - s)
            _lib.call('msq_prep_frames', _dev.ptr(frames[s:s + n]), n, H, W, _dev.ptr(bg_d), _lib.MSQ_BG_F32, _dev.ptr(roi_d),
                      y0, x0, h, w, float(cfg['min_height']), float(cfg['max_height']), flags, _dev.ptr(prep_buf),
                      _dev.ptr(invalid), None, st)
            engine.extract(prep_buf[:n], masks[s:s + n], kpts[s:s + n], **kw)

    # ---- device-resident throughput ------------------------------------------------------------------
    # the clock sampler is started BEFORE the warm-up and must have delivered a line before timing starts (nvidia-smi
    # needs a few hundred ms to come up; the timed region can be shorter than that), so its samples see the GPU under load
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        sampler.wait_first(3.0)
    barrier()
    for _ in range(args.warmup):
        resident_step()
    barrier()
    launches_before = sum(_lib.kernel_launches().values())
    _lib.kernel_timing(True)
    barrier()
    t_wall0 = time.time()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        resident_step()
    ev1.record()
    barrier()
    t_wall1 = time.time()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    _lib.kernel_timing(False)
    ktimes = _lib.kernel_timing_collect()
    gpu_launches = sum(_lib.kernel_launches().values()) - launches_before
    assert int(invalid.sum().item()) == 0

    # ---- end to end: pinned host inputs -> GPU -> pinned host results, 3-stage stream pipeline ----------------
    e2e_ms, h2d_bytes, d2h_bytes = None, 0, 0
    e2e_copy_ms = e2e_zc_ms = None
    e2e_mode = None
    if not args.no_e2e:
        n_e2e_chunks = (n_frames + chunk - 1) // chunk
        copy_in, compute, copy_out, prep_st = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
        slots = 2
        in_f = [_dev.empty((chunk, H, W), torch.int16) for _ in range(slots)]
        in_m = [_dev.empty((chunk, h, w), torch.uint8) for _ in range(slots)]
        in_k = [_dev.empty((chunk, 8, 3), torch.float32) for _ in range(slots)]
        engines = [ChunkEngine() for _ in range(slots)]
        preps = [_dev.empty((chunk, h, w), torch.uint8) for _ in range(slots)]
        invs = [_dev.empty((chunk,), torch.int32) for _ in range(slots)]
        cw, ch = cfg['crop_size']
        host_out = [{'depth_crops': torch.empty((chunk, ch, cw), dtype=torch.uint8).pin_memory(),
                     'mask_crops': torch.empty((chunk, ch, cw), dtype=torch.uint8).pin_memory(),
                     'scalars': torch.empty((_lib.NUM_SCALARS, chunk), dtype=torch.float64).pin_memory(),
                     'kpt_cols': torch.empty((_lib.NUM_KPT_COLS, chunk), dtype=torch.float64).pin_memory(),
                     'flips': torch.empty((chunk,), dtype=torch.uint8).pin_memory(),
                     'invalid': torch.empty((chunk,), dtype=torch.int32).pin_memory()} for _ in range(slots)]
        h2d_chunk = pool_frames[:chunk].numel() * 2 + pool_masks[:chunk].numel() + pool_kpts[:chunk].numel() * 4
        d2h_chunk = sum(v.numel() * v.element_size() for v in host_out[0].values())

        def e2e_step(zero_copy):
            """One pass over the session from pinned host buffers.  zero_copy=True: the prep kernel reads the ROI box
            of the raw frames straight out of pinned host memory (UVA), so only the bytes the path needs cross
            PCIe; masks / keypoints still go through cudaMemcpyAsync.  zero_copy=False: whole frames are copied."""
            ev_h2d = [None] * slots
            ev_comp = [None] * slots
            ev_d2h = [None] * slots
            for c in range(n_e2e_chunks):
                b = c % slots
                with torch.cuda.stream(copy_in):
                    if ev_comp[b] is not None:
                        copy_in.wait_event(ev_comp[b])          # input slot consumed by the previous user
                    if not zero_copy:
                        in_f[b].copy_(pool_frames[:chunk], non_blocking=True)
                    in_m[b].copy_(pool_masks[:chunk], non_blocking=True)
                    in_k[b].copy_(pool_kpts[:chunk], non_blocking=True)
                    ev_h2d[b] = torch.cuda.Event()
                    ev_h2d[b].record(copy_in)
                # prep is PCIe-bound in zero-copy mode (it pulls the ROI box from host memory) and needs few SMs; on its
                # own stream it overlaps the SM-bound extract kernels of the previous chunk
                with torch.cuda.stream(prep_st):
                    if not zero_copy:
                        prep_st.wait_event(ev_h2d[b])
                    if ev_comp[b] is not None:
                        prep_st.wait_event(ev_comp[b])          # preps[b] consumed by the previous user of the slot
                    src = pool_frames[:chunk] if zero_copy else in_f[b]
                    _lib.call('msq_prep_frames', _dev.ptr(src), chunk, H, W, _dev.ptr(bg_d), _lib.MSQ_BG_F32,
                              _dev.ptr(roi_d), y0, x0, h, w, float(cfg['min_height']), float(cfg['max_height']), flags,
                              _dev.ptr(preps[b]), _dev.ptr(invs[b]), None, _dev.stream())
                    ev_prep = torch.cuda.Event()
                    ev_prep.record(prep_st)
                with torch.cuda.stream(compute):
                    compute.wait_event(ev_h2d[b])
                    compute.wait_event(ev_prep)
                    if ev_d2h[b] is not None:
                        compute.wait_event(ev_d2h[b])           # output slot drained
                    res = engines[b].extract(preps[b], in_m[b], in_k[b], **kw)
                    ev_comp[b] = torch.cuda.Event()
                    ev_comp[b].record(compute)
                with torch.cuda.stream(copy_out):
                    copy_out.wait_event(ev_comp[b])
                    for key in ('depth_crops', 'mask_crops', 'scalars', 'kpt_cols', 'flips'):
                        host_out[b][key].copy_(res[key], non_blocking=True)
                    host_out[b]['invalid'].copy_(invs[b], non_blocking=True)
                    ev_d2h[b] = torch.cuda.Event()
                    ev_d2h[b].record(copy_out)
            torch.cuda.current_stream().wait_stream(copy_out)
            torch.cuda.current_stream().wait_stream(compute)
            torch.cuda.current_stream().wait_stream(prep_st)

        def time_e2e(zero_copy):
            for _ in range(min(args.warmup, 3)):
                e2e_step(zero_copy)
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.steps):
                e2e_step(zero_copy)
            e1.record()
            barrier()
            assert float(host_out[0]['scalars'][6].sum()) > 0          # area_px really came back
            return e0.elapsed_time(e1)

        e2e_copy_ms = time_e2e(False)
        e2e_zc_ms = time_e2e(True)
        roi_bytes_chunk = chunk * h * w * 2          # int16 ROI-box pixels the prep kernel pulls over PCIe
        small_chunk = pool_masks[:chunk].numel() + pool_kpts[:chunk].numel() * 4
        if e2e_zc_ms <= e2e_copy_ms:
            e2e_ms, e2e_mode, h2d_chunk = e2e_zc_ms, 'zero-copy', roi_bytes_chunk + small_chunk
        else:
            e2e_ms, e2e_mode, h2d_chunk = e2e_copy_ms, 'copy', pool_frames[:chunk].numel() * 2 + small_chunk
        h2d_bytes, d2h_bytes = h2d_chunk * n_e2e_chunks, d2h_chunk * n_e2e_chunks
    # ---- secondary figures (rank 0, not part of `value`): in-painting cost and the full extract with the R-CNN ------
    extras = {}
    if rank == 0 and world == 1:
        try:
            extras.update(secondary_figures(args, geom, cfg, roi, bg))
        except Exception as exc:       # never let a secondary figure break the contract line
            extras['secondary_error'] = repr(exc)[:300]

    # ---- reduce over ranks (max time) --------------------------------------------------------------------
    times = torch.tensor([ms, e2e_ms if e2e_ms is not None else 0.0], dtype=torch.float64, device='cuda')
    launches_all = torch.tensor([float(gpu_launches)], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
        dist.all_reduce(launches_all, op=dist.ReduceOp.SUM)
    ms, e2e_ms_max = float(times[0]), float(times[1])
    gpu_launches = int(launches_all[0])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    total_frames = n_frames * world * args.steps
    value = total_frames / (ms * 1e-3)

    # ---- roofline of the dominant kernel (largest share of device time in the timed region) -----------------
    peaks_path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    else:
        peak, peak_src = 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'
    A, C = h * w, cfg['crop_size'][0] * cfg['crop_size'][1]
    # algorithmic bytes per frame: only the ROI box of the raw frame is ever needed
    bytes_per_frame = {'prep_frames': 3 * A, 'clean_frames': 2 * A, 'frame_features': 2 * A + 40, 'masked_sums': 2 * A,
                       'scalars_keypoints': 113 * 8 + 24 * 4 + 8 * 1 + 64, 'crop_rotate': 2 * 2 * C + 2 * C,
                       'angles_flips_filter': 8 * 8 + 96 + 9}
    kernel_ms = {k: v[0] for k, v in ktimes.items() if v[1] > 0}
    dominant = max(kernel_ms, key=kernel_ms.get)
    total_kernel_ms = sum(kernel_ms.values())
    d_ms, d_cnt = ktimes[dominant]
    frames_per_launch_avg = n_frames * args.steps / d_cnt
    achieved = bytes_per_frame[dominant] * frames_per_launch_avg / (d_ms / d_cnt * 1e-3) / 1e9
    roofline = {
        'bound': 'hbm', 'kernel': dominant, 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
        'traffic': None, 'peak_source': peak_src, 'avg_launch_ms': d_ms / d_cnt,
        'algorithmic_bytes_per_frame': bytes_per_frame[dominant], 'frames_per_launch': frames_per_launch_avg,
        'share_of_kernel_time': d_ms / total_kernel_ms,
        'per_kernel': {k: {'ms_total': v, 'share': v / total_kernel_ms, 'launches': ktimes[k][1],
                           'GBps': bytes_per_frame.get(k, 0) * n_frames * args.steps / (v * 1e-3) / 1e9}
                       for k, v in kernel_ms.items()},
    }
    traffic_file = os.path.join(ROOT, 'profiles', 'dominant_kernel_traffic.json')
    if os.path.exists(traffic_file):
        try:
            t = json.load(open(traffic_file))
            if t.get('kernel') == dominant:
                roofline['traffic'] = t.get('dram_bytes_per_launch')
                roofline['traffic_source'] = t.get('source')
        except Exception:
            pass

    cfg_out = workload_config(args, geom, world)
    cfg_out['roi_box'] = [y0, x0, y1, x1]
    cfg_out['host_numa_binding'] = numa if numa else 'unavailable'      # rank 0's; every rank binds to its own GPU's node
    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'u8 / int16 (pixels), int64 (moments), f64 (features)', 'data': 'synthetic', 'config': cfg_out,
        'clocks': clocks, 'gpu_launches': int(gpu_launches),
        'e2e': ({'value': total_frames / (e2e_ms_max * 1e-3), 'unit': UNIT, 'h2d_bytes_per_step': int(h2d_bytes),
                 'd2h_bytes_per_step': int(d2h_bytes), 'ms_per_step': e2e_ms_max / args.steps,
                 'mode': e2e_mode,
                 'frames_per_s_full_frame_copy': n_frames * args.steps / (e2e_copy_ms * 1e-3),
                 'frames_per_s_zero_copy_roi': n_frames * args.steps / (e2e_zc_ms * 1e-3),
                 'path': 'pinned host int16 frames + u8 masks + f32 keypoints -> msq_prep_frames + msq_extract_chunk -> '
                         'pinned host crops/scalars/keypoint table/flips; 4-stream (H2D, prep, extract, D2H) double-buffered pipeline; in zero-copy '
                         'mode the prep kernel reads the ROI box of the raw frames directly from pinned host memory'}
                if e2e_ms is not None else None),
        'roofline': roofline, 'cpu_baseline': cpu_base,
    }
    line.update(extras)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def secondary_figures(args, geom, cfg, roi, bg):
    """(1) prep with Kinect-like invalid pixels (rate 0.002) incl. the GPU in-paint, (2) BASELINE configs[2]: the full
    extract path with a random-init Keypoint+Mask R-CNN R50-FPN (torchvision graph, bf16 autocast) between our kernels."""
    import numpy as np
    import torch
    from moseq2_detectron_extract_b200 import _dev, synthetic
    from moseq2_detectron_extract_b200.proc import prep_raw_frames
    out = {}

    def timed(fn, iters=3):
        fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / iters

    small = synthetic.generate_chunk(200, seed=11, geom=geom, invalid_rate=0.002)
    fr = torch.from_numpy(np.tile(small.frames, (5, 1, 1))).cuda()
    ms_fix = timed(lambda: prep_raw_frames(fr, bground_im=bg, roi=roi, vmin=0, vmax=100, fix_invalid_pixels=True))
    ms_raw = timed(lambda: prep_raw_frames(fr, bground_im=bg, roi=roi, vmin=0, vmax=100, fix_invalid_pixels=False))
    out['prep_with_invalid_pixels'] = {'frames': int(fr.shape[0]), 'invalid_rate': 0.002, 'ms_prep_only': ms_raw,
                                       'ms_prep_plus_inpaint': ms_fix, 'frames_per_s': fr.shape[0] / (ms_fix * 1e-3)}
    del fr
    # (3) use_tracking=True: the Kalman branch (a14) -- EM initialisation once, then 1000-frame chunks of one session in
    # sequence (the running state makes chunks of a session sequential; other sessions would run beside them)
    try:
        from moseq2_detectron_extract_b200.pipeline import ProcessFeaturesStep
        tcfg = dict(cfg, use_tracking=True, results_to_host=False, expected_instances=1)
        step = ProcessFeaturesStep(tcfg, 'features-tracking')
        step.initialize()
        tch = synthetic.generate_chunk(1000, seed=13, geom=geom, missing_every=97)
        tprep = prep_raw_frames(torch.from_numpy(tch.frames).cuda(), bground_im=bg, roi=roi, vmin=cfg['min_height'], vmax=cfg['max_height'])
        tmask, tkp = _dev.as_device(tch.masks), _dev.as_device(tch.keypoints, torch.float32)

        def tracked_chunk():
            step.process({'batch': 0, 'chunk': tprep, 'frame_idxs': list(range(1000)), 'offset': 0,
                          '_dense_instances': (tmask, tkp, tch.num_instances)})
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); tracked_chunk(); b.record(); torch.cuda.synchronize()      # first chunk: 2 x EM(10 iterations)
        ms_first = a.elapsed_time(b)
        ms_next = timed(tracked_chunk, iters=3)
        out['tracking_branch'] = {'workload': 'ProcessFeaturesStep(use_tracking=True), one session, 1000-frame chunks in sequence: '
                                              'clean + features + Kalman smoother (54 states) + angle filter + scalars + crops',
                                  'ms_first_chunk_with_em_init': ms_first, 'ms_per_chunk': ms_next,
                                  'frames_per_s': 1000 / (ms_next * 1e-3)}
    except Exception as exc:      # secondary figure only
        out['tracking_branch'] = {'error': repr(exc)}
    if args.rcnn_frames > 0:
        from moseq2_detectron_extract_b200.pipeline import InferenceStep, ProcessFeaturesStep
        n = args.rcnn_frames
        cfg2 = dict(cfg, batch_size=args.rcnn_batch, model='random', nframes=n, results_to_host=False, amp=True, dense_inference=True)
        infer, feats = InferenceStep(cfg2, 'infer'), ProcessFeaturesStep(cfg2, 'features')
        infer.initialize()
        feats.initialize()
        ch = synthetic.generate_chunk(min(n, 500), seed=12, geom=geom)
        raw = torch.from_numpy(np.tile(ch.frames, ((n + len(ch.frames) - 1) // len(ch.frames), 1, 1))[:n]).cuda()

        def full():
            chunk = prep_raw_frames(raw, bground_im=bg, roi=roi, vmin=cfg['min_height'], vmax=cfg['max_height'])
            data = {'batch': 0, 'chunk': chunk, 'frame_idxs': list(range(n)), 'offset': 0}
            feats.process(infer.process(data))
        ms_full = timed(full, iters=2)
        out['full_extract_rcnn'] = {
            'workload': 'configs[2]: prep -> scale/normalise/resize (one kernel) -> Keypoint+Mask R-CNN R50-FPN (random init, torchvision graph with '
                        'BatchNorm folded, cuDNN fused conv epilogues, batched heads + our NMS / RoIAlign / keypoint kernels, bf16 autocast, '
                        f'batch {args.rcnn_batch}, 100 proposals, 1 detection/frame) -> batched paste of the first instance -> features -> crops',
            'frames': n, 'ms': ms_full, 'frames_per_s': n / (ms_full * 1e-3)}
    return out


def main():
    args = parse_args()
    if args.impl == 'reference':
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()

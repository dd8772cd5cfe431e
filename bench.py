#!/usr/bin/env python
"""Benchmark of the extract hot path (BASELINE.json metric: extract frames/s + kernel HBM GB/s vs peak).

Headline workload (config.workload): BASELINE.json configs[1] -- preprocess + crop/rotate kernels only (no R-CNN) on a
30-minute synthetic session (54,000 Kinect-v2-shaped 512x424 int16 frames, ROI box 240x240, 80x80 crops, 1000-frame chunks,
use_tracking=False; instance masks + keypoints given).  One "step" = one pass over the whole session.  `value` = frames/s
with the session resident in HBM; `e2e` = the same metric with HOST (pinned) frames, bit-packed masks and keypoints copied
to the GPU and the results (crops, scalars, keypoint table, flips) copied back inside the timed region, the host frames
cycling through a pool far larger than any CPU cache.  Under torchrun the global chunk list (N sessions) is sharded
contiguously over the ranks (shard.shard_chunks; weak scaling, no collective).

Further measured workloads carried in the same JSON line (each with its own e2e / roofline):
  full_extract_rcnn        BASELINE configs[2]: prep -> the repo's own detectron2-configuration Keypoint/Mask R-CNN (bf16, the
                           reference's 1000 test proposals) -> paste -> features -> crops, >= 20 chunks
  full_extract_rcnn_topk100  the same graph exported with RPN.POST_NMS_TOPK_TEST = 100
  azure                    BASELINE configs[4] geometry: 640x576 frames, 400x400 ROI box, 128x128 crops

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
CHUNK = 1000


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--frames', type=int, default=54000, help='frames per session (per GPU)')
    ap.add_argument('--launch-chunks', type=int, default=6, help='1000-frame chunks processed per kernel launch')
    ap.add_argument('--pool-frames', type=int, default=1000, help='distinct synthetic frames generated on the host')
    ap.add_argument('--host-pool-gb', type=float, default=6.0, help='pinned host frame pool the end-to-end pass cycles through (per rank)')
    ap.add_argument('--geometry', default='kinect_v2', choices=['kinect_v2', 'azure'])
    ap.add_argument('--cpu-frames-per-worker', type=int, default=250)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--rcnn-frames', type=int, default=20000, help='frames of the full-extract (R-CNN) workloads; 0 = skip')
    ap.add_argument('--rcnn-batch', type=int, default=500)
    ap.add_argument('--azure-frames', type=int, default=12000, help='frames of the Azure-geometry workload; 0 = skip')
    ap.add_argument('--no-secondary', action='store_true', help='skip in-painting / tracking side figures')
    return ap.parse_args()


def workload_config(args, geom, n_gpus):
    return {
        'workload': f'configs[1]: prep + clean + features + angles/flips/filter + scalars/keypoints + crop/rotate, no R-CNN, '
                    f'{args.frames}-frame synthetic session per GPU',
        'frame': f'{geom.width}x{geom.height} int16', 'roi_box': None, 'crop': list(geom.crop_size), 'chunk_size': CHUNK,
        'frames_per_session': args.frames, 'sessions': n_gpus, 'frames_per_launch': args.launch_chunks * CHUNK,
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
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# one geometry's no-R-CNN workload: resident pass + end-to-end pass (used for Kinect-v2 and Azure)
# ------------------------------------------------------------------------------------------------
class ExtractWorkload:
    """The configs[1] work for one sensor geometry on this rank's GPU: synthetic session resident in HBM, and the 4-stream
    host -> GPU -> host pipeline."""

    def __init__(self, args, geom, n_frames, rank, world, host_pool_gb):
        import numpy as np
        import torch
        from moseq2_detectron_extract_b200 import _dev, _lib, synthetic
        from moseq2_detectron_extract_b200.engine import ChunkEngine
        from moseq2_detectron_extract_b200.shard import shard_chunks
        self.args, self.geom, self.n_frames, self.rank, self.world = args, geom, n_frames, rank, world
        self.torch, self._dev, self._lib, self.ChunkEngine = torch, _dev, _lib, ChunkEngine
        self.cfg = synthetic.default_config(geom)
        roi, bg = synthetic.make_roi(geom), synthetic.make_background(geom)
        self.roi_np, self.bg_np = roi, bg
        self.y0, self.x0, y1, x1 = synthetic.roi_bbox(roi)
        self.h, self.w = y1 - self.y0, x1 - self.x0
        self.H, self.W = geom.height, geom.width
        self.launch = args.launch_chunks * CHUNK
        # this rank's chunks of the global job (world sessions of n_frames each, sharded contiguously by chunk)
        per_session = (n_frames + CHUNK - 1) // CHUNK
        self.my_chunks = shard_chunks(per_session * world, rank, world)
        # ---- synthetic session: `pool` distinct frames generated on the host, cycled (with a per-chunk roll) on the GPU
        npool = min(args.pool_frames, n_frames)
        pool = synthetic.generate_chunk(npool, seed=rank, geom=geom, t0=0)
        self.npool = npool
        self.pool_frames = torch.from_numpy(pool.frames).pin_memory()
        self.pool_masks = torch.from_numpy(pool.masks).pin_memory()
        self.pool_bits = torch.from_numpy(np.packbits(pool.masks, axis=-1, bitorder='little')).pin_memory()
        self.pool_kpts = torch.from_numpy(pool.keypoints).pin_memory()
        # compulsory bytes of the masked sums: the mask, plus the frame bytes under 16-pixel groups that carry a mask byte
        groups = pool.masks.reshape(npool, -1)
        groups = groups[:, : groups.shape[1] // 16 * 16].reshape(npool, -1, 16).any(axis=2).mean()
        self.masked_group_fraction = float(groups)
        d_pool_f, d_pool_m, d_pool_k = self.pool_frames.cuda(), self.pool_masks.cuda(), self.pool_kpts.cuda()
        idx = torch.cat([(torch.arange(min(CHUNK, n_frames - c), device='cuda') + 37 * (self.my_chunks.start + c // CHUNK)) % npool
                         for c in range(0, n_frames, CHUNK)])
        self.frames = torch.empty((n_frames, self.H, self.W), dtype=torch.int16, device='cuda')
        for s in range(0, n_frames, 2000):
            self.frames[s:s + 2000] = d_pool_f[idx[s:s + 2000]]
        self.masks = d_pool_m[idx].contiguous()
        self.kpts = d_pool_k[idx].contiguous()
        del d_pool_f, d_pool_m, d_pool_k
        self.bg_d, self.roi_d = _dev.as_device(bg), _dev.as_device(roi.astype(np.uint8))
        self.prep_buf = _dev.empty((self.launch, self.h, self.w), torch.uint8)
        self.pos_buf = _dev.positive_bits_like(self.prep_buf)      # the prep kernel's bit rows for the cleaning pass
        self.invalid = _dev.empty((self.launch,), torch.int32)
        self.engine = ChunkEngine()
        self.kw = dict(chunk_size=CHUNK, min_height=self.cfg['min_height'], max_height=self.cfg['max_height'],
                       true_depth=self.cfg['true_depth'], crop_size=self.cfg['crop_size'])
        self.flags = _lib.MSQ_PREP_HAS_VMIN | _lib.MSQ_PREP_HAS_VMAX
        self.host_pool = None
        self.host_pool_gb = host_pool_gb

    def prep(self, src, n, out, invalid, positive, stream=None):
        _dev, _lib = self._dev, self._lib
        _lib.call('msq_prep_frames_bits', _dev.ptr(src), n, self.H, self.W, _dev.ptr(self.bg_d), _lib.MSQ_BG_F32, _dev.ptr(self.roi_d),
                  self.y0, self.x0, self.h, self.w, float(self.cfg['min_height']), float(self.cfg['max_height']), self.flags,
                  _dev.ptr(out), _dev.ptr(invalid), None, _dev.ptr(positive), stream if stream is not None else _dev.stream())

    def resident_step(self):
        for s in range(0, self.n_frames, self.launch):
            n = min(self.launch, self.n_frames - s)
            self.prep(self.frames[s:s + n], n, self.prep_buf, self.invalid, self.pos_buf)
            self.engine.extract(self.prep_buf[:n], self.masks[s:s + n], self.kpts[s:s + n], positive_bits=self.pos_buf[:n], **self.kw)

    # ---- end to end ---------------------------------------------------------------------------------------------
    def build_host_pool(self):
        """A pinned host frame pool far larger than any CPU cache (the chunks of an e2e pass walk through it), filled with
        rolled copies of the distinct synthetic frames."""
        torch = self.torch
        frame_bytes = self.H * self.W * 2
        n_pool_chunks = max(1, min(int(self.host_pool_gb * 1e9 / (frame_bytes * CHUNK)), (self.n_frames + CHUNK - 1) // CHUNK))
        if self.npool < CHUNK:
            n_pool_chunks = 1
        self.host_pool = torch.empty((n_pool_chunks, min(CHUNK, self.npool), self.H, self.W), dtype=torch.int16, pin_memory=True)
        for c in range(n_pool_chunks):
            self.host_pool[c].copy_(self.pool_frames[:self.host_pool.shape[1]])
        return self.host_pool.numel() * 2

    def h2d_bandwidth(self, repeats=3):
        """Plain pinned host -> device copy rate of this rank (GB/s) over the e2e pool, CUDA events."""
        torch = self.torch
        dst = torch.empty_like(self.host_pool[0], device='cuda')
        dst.copy_(self.host_pool[0], non_blocking=True)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        nbytes = 0
        a.record()
        for _ in range(repeats):
            for c in range(self.host_pool.shape[0]):
                dst.copy_(self.host_pool[c], non_blocking=True)
                nbytes += dst.numel() * 2
        b.record()
        torch.cuda.synchronize()
        return nbytes / (a.elapsed_time(b) * 1e-3) / 1e9

    def setup_e2e(self):
        import numpy as np
        torch, _dev, _lib = self.torch, self._dev, self._lib
        chunk = self.host_pool.shape[1]
        self.e2e_chunk = chunk
        self.streams = [torch.cuda.Stream() for _ in range(4)]            # H2D, prep, extract, D2H
        slots = 2
        h, w, H, W = self.h, self.w, self.H, self.W
        wb = (w + 7) // 8
        self.in_f = [_dev.empty((chunk, H, W), torch.int16) for _ in range(slots)]
        self.in_roi = [torch.zeros((chunk, h, w), dtype=torch.int16, device='cuda') for _ in range(slots)]   # ROI rows staged by strided DMA
        self.bands = _dev.roi_bands(self.roi_np[self.y0:self.y0 + h, self.x0:self.x0 + w], 16)
        self.band_bytes_per_frame = int(sum((self.bands[0][b + 1] - self.bands[0][b]) * (self.bands[2][b] - self.bands[1][b])
                                            for b in range(len(self.bands[1]))) * 2)
        self.bg_box = _dev.as_device(np.ascontiguousarray(self.bg_np[self.y0:self.y0 + h, self.x0:self.x0 + w]))
        self.roi_box = _dev.as_device(np.ascontiguousarray(self.roi_np[self.y0:self.y0 + h, self.x0:self.x0 + w].astype(np.uint8)))
        self.in_bits = [_dev.empty((chunk, h, wb), torch.uint8) for _ in range(slots)]
        self.in_m = [_dev.empty((chunk, h, w), torch.uint8) for _ in range(slots)]
        self.in_k = [_dev.empty((chunk, 8, 3), torch.float32) for _ in range(slots)]
        self.engines = [self.ChunkEngine() for _ in range(slots)]
        self.preps = [_dev.empty((chunk, h, w), torch.uint8) for _ in range(slots)]
        self.poss = [_dev.positive_bits_like(p) for p in self.preps]
        self.invs = [_dev.empty((chunk,), torch.int32) for _ in range(slots)]
        cw, ch = self.cfg['crop_size']
        self.host_out = [{'depth_crops': torch.empty((chunk, ch, cw), dtype=torch.uint8).pin_memory(),
                          'mask_crops': torch.empty((chunk, ch, cw), dtype=torch.uint8).pin_memory(),
                          'scalars': torch.empty((_lib.NUM_SCALARS, chunk), dtype=torch.float64).pin_memory(),
                          'kpt_cols': torch.empty((_lib.NUM_KPT_COLS, chunk), dtype=torch.float64).pin_memory(),
                          'flips': torch.empty((chunk,), dtype=torch.uint8).pin_memory(),
                          'invalid': torch.empty((chunk,), dtype=torch.int32).pin_memory()} for _ in range(slots)]
        self.small_chunk_bytes = self.pool_bits[:chunk].numel() + self.pool_kpts[:chunk].numel() * 4
        self.d2h_chunk_bytes = sum(v.numel() * v.element_size() for v in self.host_out[0].values())

    def copy_roi(self, src_host, dst, bands):
        _dev, _lib = self._dev, self._lib
        if src_host.shape[0] == 0:
            return
        if bands:
            _dev.copy_roi_bands(src_host, self.y0, self.x0, self.bands, dst)
        else:
            _lib.call('msq_copy_roi_rows', _dev.ptr(src_host), int(src_host.shape[0]), self.H, self.W, self.y0, self.x0, self.h, self.w,
                      _dev.ptr(dst), _dev.stream())

    def e2e_step(self, zero_copy, roi_dma=False, bands=False):
        """One pass over the session from pinned host buffers.  zero_copy=True: the prep kernel reads the ROI box of the raw
        frames straight out of pinned host memory (UVA), so only the bytes the path needs cross PCIe; bit-packed masks and
        keypoints go through cudaMemcpyAsync.  zero_copy=False: whole frames are copied -- or, with roi_dma, only the ROI box of
        every frame as ONE strided DMA transfer per chunk (msq_copy_roi_rows), prepared on the device from there."""
        torch, _dev, _lib = self.torch, self._dev, self._lib
        copy_in, prep_st, compute, copy_out = self.streams
        slots, chunk = 2, self.e2e_chunk
        n_chunks = (self.n_frames + chunk - 1) // chunk
        ev_h2d, ev_comp, ev_d2h = [None] * slots, [None] * slots, [None] * slots
        n_pool = self.host_pool.shape[0]
        for c in range(n_chunks):
            b = c % slots
            src_host = self.host_pool[c % n_pool]
            with torch.cuda.stream(copy_in):
                if ev_comp[b] is not None:
                    copy_in.wait_event(ev_comp[b])          # input slot consumed by the previous user
                if roi_dma:
                    # (measured: a second DMA queue adds nothing -- a strided transfer of 480-byte rows already runs at 47 GB/s of the
                    # 55 GB/s a plain copy reaches; rows shorter than two 256-byte PCIe payloads do not go faster either)
                    self.copy_roi(src_host, self.in_roi[b], bands)
                elif not zero_copy:
                    self.in_f[b].copy_(src_host, non_blocking=True)
                self.in_bits[b].copy_(self.pool_bits[:chunk], non_blocking=True)
                self.in_k[b].copy_(self.pool_kpts[:chunk], non_blocking=True)
                _lib.call('msq_unpack_mask_bits', _dev.ptr(self.in_bits[b]), chunk, self.h, self.w, _dev.ptr(self.in_m[b]), _dev.stream())
                ev_h2d[b] = torch.cuda.Event()
                ev_h2d[b].record(copy_in)
            # prep is PCIe-bound in zero-copy mode (it pulls the ROI box from host memory) and needs few SMs; on its own
            # stream it overlaps the SM-bound extract kernels of the previous chunk
            with torch.cuda.stream(prep_st):
                if not zero_copy:
                    prep_st.wait_event(ev_h2d[b])
                if ev_comp[b] is not None:
                    prep_st.wait_event(ev_comp[b])          # preps[b] consumed by the previous user of the slot
                if roi_dma:
                    _lib.call('msq_prep_frames_bits', _dev.ptr(self.in_roi[b]), chunk, self.h, self.w, _dev.ptr(self.bg_box), _lib.MSQ_BG_F32,
                              _dev.ptr(self.roi_box), 0, 0, self.h, self.w, float(self.cfg['min_height']), float(self.cfg['max_height']),
                              self.flags, _dev.ptr(self.preps[b]), _dev.ptr(self.invs[b]), None, _dev.ptr(self.poss[b]), _dev.stream())
                else:
                    self.prep(src_host if zero_copy else self.in_f[b], chunk, self.preps[b], self.invs[b], self.poss[b], _dev.stream())
                ev_prep = torch.cuda.Event()
                ev_prep.record(prep_st)
            with torch.cuda.stream(compute):
                compute.wait_event(ev_h2d[b])
                compute.wait_event(ev_prep)
                if ev_d2h[b] is not None:
                    compute.wait_event(ev_d2h[b])           # output slot drained
                res = self.engines[b].extract(self.preps[b], self.in_m[b], self.in_k[b], positive_bits=self.poss[b], **self.kw)
                ev_comp[b] = torch.cuda.Event()
                ev_comp[b].record(compute)
            with torch.cuda.stream(copy_out):
                copy_out.wait_event(ev_comp[b])
                for key in ('depth_crops', 'mask_crops', 'scalars', 'kpt_cols', 'flips'):
                    self.host_out[b][key].copy_(res[key], non_blocking=True)
                self.host_out[b]['invalid'].copy_(self.invs[b], non_blocking=True)
                ev_d2h[b] = torch.cuda.Event()
                ev_d2h[b].record(copy_out)
        for st in (copy_out, compute, prep_st):
            torch.cuda.current_stream().wait_stream(st)

    def frame_bytes_of_mode(self, mode):
        return {'copy': self.H * self.W * 2, 'zero-copy': self.h * self.w * 2, 'roi-dma': self.h * self.w * 2,
                'roi-dma-bands': self.band_bytes_per_frame}[mode]

    def bytes_per_frame(self):
        A, C = self.h * self.w, self.cfg['crop_size'][0] * self.cfg['crop_size'][1]
        # algorithmic bytes per frame: only the ROI box of the raw frame is ever needed; the masked sums
        # read the mask and the frame bytes under 16-pixel groups that carry a mask byte (a masked-out pixel contributes 0)
        return {'prep_frames': 3 * A, 'clean_frames': 2 * A, 'frame_features': 2 * A + 40,
                'masked_sums': A + self.masked_group_fraction * A,
                'scalars_keypoints': 113 * 8 + 24 * 4 + 8 * 1 + 64, 'crop_rotate': 2 * 2 * C + 2 * C,
                'angles_flips_filter': 8 * 8 + 96 + 9}


E2E_MODES = {'copy': dict(zero_copy=False), 'zero-copy': dict(zero_copy=True), 'roi-dma': dict(zero_copy=False, roi_dma=True),
             'roi-dma-bands': dict(zero_copy=False, roi_dma=True, bands=True)}


def measure_e2e_modes(torch, wl, steps, warmup, barrier, reduce, MAX, skip=()):
    """Time every way the raw frames can reach the GPU (max over ranks each); the end-to-end figure is the fastest."""
    names = [k for k in E2E_MODES if k not in skip]
    ms = [timed_steps(torch, (lambda kw=E2E_MODES[k]: wl.e2e_step(**kw)), steps, warmup, barrier) for k in names]
    ms = reduce(ms, MAX)
    modes = dict(zip(names, ms))
    best = min(modes, key=modes.get)
    return modes, best, modes[best]


def timed_steps(torch, fn, steps, warmup, barrier):
    for _ in range(warmup):
        fn()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    barrier()
    return e0.elapsed_time(e1)


def peak_numbers():
    peaks_path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(peaks_path):
        pk = json.load(open(peaks_path))
        return float(pk['hbm_gbs']), float(pk.get('bf16_tflops_sustained', pk.get('bf16_tflops', 1375.4))), 'measured (MEASURED_PEAKS.json)'
    return 6650.0, 1400.0, 'fallback (B200_PROFILING.md: 6.65 TB/s, ~1.4 PFLOP/s sustained)'


def kernel_roofline(wl, ktimes, n_frames_total, peak, peak_src):
    bytes_per_frame = wl.bytes_per_frame()
    kernel_ms = {k: v[0] for k, v in ktimes.items() if v[1] > 0 and k in bytes_per_frame}
    dominant = max(kernel_ms, key=kernel_ms.get)
    total_kernel_ms = sum(kernel_ms.values())
    d_ms, d_cnt = ktimes[dominant]
    frames_per_launch_avg = n_frames_total / d_cnt
    achieved = bytes_per_frame[dominant] * frames_per_launch_avg / (d_ms / d_cnt * 1e-3) / 1e9
    return {
        'bound': 'hbm', 'kernel': dominant, 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
        'traffic': None, 'peak_source': peak_src, 'avg_launch_ms': d_ms / d_cnt,
        'algorithmic_bytes_per_frame': bytes_per_frame[dominant], 'frames_per_launch': frames_per_launch_avg,
        'share_of_kernel_time': d_ms / total_kernel_ms,
        'per_kernel': {k: {'ms_total': v, 'share': v / total_kernel_ms, 'launches': ktimes[k][1],
                           'bytes_per_frame': bytes_per_frame[k],
                           'GBps': bytes_per_frame[k] * n_frames_total / (v * 1e-3) / 1e9,
                           'frac_of_peak': bytes_per_frame[k] * n_frames_total / (v * 1e-3) / 1e9 / peak}
                       for k, v in kernel_ms.items()},
    }


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

    import torch
    import torch.distributed as dist
    from moseq2_detectron_extract_b200 import _dev, _lib, synthetic

    torch.cuda.set_device(local_rank)
    _dev.require_cuda()
    from moseq2_detectron_extract_b200.shard import bind_to_gpu_numa_node
    numa = bind_to_gpu_numa_node(local_rank)      # before any pinned allocation: host buffers land on the GPU's NUMA node
    if world > 1:
        # NCCL carries only the barrier and the max / sum of a few timing scalars: there is NO data-path collective
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce(values, op):
        t = torch.tensor(values, dtype=torch.float64, device='cuda')
        if world > 1:
            dist.all_reduce(t, op=op)
        return [float(v) for v in t]

    MAX, SUM, MIN = (dist.ReduceOp.MAX, dist.ReduceOp.SUM, dist.ReduceOp.MIN)
    peak, tensor_peak, peak_src = peak_numbers()
    geom = getattr(synthetic.SessionGeometry, args.geometry)()
    wl = ExtractWorkload(args, geom, args.frames, rank, world, args.host_pool_gb)

    # ---- device-resident throughput ------------------------------------------------------------------
    # the clock sampler is started BEFORE the warm-up and must have delivered a line before timing starts (nvidia-smi
    # needs a few hundred ms to come up; the timed region can be shorter than that), so its samples see the GPU under load
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        sampler.wait_first(3.0)
    barrier()
    for _ in range(args.warmup):
        wl.resident_step()
    barrier()
    launches_before = sum(_lib.kernel_launches().values())
    _lib.kernel_timing(True)
    barrier()
    t_wall0 = time.time()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        wl.resident_step()
    ev1.record()
    barrier()
    t_wall1 = time.time()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    _lib.kernel_timing(False)
    ktimes = _lib.kernel_timing_collect()
    gpu_launches = sum(_lib.kernel_launches().values()) - launches_before
    assert int(wl.invalid.sum().item()) == 0
    roofline = kernel_roofline(wl, ktimes, args.frames * args.steps, peak, peak_src)

    # ---- end to end: pinned host inputs -> GPU -> pinned host results, 4-stream double-buffered pipeline ----------------
    e2e = None
    if not args.no_e2e:
        pool_bytes = wl.build_host_pool()
        wl.setup_e2e()
        barrier()
        h2d_gbs = wl.h2d_bandwidth()
        h2d_sum, = reduce([h2d_gbs], SUM)
        h2d_min, = reduce([h2d_gbs], MIN)
        w3 = min(args.warmup, 3)
        modes, e2e_mode, e2e_ms = measure_e2e_modes(torch, wl, args.steps, w3, barrier, reduce, MAX)
        assert float(wl.host_out[0]['scalars'][6].sum()) > 0          # area_px really came back
        n_chunks = (args.frames + wl.e2e_chunk - 1) // wl.e2e_chunk
        h2d_step = (wl.e2e_chunk * wl.frame_bytes_of_mode(e2e_mode) + wl.small_chunk_bytes) * n_chunks
        total = args.frames * world * args.steps
        e2e = {'value': total / (e2e_ms * 1e-3), 'unit': UNIT, 'h2d_bytes_per_step': int(h2d_step),
               'd2h_bytes_per_step': int(wl.d2h_chunk_bytes * n_chunks), 'ms_per_step': e2e_ms / args.steps,
               'mode': e2e_mode,
               'frames_per_s_by_mode': {k: total / (v * 1e-3) for k, v in modes.items()},
               'h2d_frame_bytes_by_mode': {k: wl.frame_bytes_of_mode(k) for k in modes},
               'h2d_GBps_sustained_all_ranks': h2d_step * args.steps / (e2e_ms * 1e-3) / 1e9 * world,
               'host_pool_bytes_per_rank': int(pool_bytes),
               'pinned_h2d_copy_GBps': {'sum_over_ranks': h2d_sum, 'min_rank': h2d_min,
                                        'how': 'cudaMemcpyAsync of the pinned frame pool, all ranks at once, CUDA events'},
               'mask_format': 'bit rows (numpy.packbits little) expanded on the GPU by msq_unpack_mask_bits',
               'path': 'pinned host int16 frames (pool >> CPU caches, cycled) + bit-packed masks + f32 keypoints -> msq_prep_frames + '
                       'msq_unpack_mask_bits + msq_extract_chunk -> pinned host crops/scalars/keypoint table/flips; 4-stream (H2D, prep, '
                       'extract, D2H) double-buffered pipeline; zero-copy mode: the prep kernel reads the ROI box of the raw frames '
                       'directly from pinned host memory; roi-dma mode: the ROI box of every frame of a chunk crosses PCIe as one strided '
                       'DMA transfer (msq_copy_roi_rows) and is prepared on the device; roi-dma-bands mode: the ROI disc as 16 horizontal '
                       'bands, one strided DMA transfer each (msq_copy_roi_bands): only the pixels the path needs, rounded to 8 columns'}
        del wl.host_pool
        wl.host_pool = None

    ms_max, = reduce([ms], MAX)
    launches_sum, = reduce([float(gpu_launches)], SUM)

    # ---- the other measured workloads: Azure geometry, full extract with the R-CNN ------------------------------------
    extras = {}
    frames_main = wl.frames
    del wl.frames, wl.masks, wl.kpts
    del frames_main
    torch.cuda.empty_cache()
    try:
        if args.azure_frames > 0 and args.geometry != 'azure':
            extras['azure'] = azure_workload(args, rank, world, barrier, reduce, (MAX, SUM, MIN), peak, peak_src)
    except Exception as exc:       # never let a further workload break the contract line
        extras['azure'] = {'error': repr(exc)[:300]}
    torch.cuda.empty_cache()
    try:
        if args.rcnn_frames > 0:
            extras.update(rcnn_workloads(args, geom, rank, world, barrier, reduce, (MAX, SUM, MIN), tensor_peak, peak_src))
    except Exception as exc:
        extras['full_extract_rcnn'] = {'error': repr(exc)[:300]}
    if rank == 0 and world == 1 and not args.no_secondary:
        try:
            extras.update(secondary_figures(args, geom, wl.cfg, wl.roi_np, wl.bg_np))
        except Exception as exc:
            extras['secondary_error'] = repr(exc)[:300]
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    total_frames = args.frames * world * args.steps
    traffic_file = os.path.join(ROOT, 'profiles', 'dominant_kernel_traffic.json')
    if os.path.exists(traffic_file):
        try:
            t = json.load(open(traffic_file))
            if t.get('kernel') == roofline['kernel']:
                roofline['traffic'] = t.get('dram_bytes_per_launch')
                roofline['traffic_source'] = t.get('source')
        except Exception:
            pass
    cfg_out = workload_config(args, geom, world)
    cfg_out['roi_box'] = [wl.y0, wl.x0, wl.y0 + wl.h, wl.x0 + wl.w]
    line = {
        'metric': METRIC, 'value': total_frames / (ms_max * 1e-3), 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': ms_max / args.steps, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'u8 / int16 (pixels), int64 (moments), f64 (features)', 'data': 'synthetic',
        'config': cfg_out, 'clocks': clocks, 'gpu_launches': int(launches_sum), 'e2e': e2e, 'roofline': roofline,
        'cpu_baseline': cpu_base,
        'collective': 'none on the data path (chunks shard; NCCL carries the timing barrier and MAX / SUM of a few scalars only)',
        'host': {'numa_binding': numa if numa else 'unavailable (NVML / sysfs report no NUMA node for the GPU)', 'cpus': os.cpu_count()},
    }
    line.update(extras)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def azure_workload(args, rank, world, barrier, reduce, ops, peak, peak_src):
    """BASELINE configs[4] geometry on this rank: 640x576 int16 frames, 400x400 ROI box, 128x128 crops; resident + end to end,
    per-kernel roofline (the multi-tile clean path and the 128-pixel crop staging run here)."""
    import torch
    from moseq2_detectron_extract_b200 import _lib, synthetic
    MAX, SUM, MIN = ops
    geom = synthetic.SessionGeometry.azure()
    sub = argparse.Namespace(**vars(args))
    sub.pool_frames = min(args.pool_frames, 500)
    sub.launch_chunks = min(args.launch_chunks, 3)
    wl = ExtractWorkload(sub, geom, args.azure_frames, rank, world, min(args.host_pool_gb, 3.0))
    steps = max(2, args.steps // 2)
    for _ in range(2):
        wl.resident_step()
    barrier()
    _lib.kernel_timing(True)
    ms = timed_steps(torch, wl.resident_step, steps, 0, barrier)
    _lib.kernel_timing(False)
    ktimes = _lib.kernel_timing_collect()
    roofline = kernel_roofline(wl, ktimes, args.azure_frames * steps, peak, peak_src)
    out = {'workload': f'configs[4] geometry: {geom.width}x{geom.height} int16, ROI box {wl.h}x{wl.w}, crops {list(geom.crop_size)}, '
                       f'{args.azure_frames} frames per GPU, batch 64 is the R-CNN batch of that config (no R-CNN in this pass)',
           'frames': args.azure_frames, 'steps': steps}
    ms_max, = reduce([ms], MAX)
    out['frames_per_s'] = args.azure_frames * world * steps / (ms_max * 1e-3)
    out['ms_per_step'] = ms_max / steps
    out['roofline'] = roofline
    if not args.no_e2e:
        wl.build_host_pool()
        wl.setup_e2e()
        modes, mode, ms_best = measure_e2e_modes(torch, wl, steps, 1, barrier, reduce, MAX, skip=('copy',))
        n_chunks = (args.azure_frames + wl.e2e_chunk - 1) // wl.e2e_chunk
        total = args.azure_frames * world * steps
        out['e2e'] = {'value': total / (ms_best * 1e-3), 'unit': UNIT, 'mode': mode,
                      'frames_per_s_by_mode': {k: total / (v * 1e-3) for k, v in modes.items()},
                      'h2d_bytes_per_step': int((wl.e2e_chunk * wl.frame_bytes_of_mode(mode) + wl.small_chunk_bytes) * n_chunks),
                      'd2h_bytes_per_step': int(wl.d2h_chunk_bytes * n_chunks)}
    return out


def rcnn_workloads(args, geom, rank, world, barrier, reduce, ops, tensor_peak, peak_src):
    """BASELINE configs[2]: the full extract with the repo's own Keypoint/Mask R-CNN graph between the pre- and post-processing
    kernels, random weights, bf16.  Measured twice: with detectron2's 1000 test proposals (the reference's configuration) and with
    the graph built for 100.  Resident (frames in HBM) and end to end (pinned host int16 frames in, pinned host crops / scalars /
    keypoint table / flips out; masks never cross PCIe on this path)."""
    import numpy as np
    import torch
    from moseq2_detectron_extract_b200 import _dev, _lib, synthetic
    from moseq2_detectron_extract_b200.engine import ChunkEngine
    from moseq2_detectron_extract_b200.model import rcnn
    from moseq2_detectron_extract_b200.model.predict import Predictor
    MAX, SUM, MIN = ops
    n = args.rcnn_frames
    B = args.rcnn_batch
    cfg = synthetic.default_config(geom)
    roi, bg = synthetic.make_roi(geom), synthetic.make_background(geom)
    y0, x0, y1, x1 = synthetic.roi_bbox(roi)
    h, w, H, W = y1 - y0, x1 - x0, geom.height, geom.width
    npool = min(1000, n)
    pool = synthetic.generate_chunk(npool, seed=12 + rank, geom=geom)
    host_frames = torch.from_numpy(pool.frames).pin_memory()
    n_pool_chunks = max(1, min(int(args.host_pool_gb * 1e9 / (H * W * 2 * npool)), (n + npool - 1) // npool))
    host_pool = torch.empty((n_pool_chunks, npool, H, W), dtype=torch.int16, pin_memory=True)
    for c in range(n_pool_chunks):
        host_pool[c].copy_(host_frames)
    dev_pool = host_frames.cuda()
    bg_d, roi_d = _dev.as_device(bg), _dev.as_device(roi.astype(np.uint8))
    flags = _lib.MSQ_PREP_HAS_VMIN | _lib.MSQ_PREP_HAS_VMAX
    kw = dict(chunk_size=CHUNK, min_height=cfg['min_height'], max_height=cfg['max_height'], true_depth=cfg['true_depth'],
              crop_size=cfg['crop_size'])
    cw, ch = cfg['crop_size']
    out = {}
    for key, topk in (('full_extract_rcnn', 1000), ('full_extract_rcnn_topk100', 100)):
        pred = Predictor.from_random_init(post_nms_topk=topk, scripted=True)
        engine = ChunkEngine()
        prep_buf = [_dev.empty((npool, h, w), torch.uint8) for _ in range(2)]
        pos_buf = [_dev.positive_bits_like(p) for p in prep_buf]
        invalid = _dev.empty((npool,), torch.int32)
        host_out = {'depth_crops': torch.empty((npool, ch, cw), dtype=torch.uint8).pin_memory(),
                    'mask_crops': torch.empty((npool, ch, cw), dtype=torch.uint8).pin_memory(),
                    'scalars': torch.empty((_lib.NUM_SCALARS, npool), dtype=torch.float64).pin_memory(),
                    'kpt_cols': torch.empty((_lib.NUM_KPT_COLS, npool), dtype=torch.float64).pin_memory(),
                    'flips': torch.empty((npool,), dtype=torch.uint8).pin_memory()}
        stage_ms = {}

        def prep_chunk(src, b, stream=None):
            _lib.call('msq_prep_frames_bits', _dev.ptr(src), npool, H, W, _dev.ptr(bg_d), _lib.MSQ_BG_F32, _dev.ptr(roi_d), y0, x0, h, w,
                      float(cfg['min_height']), float(cfg['max_height']), flags, _dev.ptr(prep_buf[b]), _dev.ptr(invalid), None,
                      _dev.ptr(pos_buf[b]), stream if stream is not None else _dev.stream())

        def infer_and_extract(b):
            chunk = prep_buf[b]
            parts = [pred.predict_dense(chunk[i:i + B], cfg['min_height'], cfg['max_height']) for i in range(0, npool, B)]
            masks, kpts = torch.cat([p[0] for p in parts]), torch.cat([p[1] for p in parts])
            return engine.extract(chunk, masks, kpts, positive_bits=pos_buf[b], **kw)

        def resident_pass():
            for c in range((n + npool - 1) // npool):
                prep_chunk(dev_pool, 0)
                infer_and_extract(0)

        prep_st, d2h_st = torch.cuda.Stream(), torch.cuda.Stream()

        def e2e_pass():
            """prep of chunk c+1 (zero-copy read of the ROI box from pinned host memory) overlaps the graph of chunk c; results of
            chunk c go back on a third stream."""
            n_chunks = (n + npool - 1) // npool
            main = torch.cuda.current_stream()
            ev_prep, ev_free, ev_out = [None, None], [None, None], None
            with torch.cuda.stream(prep_st):
                prep_chunk(host_pool[0], 0, _dev.stream())
                ev_prep[0] = torch.cuda.Event(); ev_prep[0].record(prep_st)
            for c in range(n_chunks):
                b = c % 2
                if c + 1 < n_chunks:
                    with torch.cuda.stream(prep_st):
                        if ev_free[1 - b] is not None:
                            prep_st.wait_event(ev_free[1 - b])
                        prep_chunk(host_pool[(c + 1) % n_pool_chunks], 1 - b, _dev.stream())
                        ev_prep[1 - b] = torch.cuda.Event(); ev_prep[1 - b].record(prep_st)
                main.wait_event(ev_prep[b])
                if ev_out is not None:
                    main.wait_event(ev_out)                 # the engine's result buffers were drained
                res = infer_and_extract(b)
                ev_free[b] = torch.cuda.Event(); ev_free[b].record(main)
                with torch.cuda.stream(d2h_st):
                    d2h_st.wait_event(ev_free[b])
                    for k2 in host_out:
                        host_out[k2].copy_(res[k2], non_blocking=True)
                    ev_out = torch.cuda.Event(); ev_out.record(d2h_st)
            main.wait_stream(d2h_st)
            main.wait_stream(prep_st)

        launches0 = sum(_lib.kernel_launches().values())
        ms_res = timed_steps(torch, resident_pass, 1, 1, barrier)
        launches = sum(_lib.kernel_launches().values()) - launches0
        ms_e2e = timed_steps(torch, e2e_pass, 1, 1, barrier)
        assert float(host_out['scalars'][6].sum()) >= 0
        # per-stage device times of one batch (CUDA events around each stage of the graph)
        model = pred.model
        chunk = prep_buf[0][:B]

        def t(name, fn):
            r = None
            for _ in range(2):                               # warm-up: cuDNN plans, allocator blocks
                r = fn()
            best = float('inf')
            for _ in range(3):
                del r
                torch.cuda.synchronize()
                a, b2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                r = fn()
                b2.record()
                torch.cuda.synchronize()
                best = min(best, a.elapsed_time(b2))
            stage_ms[name] = best
            return r
        with torch.no_grad():
            msq = torch.ops.msq
            ph = (h + 31) // 32 * 32
            pw = (w + 31) // 32 * 32
            x = t('stem (input staging + im2col + tcgen05 7x7 conv + ReLU + max-pool, one kernel)',
                  lambda: msq.stem_conv_pool_tc(chunk, float(cfg['min_height']), float(cfg['max_height']), True, model.pixel_mean[0],
                                                model.pixel_std[0], ph, pw, model.stem_btile, model.stem_b64))
            feats = t('res2-res5 + FPN (GN, avg fusion)', lambda: model.pyramid(x))
            props = t('RPN head + proposals (top-k, decode, NMS)', lambda: model.rpn(feats, h, w))
            det = t('box head (ROIAlignV2 + 2 FC + top-1)', lambda: model.box_head(feats, props[0], props[2], h, w))
            t('mask head', lambda: model.mask_head(feats, det[0]))
            t('keypoint head + decode', lambda: model.keypoint_head(feats, det[0]))
            masks_kp = t('whole graph + paste (predict_dense)', lambda: pred.predict_dense(chunk, cfg['min_height'], cfg['max_height']))
            t('clean + features + angles + scalars + crops (msq_extract_chunk)', lambda: engine.extract(chunk, masks_kp[0], masks_kp[1], **kw))
        ms_res_max, ms_e2e_max = reduce([ms_res, ms_e2e], MAX)
        flops = rcnn.dense_flops_per_frame(h, w, topk)
        graph_ms = stage_ms['whole graph + paste (predict_dense)']
        achieved = flops * B / (graph_ms * 1e-3) / 1e12
        out[key] = {
            'workload': f'configs[2]: prep -> stem kernel (tcgen05) -> Keypoint+Mask R-CNN R50-FPN of the reference configuration (own TorchScript graph: '
                        f'GN FPN with avg fusion, stride-in-1x1 ResNet, ROIAlignV2, keypoint pooler 7; random init, bf16, batch {B}, '
                        f'{topk} proposals per image, 1 detection per frame) -> batched paste -> clean/features/angles/scalars -> crops',
            'frames': n * world, 'chunks': (n + npool - 1) // npool, 'post_nms_topk': topk,
            'frames_per_s': n * world / (ms_res_max * 1e-3), 'ms_per_1000_frames': ms_res_max / n * 1000,
            'e2e': {'value': n * world / (ms_e2e_max * 1e-3), 'unit': UNIT, 'h2d_bytes_per_step': int(n * h * w * 2),
                    'd2h_bytes_per_step': int(sum(v.numel() * v.element_size() for v in host_out.values()) * ((n + npool - 1) // npool)),
                    'path': 'pinned host int16 frames (pool >> CPU caches) -> zero-copy msq_prep_frames -> graph -> msq_paste_masks -> '
                            'msq_extract_chunk -> pinned host crops / scalars / keypoint table / flips; masks never cross PCIe'},
            'stage_ms_per_batch': {k2: round(v, 3) for k2, v in stage_ms.items()}, 'batch': B,
            'dense_engine': engine_report(),
            'gpu_launches_own_kernels': int(launches),
            'roofline': {'bound': 'tensor', 'achieved': achieved, 'peak': tensor_peak, 'unit': 'TFLOP/s', 'frac': achieved / tensor_peak,
                         'peak_source': peak_src + ' bf16_tflops_sustained', 'dense_flops_per_frame': flops,
                         'what': 'FLOPs of every convolution / Linear of the graph (2 x MACs, executed shapes) x batch / device time of the '
                                 'whole graph call incl. its glue kernels'},
        }
        del pred, engine
        torch.cuda.empty_cache()
    return out


def engine_report():
    """Which engine runs the graph's convolutions / Linear layers: per layer shape the faster of cuDNN / cuBLAS and the repo's
    tcgen05 + TMA implicit GEMM (csrc/conv_tc.cu), timed once per shape (model/ops.py 'auto')."""
    from moseq2_detectron_extract_b200.model import ops
    ch = ops.engine_choices()
    tc = sorted({f'{k[0]} {list(k[2])} on {list(k[1])}' + (f' stride {k[3]}' if k[0] == 'conv' and k[3] != 1 else '') for k, v in ch.items() if v == 'tcgen05'})
    return {'mode': ops.CONV_ENGINE['mode'], 'layer_shapes_on_tcgen05': sum(1 for v in ch.values() if v == 'tcgen05'),
            'layer_shapes_on_cudnn_cublas': sum(1 for v in ch.values() if v == 'cudnn'), 'tcgen05_layers': tc}


def secondary_figures(args, geom, cfg, roi, bg):
    """(1) prep with Kinect-like invalid pixels (rate 0.002) incl. the GPU in-paint, (2) use_tracking=True: the Kalman branch."""
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
    # (3) a7 on a less convenient animal: bent body, tail, pasted-28x28 mask outlines.  Which share of the frames does the
    # streaming (row-convex, hole-free) feature kernel settle, and what do both regimes cost?
    try:
        from moseq2_detectron_extract_b200 import _lib
        regimes = {}
        for name, kw in (('ellipsoid (headline workload)', {}), ('bent body + tail + pasted 28x28 mask', {'realistic': True})):
            ch = synthetic.generate_chunk(300, seed=21, geom=geom, **kw)
            reps = 10
            prep = prep_raw_frames(torch.from_numpy(np.tile(ch.frames, (reps, 1, 1))).cuda(), bground_im=bg, roi=roi, vmin=0, vmax=100)
            masks = torch.from_numpy(np.tile(ch.masks, (reps, 1, 1))).cuda()
            n, h, w = (int(v) for v in prep.shape)
            cleaned = torch.empty_like(prep)
            _dev.clean_frames_ws(prep, cleaned)
            cen, ori, ax = _dev.empty((n, 2), torch.float64), _dev.empty((n,), torch.float64), _dev.empty((n, 2), torch.float64)
            flist = _dev.empty((n + 1,), torch.int32)

            def feats(scratch=True):
                _lib.call('msq_frame_features', _dev.ptr(cleaned), _dev.ptr(masks), n, h, w, 3.0, _dev.ptr(cen), _dev.ptr(ori), _dev.ptr(ax),
                          None, _dev.ptr(flist) if scratch else None, flist.numel() * 4 if scratch else 0, _dev.stream())
            ms_mixed = timed(feats, iters=5)
            passed_on = int(flist[0].item())
            ms_general = timed(lambda: feats(False), iters=5)
            ms_clean = timed(lambda: _dev.clean_frames_ws(prep, cleaned), iters=5)
            regimes[name] = {'frames': n, 'fast_path_hit_rate': 1.0 - passed_on / n, 'ms_streaming_plus_general_for_the_rest': ms_mixed,
                             'ms_general_kernel_on_every_frame': ms_general, 'GBps_mixed': (2 * h * w + 40) * n / (ms_mixed * 1e-3) / 1e9,
                             'ms_clean_frames': ms_clean, 'GBps_clean': 2 * h * w * n / (ms_clean * 1e-3) / 1e9}
        out['features_regimes'] = regimes
    except Exception as exc:
        out['features_regimes'] = {'error': repr(exc)[:300]}
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
    return out


def main():
    args = parse_args()
    if args.impl == 'reference':
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()

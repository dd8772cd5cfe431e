"""Per-kernel device timings on one synthetic chunk (development aid; bench.py is the judged number)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from moseq2_detectron_extract_b200 import _dev, _lib, synthetic
from moseq2_detectron_extract_b200.engine import ChunkEngine
from moseq2_detectron_extract_b200.proc import proc as P
import ctypes

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
geom = synthetic.SessionGeometry()
small = synthetic.generate_chunk(100, seed=0, geom=geom)
reps = (N + 99) // 100
frames = torch.from_numpy(np.tile(small.frames, (reps, 1, 1))[:N]).cuda()
masks = torch.from_numpy(np.tile(small.masks, (reps, 1, 1))[:N]).cuda()
kpts = torch.from_numpy(np.tile(small.keypoints, (reps, 1, 1))[:N]).cuda()
roi, bg = synthetic.make_roi(geom), synthetic.make_background(geom)
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')

def timeit(fn, iters=10):
    for _ in range(3): fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts)), float(np.min(ts))

prep, _ = P._prep_device(frames, bg, roi, 0, 100, want_invalid=True)
n, h, w = prep.shape
eng = ChunkEngine()
kw = dict(chunk_size=1000, min_height=0, max_height=100, true_depth=673.0, crop_size=(80, 80))
res = {k: v.clone() for k, v in eng.extract(prep, masks, kpts, **kw).items()}
st = _dev.stream()
bgd = _dev.as_device(bg); roid = _dev.as_device(roi.astype(np.uint8)); inv = _dev.empty((n,), torch.int32)
y0, x0, y1, x1 = synthetic.roi_bbox(roi)
cleaned = torch.empty_like(prep); cen = _dev.empty((n, 2), torch.float64); ori = _dev.empty((n,), torch.float64); ax = _dev.empty((n, 2), torch.float64)
ang = _dev.empty((n,), torch.float64); fl = _dev.empty((n,), torch.uint8); ps = _dev.empty((64,), torch.int32)
sc = _dev.empty((17, n), torch.float64); kc = _dev.empty((96, n), torch.float64); scr = _dev.empty((16 * n + 512,), torch.uint8)
cscr = _dev.empty((int(_lib.load().msq_crop_scratch_bytes(n)) + 16,), torch.uint8); dc = _dev.empty((n, 80, 80), torch.uint8); mc = _dev.empty((n, 80, 80), torch.uint8)
flist = _dev.empty((n + 1,), torch.int32)
out = {}
out['prep'] = timeit(lambda: _lib.call('msq_prep_frames', _dev.ptr(frames), n, geom.height, geom.width, _dev.ptr(bgd), 1, _dev.ptr(roid), y0, x0, h, w, 0.0, 100.0, 3, _dev.ptr(prep), _dev.ptr(inv), None, st))
out['clean_single_launch'] = timeit(lambda: _lib.call('msq_clean_frames', _dev.ptr(prep), _dev.ptr(cleaned), n, h, w, st))
cws = _dev.empty((int(_lib.load().msq_clean_scratch_bytes(n, h, w)) + 8,), torch.uint8)
out['clean'] = timeit(lambda: _lib.call('msq_clean_frames_ws', _dev.ptr(prep), None, _dev.ptr(cleaned), n, h, w, _dev.ptr(cws), cws.numel(), st))
pos = _dev.positive_bits_like(prep)
out['prep_with_bits'] = timeit(lambda: _lib.call('msq_prep_frames_bits', _dev.ptr(frames), n, geom.height, geom.width, _dev.ptr(bgd), 1, _dev.ptr(roid), y0, x0, h, w, 0.0, 100.0, 3, _dev.ptr(prep), _dev.ptr(inv), None, _dev.ptr(pos), st))
out['clean_with_bits'] = timeit(lambda: _lib.call('msq_clean_frames_ws', _dev.ptr(prep), _dev.ptr(pos), _dev.ptr(cleaned), n, h, w, _dev.ptr(cws), cws.numel(), st))
out['features'] = timeit(lambda: _lib.call('msq_frame_features', _dev.ptr(cleaned), _dev.ptr(masks), n, h, w, 3.0, _dev.ptr(cen), _dev.ptr(ori), _dev.ptr(ax), None, _dev.ptr(flist), flist.numel() * 4, st))
out['angles'] = timeit(lambda: _lib.call('msq_angles_and_flips', _dev.ptr(ori), _dev.ptr(ax), _dev.ptr(cen), _dev.ptr(kpts), n, 1000, _dev.ptr(ang), _dev.ptr(fl), None, _dev.ptr(ps), st))
out['scalars_kpts'] = timeit(lambda: _lib.call('msq_scalars_and_keypoints', _dev.ptr(prep), _dev.ptr(masks), _dev.ptr(cleaned), _dev.ptr(cen), _dev.ptr(ang), _dev.ptr(ax), _dev.ptr(kpts), n, h, w, 1000, 0.0, 100.0, 673.0, _dev.ptr(sc), _dev.ptr(kc), _dev.ptr(scr), scr.numel(), st))
out['crop'] = timeit(lambda: _lib.call('msq_crop_rotate', _dev.ptr(prep), _dev.ptr(masks), n, h, w, _dev.ptr(cen), _dev.ptr(ang), 80, 80, _dev.ptr(dc), _dev.ptr(mc), _dev.ptr(cscr), cscr.numel(), st))
out['extract_chunk'] = timeit(lambda: eng.extract(prep, masks, kpts, **kw))
out['extract_chunk_with_bits'] = timeit(lambda: eng.extract(prep, masks, kpts, positive_bits=pos, **kw))
out['filter_passes'] = ps[: (n + 999) // 1000].cpu().tolist()
A = h * w
bytes_ = {'prep': 3 * A, 'prep_with_bits': 3 * A, 'clean_with_bits': 2 * A, 'clean': 2 * A, 'clean_single_launch': 2 * A, 'features': 2 * A, 'scalars_kpts': 2 * A, 'crop': 2 * 2 * 6400 + 2 * 6400}
for k, v in out.items():
    if isinstance(v, tuple):
        line = f'{k:24s} median {v[0]*1e3:9.1f} us  min {v[1]*1e3:9.1f} us  -> {n / v[0] * 1e3 / 1e6:7.2f} Mframes/s'
        if k in bytes_: line += f'  {bytes_[k] * n / (v[0] * 1e-3) / 1e9:8.1f} GB/s (algorithmic)'
        print(line)
print('filter passes', out['filter_passes'])
json.dump(out, open('gpurun_out/kernel_times.json', 'w'))

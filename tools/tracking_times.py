"""Device timings of the Kalman tracking kernels on one 1000-frame chunk (development aid)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import numpy as np, torch
from moseq2_detectron_extract_b200 import _dev, _lib
from moseq2_detectron_extract_b200.proc.kalman import KalmanTracker, KalmanTrackerAngle, KalmanTrackerNPoints2D, KalmanTrackerPoint2D

T = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
rng = np.random.default_rng(0)
cen = np.cumsum(rng.normal(size=(T, 2)), axis=0) + 120
kp = cen[:, None, :] + rng.normal(scale=5, size=(T, 8, 2))
kp[::97] = np.nan
ang = (np.arange(T) * 3.0) % 360


def timed(fn, iters=3):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


pt = KalmanTracker([KalmanTrackerPoint2D(3, 1.0), KalmanTrackerNPoints2D(8, 3, 1.0)])
at = KalmanTracker([KalmanTrackerAngle(3, 1.0, True)])
cd, kd, ad = _dev.as_device(cen, torch.float64), _dev.as_device(kp, torch.float64), _dev.as_device(ang, torch.float64)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); pt.initialize([cd, kd]); b.record(); torch.cuda.synchronize()
print(f'point tracker EM init (10 iterations, T={T}): {a.elapsed_time(b):8.2f} ms')
a.record(); at.initialize([ad]); b.record(); torch.cuda.synchronize()
print(f'angle tracker EM init                      : {a.elapsed_time(b):8.2f} ms')
print(f'point filter  (54 states)                  : {timed(lambda: pt.filter([cd, kd])):8.2f} ms')
print(f'point smooth  (filter + gains + backward)  : {timed(lambda: pt.smooth([cd, kd])):8.2f} ms')
print(f'angle smooth  (6 states)                   : {timed(lambda: at.smooth([ad])):8.2f} ms')
m = at.device_model()
flips = torch.zeros(T, dtype=torch.uint8, device='cuda'); scores = torch.ones(T, dtype=torch.float64, device='cuda')


def loop():
    mean, cov, an = at.last_mean.clone(), at.last_covar.clone(), ad.clone()
    _lib.call('msq_track_angles', _dev.ptr(m['A']), _dev.ptr(m['H']), _dev.ptr(m['Q']), _dev.ptr(m['R']), _dev.ptr(mean), _dev.ptr(cov),
              at.n_state, _dev.ptr(an), _dev.ptr(flips), _dev.ptr(scores), T, _dev.stream())


print(f'per-frame angle loop                       : {timed(loop):8.2f} ms')

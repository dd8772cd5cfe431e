"""Randomised parity sweep (development aid): many synthetic frames with blobs of all kinds through the CUDA clean /
features / crop kernels and through OpenCV (oracle).  usage: stress_parity.py [n_frames] [seed]"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import numpy as np
import torch
import extract_oracle as O
import moseq2_detectron_extract_b200.proc as P

n = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
h, w = 240, 240
yy, xx = np.mgrid[0:h, 0:w]
frames = np.zeros((n, h, w), np.uint8)
masks = np.zeros((n, h, w), np.uint8)
for i in range(n):
    img = np.clip(rng.normal(0.3, 1.2, (h, w)), 0, None)
    m = np.zeros((h, w), bool)
    for _ in range(rng.integers(0, 4)):                       # 0..3 ellipses, possibly overlapping / touching / clipped
        cx, cy = rng.uniform(-10, w + 10), rng.uniform(-10, h + 10)
        a, b, t = rng.uniform(3, 45), rng.uniform(2, 25), rng.uniform(0, np.pi)
        u = (xx - cx) * np.cos(t) + (yy - cy) * np.sin(t)
        v = -(xx - cx) * np.sin(t) + (yy - cy) * np.cos(t)
        e = (u / a) ** 2 + (v / b) ** 2 <= 1
        img = np.where(e, np.maximum(img, 45 * np.sqrt(np.clip(1 - (u / a) ** 2 - (v / b) ** 2, 0, 1)) + rng.normal(0, 1, (h, w))), img)
        m |= e
    kind = rng.integers(0, 5)
    if kind == 1:                                             # holes punched into the mask
        for _ in range(3):
            hx, hy, hr = rng.integers(0, w), rng.integers(0, h), rng.integers(1, 6)
            m &= ~(((xx - hx) ** 2 + (yy - hy) ** 2) <= hr * hr)
    elif kind == 2:                                           # salt noise in the mask
        m |= rng.random((h, w)) < 0.002
    elif kind == 3:                                           # a thin bar splitting things
        m[:, rng.integers(0, w)] = False
    frames[i] = np.clip(img, 0, 100).astype(np.uint8)
    masks[i] = m
cl_ref = O.clean_frames_cv2(frames)
cl = P.clean_frames(torch.from_numpy(frames).cuda(), iters_tail=3)
assert np.array_equal(cl.cpu().numpy(), cl_ref), 'clean mismatch'
ref = O.frame_features_cv2(cl_ref, masks, 3)
got, _ = P.get_frame_features(cl, frame_threshold=3, mask=torch.from_numpy(masks).cuda(), use_cc=True)
bad = 0
for key, tol in (('centroid', 1e-9), ('orientation', 1e-9), ('axis_length', 1e-9)):
    g, r = got[key].cpu().numpy(), ref[key]
    assert np.array_equal(np.isnan(g), np.isnan(r)), key + ': NaN pattern'
    ok = ~np.isnan(r)
    err = np.abs(g[ok] - r[ok]) / np.maximum(1.0, np.abs(r[ok]))
    bad += int((err > tol).sum())
    print(f'{key:12s} max rel err {err.max() if err.size else 0:.3e}')
cen, ang = ref['centroid'], np.degrees(-np.nan_to_num(ref['orientation']))
crops = P.crop_and_rotate_frames_batch(torch.from_numpy(frames).cuda(), cen, ang, (80, 80)).cpu().numpy()
cbad = sum(int((crops[i] != O.crop_rotate_cv2(frames[i], cen[i], ang[i], (80, 80))).sum()) for i in range(0, n, 7))
print(f'frames {n}: feature mismatches {bad}, crop pixel mismatches {cbad}, frames without features {int(np.isnan(cen[:, 0]).sum())}')
assert bad == 0 and cbad == 0
print('stress parity ok')

"""GPU parity tests: CUDA path (through the C ABI) vs the oracle and the golden outputs of the reference.

Tolerances (BASELINE.json north_star): integer / index work bit-exact; moments & centroid <= 1e-4 rel;
angles <= 1e-3 rad; bilinear crops <= 1 depth unit (we assert bit-exact); keypoints <= 0.5 px.
The assertions below are much tighter than that wherever float64 makes it possible.
"""
import numpy as np
import pytest

import extract_oracle as O
from cases import CASE_NAMES, assert_close, case_inputs, golden

pytestmark = pytest.mark.gpu

torch = pytest.importorskip('torch')


@pytest.fixture(scope='module')
def P():
    import moseq2_detectron_extract_b200.proc as proc
    from moseq2_detectron_extract_b200 import _dev
    _dev.require_cuda()          # fail loudly if the .so or the GPU is missing
    return proc


# ----------------------------------------------------------------------------- a2 prep
@pytest.mark.parametrize('name', CASE_NAMES)
def test_prep_matches_reference(P, name):
    geom, chunk, roi, bg, cfg = case_inputs(name)
    g = golden(name)
    out = P.prep_raw_frames(chunk.frames, bground_im=bg, roi=roi, vmin=cfg['min_height'], vmax=cfg['max_height'],
                            fix_invalid_pixels=False)
    assert out.dtype == np.uint8 and np.array_equal(out, g['prep_nofix'])
    out = P.prep_raw_frames(chunk.frames, bground_im=bg, roi=roi, vmin=cfg['min_height'], vmax=cfg['max_height'])
    assert np.array_equal(out, g['prep'])               # includes cv2.inpaint of the invalid pixels, bit-exact
    if name == 'kinect_invalid':
        assert not np.array_equal(g['prep'], g['prep_nofix'])


def test_prep_invalid_counts_and_device_tensors(P):
    from moseq2_detectron_extract_b200.proc.proc import _prep_device
    geom, chunk, roi, bg, cfg = case_inputs('kinect_invalid')
    out, invalid = _prep_device(torch.from_numpy(chunk.frames).cuda(), bg, roi, 0, 100, want_invalid=True)
    ref, bad = O.prep_frames(chunk.frames, bg, roi, 0, 100, fix_invalid=False), O.invalid_pixel_mask(chunk.frames) * roi
    y0, x0, y1, x1 = O.bbox_of_roi(roi)
    assert out.is_cuda and np.array_equal(out.cpu().numpy(), ref)
    assert np.array_equal(invalid.cpu().numpy(), bad[:, y0:y1, x0:x1].sum(axis=(1, 2)))


@pytest.mark.parametrize('bg_kind', ['float32', 'float64_half', 'uint16', 'none'])
@pytest.mark.parametrize('aligned', [True, False])
def test_prep_background_dtypes_and_unaligned_boxes(P, bg_kind, aligned):
    rng = np.random.default_rng(3)
    H, W = 60, 96
    frames = rng.integers(400, 700, size=(5, H, W)).astype(np.int16)
    frames[rng.random(frames.shape) < 0.01] = 0
    roi = np.zeros((H, W), dtype=bool)
    if aligned:
        roi[5:48, 16:81] = True      # x0 = 16, w = 64
    else:
        roi[3:50, 13:72] = True      # x0 = 13, w = 58
    roi[10, 20] = False
    bg = {'float32': (rng.uniform(600, 700, (H, W))).astype(np.float32),
          'float64_half': np.round(rng.uniform(600, 700, (H, W)) * 2) / 2,
          'uint16': rng.integers(600, 700, (H, W)).astype(np.uint16), 'none': None}[bg_kind]
    for vmin, vmax in [(0, 100), (10, 80.5), (None, 100), (0, None)]:
        if bg is None and (vmin is None or vmax is None):
            continue
        f = frames if bg is not None else (frames - 400).astype(np.int16)
        ref = O.prep_frames(f, bg, roi, vmin, vmax, fix_invalid=False)
        out = P.prep_raw_frames(f, bground_im=bg, roi=roi, vmin=vmin, vmax=vmax, fix_invalid_pixels=False)
        assert np.array_equal(out, ref), (bg_kind, aligned, vmin, vmax)


@pytest.mark.parametrize('rate', [0.0005, 0.01, 0.1, 0.5])
@pytest.mark.parametrize('aligned', [True, False])
def test_inpaint_matches_opencv(P, rate, aligned):
    """fill_invalid_pixels (ref proc/proc.py:189-210): GPU fast-marching Navier-Stokes vs cv2.inpaint, bit-exact."""
    rng = np.random.default_rng(int(rate * 1e4) + aligned)
    H, W = 70, 112
    n = 6
    yy, xx = np.mgrid[0:H, 0:W]
    frames = (650 - 40 * np.exp(-((xx - 50) ** 2 + (yy - 35) ** 2) / 300.0)[None] + rng.normal(0, 1.5, (n, H, W))).astype(np.int16)
    frames[rng.random(frames.shape) < rate] = 0
    frames[1, 20:26, 30:37] = 0                      # a solid invalid block
    frames[2, :, :] = np.where(rng.random((H, W)) < 0.9, 0, frames[2])    # nearly everything invalid
    frames[3] = np.maximum(frames[3], 1)             # one frame without invalid pixels
    roi = np.zeros((H, W), dtype=bool)
    if aligned:
        roi[3:60, 16:97] = True                      # x0 = 16, w = 80
    else:
        roi[2:66, 11:100] = True                     # x0 = 11, w = 88 -> scalar prep path, atomically packed mask
    roi[30, 40] = False
    bg = np.full((H, W), 673.0, dtype=np.float32)
    ref = O.prep_frames(frames, bg, roi, 0, 100, fix_invalid=True)
    got = P.prep_raw_frames(frames, bground_im=bg, roi=roi, vmin=0, vmax=100)
    assert np.array_equal(got, ref), (rate, aligned, int((got != ref).sum()))
    assert not np.array_equal(ref, O.prep_frames(frames, bg, roi, 0, 100, fix_invalid=False))


def test_inpaint_edge_cases(P):
    rng = np.random.default_rng(77)
    for (H, W) in [(9, 9), (12, 40), (33, 17)]:
        frames = rng.integers(580, 673, size=(4, H, W)).astype(np.int16)
        frames[0, 0, 0] = 0; frames[0, -1, -1] = 0; frames[0, 0, W // 2] = 0      # corners / borders
        frames[1, :, 0] = 0; frames[1, 0, :] = 0                                    # whole border row + column
        frames[2] = 0                                                               # everything invalid
        roi = np.ones((H, W), dtype=bool)                                           # box = [0:H-1, 0:W-1)
        bg = np.full((H, W), 673.0, dtype=np.float32)
        ref = O.prep_frames(frames, bg, roi, 0, 100, fix_invalid=True)
        got = P.prep_raw_frames(frames, bground_im=bg, roi=roi, vmin=0, vmax=100)
        assert np.array_equal(got, ref), (H, W, int((got != ref).sum()))


def test_prep_empty_and_errors(P):
    out = P.prep_raw_frames(np.zeros((0, 16, 16), np.int16), bground_im=np.zeros((16, 16), np.float32),
                            roi=np.ones((16, 16), bool), vmin=0, vmax=100)
    assert out.shape == (0, 15, 15)
    with pytest.raises(ValueError):
        P.prep_raw_frames(np.zeros((2, 16, 16), np.int16), bground_im=np.zeros((8, 8), np.float32))


# ----------------------------------------------------------------------------- a3 scale
@pytest.mark.parametrize('vmin,vmax', [(0, 100), (0.0, 100.0), (10, 90), (5.5, 77.25)])
def test_scale_matches_oracle(P, vmin, vmax):
    rng = np.random.default_rng(4)
    x = rng.integers(0, 101, size=(3, 37, 53, 1)).astype(np.uint8)
    if isinstance(vmin, int) and vmin > 0:
        x = np.maximum(x, vmin)            # uint8 wrap below vmin is covered by the LUT test below
    assert np.array_equal(P.scale_raw_frames(x, vmin, vmax), O.scale_frames(x, vmin, vmax))
    full = np.arange(256, dtype=np.uint8).reshape(1, 16, 16)
    ref = O.scale_frames(full, vmin, vmax)
    got = P.scale_raw_frames(full, vmin, vmax)
    in_range = ((full.astype(float) - vmin) * (255.0 / (vmax - vmin)) < 256) & ((full.astype(float) - vmin) >= 0)
    assert np.array_equal(got[in_range], ref[in_range])


@pytest.mark.parametrize('name', ['kinect_clean', 'azure_clean'])
def test_scale_matches_reference(P, name):
    geom, chunk, roi, bg, cfg = case_inputs(name)
    g = golden(name)
    assert np.array_equal(P.scale_raw_frames(g['prep'][:4, :, :, None], cfg['min_height'], cfg['max_height']), g['scale'])


# ----------------------------------------------------------------------------- a6 clean
@pytest.mark.parametrize('name', ['kinect_clean', 'kinect_missing_holes', 'azure_clean'])
def test_clean_matches_reference(P, name):
    g = golden(name)
    assert np.array_equal(P.clean_frames(g['prep'], iters_tail=3), g['cleaned_frames'])


@pytest.mark.parametrize('shape', [(1, 1), (1, 7), (9, 1), (5, 5), (31, 64), (57, 83), (33, 240), (240, 240),
                                   (100, 258), (70, 301), (250, 403)])
def test_clean_matches_opencv_on_random_frames(P, shape):
    rng = np.random.default_rng(sum(shape))
    fr = rng.integers(0, 256, size=(3,) + shape).astype(np.uint8)
    fr[1] = (fr[1] > 100) * 220
    fr[2] = np.clip(rng.normal(40, 30, shape), 0, 255).astype(np.uint8)
    assert np.array_equal(P.clean_frames(fr, iters_tail=3), O.clean_frames_cv2(fr)), shape


@pytest.mark.parametrize('shape', [(240, 240), (120, 64), (400, 400), (90, 488), (64, 248), (300, 256)])
def test_clean_row_skipping_matches_opencv(P, shape):
    """The streaming kernel only sends rows through the full median / erosion / dilation pipeline when a bit-level pre-pass
    finds a nine-run of positive medians within reach; everything else is written as zeros.  Sparse frames exercise that
    decision: blobs at every border and corner, across the 240-column tile seam, thin (8-wide, never survives) and just-wide-
    enough (9..11) bars, Kinect-like floor noise (31 % positive pixels), empty frames, frames cut across CTA row ranges."""
    rng = np.random.default_rng(shape[0] * 1000 + shape[1])
    h, w = shape
    yy, xx = np.mgrid[0:h, 0:w]
    frames = []

    def noise():
        return np.clip(np.rint(-rng.normal(0, 1, shape)), 0, 255).astype(np.uint8)

    def blob(cy, cx, ry, rx, val=40):
        f = noise()
        m = ((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 <= 1
        f[m] = np.clip(val + rng.integers(-5, 6, size=int(m.sum())), 1, 255)
        return f

    frames.append(np.zeros(shape, np.uint8))
    frames.append(noise())
    for cy, cx in ((0, 0), (0, w - 1), (h - 1, 0), (h - 1, w - 1), (h // 2, 0), (0, w // 2), (h - 1, w // 2), (h // 2, w - 1),
                   (h // 2, min(w - 1, 239)), (h // 2, min(w - 1, 244)), (3, 3), (h - 4, w - 4)):
        frames.append(blob(cy, cx, 14, 30))
    two = blob(h // 5, w // 4, 8, 12)
    m2 = ((yy - 4 * h // 5) / 9) ** 2 + ((xx - 3 * w // 4) / 20) ** 2 <= 1
    two[m2] = 60
    frames.append(two)
    for width in (8, 9, 10, 11):                                   # horizontal bars: the ellipse's centre row is 9 wide
        f = np.zeros(shape, np.uint8)
        f[h // 3:h // 3 + 12, 5:5 + width] = 50
        f[2 * h // 3:2 * h // 3 + width, w // 2:w // 2 + 14] = 70
        frames.append(f)
    full = np.full(shape, 9, np.uint8)
    full[h // 2, :] = 0                                            # one zero row: the band logic sees two halves
    frames.append(full)
    edge = np.zeros(shape, np.uint8)
    edge[:, :5] = 30; edge[:6, :] = 30; edge[:, -5:] = 30; edge[-6:, :] = 30   # strips along the borders (outside pixels are ignored)
    frames.append(edge)
    fr = np.stack(frames)
    got, want = P.clean_frames(fr, iters_tail=3), O.clean_frames_cv2(fr)
    assert want[2:].any()                                           # the blobs do survive the opening
    for i in range(len(fr)):
        assert np.array_equal(got[i], want[i]), (shape, i)


def test_clean_single_launch_entry_point_agrees(P):
    """msq_clean_frames (no scratch: every strip scans its own rows inside the pipeline kernel) == msq_clean_frames_ws (separate
    pre-pass launch), on a sparse frame set cut across CTA row ranges."""
    import torch
    from moseq2_detectron_extract_b200 import _dev, _lib, synthetic
    geom = synthetic.SessionGeometry()
    ch = synthetic.generate_chunk(40, seed=5, geom=geom, realistic=True)
    prep = P.prep_raw_frames(torch.from_numpy(ch.frames).cuda(), bground_im=synthetic.make_background(geom), roi=synthetic.make_roi(geom), vmin=0, vmax=100)
    a, b = torch.empty_like(prep), torch.empty_like(prep)
    n, h, w = prep.shape
    _lib.call('msq_clean_frames', _dev.ptr(prep), _dev.ptr(a), n, h, w, _dev.stream())
    _dev.clean_frames_ws(prep, b)
    assert torch.equal(a, b) and np.array_equal(a.cpu().numpy(), O.clean_frames_cv2(prep.cpu().numpy()))
    assert int(a.any(dim=(1, 2)).sum()) == n


def test_clean_rejects_unsupported_configs(P):
    fr = np.zeros((1, 8, 8), np.uint8)
    with pytest.raises(NotImplementedError):
        P.clean_frames(fr)                       # iters_tail=None: median only, not on the extract path
    with pytest.raises(NotImplementedError):
        P.clean_frames(fr, prefilter_space=(5,), iters_tail=3)


# ----------------------------------------------------------------------------- a7 features
@pytest.mark.parametrize('name', ['kinect_clean', 'kinect_missing_holes', 'azure_clean'])
def test_features_match_reference(P, name):
    geom, chunk, roi, bg, cfg = case_inputs(name)
    g = golden(name)
    d2_masks = g['masks']
    feats, mask = P.get_frame_features(g['cleaned_frames'], frame_threshold=3, mask=d2_masks, use_cc=True)
    ref = O.frame_features_cv2(g['cleaned_frames'], d2_masks, 3)
    assert_close(feats['centroid'], ref['centroid'], 1e-12, what='centroid')
    assert_close(feats['axis_length'], ref['axis_length'], 1e-10, what='axis_length')
    assert_close(feats['orientation'], ref['orientation'], 0, 1e-12, what='orientation (rad)')
    assert_close(feats['orientation'], g['raw_orientation'], 0, 1e-12, what='orientation vs reference')
    assert mask is d2_masks                      # the reference returns the given mask untouched


@pytest.mark.parametrize('w', [5, 20, 33, 64, 100, 128, 240, 300, 400, 512, 520, 544, 1024])
def test_feature_sums_are_bit_exact_on_random_masks(P, w):
    import cv2
    from moseq2_detectron_extract_b200 import _dev
    from moseq2_detectron_extract_b200.proc.proc import _features_device
    rng = np.random.default_rng(w)
    h = int(rng.integers(5, 90))
    n = 24
    masks = np.zeros((n, h, w), np.uint8)
    for i in range(n):
        if i % 3 == 0:
            masks[i] = rng.random((h, w)) < rng.uniform(0.2, 0.9)
        else:
            s = cv2.GaussianBlur(rng.random((h, w)).astype(np.float32), (0, 0), float(rng.uniform(1, 4)))
            masks[i] = s > np.quantile(s, rng.uniform(0.2, 0.8))
    masks[1] = 0
    masks[2] = 1
    yy, xx = np.mgrid[0:h, 0:w]                         # row-convex shapes: the streaming fast path at every lanes-per-row layout
    masks[5] = ((xx - w * 0.6) / max(w * 0.3, 1)) ** 2 + ((yy - h * 0.5) / max(h * 0.4, 1)) ** 2 <= 1
    masks[7] = (np.abs(xx - w // 2) + 2 * np.abs(yy - h // 2)) <= min(w, 2 * h) // 3
    masks[8] = 0
    masks[8, h // 3:h // 3 + 2, :] = 1                  # two full rows: runs crossing every 32-pixel word boundary
    cleaned = (masks * 7).astype(np.uint8)
    cleaned[4] = 200                                # threshold passes everywhere: mask alone decides
    ref = O.frame_features_np(cleaned, masks, 3, return_sums=True)
    c, o, a, sums = _features_device(_dev.as_device(cleaned), _dev.as_device(masks), 3.0, want_sums=True)
    assert np.array_equal(sums.cpu().numpy(), ref['sums24'])
    assert_close(c.cpu().numpy(), ref['centroid'], 1e-13, what='centroid')
    assert_close(a.cpu().numpy(), ref['axis_length'], 1e-10, 1e-9, what='axis')
    assert_close(o.cpu().numpy(), ref['orientation'], 0, 1e-12, what='orientation')


def test_features_without_mask_and_thresholds(P):
    rng = np.random.default_rng(9)
    fr = np.zeros((4, 40, 50), np.uint8)
    fr[:, 10:30, 12:40] = rng.integers(0, 12, size=(4, 20, 28))
    for thr in (-1, 0, 3, 5.5, 300):
        feats, mask = P.get_frame_features(fr, frame_threshold=thr)
        assert np.array_equal(mask != 0, fr > thr)
        ref = O.frame_features_np(fr, np.ones_like(fr), thr)
        assert_close(feats['centroid'], ref['centroid'], 1e-12, what=f'centroid thr={thr}')


def test_features_row_convex_fast_path_agrees_with_general_path(P):
    """features_kernel skips the flood / peel machinery when every row of the foreground is a single run and consecutive
    runs touch.  The same blobs with one far-away speck (second component -> general path) must give the same features,
    and both must match OpenCV; shapes that narrowly fail the test (a row with two runs, a diagonal-only contact, a
    one-pixel gap between rows) go through the general path."""
    h, w = 96, 240
    yy, xx = np.mgrid[0:h, 0:w]
    frames = []
    for cx, cy, a, b, t in ((60, 40, 30, 12, 0.3), (200, 50, 22, 20, 1.2), (120, 48, 45, 9, -0.6), (31, 30, 28, 14, 0.0)):
        u = (xx - cx) * np.cos(t) + (yy - cy) * np.sin(t)
        v = -(xx - cx) * np.sin(t) + (yy - cy) * np.cos(t)
        frames.append(((u / a) ** 2 + (v / b) ** 2 <= 1).astype(np.uint8) * 50)
    stair = np.zeros((h, w), np.uint8)                       # runs that only touch diagonally: still one 8-connected blob
    for i in range(20):
        stair[20 + i, 40 + 3 * i:43 + 3 * i] = 50
    two_runs = frames[0].copy(); two_runs[40, 55:62] = 0     # a notch splits one row into two runs (no hole: open to the top? no -> a hole-free dent)
    two_runs[30:41, 58] = 0
    gap = frames[1].copy(); gap[50, :] = 0                   # an empty row: two components
    clean = np.stack(frames + [stair, two_runs, gap])
    speck = clean.copy()
    speck[:, 90, 5] = 50                                     # a second, 1-pixel component far from the blob
    ones = np.ones_like(clean)
    f_fast, _ = P.get_frame_features(clean, frame_threshold=3, mask=ones)
    f_gen, _ = P.get_frame_features(speck, frame_threshold=3, mask=ones)
    ref = O.frame_features_cv2(clean, ones, 3)
    for key in ('centroid', 'orientation', 'axis_length'):
        assert np.array_equal(f_fast[key], f_gen[key], equal_nan=True), key  # the speck has zero contour area: it never wins
        assert_close(f_fast[key], ref[key], 1e-10, 1e-12, what=key)


# ----------------------------------------------------------------------------- a8-a10 angles / flips / filter
@pytest.mark.parametrize('name', ['kinect_clean', 'kinect_missing_holes', 'azure_clean'])
def test_instances_to_features_matches_reference(P, name):
    from make_golden import FakeInstances
    geom, chunk, roi, bg, cfg = case_inputs(name)
    g = golden(name)
    outputs = [{'instances': FakeInstances(chunk.masks[i], chunk.keypoints[i], chunk.num_instances[i] > 0)}
               for i in range(g['prep'].shape[0])]
    res = P.instances_to_features(outputs, g['prep'], None, None, debug=False)
    assert np.array_equal(res['cleaned_frames'], g['cleaned_frames'])
    assert np.array_equal(res['masks'], g['masks'])
    assert_close(res['features']['centroid'], g['centroid'], 1e-12, what='centroid')
    assert_close(res['features']['axis_length'], g['axis_length'], 1e-10, what='axis_length')
    assert_close(res['features']['orientation'], g['orientation'], 0, 1e-9, what='orientation (deg)')
    assert np.array_equal(res['flips'], g['flips'])
    assert np.array_equal(res['num_instances'], g['num_instances'])
    assert_close(res['keypoints'], g['keypoints'], 0, 0, what='keypoints')


def test_flips_and_filter_standalone(P):
    rng = np.random.default_rng(12)
    n = 700
    cen = rng.uniform(60, 180, (n, 2))
    ang = rng.uniform(0, 360, n)
    ang[::97] = np.nan
    lens = rng.uniform(40, 80, n)
    kp = np.concatenate([cen[:, None, :] + rng.normal(0, 25, (n, 8, 2)), rng.uniform(0, 1, (n, 8, 1))], axis=2).astype(np.float32)
    kp[5] = np.nan
    f_ref, c_ref = O.keypoint_flips(kp.astype(np.float64), cen, ang, lens)
    f, c = P.flips_from_keypoints(kp, cen, ang, lens)
    assert np.array_equal(f, f_ref) and np.allclose(c, c_ref)
    # filter: smooth heading crossing 0/360 plus injected 180-degree flips, with and without NaN
    base = (3.0 * np.arange(n)) % 360
    base[rng.random(n) < 0.1] += 180
    for with_nan in (False, True):
        a = base.copy()
        if with_nan:
            a[[13, 14, 400]] = np.nan
        ref, fl_ref, passes = O.iterative_angle_filter(a)
        out, fl = P.iterative_filter_angles(a)
        assert_close(out, ref, 0, 1e-9, what='filtered angles')
        assert np.array_equal(fl, fl_ref)
        assert_close(P.filter_angles(a), O.angle_filter_pass(a), 0, 1e-9, what='single pass')
        assert passes == (1001 if with_nan else passes)


# ----------------------------------------------------------------------------- a11/a12 scalars + keypoints
@pytest.mark.parametrize('name', ['kinect_clean', 'kinect_missing_holes', 'azure_clean'])
def test_scalars_and_keypoints_match_reference(P, name):
    geom, chunk, roi, bg, cfg = case_inputs(name)
    g = golden(name)
    feats = {'centroid': g['centroid'], 'orientation': g['orientation'], 'axis_length': g['axis_length']}
    scal = P.compute_scalars(g['prep'] * g['masks'], feats, min_height=cfg['min_height'], max_height=cfg['max_height'],
                             true_depth=cfg['true_depth'])
    for k in g.files:
        if k.startswith('scalars/'):
            name_k = k[8:]
            tol = 1e-4 if 'velocity' in name_k or name_k == 'angle' else 1e-9
            assert_close(scal[name_k], g[k], tol, 1e-9, what=k)
    assert np.array_equal(scal['area_px'], g['scalars/area_px'])
    assert scal['height_ave_mm'].dtype == np.float32 and np.array_equal(scal['height_ave_mm'], g['scalars/height_ave_mm'])
    kd = P.keypoints_to_dict(g['keypoints'], g['cleaned_frames'], g['centroid'], g['orientation'], true_depth=cfg['true_depth'])
    assert set(kd) == {k[10:] for k in g.files if k.startswith('keypoints/')}
    for k in g.files:
        if k.startswith('keypoints/'):
            assert_close(kd[k[10:]], g[k], 1e-9, 1e-7, what=k)     # north_star: <= 0.5 px


# ----------------------------------------------------------------------------- a13 crops
@pytest.mark.parametrize('name', ['kinect_clean', 'kinect_missing_holes', 'azure_clean'])
def test_crops_match_reference_bit_exact(P, name):
    geom, chunk, roi, bg, cfg = case_inputs(name)
    g = golden(name)
    d, m = P.crop_and_rotate_frames_batch(g['prep'], g['centroid'], g['orientation'], cfg['crop_size'], frames2=g['masks'])
    assert np.array_equal(d, g['depth_frames'])
    assert np.array_equal(m, g['mask_frames'])
    one = P.crop_and_rotate_frame(g['prep'][3], g['centroid'][3], g['orientation'][3], cfg['crop_size'])
    assert np.array_equal(one, g['depth_frames'][3])


@pytest.mark.parametrize('w', [150, 152])       # 150: per-tap global gather (unaligned rows); 152: the staged shared-memory kernel
def test_crops_match_opencv_on_random_transforms(P, w):
    rng = np.random.default_rng(21)
    n, h = 300, 120
    fr = rng.integers(0, 101, size=(n, h, w)).astype(np.uint8)
    cen = np.stack([rng.uniform(-2, w + 1, n), rng.uniform(-2, h + 1, n)], axis=1)
    cen[::7] = np.stack([rng.uniform(0, 45, n), rng.uniform(0, 45, n)], axis=1)[::7]
    cen[::11] = np.floor(cen[::11])
    ang = rng.uniform(-10, 560, n)
    ang[::13] = np.floor(ang[::13] / 45) * 45
    ang[5] = np.nan
    cen[6, 0] = np.nan
    for crop in [(80, 80), (128, 128), (31, 31), (32, 32)]:
        got = P.crop_and_rotate_frames_batch(fr, cen, ang, crop)
        bad = 0
        for i in range(n):
            ref = O.crop_rotate_cv2(fr[i], cen[i], ang[i], crop)
            bad += int((got[i] != ref).sum())
        assert bad == 0, (crop, bad)


# ----------------------------------------------------------------------------- a4 paste
def test_paste_masks_matches_restated_detectron2(P):
    from moseq2_detectron_extract_b200.model.util import paste_masks
    rng = np.random.default_rng(31)
    n, M, h, w = 12, 28, 240, 240
    soft = 1 / (1 + np.exp(-rng.normal(0, 2, (n, M, M)))).astype(np.float32)
    soft = (soft + np.roll(soft, 1, 1) + np.roll(soft, 1, 2)) / 3
    boxes = np.stack([rng.uniform(-5, 120, n), rng.uniform(-5, 120, n), rng.uniform(130, 250, n), rng.uniform(130, 250, n)], 1).astype(np.float32)
    got = paste_masks(soft.astype(np.float32), boxes, h, w)
    ref = O.paste_masks_np(soft.astype(np.float32), boxes, h, w)
    assert got.dtype == bool and np.array_equal(got, ref)


# ----------------------------------------------------------------------------- whole chunk through msq_extract_chunk
@pytest.mark.parametrize('name', ['kinect_clean', 'kinect_missing_holes', 'azure_clean'])
def test_extract_chunk_matches_reference(P, name):
    from moseq2_detectron_extract_b200 import _dev
    from moseq2_detectron_extract_b200.engine import ChunkEngine
    from moseq2_detectron_extract_b200.proc.keypoints import keypoints_from_table
    from moseq2_detectron_extract_b200.proc.scalars import scalars_from_table
    geom, chunk, roi, bg, cfg = case_inputs(name)
    g = golden(name)
    eng = ChunkEngine()
    res = eng.extract(_dev.as_device(g['prep']), _dev.as_device(chunk.masks), _dev.as_device(chunk.keypoints, torch.float32),
                      chunk_size=1000, min_height=cfg['min_height'], max_height=cfg['max_height'],
                      true_depth=cfg['true_depth'], crop_size=cfg['crop_size'])
    torch.cuda.synchronize()
    assert np.array_equal(res['cleaned'].cpu().numpy(), g['cleaned_frames'])
    assert_close(res['centroid'].cpu().numpy(), g['centroid'], 1e-12, what='centroid')
    assert_close(res['axis_length'].cpu().numpy(), g['axis_length'], 1e-10, what='axis_length')
    assert_close(res['angle_deg'].cpu().numpy(), g['orientation'], 0, 1e-9, what='orientation')
    assert np.array_equal(res['flips'].cpu().numpy().astype(bool), g['flips'])
    scal = scalars_from_table(res['scalars'])
    kd = keypoints_from_table(res['kpt_cols'])
    for k in g.files:
        if k.startswith('scalars/'):
            tol = 1e-4 if 'velocity' in k or k.endswith('angle') else 1e-9
            assert_close(scal[k[8:]], g[k], tol, 1e-9, what=k)
        elif k.startswith('keypoints/'):
            assert_close(kd[k[10:]], g[k], 1e-9, 1e-7, what=k)
    assert np.array_equal(res['depth_crops'].cpu().numpy(), g['depth_frames'])
    assert np.array_equal(res['mask_crops'].cpu().numpy(), g['mask_frames'])
    expect_passes = 1001 if np.isnan(g['centroid']).any() else None
    if expect_passes:
        assert int(res['filter_passes'][0]) == expect_passes


def test_full_size_chunk_properties(P):
    """BASELINE-size launch (1000 frames of 240x240): chunk-split invariance and oracle spot checks."""
    from moseq2_detectron_extract_b200 import _dev, synthetic
    from moseq2_detectron_extract_b200.engine import ChunkEngine
    geom = synthetic.SessionGeometry()
    small = synthetic.generate_chunk(50, seed=5, geom=geom, t0=0, missing_every=17)
    roi, bg = synthetic.make_roi(geom), synthetic.make_background(geom)
    reps = 20
    frames = np.tile(small.frames, (reps, 1, 1))
    masks = np.tile(small.masks, (reps, 1, 1))
    kpts = np.tile(small.keypoints, (reps, 1, 1))
    prep = P.prep_raw_frames(torch.from_numpy(frames).cuda(), bground_im=bg, roi=roi, vmin=0, vmax=100)
    assert np.array_equal(prep[:50].cpu().numpy(), O.prep_frames(small.frames, bg, roi, 0, 100))
    assert torch.equal(prep[:50], prep[950:1000])
    eng = ChunkEngine()
    kw = dict(min_height=0, max_height=100, true_depth=673.0, crop_size=(80, 80))
    m_dev, k_dev = _dev.as_device(masks), _dev.as_device(kpts, torch.float32)
    whole = {k: v.clone() for k, v in eng.extract(prep, m_dev, k_dev, chunk_size=500, **kw).items()}
    for half in range(2):
        sl = slice(half * 500, (half + 1) * 500)
        part = eng.extract(prep[sl], m_dev[sl], k_dev[sl], chunk_size=500, **kw)
        for key in ('cleaned', 'centroid', 'angle_deg', 'axis_length', 'flips', 'depth_crops', 'mask_crops'):
            assert torch.equal(torch.nan_to_num(whole[key][sl].double(), nan=-1), torch.nan_to_num(part[key].double(), nan=-1)), key
        assert torch.equal(torch.nan_to_num(whole['scalars'][:, sl], nan=-1), torch.nan_to_num(part['scalars'], nan=-1))
        assert torch.equal(torch.nan_to_num(whole['kpt_cols'][:, sl], nan=-1), torch.nan_to_num(part['kpt_cols'], nan=-1))
    # oracle on the first 50 frames as their own chunk
    ref = O.extract_chunk(prep[:50].cpu().numpy(), small.masks, small.keypoints, small.num_instances)
    part = eng.extract(prep[:50], m_dev[:50], k_dev[:50], chunk_size=50, **kw)
    assert np.array_equal(part['cleaned'].cpu().numpy(), ref['cleaned_frames'])
    assert_close(part['angle_deg'].cpu().numpy(), ref['features']['orientation'], 0, 1e-9, what='angles')
    assert np.array_equal(part['depth_crops'].cpu().numpy(), ref['depth_frames'])
    assert np.array_equal(part['mask_crops'].cpu().numpy(), ref['mask_frames'])


# ----------------------------------------------------------------------------- f4 get_bground_im (session setup)
@pytest.mark.parametrize('n,shape,scale,dtype', [(9, (48, 64), 5, np.int16), (10, (37, 53), 5, np.int16), (4, (16, 24), 3, np.uint16),
                                                   (1, (20, 20), 5, np.int16), (216, (424, 512), 5, np.int16)])
def test_bground_im_matches_opencv_and_numpy(P, n, shape, scale, dtype):
    """ref proc/roi.py:293-307: cv2.medianBlur(frame, med_scale) per frame then np.median over frames -- bit-exact, odd and
    even frame counts, signed (negative values included) and unsigned 16-bit, the full-size case io/session.py:217 produces."""
    rng = np.random.default_rng(n * 7 + scale)
    if n == 216:
        from moseq2_detectron_extract_b200 import synthetic
        geom = synthetic.SessionGeometry()
        base = synthetic.generate_chunk(8, seed=3, geom=geom, t0=0, invalid_rate=0.002).frames
        frames = base[rng.integers(0, 8, size=n)].astype(dtype)
        frames = (frames + rng.integers(-2, 3, size=frames.shape)).astype(dtype)
    elif dtype == np.uint16:
        frames = rng.integers(0, 65536, size=(n,) + shape).astype(np.uint16)
    else:
        frames = rng.integers(-300, 3000, size=(n,) + shape).astype(np.int16)
        frames[rng.random(frames.shape) < 0.05] = 0
    keep = frames.copy()
    got = P.get_bground_im(frames, med_scale=scale)
    assert got.dtype == np.float64 and got.shape == shape
    assert np.array_equal(frames, keep)                                  # the input is left alone (the reference blurs in place)
    assert np.array_equal(got, O.bground_im(keep, scale))
    dev = P.get_bground_im(torch.from_numpy(keep.view(np.int16)).cuda(), med_scale=scale) if dtype == np.int16 else None
    if dev is not None:
        assert dev.is_cuda and np.array_equal(dev.cpu().numpy(), got)


def test_bground_im_rejects_unsupported(P):
    from moseq2_detectron_extract_b200 import _lib
    frames = np.zeros((3, 8, 8), dtype=np.int16)
    with pytest.raises(_lib.MoseqB200Error, match='med_scale'):
        P.get_bground_im(frames, med_scale=7)                            # cv2.medianBlur: 16-bit only for ksize 3 and 5
    with pytest.raises(TypeError):
        P.get_bground_im(frames.astype(np.float32))
    with pytest.raises(ValueError):
        P.get_bground_im(frames[0])


def test_unpack_mask_bits_matches_numpy():
    """msq_unpack_mask_bits: bit rows (numpy.packbits little) -> {0,1} bytes, widths that are and are not multiples of 8."""
    import numpy as np
    import torch
    from moseq2_detectron_extract_b200 import _dev, _lib
    rng = np.random.default_rng(3)
    for h, w in ((240, 240), (17, 43), (400, 400), (5, 8)):
        mask = (rng.random((6, h, w)) < 0.3).astype(np.uint8)
        bits = np.packbits(mask, axis=-1, bitorder='little')
        d_bits = torch.from_numpy(bits).cuda()
        out = torch.full((6, h, w), 7, dtype=torch.uint8, device='cuda')
        _lib.call('msq_unpack_mask_bits', _dev.ptr(d_bits), 6, h, w, _dev.ptr(out), _dev.stream())
        assert np.array_equal(out.cpu().numpy(), mask)


def test_roi_box_dma_then_prep_equals_prep_of_full_frames(P):
    """msq_copy_roi_rows (one strided DMA of the ROI box per chunk from pinned host memory) + msq_prep_frames on the dense box
    with the cropped background / ROI == msq_prep_frames on the full frames (and the oracle)."""
    import torch
    from moseq2_detectron_extract_b200 import _dev, _lib, synthetic
    for geom in (synthetic.SessionGeometry(), synthetic.SessionGeometry.azure()):
        ch = synthetic.generate_chunk(9, seed=2, geom=geom, invalid_rate=0.001)
        roi, bg = synthetic.make_roi(geom), synthetic.make_background(geom)
        y0, x0, y1, x1 = synthetic.roi_bbox(roi)
        h, w = y1 - y0, x1 - x0
        host = torch.from_numpy(ch.frames).pin_memory()
        box = torch.empty((9, h, w), dtype=torch.int16, device='cuda')
        _lib.call('msq_copy_roi_rows', _dev.ptr(host), 9, geom.height, geom.width, y0, x0, h, w, _dev.ptr(box), _dev.stream())
        torch.cuda.synchronize()
        assert np.array_equal(box.cpu().numpy(), ch.frames[:, y0:y1, x0:x1])
        out = torch.empty((9, h, w), dtype=torch.uint8, device='cuda')
        inv = torch.empty((9,), dtype=torch.int32, device='cuda')
        bg_box = _dev.as_device(np.ascontiguousarray(bg[y0:y1, x0:x1]))      # named: the device arrays must outlive the call
        roi_box = _dev.as_device(np.ascontiguousarray(roi[y0:y1, x0:x1].astype(np.uint8)))
        _lib.call('msq_prep_frames', _dev.ptr(box), 9, h, w, _dev.ptr(bg_box), _lib.MSQ_BG_F32, _dev.ptr(roi_box), 0, 0, h, w, 0.0, 100.0,
                  _lib.MSQ_PREP_HAS_VMIN | _lib.MSQ_PREP_HAS_VMAX, _dev.ptr(out), _dev.ptr(inv), None, _dev.stream())
        want = O.prep_frames(ch.frames, bg, roi, 0, 100, fix_invalid=False)
        assert np.array_equal(out.cpu().numpy(), want)
        assert int(inv.sum()) == int(((ch.frames[:, y0:y1, x0:x1] == 0) & roi[y0:y1, x0:x1]).sum())


def test_roi_band_dma_then_prep_equals_prep_of_full_frames(P):
    """msq_copy_roi_bands (the ROI disc as 16 / 5 / 1 horizontal bands, one strided DMA each; pixels outside the bands are never
    written and hold junk) + msq_prep_frames on the dense box == the oracle's prep of the full frames, bit for bit."""
    import torch
    from moseq2_detectron_extract_b200 import _dev, _lib, synthetic
    for geom in (synthetic.SessionGeometry(), synthetic.SessionGeometry.azure()):
        ch = synthetic.generate_chunk(7, seed=4, geom=geom, invalid_rate=0.001)
        roi, bg = synthetic.make_roi(geom), synthetic.make_background(geom)
        y0, x0, y1, x1 = synthetic.roi_bbox(roi)
        h, w = y1 - y0, x1 - x0
        host = torch.from_numpy(ch.frames).pin_memory()
        bg_box = _dev.as_device(np.ascontiguousarray(bg[y0:y1, x0:x1]))
        roi_box = _dev.as_device(np.ascontiguousarray(roi[y0:y1, x0:x1].astype(np.uint8)))
        want = O.prep_frames(ch.frames, bg, roi, 0, 100, fix_invalid=False)
        for n_bands in (16, 5, 1):
            bands = _dev.roi_bands(roi[y0:y1, x0:x1], n_bands)
            covered = np.zeros((h, w), bool)
            for b in range(len(bands[1])):
                covered[bands[0][b]:bands[0][b + 1], bands[1][b]:bands[2][b]] = True
            assert not (roi[y0:y1, x0:x1] & ~covered).any()
            if n_bands == 16:
                assert covered.mean() < 0.87
            box = torch.full((7, h, w), -12345, dtype=torch.int16, device='cuda')      # junk where no band writes
            _dev.copy_roi_bands(host, y0, x0, bands, box)
            torch.cuda.synchronize()
            got = box.cpu().numpy()
            assert np.array_equal(got[:, covered], ch.frames[:, y0:y1, x0:x1][:, covered])
            assert (got[:, ~covered] == -12345).all()
            out = torch.empty((7, h, w), dtype=torch.uint8, device='cuda')
            inv = torch.empty((7,), dtype=torch.int32, device='cuda')
            _lib.call('msq_prep_frames', _dev.ptr(box), 7, h, w, _dev.ptr(bg_box), _lib.MSQ_BG_F32, _dev.ptr(roi_box), 0, 0, h, w, 0.0, 100.0,
                      _lib.MSQ_PREP_HAS_VMIN | _lib.MSQ_PREP_HAS_VMAX, _dev.ptr(out), _dev.ptr(inv), None, _dev.stream())
            assert np.array_equal(out.cpu().numpy(), want)
            assert int(inv.sum()) == int(((ch.frames[:, y0:y1, x0:x1] == 0) & roi[y0:y1, x0:x1]).sum())
    with pytest.raises(_lib.MoseqB200Error):
        bad = (np.array([0, h + 1], np.int32), np.array([0], np.int32), np.array([w], np.int32))
        _dev.copy_roi_bands(host, y0, x0, bad, box)

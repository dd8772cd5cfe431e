"""CPU: pin oracle/extract_oracle.py against outputs of the UNMODIFIED reference (tests/golden/*.npz,
made by oracle/make_golden.py) and its numpy restatements against OpenCV itself."""
import os

import numpy as np
import pytest

import extract_oracle as O
from cases import CASE_NAMES, assert_close, case_inputs, golden

cv2 = pytest.importorskip('cv2')


@pytest.mark.parametrize('name', CASE_NAMES)
def test_prep_matches_reference(name):
    geom, chunk, roi, bg, cfg = case_inputs(name)
    g = golden(name)
    out = O.prep_frames(chunk.frames, bg, roi, cfg['min_height'], cfg['max_height'], fix_invalid=True)
    assert out.dtype == np.uint8 and np.array_equal(out, g['prep'])
    nofix = O.prep_frames(chunk.frames, bg, roi, cfg['min_height'], cfg['max_height'], fix_invalid=False)
    assert np.array_equal(nofix, g['prep_nofix'])
    if name == 'kinect_invalid':
        assert not np.array_equal(g['prep'], g['prep_nofix'])    # the inpaint branch really ran


@pytest.mark.parametrize('name', CASE_NAMES)
def test_scale_matches_reference(name):
    geom, chunk, roi, bg, cfg = case_inputs(name)
    g = golden(name)
    out = O.scale_frames(g['prep'][:4, :, :, None], cfg['min_height'], cfg['max_height'])
    assert np.array_equal(out, g['scale'])
    lut = O.scale_lut(cfg['min_height'], cfg['max_height'])
    assert np.array_equal(lut[g['prep'][:4, :, :, None]], g['scale'])


@pytest.mark.parametrize('name', CASE_NAMES)
@pytest.mark.parametrize('use_cv2', [True, False])
def test_extract_chunk_matches_reference(name, use_cv2):
    if name == 'kinect_invalid' and not use_cv2:
        pytest.skip('tiny case, covered with cv2')
    geom, chunk, roi, bg, cfg = case_inputs(name)
    g = golden(name)
    res = O.extract_chunk(g['prep'], chunk.masks, chunk.keypoints, chunk.num_instances, cfg['min_height'],
                          cfg['max_height'], cfg['true_depth'], cfg['crop_size'], use_cv2=use_cv2)
    assert np.array_equal(res['cleaned_frames'], g['cleaned_frames'])
    assert np.array_equal(res['masks'], g['masks'])
    assert_close(res['features']['centroid'], g['centroid'], 1e-12, what='centroid')
    assert_close(res['features']['axis_length'], g['axis_length'], 1e-11, what='axis_length')
    assert_close(res['features']['orientation'], g['orientation'], 0, 1e-9, what='orientation (deg)')
    assert np.array_equal(res['flips'], g['flips'])
    for k in g.files:
        if k.startswith('scalars/'):
            assert_close(res['scalars'][k[8:]], g[k], 1e-9, 1e-9, what=k)
        elif k.startswith('keypoints/'):
            assert_close(res['keypoint_table'][k[10:]], g[k], 1e-9, 1e-9, what=k)
    assert np.array_equal(res['depth_frames'], g['depth_frames'])
    assert np.array_equal(res['mask_frames'], g['mask_frames'])


def test_missing_case_has_nans_and_runs_all_passes():
    geom, chunk, roi, bg, cfg = case_inputs('kinect_missing_holes')
    g = golden('kinect_missing_holes')
    assert np.isnan(g['centroid']).any() and (chunk.num_instances == 0).any()
    res = O.extract_chunk(g['prep'], chunk.masks, chunk.keypoints, chunk.num_instances)
    assert res['filter_passes'] == 1001                        # SURVEY trap 16: NaN defeats allclose


def test_median_and_open_restatement_equals_opencv():
    rng = np.random.default_rng(5)
    for shape in [(1, 1), (1, 7), (9, 1), (5, 5), (31, 64), (57, 83), (240, 240)]:
        fr = rng.integers(0, 256, size=(2,) + shape).astype(np.uint8)
        fr[1] = (fr[1] > 128) * 200
        assert np.array_equal(O.clean_frames_cv2(fr), O.clean_frames_np(fr)), shape


def test_polygon_cells_equal_opencv_contour_moments():
    rng = np.random.default_rng(6)
    for t in range(150):
        h, w = int(rng.integers(3, 40)), int(rng.integers(3, 40))
        if t % 3 == 0:
            m = (rng.random((h, w)) < rng.uniform(0.2, 0.9)).astype(np.uint8)
        else:
            s = cv2.GaussianBlur(rng.random((h, w)).astype(np.float32), (0, 0), float(rng.uniform(1, 3)))
            m = (s > np.quantile(s, rng.uniform(0.2, 0.8))).astype(np.uint8)
        if t % 10 == 0:
            m[:] = 0
        if t % 10 == 1:
            m[:] = 1
        cl = (m * 9).astype(np.uint8)[None]
        a, b = O.frame_features_cv2(cl, m[None]), O.frame_features_np(cl, m[None])
        for k in a:
            assert_close(b[k], a[k], 1e-12, 1e-12, what=f'{k} trial {t}')


def test_fixed_point_warp_restatement_equals_opencv():
    rng = np.random.default_rng(7)
    for t in range(120):
        h, w = (240, 240) if t % 2 else (int(rng.integers(60, 200)), int(rng.integers(60, 200)))
        fr = rng.integers(0, 101, size=(h, w)).astype(np.uint8)
        c = np.array([rng.uniform(-2, w + 1), rng.uniform(-2, h + 1)])
        if t % 7 == 0:
            c = np.array([rng.uniform(0, 45), rng.uniform(0, 45)])
        if t % 11 == 0:
            c = np.array([float(rng.integers(0, w)), float(rng.integers(0, h))])
        ang = float(rng.uniform(-10, 560)) if t % 13 else float(rng.integers(0, 8) * 45)
        if t % 17 == 0:
            ang = float('nan')
        crop = (80, 80) if t % 3 else (128, 128)
        assert np.array_equal(O.crop_rotate_cv2(fr, c, ang, crop), O.crop_rotate_np(fr, c, ang, crop)), (t, c, ang)


def test_trailing_median_semantics():
    a = np.array([1.0, np.nan, 5.0, 2.0, np.nan, np.nan, np.nan, 7.0])
    m = O.trailing_median3(a)
    assert np.array_equal(m, np.array([1.0, 1.0, 3.0, 3.5, 3.5, 2.0, np.nan, 7.0]), equal_nan=True)


def test_paste_restatement_equals_torch_grid_sample():
    """detectron2 is not installed; its _do_paste_mask is grid_sample on pixel centres mapped into the box.
    Pin the numpy restatement (what csrc/paste.cu follows) against torch's own grid_sample arithmetic."""
    torch = pytest.importorskip('torch')
    F = torch.nn.functional
    rng = np.random.default_rng(3)
    n, M, h, w = 6, 28, 120, 150
    soft = torch.sigmoid(torch.from_numpy(rng.normal(0, 2, (n, M, M)).astype(np.float32)))
    soft = F.avg_pool2d(soft[:, None], 3, 1, 1)[:, 0]
    boxes = np.stack([rng.uniform(-5, 60, n), rng.uniform(-5, 50, n), rng.uniform(70, 155, n), rng.uniform(60, 125, n)], 1).astype(np.float32)
    bt = torch.from_numpy(boxes)
    x0, y0, x1, y1 = torch.split(bt, 1, dim=1)
    img_y = (torch.arange(0, h, dtype=torch.float32) + 0.5 - y0) / (y1 - y0) * 2 - 1
    img_x = (torch.arange(0, w, dtype=torch.float32) + 0.5 - x0) / (x1 - x0) * 2 - 1
    gx = img_x[:, None, :].expand(n, h, w)
    gy = img_y[:, :, None].expand(n, h, w)
    ref = F.grid_sample(soft[:, None], torch.stack([gx, gy], dim=3), align_corners=False)[:, 0] >= 0.5
    got = O.paste_masks_np(soft.numpy(), boxes, h, w)
    assert np.array_equal(got, ref.numpy())


def test_warp_emulation_of_feature_kernel_matches_oracle():
    """Lane-level emulation of csrc/features.cu (bit-row flood with carry-lookahead, blob peeling, popcount cell
    sums) against the oracle's exact integer sums -- algorithm check that runs without a GPU."""
    from warp_emulation import emulate_frame
    rng = np.random.default_rng(11)
    for t in range(60):
        h, w = int(rng.integers(3, 50)), int(rng.integers(3, 110))
        if t % 3 == 0:
            m = rng.random((h, w)) < rng.uniform(0.2, 0.9)
        else:
            s = cv2.GaussianBlur(rng.random((h, w)).astype(np.float32), (0, 0), float(rng.uniform(1, 4)))
            m = s > np.quantile(s, rng.uniform(0.2, 0.8))
        if t % 20 == 0:
            m[:] = True
        if t % 20 == 1:
            m[:] = False
        ref = O.frame_features_np((m * 9).astype(np.uint8)[None], m.astype(np.uint8)[None], return_sums=True)['sums24'][0]
        assert list(ref) == list(emulate_frame(m)), (t, h, w)


def test_inpaint_restatement_equals_opencv():
    """oracle/inpaint_ns.py (the algorithm csrc/inpaint.cu follows) against cv2.inpaint(.., 3, INPAINT_NS)."""
    from inpaint_ns import inpaint_ns
    rng = np.random.default_rng(5)
    for t in range(40):
        h, w = int(rng.integers(6, 30)), int(rng.integers(6, 30))
        img = rng.integers(0, 256, (h, w)).astype(np.uint8) if t % 3 else cv2.GaussianBlur(rng.integers(0, 101, (h, w)).astype(np.uint8), (0, 0), 1.5)
        m = (rng.random((h, w)) < [0.002, 0.02, 0.1, 0.3, 0.6][t % 5]).astype(np.uint8)
        if t % 4 == 0:
            y, x = int(rng.integers(0, h - 3)), int(rng.integers(0, w - 3))
            m[y:y + int(rng.integers(2, 6)), x:x + int(rng.integers(2, 6))] = 1
        if t % 11 == 0:
            m[:] = 1
        if t % 13 == 0:
            m[0, :] = 1
            m[:, 0] = 1
        assert np.array_equal(inpaint_ns(img, m, 3), cv2.inpaint(img, m, 3, cv2.INPAINT_NS)), t


# ----------------------------------------------------------------------------- a14 tracking branch
def test_tracking_oracle_matches_reference_golden():
    """oracle/tracking_oracle.py (restated KalmanTracker + tracking strategy over the restated pykalman filter) against
    the outputs of the UNMODIFIED reference proc/proc.py:730-826 + proc/kalman.py run on the same stand-in."""
    import make_golden
    import tracking_oracle as TO
    from moseq2_detectron_extract_b200 import synthetic
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'kinect_tracking.npz'))
    geom = synthetic.SessionGeometry.kinect_v2()
    kw = {k: v for k, v in make_golden.TRACKING_CASE.items() if k != 'chunks'}
    chunk = synthetic.generate_chunk(geom=geom, **kw)
    roi, bg = synthetic.make_roi(geom), synthetic.make_background(geom)
    prep = O.prep_frames(chunk.frames, bg, roi, 0, 100)
    pt, at = TO.Tracker(18), TO.Tracker(2)
    start = 0
    for c, n in enumerate(make_golden.TRACKING_CASE['chunks']):
        sl = slice(start, start + n)
        feats = O.frame_features_cv2(O.clean_frames_cv2(prep[sl]), chunk.masks[sl])
        cen, kp, ang, fl = TO.track_chunk(feats['centroid'], feats['orientation'], feats['axis_length'], chunk.keypoints[sl], pt, at)
        np.testing.assert_allclose(cen, g[f'c{c}/centroid'], rtol=0, atol=1e-10)
        np.testing.assert_allclose(kp, g[f'c{c}/keypoints'], rtol=0, atol=1e-10)
        np.testing.assert_allclose(ang, g[f'c{c}/orientation'], rtol=0, atol=1e-9)
        assert np.array_equal(fl, g[f'c{c}/flips'])
        assert not np.isnan(ang).any() and np.isnan(g[f'c{c}/keypoints']).any()     # missing frames get the tracked angle
        start += n
    np.testing.assert_allclose(pt.kf.transition_covariance, g['point_transition_covariance'], rtol=1e-10, atol=1e-12)


def test_restated_pykalman_conventions():
    """The documented pykalman behaviours the tracking branch leans on (oracle/pykalman_standin.py)."""
    from numpy import ma
    from pykalman_standin import KalmanFilter
    import tracking_oracle as TO
    tr = TO.Tracker(2)
    kf = KalmanFilter(transition_matrices=tr.A, observation_matrices=tr.H, initial_state_mean=np.arange(6.0))
    # sample(1, x) returns x itself: `KalmanTracker.sample(1)` is deterministic (reference proc/kalman.py:370-377)
    states, _ = kf.sample(1, np.arange(6.0) + 1)
    assert np.array_equal(states[0], np.arange(6.0) + 1)
    # the first predicted state is the prior itself; a partly masked observation is skipped entirely
    z = ma.masked_invalid(np.array([[1.0, np.nan], [2.0, 3.0]]))
    means, covs = kf.filter(z)
    assert np.array_equal(means[0], np.arange(6.0)) and np.array_equal(covs[0], np.eye(6))
    m1, c1 = kf.filter_update(np.arange(6.0), np.eye(6), z[0])
    assert np.allclose(m1, tr.A @ np.arange(6.0)) and np.allclose(c1, tr.A @ tr.A.T + np.eye(6))


def test_bground_oracle_twins_agree():
    """cv2.medianBlur(16-bit, 5) == middle of the replicate-padded sorted window; np.median averages the two middle frames."""
    rng = np.random.default_rng(4)
    for n, dtype in ((6, np.int16), (7, np.uint16)):
        frames = rng.integers(0, 4000, size=(n, 30, 41)).astype(dtype)
        if dtype == np.int16:
            frames -= 500
        for scale in (3, 5):
            a, b = O.bground_im(frames, scale), O.bground_im_np(frames, scale)
            assert a.dtype == np.float64 and np.array_equal(a, b)
    assert (O.bground_im(np.stack([np.full((8, 8), v, np.int16) for v in (1, 2, 4, 9)]), 5) == 3.0).all()


def test_roi_oracle_matches_reference_golden():
    """oracle/roi_oracle.py against the reference's get_roi outputs (tests/golden/session_roi.npz)."""
    import cv2
    import make_golden
    import roi_oracle
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'session_roi.npz'))
    for name, (bg_kw, seed, kw) in make_golden.ROI_CASES.items():
        bg = roi_oracle.synthetic_bground(**bg_kw)
        np.random.seed(seed)
        rois, plane, bboxes, label_im, ranks, shape_index = roi_oracle.get_roi(bg, **kw(cv2))
        assert np.array_equal(plane, g[name + '/plane'])
        assert np.array_equal(label_im, g[name + '/label_im'])
        assert np.array_equal(ranks, g[name + '/ranks'])
        assert np.array_equal(shape_index, g[name + '/shape_index'])
        assert np.array_equal(np.stack(bboxes), g[name + '/bboxes'])
        want = np.unpackbits(g[name + '/rois'], axis=-1)[..., :bg.shape[1]].astype(bool)
        assert len(rois) == len(shape_index) > make_golden.ROI_KEEP
        for i in range(make_golden.ROI_KEEP):
            assert np.array_equal(rois[i], want[i])
        # the first-ranked region of the default case is the bucket floor with its closed hole filled
        if name == 'default':
            cy, cx = bg.shape[0] // 2, bg.shape[1] // 2
            assert rois[0][cy - 25, cx + 25] and not g[name + '/label_im'][cy - 25, cx + 25]


def _reference_function(relpath, name):
    """Compile ONE function of the reference straight from its source file (this container only; the package around it
    needs detectron2 / norfair to import)."""
    import ast
    import ref_import
    path = os.path.join(ref_import.REFERENCE_ROOT, relpath)
    if not os.path.exists(path):
        pytest.skip('reference tree not present')
    tree = ast.parse(open(path).read())
    node = next(n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef) and n.name == name)
    node.returns = None
    for arg in node.args.args:
        arg.annotation = None
    import torch
    import cv2
    scope = {'np': np, 'torch': torch, 'cv2': cv2}
    exec(compile(ast.Module(body=[node], type_ignores=[]), path, 'exec'), scope)
    return scope[name]


@pytest.mark.filterwarnings('ignore::DeprecationWarning')          # the reference's np.matmul on torch tensors under NumPy 2
def test_mask_nms_restatement_equals_reference_method():
    """oracle nms_mask_instances against ProcessFeaturesStep.__nms_mask_instances (process_features_step.py:63-113) on random
    overlapping boxes, empty masks and the chain case where the reference differs from textbook greedy NMS."""
    import torch
    ref_nms = _reference_function('moseq2_detectron_extract/pipeline/process_features_step.py', '__nms_mask_instances')

    class Fake:                                   # the slice of detectron2's Instances the method touches
        def __init__(self, masks, scores):
            self.pred_masks, self.scores = masks, scores

        def __len__(self):
            return len(self.scores)

        def __getitem__(self, k):
            k = torch.tensor(k, dtype=torch.long) if isinstance(k, list) else k
            return Fake(self.pred_masks[k], self.scores[k])

    def run(masks, scores):
        out = ref_nms(None, Fake(torch.from_numpy(masks), torch.from_numpy(scores)))
        want = [masks[i] for i in O.nms_mask_instances(masks, scores)]
        got = list(out.pred_masks.numpy())
        assert len(got) == len(want) and all(np.array_equal(a, b) for a, b in zip(got, want))
        return len(got)

    def box(x0, x1):
        m = np.zeros((40, 40), bool)
        m[10:30, x0:x1] = True
        return m
    # A-B and B-C overlap (IoU 0.6), A-C do not (0.33): textbook NMS keeps A and C, the reference keeps only A
    assert run(np.stack([box(0, 20), box(5, 25), box(10, 30)]), np.array([.9, .8, .7], np.float32)) == 1
    rng = np.random.default_rng(0)
    suppressed = 0
    for trial in range(200):
        n = int(rng.integers(2, 8))
        masks = np.zeros((n, 32, 32), bool)
        for i in range(n):
            x0, y0 = int(rng.integers(0, 14)), int(rng.integers(0, 14))
            masks[i, y0:y0 + int(rng.integers(6, 18)), x0:x0 + int(rng.integers(6, 18))] = True
        if trial % 5 == 0:
            masks[int(rng.integers(0, n))] = False
        suppressed += run(masks, rng.random(n).astype(np.float32)) < n
    assert suppressed > 50


def test_chunk_ranges_equal_reference_gen_batch_sequence():
    """shard.chunk_ranges against the reference's gen_batch_sequence (io/util.py:24-35, offset 0) compiled from its source."""
    from moseq2_detectron_extract_b200.shard import chunk_ranges
    gen = _reference_function('moseq2_detectron_extract/io/util.py', 'gen_batch_sequence')
    for nframes in (0, 1, 9, 10, 11, 25, 70, 1000, 54000):
        for chunk_size, overlap in ((10, 0), (10, 3), (30, 10), (1000, 0), (1000, 100), (7, 6)):
            want = [list(r) for r in gen(nframes, chunk_size, overlap)]
            got = [list(r) for r in chunk_ranges(nframes, chunk_size, overlap)]
            assert got == want, (nframes, chunk_size, overlap)


def test_roi_host_helpers_equal_reference_functions():
    """proc.plane_fit3 / proc.select_strel (host helpers of the session ROI set-up) against the reference's functions compiled
    from source (proc/roi.py:106-130, proc/util.py:9-26)."""
    from moseq2_detectron_extract_b200.proc.roi import plane_fit3
    from moseq2_detectron_extract_b200.proc.util import select_strel
    ref_fit = _reference_function('moseq2_detectron_extract/proc/roi.py', 'plane_fit3')
    ref_strel = _reference_function('moseq2_detectron_extract/proc/util.py', 'select_strel')
    rng = np.random.default_rng(0)
    for _ in range(200):
        pts = np.column_stack([rng.integers(0, 512, 3), rng.integers(0, 424, 3), rng.integers(600, 720, 3)]).astype(np.float64)
        np.testing.assert_allclose(plane_fit3(pts), ref_fit(pts), rtol=1e-13, atol=1e-13, equal_nan=True)
    line = np.array([[0., 0., 1.], [1., 1., 2.], [2., 2., 3.]])
    assert np.isnan(plane_fit3(line)).all() and np.isnan(ref_fit(line)).all()
    for shape in ('ellipse', 'rect', 'e', 'r', 'other'):
        for size in ((10, 10), (15, 15), (7, 3), (1, 1), (4, 9)):
            assert np.array_equal(select_strel(shape, size), ref_strel(shape, size)), (shape, size)

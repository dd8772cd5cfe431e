"""GPU: the step classes end to end on a synthetic session, checked chunk by chunk against the oracle."""
import numpy as np
import pytest

import extract_oracle as O

pytestmark = pytest.mark.gpu
torch = pytest.importorskip('torch')


class Collect:
    pass


def _run_pipeline(nframes, chunk_size, **gen):
    from moseq2_detectron_extract_b200 import synthetic
    from moseq2_detectron_extract_b200.pipeline import (Pipeline, PipelineStep, ProcessFeaturesStep, ProduceFramesStep,
                                                        SyntheticInferenceStep)
    sess = synthetic.SyntheticSession(nframes, seed=4, **gen)
    cfg = synthetic.default_config(sess.geom)
    cfg.update(chunk_size=chunk_size, nframes=nframes)
    results = []

    class Sink(PipelineStep):
        def process(self, data):
            results.append(data)
            return data

    pipe = Pipeline()
    steps = [pipe.add_step(ProduceFramesStep(sess, cfg, 'produce')), pipe.add_step(SyntheticInferenceStep(cfg, 'infer')),
             pipe.add_step(ProcessFeaturesStep(cfg, 'features')), pipe.add_step(Sink(cfg, 'sink'))]
    for a, b in zip(steps[:-1], steps[1:]):
        pipe.link(a, b)
    pipe.run()
    return sess, cfg, results


def test_steps_match_oracle_chunk_by_chunk():
    from moseq2_detectron_extract_b200 import synthetic
    sess, cfg, results = _run_pipeline(130, 50, missing_every=19, mask_holes=True)
    assert [r['batch'] for r in results] == [0, 1, 2]
    assert sum(len(r['frame_idxs']) for r in results) == 130
    roi, bg = sess.roi, sess.bground_im
    for r in results:
        idxs = r['frame_idxs']
        ch = synthetic.generate_chunk(len(idxs), seed=4, geom=sess.geom, t0=idxs[0], missing_every=19, mask_holes=True)
        prep = O.prep_frames(ch.frames, bg, roi, cfg['min_height'], cfg['max_height'])
        ref = O.extract_chunk(prep, ch.masks, ch.keypoints, ch.num_instances, cfg['min_height'], cfg['max_height'],
                              cfg['true_depth'], cfg['crop_size'])
        assert np.array_equal(r['chunk'].cpu().numpy(), prep)
        f = r['features']
        assert np.array_equal(f['cleaned_frames'], ref['cleaned_frames'])
        assert np.array_equal(f['masks'], ch.masks)
        assert np.allclose(f['features']['centroid'], ref['features']['centroid'], rtol=1e-12, atol=0, equal_nan=True)
        assert np.allclose(f['features']['orientation'], ref['features']['orientation'], rtol=0, atol=1e-9, equal_nan=True)
        assert np.array_equal(f['flips'], ref['flips'])
        assert np.array_equal(f['num_instances'], ch.num_instances)
        assert set(r['scalars']) == set(ref['scalars']) and set(r['keypoints']) == set(ref['keypoint_table'])
        for k, v in ref['scalars'].items():
            tol = 1e-4 if 'velocity' in k else 1e-9
            assert np.allclose(r['scalars'][k], v, rtol=tol, atol=1e-9, equal_nan=True), k
        assert r['scalars']['area_px'].dtype == np.int64 and r['scalars']['height_ave_mm'].dtype == np.float32
        for k, v in ref['keypoint_table'].items():
            assert np.allclose(r['keypoints'][k], v, rtol=1e-9, atol=1e-7, equal_nan=True), k
        assert np.array_equal(r['depth_frames'], ref['depth_frames'])
        assert np.array_equal(r['mask_frames'], ref['mask_frames'])


def test_tracking_step_builds_trackers_and_invalid_pixels_are_inpainted():
    from moseq2_detectron_extract_b200 import synthetic
    from moseq2_detectron_extract_b200.pipeline import ProcessFeaturesStep
    cfg = synthetic.default_config()
    cfg['use_tracking'] = True
    step = ProcessFeaturesStep(cfg, 'features')
    step.initialize()             # ref process_features_step.py:41-50: centroid + 8 keypoints, and the angle, at order 3
    assert step.point_tracker.n_state == 54 and step.angle_tracker.n_state == 6 and not step.point_tracker.is_initialized
    sess, cfg, results = _run_pipeline(20, 10, invalid_rate=0.002)     # Kinect-like invalid pixels inside the ROI
    for r in results:
        idxs = r['frame_idxs']
        ch = synthetic.generate_chunk(len(idxs), seed=4, geom=sess.geom, t0=idxs[0], invalid_rate=0.002)
        prep = O.prep_frames(ch.frames, sess.bground_im, sess.roi, cfg['min_height'], cfg['max_height'])   # cv2.inpaint inside
        assert np.array_equal(r['chunk'].cpu().numpy(), prep)
        ref = O.extract_chunk(prep, ch.masks, ch.keypoints, ch.num_instances, cfg['min_height'], cfg['max_height'],
                              cfg['true_depth'], cfg['crop_size'])
        assert np.array_equal(r['depth_frames'], ref['depth_frames'])


def test_raw_file_session_zero_copy_prep(tmp_path):
    """a1 + a2: raw .dat file -> pinned host buffer -> prep kernel reading host memory directly -> oracle parity."""
    from moseq2_detectron_extract_b200 import synthetic
    from moseq2_detectron_extract_b200.io import RawDepthSession, read_frames_raw
    from moseq2_detectron_extract_b200.pipeline import Pipeline, PipelineStep, ProduceFramesStep
    from moseq2_detectron_extract_b200.proc import prep_raw_frames
    geom = synthetic.SessionGeometry()
    ch = synthetic.generate_chunk(45, seed=6, geom=geom, invalid_rate=0.001)
    path = str(tmp_path / 'depth.dat')
    ch.frames.astype('<i2').tofile(path)
    roi, bg = synthetic.make_roi(geom), synthetic.make_background(geom)
    ref = O.prep_frames(ch.frames, bg, roi, 0, 100)
    pinned = read_frames_raw(path, frame_dims=(geom.width, geom.height), pinned=True, as_tensor=True)
    assert pinned.is_pinned() and np.array_equal(pinned.numpy(), ch.frames)
    out = prep_raw_frames(pinned, bground_im=bg, roi=roi, vmin=0, vmax=100)           # zero-copy path
    assert np.array_equal(out.cpu().numpy() if hasattr(out, 'cpu') else out, ref)
    sess = RawDepthSession(path, bg, roi, geom.floor_depth, frame_dims=(geom.width, geom.height))
    cfg = synthetic.default_config(geom)
    cfg.update(chunk_size=20, nframes=45)
    got = []

    class Sink(PipelineStep):
        def process(self, data):
            got.append(data['chunk'].cpu().numpy())
            return data

    pipe = Pipeline()
    a, b = pipe.add_step(ProduceFramesStep(sess, cfg, 'produce')), pipe.add_step(Sink(cfg, 'sink'))
    pipe.link(a, b)
    pipe.run()
    assert np.array_equal(np.concatenate(got), ref)


def test_raw_session_background_from_file(tmp_path):
    """RawDepthSession.compute_bground: every k-th frame of a raw .dat file -> GPU median blur + temporal median, equal to
    the reference's get_bground_im on the same frames (ref io/session.py:217-218, proc/roi.py:293-307)."""
    from moseq2_detectron_extract_b200 import synthetic
    from moseq2_detectron_extract_b200.io.video import RawDepthSession
    geom = synthetic.SessionGeometry()
    ch = synthetic.generate_chunk(23, seed=8, geom=geom, invalid_rate=0.002)
    path = tmp_path / 'depth.dat'
    ch.frames.astype('<i2').tofile(path)
    sess = RawDepthSession(str(path), bground_im=None, roi=synthetic.make_roi(geom), true_depth=673.0, pinned=False)
    assert sess.nframes == 23
    bg = sess.compute_bground(frame_stride=3)
    assert bg.dtype == np.float64 and bg.shape == (geom.height, geom.width)
    assert np.array_equal(bg, O.bground_im(ch.frames[::3].copy(), 5)) and sess.bground_im is bg


def test_raw_session_find_roi_from_file(tmp_path):
    """RawDepthSession.find_roi: background + RANSAC floor plane + ranked, dilated, hole-filled region -> roi and true
    depth, equal to the oracle's get_roi on the oracle's background (ref io/session.py:181-264)."""
    import cv2
    import roi_oracle
    from moseq2_detectron_extract_b200 import synthetic
    from moseq2_detectron_extract_b200.io.video import RawDepthSession
    from moseq2_detectron_extract_b200.proc.util import select_strel
    geom = synthetic.SessionGeometry()
    ch = synthetic.generate_chunk(21, seed=12, geom=geom)
    path = tmp_path / 'depth.dat'
    ch.frames.astype('<i2').tofile(path)
    assert np.array_equal(select_strel('ellipse', (10, 10)), cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (10, 10)))
    bg = O.bground_im(ch.frames[::2].copy(), 5)
    np.random.seed(4)
    rois, plane, _, _, _, _ = roi_oracle.get_roi(bg, strel_dilate=cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (10, 10)))
    for plane_bg in (False, True):
        sess = RawDepthSession(str(path), pinned=False)
        np.random.seed(4)
        first, bground, roi, true_depth = sess.find_roi(frame_stride=2, use_plane_bground=plane_bg)
        assert np.array_equal(np.asarray(first)[0], ch.frames[0])
        assert roi.dtype == bool and np.array_equal(roi, rois[0]) and sess.roi is roi
        yy, xx = np.mgrid[:geom.height, :geom.width]
        want_bg = (xx * plane[0] + yy * plane[1] + plane[3]) / -plane[2] if plane_bg else bg
        np.testing.assert_allclose(bground, want_bg, rtol=1e-14)
        assert true_depth == pytest.approx(float(np.median(want_bg[rois[0]])), rel=1e-14) and sess.true_depth == true_depth
        assert abs(true_depth - 673.0) < 2.0
        # the detected arena covers the synthetic bucket floor the generator masks with
        assert (roi & synthetic.make_roi(geom)).sum() > 0.95 * synthetic.make_roi(geom).sum()
    # the tiff cache (ref io/session.py:194-257): first call writes first_frame / bground / roi_00, the second one loads them
    cache = str(tmp_path / 'cache')
    sess = RawDepthSession(str(path), pinned=False)
    np.random.seed(4)
    _, bg1, roi1, depth1 = sess.find_roi(frame_stride=2, cache_dir=cache)
    for name in ('first_frame.tiff', 'bground.tiff', 'roi_00.tiff'):
        assert (tmp_path / 'cache' / name).exists(), name
    sess2 = RawDepthSession(str(path), pinned=False)
    first2, bg2, roi2, depth2 = sess2.find_roi(frame_stride=2, cache_dir=cache)
    assert np.array_equal(roi2, roi1) and bg2.dtype == np.uint16 and np.array_equal(bg2, bg1.astype('uint16'))   # SURVEY trap 8
    assert abs(depth2 - depth1) <= 1.0 and np.abs(first2[0].astype(float) - np.clip(ch.frames[0], 650, 750)).max() <= 1.0


def test_pipeline_with_result_writer(tmp_path):
    """Produce -> (synthetic) inference -> features -> ResultWriterStep over chunks with overlap: the stored crops, masks,
    scalars and flips are the oracle's, every frame exactly once (ref pipeline/write_results_step.py, io/result.py:105-130)."""
    from moseq2_detectron_extract_b200 import synthetic
    from moseq2_detectron_extract_b200.pipeline import (Pipeline, ProcessFeaturesStep, ProduceFramesStep, ResultWriterStep,
                                                        SyntheticInferenceStep)
    nframes, chunk_size = 70, 30
    sess = synthetic.SyntheticSession(nframes, seed=6)
    cfg = synthetic.default_config(sess.geom)
    cfg.update(chunk_size=chunk_size, nframes=nframes, output_dir=str(tmp_path), bg_roi_index=0, roi=sess.roi,
               bground_im=sess.bground_im, timestamps=np.arange(nframes) / 30.0)
    pipe = Pipeline()
    steps = [pipe.add_step(ProduceFramesStep(sess, cfg, 'produce')), pipe.add_step(SyntheticInferenceStep(cfg, 'infer')),
             pipe.add_step(ProcessFeaturesStep(cfg, 'features')), pipe.add_step(ResultWriterStep(cfg, 'writer'))]
    for a, b in zip(steps[:-1], steps[1:]):
        pipe.link(a, b)
    pipe.run()
    store = steps[-1].store
    assert store.closed
    if not store.path.endswith('.npz'):
        pytest.skip('h5py present: layout checked by the CPU test')
    out = np.load(store.path)
    roi, bg = sess.roi, sess.bground_im
    for start in range(0, nframes, chunk_size):
        idxs = list(range(start, min(start + chunk_size, nframes)))
        ch = synthetic.generate_chunk(len(idxs), seed=6, geom=sess.geom, t0=idxs[0])
        prep = O.prep_frames(ch.frames, bg, roi, cfg['min_height'], cfg['max_height'])
        ref = O.extract_chunk(prep, ch.masks, ch.keypoints, ch.num_instances, cfg['min_height'], cfg['max_height'],
                              cfg['true_depth'], cfg['crop_size'])
        assert np.array_equal(out['frames'][idxs], ref['depth_frames'])
        assert np.array_equal(out['frames_mask'][idxs], ref['mask_frames'] != 0)
        assert np.array_equal(out['metadata/extraction/flips'][idxs], ref['flips'])
        assert np.allclose(out['scalars/centroid_x_px'][idxs], ref['scalars']['centroid_x_px'].astype(np.float32), equal_nan=True)
    tsv = open(steps[-1].keypoint_data_dest).read().strip().split('\n')
    assert len(tsv) == 1 + nframes and tsv[-1].split('\t')[0] == str(nframes - 1)


def test_mask_iou_suppression_matches_reference_semantics():
    """ProcessFeaturesStep._nms_mask_instances against the oracle restatement of the reference's method (pinned against the
    reference source in tests/test_oracle_vs_golden.py): random overlapping instances, empty masks, and the chain case in
    which the reference keeps fewer instances than textbook greedy NMS."""
    from moseq2_detectron_extract_b200.model.instances import Boxes, Instances
    from moseq2_detectron_extract_b200.pipeline import ProcessFeaturesStep

    def run(masks, scores):
        n, h, w = masks.shape
        inst = Instances((h, w), pred_boxes=Boxes(torch.zeros((n, 4), device='cuda')), scores=torch.from_numpy(scores).cuda(),
                         pred_classes=torch.zeros((n,), dtype=torch.int64, device='cuda'), pred_masks=torch.from_numpy(masks).cuda(),
                         pred_keypoints=torch.arange(n, device='cuda', dtype=torch.float32)[:, None, None].expand(n, 8, 3).contiguous())
        out = ProcessFeaturesStep._nms_mask_instances(inst)
        want = O.nms_mask_instances(masks, scores)
        if n <= 1:
            assert len(out) == n
            return len(out)
        assert len(out) == len(want)
        assert np.array_equal(out.pred_masks.cpu().numpy(), masks[want])
        assert out.pred_keypoints[:, 0, 0].cpu().numpy().astype(int).tolist() == want           # every field follows the pick
        return len(out)

    def box(x0, x1):
        m = np.zeros((40, 40), bool)
        m[10:30, x0:x1] = True
        return m
    assert run(np.stack([box(0, 20), box(5, 25), box(10, 30)]), np.array([.9, .8, .7], np.float32)) == 1
    assert run(np.stack([box(0, 20)]), np.array([.5], np.float32)) == 1
    rng = np.random.default_rng(3)
    for trial in range(60):
        n = int(rng.integers(2, 9))
        masks = np.zeros((n, 48, 40), bool)
        for i in range(n):
            x0, y0 = int(rng.integers(0, 18)), int(rng.integers(0, 22))
            masks[i, y0:y0 + int(rng.integers(6, 22)), x0:x0 + int(rng.integers(6, 20))] = True
        if trial % 4 == 0:
            masks[int(rng.integers(0, n))] = False
        run(masks, rng.random(n).astype(np.float32))


def test_select_instances_keeps_the_oldest_tracked_identity():
    """ProcessFeaturesStep._select_instances (ref: process_features_step.py:132-160): after the mask-IoU suppression the SORT
    tracker decides -- with expected_instances = 1 the OLDEST identity that is live in the frame is kept, not the best score."""
    from moseq2_detectron_extract_b200 import synthetic
    from moseq2_detectron_extract_b200.model.instances import Boxes, Instances
    from moseq2_detectron_extract_b200.pipeline import ProcessFeaturesStep
    cfg = dict(synthetic.default_config(), expected_instances=1, results_to_host=False)
    step = ProcessFeaturesStep(cfg, 'features')
    step.initialize()
    h = w = 120

    def disk(cy, cx, r=9):
        yy, xx = np.mgrid[0:h, 0:w]
        return (yy - cy) ** 2 + (xx - cx) ** 2 <= r * r

    def inst(masks, scores, tags):
        n = len(masks)
        return Instances((h, w), pred_boxes=Boxes(torch.zeros((n, 4), device='cuda')), scores=torch.tensor(scores, device='cuda'),
                         pred_classes=torch.zeros((n,), dtype=torch.int64, device='cuda'),
                         pred_masks=torch.from_numpy(np.stack(masks)).cuda() if n else torch.zeros((0, h, w), dtype=torch.bool, device='cuda'),
                         pred_keypoints=torch.tensor(tags, device='cuda', dtype=torch.float32)[:, None, None].expand(n, 8, 3).contiguous())
    frames = []
    for t in range(8):
        a = disk(30 + 3 * t, 30 + 2 * t)                           # animal A from the first frame on (tag 1)
        if t < 3:
            frames.append({'instances': inst([a], [0.6], [1.0])})
        elif t == 6:
            frames.append({'instances': inst([disk(90, 90 - 4 * t)], [0.99], [2.0])})      # A missed by the detector once
        else:
            b = disk(90, 90 - 4 * t)                                # animal B appears later with the better score (tag 2)
            frames.append({'instances': inst([b, a], [0.99, 0.6], [2.0, 1.0])})
    data = step._select_instances({'inference': frames, 'frame_idxs': list(range(8))})
    tags = [int(f['instances'].pred_keypoints[0, 0, 0]) if len(f['instances']) else 0 for f in data['inference']]
    counts = [len(f['instances']) for f in data['inference']]
    assert counts == [1] * 8
    # frames 3-5: both alive -> the older A; frame 6: only B is detected, A is alive but B is what the frame holds ... the
    # reference takes the oldest LIVE object's last detection, which is A's detection of frame 5; frame 7: A again
    assert tags[:6] == [1, 1, 1, 1, 1, 1] and tags[7] == 1 and tags[6] == 1


def test_pinned_multi_chunk_session_without_inpainting(tmp_path):
    """The zero-copy producer path over several chunks with fix_invalid_pixels=False (no synchronising step after the prep
    launch): every pinned chunk buffer must stay alive until its kernel has read it.  Frames carry their index so that a
    recycled buffer would show."""
    from moseq2_detectron_extract_b200 import synthetic
    from moseq2_detectron_extract_b200.io import RawDepthSession
    from moseq2_detectron_extract_b200.pipeline import Pipeline, PipelineStep, ProduceFramesStep
    geom = synthetic.SessionGeometry()
    n, chunk = 96, 8
    ch = synthetic.generate_chunk(n, seed=8, geom=geom, invalid_rate=0.001)
    path = str(tmp_path / 'depth.dat')
    ch.frames.astype('<i2').tofile(path)
    roi, bg = synthetic.make_roi(geom), synthetic.make_background(geom)
    sess = RawDepthSession(path, bg, roi, geom.floor_depth, frame_dims=(geom.width, geom.height), pinned=True)
    cfg = dict(synthetic.default_config(geom), chunk_size=chunk, chunk_overlap=0, nframes=n, fix_invalid_pixels=False)
    got = {}

    class Sink(PipelineStep):
        def process(self, data):
            got[data['batch']] = data['chunk']             # keep device tensors; look at them only after the whole run
            return data

    pipe = Pipeline(queue_depth=8)
    a, b = pipe.add_step(ProduceFramesStep(sess, cfg, 'produce')), pipe.add_step(Sink(cfg, 'sink'))
    pipe.link(a, b)
    pipe.run()
    torch.cuda.synchronize()
    assert sorted(got) == list(range(n // chunk))
    want = O.prep_frames(ch.frames, bg, roi, 0, 100, fix_invalid=False)
    assert (ch.frames == 0).any()
    for i in range(n // chunk):
        assert np.array_equal(got[i].cpu().numpy(), want[i * chunk:(i + 1) * chunk]), i


def test_explicit_engine_equals_library_held_engine():
    """msq_extract_chunk_engine with a caller-owned msq_engine == msq_extract_chunk with the library's per-thread one."""
    import ctypes
    from moseq2_detectron_extract_b200 import _dev, _lib, synthetic
    from moseq2_detectron_extract_b200.engine import ChunkEngine
    from moseq2_detectron_extract_b200.proc import prep_raw_frames
    geom = synthetic.SessionGeometry()
    ch = synthetic.generate_chunk(24, seed=2, geom=geom, missing_every=7, mask_holes=True)
    prep = prep_raw_frames(torch.from_numpy(ch.frames).cuda(), bground_im=synthetic.make_background(geom), roi=synthetic.make_roi(geom), vmin=0, vmax=100)
    masks, kpts = _dev.as_device(ch.masks), _dev.as_device(ch.keypoints, torch.float32)
    eng = ChunkEngine()
    kw = dict(chunk_size=1000, min_height=0, max_height=100, true_depth=673.0, crop_size=(80, 80))
    a = {k: v.clone() for k, v in eng.extract(prep, masks, kpts, **kw).items()}
    b = eng._buf
    n, h, w = prep.shape
    outs = _lib.ChunkOutputs(*(b[k].data_ptr() for k in ('cleaned', 'centroid', 'angle_deg', 'axis_length', 'flips', 'scalars', 'kpt_cols',
                                                         'depth_crops', 'mask_crops', 'filter_passes')))
    for key in ('depth_crops', 'scalars', 'centroid'):
        b[key].zero_()
    _lib.call('msq_extract_chunk', _dev.ptr(prep), _dev.ptr(masks), _dev.ptr(kpts), n, h, w, 1000, 0.0, 100.0, 673.0, 80, 80, ctypes.byref(outs),
              _dev.ptr(b['scratch']), b['scratch'].numel(), _dev.stream())
    torch.cuda.synchronize()
    assert torch.equal(b['depth_crops'][:n], a['depth_crops']) and torch.equal(b['cleaned'][:n], a['cleaned'])
    assert torch.allclose(b['centroid'][:n], a['centroid'], rtol=0, atol=0, equal_nan=True)
    with pytest.raises(_lib.MoseqB200Error):
        _lib.call('msq_extract_chunk_engine', None, _dev.ptr(prep), None, _dev.ptr(masks), _dev.ptr(kpts), n, h, w, 1000, 0.0, 100.0, 673.0, 80, 80,
                  ctypes.byref(outs), _dev.ptr(b['scratch']), b['scratch'].numel(), _dev.stream())


@pytest.mark.parametrize('w', [240, 496])
def test_band_limited_feature_pass_equals_full_frame_pass(w):
    """msq_extract_chunk hands the cleaning pass's row bands to the streaming feature kernel, which then reads only the chunks of
    rows that can be non-zero.  Same centroid / axes / cleaned frames as clean + msq_frame_features over whole frames, on the
    cases that stress the row arithmetic: animal against the top / bottom edge, empty and all-noise frames, a body spanning
    every row, two bodies, and (w = 496) frames wider than one 240-column cleaning tile."""
    from moseq2_detectron_extract_b200 import _dev
    from moseq2_detectron_extract_b200.engine import ChunkEngine
    h = 240
    rng = np.random.default_rng(11)
    yy, xx = np.mgrid[0:h, 0:w]

    def blob(cy, cx, ry, rx, height=40):
        return np.where(((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 <= 1.0, height, 0)

    bodies = [blob(10, 60, 18, 30), blob(h - 8, w - 70, 16, 28), np.zeros((h, w), int), blob(120, w // 2, 14, 40),
              blob(120, 100, 140, 20), blob(40, 50, 15, 25) + blob(190, w - 60, 15, 25), blob(0, 0, 12, 12), blob(h - 1, w - 1, 13, 13),
              blob(6, w // 2, 5.5, 30), blob(h - 6, 90, 5.5, 30), blob(12, 200, 11, 11), blob(11.5, 30, 12.4, 12.4)]
    frames = np.stack(bodies * 3).astype(np.int64)
    noise = rng.integers(-3, 4, size=frames.shape)
    frames = np.clip(frames + noise, 0, 100).astype(np.uint8)
    frames[5] = rng.integers(0, 3, size=(h, w))                       # nothing but floor noise
    masks = (np.stack(bodies * 3) > 0).astype(np.uint8)
    masks[7] = 1                                                       # mask wider than the animal
    n = len(frames)
    kpts = np.full((n, 8, 3), np.nan, np.float32)
    chunk, m, k = _dev.as_device(frames), _dev.as_device(masks), _dev.as_device(kpts, torch.float32)
    eng = ChunkEngine()
    got = eng.extract(chunk, m, k, chunk_size=1000, min_height=0, max_height=100, true_depth=673.0, crop_size=(80, 80))
    want = eng.clean_and_features(chunk, m)
    assert torch.equal(got['cleaned'], want['cleaned'])
    for key in ('centroid', 'axis_length'):
        assert torch.allclose(got[key], want[key], rtol=0, atol=0, equal_nan=True), key
    assert (want['cleaned'].flatten(1).max(1).values > 0).sum() >= n - 9     # most frames do carry an animal


@pytest.mark.parametrize('geom_name', ['kinect_v2', 'azure'])
def test_positive_bit_rows_from_prep_give_the_same_chunk_outputs(geom_name):
    """msq_prep_frames_bits writes, beside the prepared frames, one bit per pixel that is > 0; msq_extract_chunk_engine /
    msq_clean_frames_ws given those rows return exactly what they return without them (and the rows are what they say).  Also
    after in-painting (the rows are then a superset), with an odd box width (scalar prep path) and with all-zero bit rows
    replaced by all-ones rows (the loosest superset)."""
    from moseq2_detectron_extract_b200 import _dev, synthetic
    from moseq2_detectron_extract_b200.engine import ChunkEngine
    from moseq2_detectron_extract_b200.proc import prep_raw_frames
    geom = synthetic.SessionGeometry() if geom_name == 'kinect_v2' else synthetic.SessionGeometry.azure()
    ch = synthetic.generate_chunk(14, seed=5, geom=geom, invalid_rate=0.002, missing_every=5)
    bg, roi = synthetic.make_background(geom), synthetic.make_roi(geom)
    eng = ChunkEngine()
    kw = dict(chunk_size=1000, min_height=0, max_height=100, true_depth=673.0, crop_size=(80, 80))
    for fix in (False, True):
        bits = []
        prep = prep_raw_frames(torch.from_numpy(ch.frames).cuda(), bground_im=bg, roi=roi, vmin=0, vmax=100, fix_invalid_pixels=fix,
                               positive_bits_out=bits)
        pos = bits[0]
        n, h, w = prep.shape
        assert pos.shape == (n, h, (w + 31) // 32) and pos.dtype == torch.int32
        got_bits = np.unpackbits(pos.cpu().numpy().view(np.uint8).reshape(n, h, -1), axis=-1, bitorder='little')[:, :, :w].astype(bool)
        truth = prep.cpu().numpy() > 0
        if fix:
            assert not (truth & ~got_bits).any()             # superset after in-painting
        else:
            assert np.array_equal(got_bits, truth)
        masks, kpts = _dev.as_device(ch.masks), _dev.as_device(ch.keypoints, torch.float32)
        a = {k: v.clone() for k, v in eng.extract(prep, masks, kpts, **kw).items()}
        b = {k: v.clone() for k, v in eng.extract(prep, masks, kpts, positive_bits=pos, **kw).items()}
        c = {k: v.clone() for k, v in eng.extract(prep, masks, kpts, positive_bits=torch.full_like(pos, -1), **kw).items()}
        for key in a:
            assert torch.allclose(a[key].double(), b[key].double(), rtol=0, atol=0, equal_nan=True), key
            assert torch.allclose(a[key].double(), c[key].double(), rtol=0, atol=0, equal_nan=True), key
        c1, c2 = torch.empty_like(prep), torch.empty_like(prep)
        _dev.clean_frames_ws(prep, c1)
        _dev.clean_frames_ws(prep, c2, pos)
        assert torch.equal(c1, c2) and torch.equal(c1, a['cleaned'])
    # an odd box (scalar prep kernel; the streaming clean kernel is not used for it): the rows are still exact
    roi2 = roi.copy()
    ys, xs = np.nonzero(roi2)
    roi2[:, xs.max() - 2:] = False
    bits = []
    prep = prep_raw_frames(torch.from_numpy(ch.frames).cuda(), bground_im=bg, roi=roi2, vmin=0, vmax=100, fix_invalid_pixels=False,
                           positive_bits_out=bits)
    n, h, w = prep.shape
    assert w % 8 != 0
    got_bits = np.unpackbits(bits[0].cpu().numpy().view(np.uint8).reshape(n, h, -1), axis=-1, bitorder='little')[:, :, :w].astype(bool)
    assert np.array_equal(got_bits, prep.cpu().numpy() > 0)
    masks = _dev.as_device(np.ascontiguousarray(ch.masks[:, :, :w]))
    a = eng.extract(prep, masks, kpts, **kw)['cleaned'].clone()
    b = eng.extract(prep, masks, kpts, positive_bits=bits[0], **kw)['cleaned'].clone()
    assert torch.equal(a, b)


def test_bench_launch_size_properties():
    """BASELINE configs[1] launches 6 000 frames at a time (six 1000-frame chunks per call).  The oracle cannot follow at that
    size, so the full-size launch is checked through properties that do not depend on it: (1) every output of the 6 000-frame
    call equals, bit for bit, the outputs of the same frames processed as six separate 1000-frame calls; (2) the session is the
    same 1000 distinct frames rolled by 37 per chunk, and every frame-local output follows the roll; (3) a 40-frame window from
    the middle of the fourth chunk equals the oracle on those frames; (4) with and without the prep kernel's bit rows."""
    from moseq2_detectron_extract_b200 import _dev, synthetic
    from moseq2_detectron_extract_b200.engine import ChunkEngine
    from moseq2_detectron_extract_b200.proc import prep_raw_frames
    geom = synthetic.SessionGeometry()
    pool = synthetic.generate_chunk(1000, seed=21, geom=geom, realistic=True, missing_every=97)
    bg, roi = synthetic.make_background(geom), synthetic.make_roi(geom)
    n_chunks, C = 6, 1000
    idx = np.concatenate([(np.arange(C) + 37 * c) % C for c in range(n_chunks)])
    frames = torch.from_numpy(pool.frames).cuda()[torch.from_numpy(idx).cuda()]
    masks = _dev.as_device(pool.masks)[torch.from_numpy(idx).cuda()].contiguous()
    kpts = _dev.as_device(pool.keypoints, torch.float32)[torch.from_numpy(idx).cuda()].contiguous()
    kw = dict(chunk_size=C, min_height=0, max_height=100, true_depth=673.0, crop_size=(80, 80))
    eng = ChunkEngine()
    bits = []
    prep = prep_raw_frames(frames, bground_im=bg, roi=roi, vmin=0, vmax=100, fix_invalid_pixels=False, positive_bits_out=bits)
    whole = {k: v.clone() for k, v in eng.extract(prep, masks, kpts, positive_bits=bits[0], **kw).items()}
    n = n_chunks * C
    # (4) bit rows or not
    plain = eng.extract(prep, masks, kpts, **kw)
    for k in whole:
        assert torch.allclose(whole[k].double(), plain[k].double(), rtol=0, atol=0, equal_nan=True), k
    # (1) one launch == six launches
    per_frame = ('cleaned', 'centroid', 'angle_deg', 'axis_length', 'flips', 'depth_crops', 'mask_crops')
    for c in range(n_chunks):
        s = slice(c * C, (c + 1) * C)
        part = eng.extract(prep_raw_frames(frames[s], bground_im=bg, roi=roi, vmin=0, vmax=100, fix_invalid_pixels=False),
                           masks[s].contiguous(), kpts[s].contiguous(), **kw)
        for k in per_frame:
            assert torch.allclose(whole[k][s].double(), part[k].double(), rtol=0, atol=0, equal_nan=True), (k, c)
        for k in ('scalars', 'kpt_cols'):
            assert torch.allclose(whole[k][:, s], part[k], rtol=0, atol=0, equal_nan=True), (k, c)
        assert int(whole['filter_passes'][c]) == int(part['filter_passes'][0])
    # (2) frame-local outputs follow the roll
    for c in range(1, n_chunks):
        src = torch.from_numpy((np.arange(C) + 37 * c) % C).cuda()
        for k in ('cleaned', 'centroid', 'axis_length'):
            assert torch.allclose(whole[k][c * C:(c + 1) * C].double(), whole[k][:C][src].double(), rtol=0, atol=0, equal_nan=True), (k, c)
    # (3) a window of the fourth chunk against the oracle
    lo = 3 * C + 480
    sel = idx[lo:lo + 40]
    want_prep = O.prep_frames(pool.frames[sel], bg, roi, 0, 100, fix_invalid=False)
    assert np.array_equal(prep[lo:lo + 40].cpu().numpy(), want_prep)
    cleaned = O.clean_frames_cv2(want_prep)
    feats = O.frame_features_cv2(cleaned, pool.masks[sel])
    assert np.array_equal(whole['cleaned'][lo:lo + 40].cpu().numpy(), cleaned)
    assert np.allclose(whole['centroid'][lo:lo + 40].cpu().numpy(), feats['centroid'], rtol=1e-12, atol=0, equal_nan=True)
    assert np.allclose(whole['axis_length'][lo:lo + 40].cpu().numpy(), feats['axis_length'], rtol=1e-10, atol=0, equal_nan=True)
    assert n == prep.shape[0] and int((whole['cleaned'].flatten(1).amax(1) > 0).sum()) > 0.9 * n

"""CPU tests of the host-side logic: C-ABI surface, containers, pipeline runtime, sharding helpers."""
import ctypes
import os
import queue
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def lib():
    from moseq2_detectron_extract_b200 import _lib, build
    build.build()                     # nvcc cross-compiles for sm_100a without a GPU
    return _lib


def _declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'moseq_b200.h')).read()
    return sorted(set(re.findall(r'MSQ_API[^;(]*?\b(msq_\w+)\s*\(', text)))


def test_library_exports_every_declared_symbol(lib):
    declared = _declared_symbols()
    assert len(declared) >= 25
    handle = ctypes.CDLL(lib.LIB_PATH)
    missing = [s for s in declared if not hasattr(handle, s)]
    assert not missing, f'declared in include/moseq_b200.h but not exported: {missing}'
    unbound = [s for s in declared if s not in lib.SIGNATURES]
    assert not unbound, f'declared but not bound in _lib.SIGNATURES: {unbound}'
    assert sorted(lib.SIGNATURES) == declared


def test_library_metadata_without_gpu(lib):
    l = lib.load()
    assert l.msq_version() == 100
    names = lib.scalar_names()
    assert names[0] == 'centroid_x_px' and names[-1] == 'velocity_theta' and len(names) == 17
    cols = lib.keypoint_col_names()
    assert len(cols) == 96 and cols[0] == 'reference/Nose_x_px' and cols[-1] == 'rotated/TailTip_z_mm'
    from moseq2_detectron_extract_b200.proc.keypoints import keypoint_attributes
    from moseq2_detectron_extract_b200.proc.scalars import scalar_attributes
    assert set(cols) == set(keypoint_attributes()) and set(names) == set(scalar_attributes())
    assert lib.kernel_names()[0] == 'prep_frames'
    assert l.msq_frame_features_scratch_bytes(10, 240, 240) == 44          # count + one index per frame
    assert l.msq_extract_scratch_bytes(1000, 240, 240) >= 1000 * (8 + 8 + 64)


def test_argument_validation_returns_error_codes(lib):
    l = lib.load()
    assert l.msq_clean_frames(None, None, 5, 8, 8, None) == -1
    assert 'null' in lib.last_error()
    assert l.msq_prep_frames(None, 0, 8, 8, None, 0, None, 0, 0, 8, 8, 0.0, 1.0, 0, None, None, None, None) == 0   # n = 0 is a no-op
    assert l.msq_prep_frames(None, 1, 8, 8, None, 0, None, 4, 4, 8, 8, 0.0, 1.0, 0, None, None, None, None) == -1
    assert l.msq_inpaint_frames(None, None, None, 0, 8, 8, 3, None, 0, None) == 0
    assert l.msq_inpaint_frames(None, None, None, 2, 8, 8, 9, None, 0, None) == -1
    assert l.msq_inpaint_scratch_bytes(3, 240, 240) >= 3 * 240 * 240 * 16
    with pytest.raises(lib.MoseqB200Error):
        lib.call('msq_scale_frames', None, None, 16, 1.0, 1.0, 0, None)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip('CUDA present')
    from moseq2_detectron_extract_b200 import _dev
    import moseq2_detectron_extract_b200.proc as P
    for call in (lambda: P.clean_frames(np.zeros((1, 8, 8), np.uint8), iters_tail=3),
                 lambda: P.prep_raw_frames(np.zeros((1, 8, 8), np.int16)),
                 lambda: P.scale_raw_frames(np.zeros((1, 8, 8), np.uint8), 0, 100),
                 lambda: P.crop_and_rotate_frames_batch(np.zeros((1, 8, 8), np.uint8), np.zeros((1, 2)), np.zeros(1))):
        with pytest.raises(_dev.CudaRequiredError):
            call()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'moseq2_detectron_extract_b200')
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                text = open(os.path.join(base, f)).read()
                assert 'extract_oracle' not in text and 'oracle/' not in text and 'import oracle' not in text, f


def test_host_helpers_match_oracle():
    import cv2
    import extract_oracle as O
    from moseq2_detectron_extract_b200.proc import convert_pxs_to_mm, im_moment_features, rotate_points, rotate_points_batch
    rng = np.random.default_rng(0)
    pts = rng.uniform(0, 240, (50, 2))
    assert np.array_equal(convert_pxs_to_mm(pts, true_depth=673.0), O.px_to_mm(pts, 673.0))
    kp = rng.uniform(0, 240, (6, 8, 3))
    cen = rng.uniform(50, 200, (6, 2))
    ang = rng.uniform(0, 360, 6)
    ref = O.rotate_about(kp[..., :2], cen, ang)
    got = rotate_points_batch(kp.copy(), cen, ang)
    assert np.allclose(got[..., :2], ref, rtol=0, atol=1e-12) and np.array_equal(got[..., 2], kp[..., 2])
    assert np.allclose(rotate_points(kp[0, :, :2], cen[0], ang[0]), ref[0], atol=1e-12)
    m = np.zeros((40, 50), np.uint8)
    cv2.ellipse(m, (25, 20), (15, 8), 30, 0, 360, 1, -1)
    cnt, _ = cv2.findContours(m, cv2.RETR_TREE, cv2.CHAIN_APPROX_SIMPLE)
    mine = im_moment_features(cnt[0])
    mo = cv2.moments(cnt[0])
    assert np.allclose(mine['centroid'], [mo['m10'] / mo['m00'], mo['m01'] / mo['m00']], rtol=1e-12)
    ref_f = O.frame_features_cv2((m * 9)[None], m[None])
    assert np.isclose(mine['orientation'], ref_f['orientation'][0], atol=1e-12)
    assert np.allclose(mine['axis_length'], ref_f['axis_length'][0], rtol=1e-10)


def test_instances_container():
    import torch
    from moseq2_detectron_extract_b200.model import Boxes, Instances, create_empty_instances
    inst = Instances((240, 240), pred_boxes=Boxes(torch.tensor([[0., 0, 10, 10], [5, 5, 20, 30], [1, 1, 2, 2]])),
                     scores=torch.tensor([0.9, 0.8, 0.1]), pred_masks=torch.zeros((3, 240, 240), dtype=torch.bool),
                     pred_keypoints=torch.zeros((3, 8, 3)))
    assert len(inst) == 3 and inst.image_size == (240, 240)
    assert len(inst[1]) == 1 and torch.equal(inst[1].scores, torch.tensor([0.8]))
    assert len(inst[torch.tensor([True, False, True])]) == 2
    assert len(inst[[2, 0]]) == 2 and float(inst[[2, 0]].scores[0]) == pytest.approx(0.1)
    assert torch.allclose(inst.pred_boxes.get_centers()[1], torch.tensor([12.5, 17.5]))
    both = Instances.cat([inst[0], inst[2]])
    assert len(both) == 2 and both.pred_masks.shape == (2, 240, 240)
    empty = create_empty_instances(240, 240, 8)
    assert len(empty) == 0 and empty.pred_keypoints.shape == (0, 8, 3) and empty.to('cpu').image_size == (240, 240)
    with pytest.raises(AttributeError):
        _ = inst.nonexistent


def test_pipeline_runtime_propagates_data_sentinel_and_errors():
    from moseq2_detectron_extract_b200.pipeline import Pipeline, PipelineStep, ProducerPipelineStep, WorkerError

    class Source(ProducerPipelineStep):
        def initialize(self):
            self.i = 0

        def process(self, data):
            self.i += 1
            return {'batch': self.i, 'x': self.i * 2} if self.i <= 5 else None

    class Double(PipelineStep):
        def process(self, data):
            data['x'] *= 2
            self.update_progress(1)
            return data

    class Sink(PipelineStep):
        def initialize(self):
            self.seen = []

        def process(self, data):
            self.seen.append((data['batch'], data['x']))
            return data

    cfg = {'nframes': 5}
    pipe = Pipeline()
    src, dbl, sink = pipe.add_step(Source(cfg, 'src')), pipe.add_step(Double(cfg, 'dbl')), pipe.add_step(Sink(cfg, 'sink'))
    pipe.link(src, dbl)
    pipe.link(dbl, sink)
    pipe.run()
    assert sink.seen == [(i, i * 4) for i in range(1, 6)]
    assert all(s.is_complete.is_set() for s in (src, dbl, sink))

    class Boom(PipelineStep):
        def process(self, data):
            raise RuntimeError('boom')

    pipe = Pipeline()
    src, boom = pipe.add_step(Source(cfg, 'src')), pipe.add_step(Boom(cfg, 'boom'))
    pipe.link(src, boom)
    with pytest.raises(WorkerError, match='boom'):
        pipe.run()
    msgs = []
    while True:
        try:
            msgs.append(pipe.progress.get_nowait())
        except queue.Empty:
            break
    assert any(m.get('raise') for m in msgs if 'message' in m)


def test_synthetic_session_iterator_and_chunking():
    from moseq2_detectron_extract_b200 import shard, synthetic
    sess = synthetic.SyntheticSession(25, seed=3)
    it = sess.iterate(10, 0)
    seen = []
    it.attach_filter(filterer=lambda f: f[:, :2, :2])
    for idxs, frames in it:
        seen.append((idxs[0], idxs[-1], frames.shape))
    assert seen == [(0, 9, (10, 2, 2)), (10, 19, (10, 2, 2)), (20, 24, (5, 2, 2))]
    a = synthetic.generate_chunk(6, seed=1, t0=100)
    b = synthetic.generate_chunk(6, seed=1, t0=100)
    assert np.array_equal(a.frames, b.frames) and np.array_equal(a.keypoints, b.keypoints, equal_nan=True)
    assert [list(r) for r in shard.chunk_ranges(25, 10)] == [list(range(0, 10)), list(range(10, 20)), list(range(20, 25))]
    # overlap: the reference's gen_batch_sequence strides by chunk_size - overlap (io/util.py:33-35)
    assert shard.chunk_ranges(25, 10, 3) == [range(0, 10), range(7, 17), range(14, 24), range(21, 25)]
    assert shard.chunk_ranges(20, 10, 0) == [range(0, 10), range(10, 20)] and shard.chunk_ranges(0, 10) == []
    roi = synthetic.make_roi(synthetic.SessionGeometry())
    y0, x0, y1, x1 = synthetic.roi_bbox(roi)
    assert (y1 - y0, x1 - x0) == (240, 240) and x0 % 8 == 0


@pytest.mark.parametrize('n_chunks,world', [(54, 1), (54, 2), (54, 4), (54, 8), (5, 8), (0, 3), (864, 8)])
def test_shard_chunks_partition(n_chunks, world):
    from moseq2_detectron_extract_b200.shard import shard_chunks
    owned = [shard_chunks(n_chunks, r, world) for r in range(world)]
    flat = [c for rng in owned for c in rng]
    assert flat == list(range(n_chunks))                 # contiguous, ordered, exactly once
    sizes = [len(r) for r in owned]
    assert max(sizes) - min(s for s in sizes if s or True) <= -(-n_chunks // world)
    with pytest.raises(ValueError):
        shard_chunks(10, 3, 3)


def test_read_frames_raw_matches_reference_semantics(tmp_path):
    """a1 decode-to-tensor (ref io/video.py:67-127): headerless little-endian int16, plain file and tar member,
    contiguous / unordered / repeated indices."""
    import tarfile
    from moseq2_detectron_extract_b200.io import RawDepthSession, get_raw_info, read_frames_raw
    rng = np.random.default_rng(0)
    W, H, N = 32, 20, 17
    frames = rng.integers(-5, 4000, size=(N, H, W)).astype('<i2')
    path = tmp_path / 'depth.dat'
    frames.tofile(path)
    info = get_raw_info(str(path), frame_dims=(W, H))
    assert info == {'bytes': N * H * W * 2, 'nframes': N, 'dims': (W, H), 'bytes_per_frame': H * W * 2}
    assert np.array_equal(read_frames_raw(str(path), frame_dims=(W, H)), frames)
    assert np.array_equal(read_frames_raw(str(path), range(3, 11), frame_dims=(W, H)), frames[3:11])
    assert np.array_equal(read_frames_raw(str(path), 5, frame_dims=(W, H)), frames[5:6])
    pick = [9, 2, 3, 4, 16, 2, 0]
    got = read_frames_raw(str(path), pick, frame_dims=(W, H))
    assert got.dtype == np.dtype('<i2') and np.array_equal(got, frames[pick])
    with tarfile.open(tmp_path / 'session.tar.gz', 'w:gz') as tar:
        tar.add(path, arcname='depth.dat')
    with tarfile.open(tmp_path / 'session.tar.gz', 'r:gz') as tar:
        member = tar.getmember('depth.dat')
        assert np.array_equal(read_frames_raw(member, range(4, 9), frame_dims=(W, H), tar_object=tar), frames[4:9])
        assert np.array_equal(read_frames_raw(member, [8, 1], frame_dims=(W, H), tar_object=tar), frames[[8, 1]])
    sess = RawDepthSession(str(path), np.zeros((H, W), np.float32), np.ones((H, W), bool), 673.0, frame_dims=(W, H), pinned=False)
    chunks = list(sess.iterate(7, 0))
    assert [c[0][0] for c in chunks] == [0, 7, 14] and np.array_equal(np.concatenate([c[1] for c in chunks]), frames)
    with pytest.raises(EOFError):
        read_frames_raw(str(path), range(10, 30), frame_dims=(W, H))


def test_kalman_items_build_the_reference_model_blocks():
    """Model blocks of the tracker items (host side, no GPU): constant-jerk transition per coordinate, position-only
    observation, same layout as reference proc/kalman.py:146-278 (restated in oracle/tracking_oracle.py)."""
    import tracking_oracle as TO
    from moseq2_detectron_extract_b200.proc import kalman as K
    for item, n_coords in ((K.KalmanTrackerPoint1D(3, 1.0), 1), (K.KalmanTrackerPoint2D(3, 1.0), 2),
                           (K.KalmanTrackerAngle(3, 1.0), 2), (K.KalmanTrackerNPoints2D(8, 3, 1.0), 16),
                           (K.KalmanTrackerPoint2D(4, 0.5), 2)):
        ref = TO.Tracker(n_coords, item.order)
        if item.delta_t == 1.0:
            assert np.array_equal(item.build_trans_mat(), ref.A)
        assert np.array_equal(item.build_observ_mat(), ref.H)
        assert item.state_size == n_coords * item.order and item.obs_size == n_coords
    blk = K.KalmanTrackerPoint1D(4, 0.5).build_trans_mat()
    assert np.allclose(blk, [[1, .5, .125, .125 / 6], [0, 1, .5, .125], [0, 0, 1, .5], [0, 0, 0, 1]])
    with pytest.raises(ValueError):
        K.KalmanTracker([])
    with pytest.raises(ValueError):
        K.KalmanTracker([K.KalmanTrackerPoint2D(3, 1.0), K.KalmanTrackerPoint2D(3, 2.0)])
    tr = K.KalmanTracker([K.KalmanTrackerPoint2D(3, 1.0), K.KalmanTrackerNPoints2D(8, 3, 1.0)])
    assert (tr.n_state, tr.n_obs, tr.is_initialized) == (54, 18, False)
    assert np.allclose(K.angle_difference(np.array([350.0, 10.0, 0.0]), np.array([10.0, 350.0, 180.0])), [20.0, -20.0, 180.0])


def test_numa_binding_helpers():
    from moseq2_detectron_extract_b200.shard import bind_to_gpu_numa_node, parse_cpulist
    assert parse_cpulist('0-3,8,10-11\n') == [0, 1, 2, 3, 8, 10, 11] and parse_cpulist('') == [] and parse_cpulist('5') == [5]
    before = os.sched_getaffinity(0)
    out = bind_to_gpu_numa_node(0)                 # no GPU / no NVML here: must degrade to None and leave the affinity alone
    assert out is None or set(os.sched_getaffinity(0)) <= set(before)
    os.sched_setaffinity(0, before)


def test_result_writer_step_layout_and_offsets(tmp_path):
    """ResultWriterStep (ref pipeline/write_results_step.py:14-73, io/result.py:14-130): dataset paths, dtypes, chunk offsets
    (overlap frames dropped) and the keypoint TSV columns, through the pipeline's queue / sentinel machinery.  No GPU."""
    import queue as _queue
    import threading
    from moseq2_detectron_extract_b200.pipeline.write_results_step import ResultWriterStep
    from moseq2_detectron_extract_b200.proc.keypoints import keypoint_attributes
    from moseq2_detectron_extract_b200.proc.scalars import scalar_attributes
    n, crop = 7, (8, 8)
    cfg = {'nframes': n, 'crop_size': crop, 'output_dir': str(tmp_path), 'bg_roi_index': 0, 'true_depth': 673.0,
           'roi': np.ones((4, 4), bool), 'bground_im': np.full((4, 4), 673.0), 'timestamps': np.arange(n) * 33.3}

    def chunk(idxs, offset):
        m = len(idxs)
        return {'batch': 0, 'chunk': np.zeros((m, 4, 4), np.uint8), 'frame_idxs': idxs, 'offset': offset,
                'scalars': {k: np.array(idxs, dtype=float) + 0.5 for k in scalar_attributes()},
                'keypoints': {k: np.array(idxs, dtype=float) * 2 for k in keypoint_attributes()},
                'features': {'flips': np.array([i % 2 == 0 for i in idxs]),
                             'features': {'centroid': np.stack([np.array(idxs, float), np.array(idxs, float) + 1], 1),
                                          'orientation': np.array(idxs, float) * 10}},
                'depth_frames': np.stack([np.full(crop, i, np.uint8) for i in idxs]),
                'mask_frames': np.stack([np.full(crop, i % 2, np.uint8) for i in idxs])}

    step = ResultWriterStep(cfg, 'writer')
    step.shutdown_event = threading.Event()
    step.in_queue = _queue.Queue()
    step.in_queue.put(chunk([0, 1, 2, 3], 0))
    step.in_queue.put(chunk([3, 4, 5, 6], 1))          # one frame of overlap with the previous chunk
    step.in_queue.put(None)
    step.run()
    assert step.error is None and step.is_complete.is_set()
    out = np.load(step.store.path) if step.store.path.endswith('.npz') else None
    if out is not None:
        assert out['frames'].shape == (n, 8, 8) and out['frames'].dtype == np.uint8 and out['frames_mask'].dtype == bool
        assert [int(f[0, 0]) for f in out['frames']] == list(range(n))
        assert np.array_equal(out['scalars/centroid_x_px'], np.arange(n, dtype=np.float32) + 0.5)
        assert out['scalars/area_px'].dtype == np.float32 and len([k for k in out.files if k.startswith('keypoints/')]) == 96
        assert np.array_equal(out['metadata/extraction/flips'], np.arange(n) % 2 == 0)
        assert float(out['metadata/extraction/true_depth']) == 673.0 and out['metadata/extraction/background'].shape == (4, 4)
    lines = open(step.keypoint_data_dest).read().strip().split('\n')
    header = lines[0].split('\t')
    assert header[:5] == ['Frame_Idx', 'Flip', 'Centroid_X', 'Centroid_Y', 'Angle'] and len(header) == 5 + 96
    assert len(lines) == 1 + 8 and lines[1].split('\t')[:3] == ['0', 'True', '0.0']
    # uuid / parameters / acquisition metadata (ref io/result.py:31,88-102) and the status file with its `complete` flag
    import yaml
    cfg2 = dict(cfg, output_dir=str(tmp_path / 'second'),
                status_dict={'uuid': 'abc-123', 'parameters': {'chunk_size': 1000, 'crop_size': (80, 80), 'fn': None},
                             'metadata': {'SubjectName': 'mouse1', 'DepthResolution': [512, 424], 'Empty': None}})
    step2 = ResultWriterStep(cfg2, 'writer')
    step2.shutdown_event = threading.Event()
    step2.in_queue = _queue.Queue()
    step2.in_queue.put(chunk([0, 1, 2, 3], 0))
    step2.in_queue.put(None)
    step2.run()
    assert step2.store.path.endswith(('results_00.npz', 'results_00.h5'))
    status = yaml.safe_load(open(step2.status_filename))
    assert status['complete'] is True and status['uuid'] == 'abc-123' and status['frames_written'] == 4
    if step2.store.path.endswith('.npz'):
        out2 = np.load(step2.store.path)
        assert str(out2['metadata/uuid']) == 'abc-123' and int(out2['metadata/extraction/parameters/chunk_size']) == 1000
        assert out2['metadata/acquisition/DepthResolution'].tolist() == [512, 424] and str(out2['metadata/acquisition/SubjectName']) == 'mouse1'
    # a run that is shut down before the end-of-stream sentinel must not leave a file that looks finished
    cfg3 = dict(cfg, output_dir=str(tmp_path / 'third'))
    step3 = ResultWriterStep(cfg3, 'writer')
    step3.shutdown_event = threading.Event()
    step3.in_queue = _queue.Queue()
    step3.in_queue.put(chunk([0, 1, 2, 3], 0))
    t = threading.Thread(target=step3.run)
    t.start()
    import time as _time
    while step3.in_queue.qsize() > 0:
        _time.sleep(0.01)
    _time.sleep(0.2)
    step3.shutdown_event.set()                          # an upstream failure shuts the pipeline down
    t.join(timeout=10)
    assert '.partial.' in os.path.basename(step3.store.path) and not os.path.exists(os.path.join(str(tmp_path / 'third'), 'results_00.npz'))
    assert yaml.safe_load(open(step3.status_filename))['complete'] is False


def test_tiff_cache_round_trip_and_find_roi_cache_files(tmp_path):
    """io/image.py (ref io/image.py:13-103): scaled uint16 / uint8 TIFFs with the scale factor in the ImageDescription, readable by
    an independent decoder (OpenCV); the three cache files of find_roi (ref io/session.py:194-257) carry those conventions."""
    import cv2
    from moseq2_detectron_extract_b200.io.image import read_tiff_image, write_image
    rng = np.random.default_rng(0)
    bg = rng.uniform(600, 700, (48, 64))
    path = str(tmp_path / 'cache' / 'bground.tiff')
    write_image(path, bg, scale=True)
    back = read_tiff_image(path, scale=True)
    assert back.dtype == np.uint16 and np.array_equal(back, bg.astype('uint16'))          # the uint16 background of SURVEY trap 8
    raw = cv2.imread(path, cv2.IMREAD_UNCHANGED)
    factor = int(65535 / bg.astype('uint16').max())
    assert raw.dtype == np.uint16 and np.array_equal(raw, bg.astype('uint16') * factor)
    ff = rng.integers(500, 800, (48, 64)).astype('int16')
    write_image(str(tmp_path / 'ff.tiff'), ff, scale=True, scale_factor=(650, 750))
    got = read_tiff_image(str(tmp_path / 'ff.tiff')).astype(float)
    assert np.abs(got - np.clip(ff, 650, 750)).max() <= 1.0
    roi = rng.random((48, 64)) > 0.5
    write_image(str(tmp_path / 'roi_00.tiff'), roi, scale=True, dtype='uint8')
    assert np.array_equal(read_tiff_image(str(tmp_path / 'roi_00.tiff'), scale=True) > 0, roi)
    with pytest.raises(NotImplementedError):
        write_image(str(tmp_path / 'x.tiff'), roi, compress=3)


def test_select_strel_and_sobel_kernels_equal_opencv():
    import cv2
    from moseq2_detectron_extract_b200.proc.roi import plane_fit3, sobel_kernels
    from moseq2_detectron_extract_b200.proc.util import select_strel
    for w in range(1, 24):
        for h in range(1, 24):
            assert np.array_equal(select_strel('ellipse', (w, h)), cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (w, h))), (w, h)
            assert np.array_equal(select_strel('rect', (w, h)), cv2.getStructuringElement(cv2.MORPH_RECT, (w, h)))
    assert np.array_equal(select_strel('x', (10, 10)), select_strel('e', (10, 10)))      # unknown shapes mean ellipse
    for k in (1, 3, 5, 7, 9, 31):
        kx, ky = cv2.getDerivKernels(1, 0, k, normalize=False, ktype=cv2.CV_64F)
        deriv, smooth = sobel_kernels(k)
        assert np.array_equal(deriv, kx.ravel()) and np.array_equal(smooth, ky.ravel())
    with pytest.raises(ValueError):
        sobel_kernels(4)
    plane = plane_fit3(np.array([[0., 0., 2.], [4., 0., 2.], [0., 3., 2.]]))
    assert np.allclose(plane, [0, 0, 1, -2])
    assert np.isnan(plane_fit3(np.zeros((3, 3)))).all()


def test_raw_session_archive_metadata_timestamps_and_trim(tmp_path):
    """RawDepthSession over a .tar.gz session archive (ref io/session.py:50-178): depth.dat member, DepthResolution from
    metadata.json, depth_ts.txt / timestamps.csv, frame_trim like __trim_frames."""
    import json
    import tarfile
    from moseq2_detectron_extract_b200.io import RawDepthSession
    rng = np.random.default_rng(1)
    W, H, N = 24, 18, 23
    frames = rng.integers(0, 3000, size=(N, H, W)).astype('<i2')
    (tmp_path / 'plain').mkdir()
    frames.tofile(tmp_path / 'plain' / 'depth.dat')
    (tmp_path / 'plain' / 'metadata.json').write_text(json.dumps({'DepthResolution': [W, H], 'SubjectName': 'm1'}))
    stamps = np.arange(N) * 33.3 + 1000.0
    (tmp_path / 'plain' / 'depth_ts.txt').write_text(''.join(f'{t:.3f} {i}\n' for i, t in enumerate(stamps)))
    with tarfile.open(tmp_path / 'session_20260101.tar.gz', 'w:gz') as tar:
        for name in ('depth.dat', 'metadata.json', 'depth_ts.txt'):
            tar.add(tmp_path / 'plain' / name, arcname=name)

    for path, compressed in ((str(tmp_path / 'plain' / 'depth.dat'), False), (str(tmp_path / 'session_20260101.tar.gz'), True)):
        sess = RawDepthSession(path, frame_dims=None, pinned=False)
        assert sess.is_compressed == compressed and sess.frame_dims == (W, H) and sess.nframes == N
        assert sess.load_metadata()['SubjectName'] == 'm1'
        np.testing.assert_allclose(sess.load_timestamps(), np.round(stamps, 3))
        chunks = list(sess.iterate(10, 0))
        assert [list(c[0]) for c in chunks] == [list(range(0, 10)), list(range(10, 20)), list(range(20, 23))]
        assert np.array_equal(np.concatenate([c[1] for c in chunks]), frames)
        assert np.array_equal(sess.read_frames([22, 3]), frames[[22, 3]])
        # trimming: 4 frames off the head, 5 off the tail; session indices restart at 0
        trimmed = RawDepthSession(path, frame_dims=(W, H), pinned=False, frame_trim=(4, 5))
        assert (trimmed.first_frame_idx, trimmed.last_frame_idx, trimmed.nframes) == (4, N - 5, N - 9)
        assert np.array_equal(np.concatenate([c[1] for c in trimmed.iterate(6, 0)]), frames[4:N - 5])
        np.testing.assert_allclose(trimmed.load_timestamps(), np.round(stamps, 3)[4:N - 5])
        # a trim that would leave nothing is ignored, like the reference
        assert RawDepthSession(path, frame_dims=(W, H), pinned=False, frame_trim=(40, 0)).nframes == N
    assert RawDepthSession(str(tmp_path / 'session_20260101.tar.gz'), frame_dims=(W, H), pinned=False).session_id == 'session_20260101'
    # timestamps.csv fallback is in seconds -> milliseconds
    (tmp_path / 'plain' / 'depth_ts.txt').unlink()
    (tmp_path / 'plain' / 'timestamps.csv').write_text(''.join(f'{t / 1000.0:.6f}\n' for t in stamps))
    np.testing.assert_allclose(RawDepthSession(str(tmp_path / 'plain' / 'depth.dat'), frame_dims=(W, H), pinned=False).load_timestamps(),
                               stamps, rtol=1e-9)
    (tmp_path / 'plain' / 'timestamps.csv').unlink()
    with pytest.raises(ValueError):
        RawDepthSession(str(tmp_path / 'plain' / 'depth.dat'), frame_dims=(W, H), pinned=False).load_timestamps()


def test_ransac_triples_reproduce_the_reference_draws():
    """One randint call for all RANSAC iterations == the reference's per-iteration np.random.choice(n, 3, replace=True)
    (proc/roi.py:170): same numbers, same generator state afterwards."""
    from moseq2_detectron_extract_b200.proc.roi import _ransac_triples
    for npoints in (1, 2, 7, 45225, 217088, 368640, 5_000_000):
        for seed, iters in ((0, 1), (3, 5), (11, 1000)):
            np.random.seed(seed)
            want = np.stack([np.random.choice(npoints, 3, replace=True) for _ in range(iters)])
            state_want = np.random.get_state()
            np.random.seed(seed)
            got = _ransac_triples(npoints, iters)
            state_got = np.random.get_state()
            assert got.dtype == want.dtype and np.array_equal(got, want)
            assert np.array_equal(state_got[1], state_want[1]) and state_got[2] == state_want[2]
    assert _ransac_triples(10, 0).shape == (0, 3)


def test_rcnn_graph_scripts_and_keeps_export_contract(tmp_path):
    """model/rcnn.py: the graph is TorchScript-scriptable, survives save / load, carries the reference's export contract
    (ref model/deploy.py:91-97: forward(List[Dict[str, Tensor]]) -> List[Dict[str, Tensor]]) plus the batched `forward_dense`
    entry, and fails loudly on CPU tensors (no CPU implementation of torch.ops.msq.*)."""
    import torch
    from moseq2_detectron_extract_b200.model import rcnn
    model = rcnn.finalize(rcnn.MoseqRCNN(post_nms_topk=100), torch.float32, 'cpu')
    path = str(tmp_path / 'model.ts')
    rcnn.export_torchscript(model, path)
    loaded = torch.jit.load(path)
    assert int(loaded.graph_version) == rcnn.GRAPH_VERSION and loaded.input_format == 'RGB' and int(loaded.post_nms_topk) == 100
    schema = str(loaded.forward.schema)
    assert 'Dict(str, Tensor)[] inputs' in schema and schema.endswith('-> Dict(str, Tensor)[]')
    assert hasattr(loaded, 'forward_dense')
    for op in ('conv2d', 'group_norm_nhwc', 'rpn_proposals', 'roi_align_v2', 'fastrcnn_top1', 'keypoints_from_heatmaps_d2', 'linear'):
        assert f'msq::{op}' in str(loaded.inlined_graph) or f'msq::{op}' in str(loaded.forward_dense.inlined_graph), op
    with pytest.raises(Exception):
        loaded([{'image': torch.zeros((3, 64, 64), dtype=torch.uint8)}])
    with pytest.raises(NotImplementedError):
        rcnn.MoseqRCNN(detections_per_image=2)


def test_detectron2_state_dict_rewrites_are_exact_on_cpu():
    """from_detectron2_state_dict's rewrites checked with plain torch on the CPU: FrozenBN folding == conv + FrozenBN, the fc1
    column permutation == flatten of (C, 7, 7) vs (7, 7, C), the merged RPN predictor == the two 1x1 convolutions."""
    import torch
    import torch.nn.functional as F
    import d2_rcnn_oracle as D
    from moseq2_detectron_extract_b200.model import rcnn
    st = D.make_random_state(1)
    model = rcnn.from_detectron2_state_dict(st, dtype=torch.float32, device='cpu')
    g = torch.Generator().manual_seed(0)
    x = torch.randn((2, 64, 20, 20), generator=g)
    blk = model.res2[0]
    p = 'backbone.bottom_up.res2.0.'
    for conv, name, stride, pad in ((blk.conv1, 'conv1', 1, 0), (blk.shortcut, 'shortcut', 1, 0)):
        want = D.conv_frozen_bn(x, st, p + name, stride, pad)
        got = F.conv2d(x, conv.weight, conv.bias, stride, pad)
        assert float((want - got).abs().max()) <= 1e-5 * float(want.abs().max())
    assert model.res3[0].conv1.stride == 2 and model.res3[0].conv2.stride == 1 and model.res3[0].shortcut.stride == 2    # STRIDE_IN_1X1
    roi = torch.randn((5, 256, 7, 7), generator=g)
    want = F.linear(roi.flatten(1), st['roi_heads.box_head.fc1.weight'], st['roi_heads.box_head.fc1.bias'])
    got = F.linear(roi.permute(0, 2, 3, 1).reshape(5, -1), model.fc1_w, model.fc1_b)
    assert float((want - got).abs().max()) <= 1e-4 * float(want.abs().max())
    t = torch.randn((2, 256, 6, 6), generator=g)
    rp = 'proposal_generator.rpn_head.'
    both = F.conv2d(t, model.rpn_pred.weight, model.rpn_pred.bias)
    assert torch.allclose(both[:, :3], F.conv2d(t, st[rp + 'objectness_logits.weight'], st[rp + 'objectness_logits.bias']), atol=1e-6)
    assert torch.allclose(both[:, 3:15], F.conv2d(t, st[rp + 'anchor_deltas.weight'], st[rp + 'anchor_deltas.bias']), atol=1e-6)
    assert float(both[:, 15].abs().max()) == 0
    pred = F.linear(torch.randn((3, 1024), generator=g), model.box_pred_w, model.box_pred_b)
    assert pred.shape == (3, 8) and float(pred[:, 6:].abs().max()) == 0
    assert model.pixel_mean == pytest.approx([1.12] * 3) and model.pixel_std == pytest.approx([5.79] * 3)
    assert tuple(model.kp_deconv_w.shape) == (512, 8, 4, 4) and model.keypoint_pooler == 7 and model.post_nms_topk == 1000


def test_sort_tracker_follows_norfair_semantics():
    """proc/sort_tracker.py (a15; norfair's Tracker for the reference's configuration, ref process_features_step.py:35-38):
    greedy nearest matching below the threshold, hit counters (+2 on a hit capped at 3, -1 per frame, dropped below 0), ages,
    live points, constant-velocity prediction, identity kept through a crossing and through a short occlusion."""
    from moseq2_detectron_extract_b200.proc.sort_tracker import Detection, TrackedObject, Tracker
    TrackedObject._next_id = 0
    trk = Tracker(distance_function='euclidean', distance_threshold=50, initialization_delay=0, hit_counter_max=3)
    assert trk.update(detections=[]) == []
    # frame 0: one animal -> initialised at once (initialization_delay = 0), returned as active
    out = trk.update(detections=[Detection(np.array([100.0, 100.0]), data='a0')])
    assert len(out) == 1 and out[0].id == 0 and out[0].age == 0 and out[0].hit_counter == 1 and out[0].live_points.all()
    # frames 1..3: it moves 5 px per frame; a second animal appears at frame 2 far away
    for t in range(1, 4):
        dets = [Detection(np.array([100.0 + 5 * t, 100.0]), data=f'a{t}')]
        if t >= 2:
            dets.insert(0, Detection(np.array([30.0, 200.0]), data=f'b{t}'))
        out = trk.update(detections=dets)
    assert sorted(o.id for o in out) == [0, 1]
    a = next(o for o in out if o.id == 0)
    b = next(o for o in out if o.id == 1)
    assert a.age == 3 and b.age == 1 and a.hit_counter == 3 and a.last_detection.data == 'a3' and b.last_detection.data == 'b3'
    assert 108.0 < a.estimate[0, 0] <= 115.0 and a.estimate[0, 1] == 100.0      # the filter follows the motion (it starts at rest)
    # the reference keeps the OLDEST live object: sorted by age, pop from the end
    oldest = sorted((o for o in out if o.live_points.any()), key=lambda o: o.age).pop()
    assert oldest.id == 0
    # a detection beyond the threshold starts a new object instead of moving an old one
    out = trk.update(detections=[Detection(np.array([120.0, 100.0])), Detection(np.array([30.0, 200.0])), Detection(np.array([230.0, 20.0]))])
    assert sorted(o.id for o in out) == [0, 1, 2]
    # occlusion: object 0 unseen for 3 frames keeps its identity (hit counter 3 -> 2 -> 1 -> 0), the 4th miss ends it
    for t in range(3):
        out = trk.update(detections=[Detection(np.array([30.0, 200.0]))])
        assert 0 in [o.id for o in out]
    a = next(o for o in out if o.id == 0)
    assert a.hit_counter == 0 and int(a.point_hit_counter[0]) == 1 and a.live_points.any()   # point counters are capped at 4, object counters at 3
    out = trk.update(detections=[Detection(np.array([30.0, 200.0]))])
    assert 0 not in [o.id for o in out]
    # greedy matching: smallest distance first, one detection per object
    d = np.array([[10.0, 12.0], [11.0, 60.0]])
    det_idxs, obj_idxs = Tracker.match_dets_and_objs(d, 50.0)
    assert (det_idxs, obj_idxs) == ([0], [0])
    det_idxs, obj_idxs = Tracker.match_dets_and_objs(np.array([[10.0, 12.0], [11.0, 40.0]]), 50.0)
    assert (det_idxs, obj_idxs) == ([0, 1], [0, 1])
    with pytest.raises(ValueError):
        trk._update_objects_in_place(trk.tracked_objects[:1], [Detection(np.array([np.nan, 1.0]))], 1)


def test_every_declared_entry_point_is_documented_in_integration_md():
    """include/moseq_b200.h is the boundary; INTEGRATION.md names, for every entry point, the reference function it replaces."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    header = open(os.path.join(root, 'include', 'moseq_b200.h')).read()
    names = re.findall(r'MSQ_API\s+[\w\s\*]+?\b(msq_\w+)\s*\(', header)
    assert len(names) >= 70
    text = open(os.path.join(root, 'INTEGRATION.md')).read()
    assert [n for n in names if n not in text] == []


def test_roi_bands_cover_the_roi_and_stay_inside_the_box():
    """_dev.roi_bands (host side of msq_copy_roi_bands): every band covers the ROI pixels of its rows, columns are rounded outwards
    to the alignment and clipped to the box, bands tile the rows without gaps; a disc costs ~15 % fewer bytes than its box."""
    from moseq2_detectron_extract_b200 import _dev, synthetic
    rng = np.random.default_rng(0)
    geom = synthetic.SessionGeometry()
    roi = synthetic.make_roi(geom)
    y0, x0, y1, x1 = synthetic.roi_bbox(roi)
    disc = roi[y0:y1, x0:x1]
    ragged = rng.random((37, 50)) > 0.7
    ragged[5:9] = False                                               # rows without any ROI pixel
    for box, n_bands, align in ((disc, 16, 8), (disc, 1, 8), (disc, 1000, 4), (ragged, 7, 8), (np.zeros((10, 16), bool), 3, 8)):
        by, bx0, bx1 = _dev.roi_bands(box, n_bands, align)
        h, w = box.shape
        assert by[0] == 0 and by[-1] == h and np.all(np.diff(by) >= 0) and len(bx0) == len(bx1) == len(by) - 1
        assert by.dtype == bx0.dtype == bx1.dtype == np.int32
        covered = np.zeros_like(box)
        for b in range(len(bx0)):
            assert 0 <= bx0[b] <= bx1[b] <= w
            assert bx0[b] % align == 0 and (bx1[b] % align == 0 or bx1[b] == w)
            covered[by[b]:by[b + 1], bx0[b]:bx1[b]] = True
        assert not (box & ~covered).any()
    by, bx0, bx1 = _dev.roi_bands(disc, 16, 8)
    area = sum((by[b + 1] - by[b]) * (bx1[b] - bx0[b]) for b in range(16))
    assert 0.80 < area / disc.size < 0.87

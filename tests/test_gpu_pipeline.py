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


def test_predictor_random_init_smoke():
    """R-CNN path (BASELINE configs[2]): random-init Keypoint+Mask R-CNN, outputs only checked for structure."""
    pytest.importorskip('torchvision')
    from moseq2_detectron_extract_b200 import synthetic
    from moseq2_detectron_extract_b200.model.predict import Predictor
    from moseq2_detectron_extract_b200.proc import prep_raw_frames
    geom = synthetic.SessionGeometry()
    ch = synthetic.generate_chunk(4, seed=9, geom=geom)
    prep = prep_raw_frames(torch.from_numpy(ch.frames).cuda(), bground_im=synthetic.make_background(geom),
                           roi=synthetic.make_roi(geom), vmin=0, vmax=100)
    pred = Predictor.from_random_init(detections_per_img=2)
    out = pred.predict_prepared(prep, 0, 100)
    assert len(out) == 4
    for o in out:
        inst = o['instances']
        assert inst.image_size == (240, 240)
        k = len(inst)
        assert inst.pred_masks.shape == (k, 240, 240) and inst.pred_masks.dtype == torch.bool
        assert inst.pred_keypoints.shape == (k, 8, 3) and inst.pred_boxes.tensor.shape == (k, 4)
    # reference-shaped call: (N, H, W, 1) uint8 numpy, scaled like InferenceStep does
    from moseq2_detectron_extract_b200.proc import scale_raw_frames
    out2 = pred(scale_raw_frames(prep.cpu().numpy()[:, :, :, None], 0, 100))
    assert len(out2) == 4 and out2[0]['instances'].image_size == (240, 240)


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


def test_dense_inference_matches_per_image_instances():
    """Predictor.predict_dense (batched detector_postprocess + first-instance gather, one paste launch) against the
    reference-shaped path: predict_prepared -> outputs_to_instances -> mask_and_keypoints_from_model_output."""
    pytest.importorskip('torchvision')
    from moseq2_detectron_extract_b200 import synthetic
    from moseq2_detectron_extract_b200.model.predict import Predictor
    from moseq2_detectron_extract_b200.proc import prep_raw_frames
    from moseq2_detectron_extract_b200.proc.proc import _gather_instances
    geom = synthetic.SessionGeometry()
    ch = synthetic.generate_chunk(6, seed=10, geom=geom)
    prep = prep_raw_frames(torch.from_numpy(ch.frames).cuda(), bground_im=synthetic.make_background(geom),
                           roi=synthetic.make_roi(geom), vmin=0, vmax=100)
    pred = Predictor.from_random_init(detections_per_img=1)
    ref_masks, ref_kpts, ref_n = _gather_instances(pred.predict_prepared(prep, 0, 100))
    masks, kpts, ninst = pred.predict_dense(prep, 0, 100)
    assert masks.shape == ref_masks.shape and masks.dtype == torch.uint8
    assert np.array_equal(ninst.cpu().numpy(), ref_n)
    assert torch.equal(masks, ref_masks)
    assert torch.allclose(kpts, ref_kpts, rtol=0, atol=0, equal_nan=True)


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


def test_batched_rcnn_heads_match_torchvision():
    """model/batched_heads.py (batched proposal filtering with ONE segmented NMS launch, arg-max detections) against
    torchvision's own per-image RegionProposalNetwork.filter_proposals / RoIHeads.postprocess_detections: same outputs."""
    pytest.importorskip('torchvision')
    from moseq2_detectron_extract_b200 import _dev, _lib, synthetic
    from moseq2_detectron_extract_b200.model.batched_heads import disable_batched_heads, enable_batched_heads
    from moseq2_detectron_extract_b200.model.predict import Predictor
    from moseq2_detectron_extract_b200.proc import prep_raw_frames
    geom = synthetic.SessionGeometry()
    ch = synthetic.generate_chunk(5, seed=12, geom=geom)
    prep = prep_raw_frames(torch.from_numpy(ch.frames).cuda(), bground_im=synthetic.make_background(geom),
                           roi=synthetic.make_roi(geom), vmin=0, vmax=100)
    pred = Predictor.from_random_init(detections_per_img=1, amp=True)          # bf16 autocast, like the bench
    net = pred.model.model
    assert hasattr(net.rpn, '_msq_filter_proposals')
    fast = pred.predict_dense(prep, 0, 100)
    disable_batched_heads(net)
    slow = pred.predict_dense(prep, 0, 100)
    enable_batched_heads(net)
    assert torch.equal(fast[2], slow[2])
    # masks: the RoI pooling kernel and torchvision's roi_align agree to an ulp of float32 (fused multiply-adds), which bf16
    # rounding can turn into a flipped pixel on a mask boundary
    assert float((fast[0] != slow[0]).float().mean()) <= 2e-4 * max(1.0, float(slow[0].float().mean()) * 100)
    # keypoints: our bicubic decode evaluates the same formula as torch's upsample kernel but without its fused
    # multiply-adds, so an arg-max can land on a neighbouring pixel when two values agree to the last bits
    diff = (fast[1] - slow[1]).abs()
    assert torch.equal(torch.isnan(fast[1]), torch.isnan(slow[1]))
    # ... and with a random-init network the heat maps are nearly flat, so the ulp-level differences of the pooled features
    # (see above) can move a maximum anywhere: most keypoints agree, the decode itself is checked exactly below
    near = (torch.nan_to_num(diff[..., :2]).amax(dim=-1) <= 1.5).float().mean()
    assert float(near) >= 0.75 and float(torch.nan_to_num(diff[..., :2]).median()) <= 1e-3, float(near)
    # the decode on its own against torchvision's per-RoI loop, float32 and bfloat16 heatmaps
    from torchvision.models.detection.roi_heads import heatmaps_to_keypoints
    from moseq2_detectron_extract_b200.model.batched_heads import keypoints_from_heatmaps
    gen0 = torch.Generator(device='cuda').manual_seed(1)
    maps = torch.randn((7, 8, 56, 56), device='cuda', generator=gen0)
    maps = torch.nn.functional.avg_pool2d(maps, 5, stride=1, padding=2) * 4           # smooth: well separated maxima
    rois = torch.tensor([[10.2, 20.7, 90.1, 140.9], [0, 0, 240, 240], [100.5, 50.5, 101.0, 51.0], [5, 5, 35.5, 200],
                         [30, 40, 230.3, 60.8], [0.5, 0.5, 239.5, 239.5], [120, 120, 180, 181]], device='cuda')
    for dtype in (torch.float32, torch.bfloat16):
        ref_xy, ref_s = heatmaps_to_keypoints(maps.to(dtype), rois)
        got_xy, got_s = keypoints_from_heatmaps(maps.to(dtype), rois)
        close = ((got_xy - ref_xy.float()).abs().amax(dim=-1) <= 1e-3)
        assert float(close.float().mean()) >= 0.97, (dtype, float(close.float().mean()))
        assert float((got_s - ref_s.float()).abs()[close].max()) <= (1e-4 if dtype == torch.float32 else 1e-1)
    # the NMS entry point on its own: random boxes in score order vs torchvision.ops.nms
    import torchvision
    gen = torch.Generator(device='cuda').manual_seed(0)
    n, K = 3, 400
    xy = torch.rand((n, K, 2), device='cuda', generator=gen) * 200
    wh = torch.rand((n, K, 2), device='cuda', generator=gen) * 60 + 1
    boxes = torch.cat([xy, xy + wh], dim=-1).contiguous()
    valid = (torch.rand((n, K), device='cuda', generator=gen) > 0.1)
    keep = torch.empty((n, 50), dtype=torch.int32, device='cuda')
    count = torch.empty((n,), dtype=torch.int32, device='cuda')
    v8 = valid.to(torch.uint8).contiguous()
    _lib.call('msq_nms_sorted', _dev.ptr(boxes), _dev.ptr(v8), n, K, 0.5, 50, _dev.ptr(keep), _dev.ptr(count), _dev.stream())
    scores = torch.arange(K, 0, -1, device='cuda', dtype=torch.float32)          # already in descending order
    for i in range(n):
        idx = torch.nonzero(valid[i])[:, 0]
        ref = idx[torchvision.ops.nms(boxes[i, idx], scores[idx], 0.5)][:50]
        c = int(count[i])
        assert c == len(ref) and torch.equal(keep[i, :c].long(), ref)


@pytest.mark.parametrize('amp', [False, True])
def test_fused_conv_epilogues_match_eager_backbone_and_heads(amp):
    """model/fused_convs.py (cuDNN conv+bias+ReLU / conv+bias+residual+ReLU) against the eager modules it replaces, on the
    FPN features and the keypoint / mask / RPN head outputs of the same random-init network."""
    pytest.importorskip('torchvision')
    from moseq2_detectron_extract_b200.model.predict import Predictor
    plain = Predictor.from_random_init(amp=amp, fused_convs=False).model.model
    fused = Predictor.from_random_init(amp=amp, fused_convs=True).model.model
    assert any(type(m).__name__ == 'FusedConvReLU' for m in fused.modules())
    g = torch.Generator(device='cuda').manual_seed(1)
    x = torch.rand((4, 3, 256, 256), device='cuda', generator=g)
    if amp:
        x = x.contiguous(memory_format=torch.channels_last)
    tol = 3e-2 if amp else 5e-3                                   # bf16 rounding of different fusion orders / TF32
    with torch.no_grad(), torch.autocast('cuda', dtype=torch.bfloat16, enabled=amp):
        want, got = plain.backbone(x), fused.backbone(x)
        for k in want:
            scale = float(want[k].float().abs().max())
            assert float((want[k].float() - got[k].float()).abs().max()) <= tol * scale, k
        feat = want['0'].float()
        pooled = torch.rand((6, 256, 14, 14), device='cuda', generator=g)
        for a, b, inp in ((plain.roi_heads.keypoint_head, fused.roi_heads.keypoint_head, pooled),
                          (plain.roi_heads.mask_head, fused.roi_heads.mask_head, pooled),
                          (plain.rpn.head.conv, fused.rpn.head.conv, feat)):
            ya, yb = a(inp).float(), b(inp).float()
            assert float((ya - yb).abs().max()) <= tol * float(ya.abs().max())


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('pooled,sampling', [(7, 2), (14, 2), (5, 3)])
def test_multilevel_roi_align_matches_torchvision(dtype, pooled, sampling):
    """msq_roi_align_levels (one launch, channels-last features) against torchvision's per-level roi_align + index_put,
    RoIs on every pyramid level, partly or wholly outside the image, degenerate and sub-pixel ones included."""
    pytest.importorskip('torchvision')
    from torchvision.ops import poolers as tv_poolers
    from moseq2_detectron_extract_b200.model import batched_heads as bh
    g = torch.Generator(device='cuda').manual_seed(pooled)
    n_img, ch, size = 3, 64, 256
    feats = [torch.randn((n_img, ch, size // s, size // s), device='cuda', generator=g).to(dtype).contiguous(memory_format=torch.channels_last)
             for s in (4, 8, 16, 32)]
    boxes = []
    for i in range(n_img):
        c = torch.rand((40, 2), device='cuda', generator=g) * size
        wh = torch.exp(torch.rand((40, 2), device='cuda', generator=g) * 6.5 - 1.0)          # 0.4 .. 245 px: all four levels
        b = torch.cat([c - wh / 2, c + wh / 2], dim=1)
        b[0] = torch.tensor([-30., -20., 10., 15.])                                             # sticks out of the image
        b[1] = torch.tensor([300., 300., 340., 330.])                                           # wholly outside
        b[2] = torch.tensor([50., 60., 50., 60.])                                               # zero size
        b[3] = torch.tensor([0., 0., float(size), float(size)])                                 # the whole image
        b[4] = torch.tensor([-200., -200., 456., 456.])                                         # coarsest level, mostly outside
        boxes.append(b)
    scales = [1 / 4, 1 / 8, 1 / 16, 1 / 32]
    mapper = tv_poolers.LevelMapper(2, 5)
    original = getattr(bh._multiscale_roi_align, 'original', None) or tv_poolers._multiscale_roi_align
    if original is bh._multiscale_roi_align:
        pytest.skip('torchvision pooler already replaced and the original is unknown')
    bh._multiscale_roi_align.original = original
    with torch.no_grad():
        want = original([f.float() for f in feats], boxes, (pooled, pooled), sampling, scales, mapper)
        got = bh._multiscale_roi_align(feats, boxes, (pooled, pooled), sampling, scales, mapper)
        one = bh._multiscale_roi_align(feats[:1], boxes, (pooled, pooled), sampling, scales[:1], None)
        want_one = original([feats[0].float()], boxes, (pooled, pooled), sampling, scales[:1], mapper)
    assert got.dtype == dtype and got.shape == want.shape == (n_img * 40, ch, pooled, pooled)
    assert len(set(mapper(boxes).tolist())) == 4
    for a, b in ((got, want), (one, want_one)):
        # torchvision's kernel is compiled with fused multiply-adds, ours without: sample positions differ by an ulp of
        # float32 (~1e-5 px), which moves the bilinear weights of white-noise features by as much
        scale = max(1.0, float(b.abs().max()))
        if dtype == torch.float32:
            diff = float((a - b).abs().max())
            assert diff <= 2e-4 * scale, diff
            assert float(((a - b).abs() <= 2e-6 * scale).float().mean()) > 0.99
        else:                                                        # same float32 arithmetic, then one rounding to bf16
            ref = b.to(torch.bfloat16)
            same = float((a == ref).float().mean())
            assert same > 0.99, same
            diff = float((a.float() - ref.float()).abs().max())
            assert diff <= 2 ** -7 * scale, diff


@pytest.mark.parametrize('amp', [False, True])
@pytest.mark.parametrize('h,w,vmax', [(240, 240, 100), (250, 250, 100), (200, 236, 80)])
def test_fused_detector_input_matches_torch_transform(amp, h, w, vmax):
    """msq_detector_input (scale + 3 channels + normalise + bilinear resize + pad, one kernel, channels-last) against
    msq_scale_frames_chw3_f32 followed by the detector's own GeneralizedRCNNTransform."""
    pytest.importorskip('torchvision')
    from moseq2_detectron_extract_b200 import _dev, _lib
    from moseq2_detectron_extract_b200.model.predict import Predictor
    pred = Predictor.from_random_init(amp=amp)
    g = torch.Generator(device='cuda').manual_seed(h + w)
    chunk = torch.randint(0, 140, (5, h, w), dtype=torch.uint8, device='cuda', generator=g)
    chw = torch.empty((5, 3, h, w), dtype=torch.float32, device='cuda')
    _lib.call('msq_scale_frames_chw3_f32', _dev.ptr(chunk), _dev.ptr(chw), 5, h, w, 0.0, float(vmax), 1, _dev.stream())
    with torch.no_grad():
        images, _ = pred.model.model.transform(list(chw.unbind(0)))
    got, size = pred.detector_input(chunk, 0, vmax)
    assert tuple(got.shape) == tuple(images.tensors.shape) and [tuple(s) for s in images.image_sizes] == [size] * 5
    assert got.is_contiguous(memory_format=torch.channels_last) and got.dtype == (torch.bfloat16 if amp else torch.float32)
    want = images.tensors
    if amp:
        ref = want.to(torch.bfloat16)
        assert float((got == ref).float().mean()) > 0.995              # a float32 ulp before the rounding to bf16
        assert float((got.float() - ref.float()).abs().max()) <= 2 ** -7 * float(want.abs().max())
    else:
        # torch's kernel forms the source coordinate with a fused multiply-add: an ulp of a coordinate near 240 moves the
        # interpolation weight by ~1e-5, times the local contrast of white noise
        diff = (got - want).abs()
        assert float(diff.max()) <= 2e-4 * float(want.abs().max()), float(diff.max())
        assert float(diff.mean()) <= 2e-6 * float(want.abs().max()), float(diff.mean())
    assert float(got[:, :, size[0]:, :].abs().max() if got.shape[2] > size[0] else 0.0) == 0.0      # padding is zero


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

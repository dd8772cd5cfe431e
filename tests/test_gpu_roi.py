"""GPU: session ROI detection (get_roi / plane_ransac, ref proc/roi.py:14-212) against oracle/roi_oracle.py, the committed
reference outputs (tests/golden/session_roi.npz) and SciPy / OpenCV directly.  Everything but the plane's mean-distance
tie-break is integer work: labels, region features, ranks, masks and boxes must be identical."""
import os

import cv2
import numpy as np
import pytest
import scipy.ndimage as ndi
import torch

import make_golden
import roi_oracle

pytestmark = pytest.mark.gpu


def _proc():
    from moseq2_detectron_extract_b200 import proc
    return proc


def _same_result(got, want, n_masks=None):
    rois, plane, bboxes, label_im, ranks, shape_index = got
    o_rois, o_plane, o_bboxes, o_label, o_ranks, o_index = want
    np.testing.assert_allclose(plane, o_plane, rtol=0, atol=1e-12)
    assert np.array_equal(np.asarray(label_im), o_label)
    assert np.array_equal(ranks, o_ranks) and np.array_equal(shape_index, o_index)
    assert len(rois) == len(o_rois) and len(bboxes) == len(o_bboxes)
    for a, b in zip(bboxes, o_bboxes):
        assert (a is None and b is None) or np.array_equal(a, b)
    for a, b in list(zip(rois, o_rois))[:n_masks]:
        assert a.dtype == bool and np.array_equal(a, b)


@pytest.mark.parametrize('name', list(make_golden.ROI_CASES))
def test_get_roi_matches_oracle_and_reference_golden(name):
    bg_kw, seed, kw = make_golden.ROI_CASES[name]
    bg = roi_oracle.synthetic_bground(**bg_kw)
    np.random.seed(seed)
    want = roi_oracle.get_roi(bg, **kw(cv2))
    np.random.seed(seed)
    got = _proc().get_roi(bg, progress_bar=False, **kw(cv2))
    _same_result(got, want)
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'session_roi.npz'))
    assert np.array_equal(got[1], g[name + '/plane'])              # same candidate wins, same float64 formula
    assert np.array_equal(got[3], g[name + '/label_im'])
    assert np.array_equal(np.stack(got[2]), g[name + '/bboxes'])
    ref_masks = np.unpackbits(g[name + '/rois'], axis=-1)[..., :bg.shape[1]].astype(bool)
    for i in range(make_golden.ROI_KEEP):
        assert np.array_equal(got[0][i], ref_masks[i])


@pytest.mark.parametrize('h,w,seed', [(576, 640, 2), (97, 130, 4), (424, 512, 7), (200, 1024, 8)])
def test_get_roi_other_sizes_and_options(h, w, seed):
    bg = roi_oracle.synthetic_bground(h=h, w=w, seed=seed, speckles=25)
    kw = dict(noise_tolerance=25, iters=200, strel_dilate=cv2.getStructuringElement(cv2.MORPH_RECT, (11, 5)),
              strel_erode=cv2.getStructuringElement(cv2.MORPH_CROSS, (3, 7)) if seed % 2 else None)
    overlap = np.zeros((h, w), bool)
    overlap[: h // 6, : w // 3] = True                              # drops the region that covers the ledge corner most
    for extra in (dict(), dict(overlap_roi=overlap)):
        np.random.seed(seed)
        want = roi_oracle.get_roi(bg, **kw, **extra)
        np.random.seed(seed)
        got = _proc().get_roi(bg, **kw, **extra)
        _same_result(got, want)


def test_get_roi_device_tensor_in_device_tensors_out():
    bg = roi_oracle.synthetic_bground(seed=5)
    np.random.seed(1)
    want = roi_oracle.get_roi(bg)
    np.random.seed(1)
    rois, plane, bboxes, label_im, ranks, shape_index = _proc().get_roi(torch.from_numpy(bg).cuda())
    assert rois[0].is_cuda and rois[0].dtype == torch.bool and label_im.is_cuda
    _same_result(([r.cpu().numpy() for r in rois], plane, bboxes, label_im.cpu().numpy(), ranks, shape_index), want)


def test_plane_ransac_matches_oracle_and_plane_fit3():
    proc = _proc()
    bg = roi_oracle.synthetic_bground(seed=9)
    mask = np.ones(bg.shape, bool)
    mask[:, :200] = False
    for m in (None, mask):
        np.random.seed(2)
        o_plane, o_dist = roi_oracle.plane_ransac(bg, iters=150, mask=m)
        np.random.seed(2)
        plane, dist = proc.plane_ransac(bg, iters=150, mask=m, progress_bar=False)
        assert np.array_equal(plane, o_plane)
        assert dist.shape == (bg.size,) and np.array_equal(dist, o_dist)
    pts = np.array([[0., 0., 5.], [10., 0., 5.], [0., 10., 7.]])
    assert np.array_equal(proc.plane_fit3(pts), roi_oracle.plane_from_points(pts))
    assert np.isnan(proc.plane_fit3(np.array([[0., 0., 0.], [1., 1., 1.], [2., 2., 2.]]))).all()
    with pytest.raises(RuntimeError):
        np.random.seed(0)
        proc.plane_ransac(bg, depth_range=(400, 900), noise_tolerance=0.001, iters=20)      # nothing reaches in_ratio


def _label_regions(binary):
    from moseq2_detectron_extract_b200 import _dev, _lib
    h, w = binary.shape
    b = torch.from_numpy(binary.astype(np.uint8)).cuda()
    labels = torch.empty((h, w), dtype=torch.int32, device='cuda')
    count = torch.empty((1,), dtype=torch.int32, device='cuda')
    scratch = torch.empty((int(_lib.load().msq_label_scratch_bytes(h, w)) + 16,), dtype=torch.uint8, device='cuda')
    _lib.call('msq_label_regions', _dev.ptr(b), h, w, _dev.ptr(labels), _dev.ptr(count), _dev.ptr(scratch), scratch.numel(), _dev.stream())
    return labels, int(count.item())


@pytest.mark.parametrize('h,w,density', [(424, 512, 0.5), (424, 512, 0.35), (61, 1000, 0.6), (300, 33, 0.45), (1, 70, 0.5), (50, 1, 0.5),
                                          (64, 64, 1.0), (64, 64, 0.0)])
def test_labels_and_region_props_equal_scipy(h, w, density):
    """Random noise near the 8-connectivity percolation threshold: thousands of regions and long winding ones."""
    from moseq2_detectron_extract_b200 import _dev, _lib
    rng = np.random.default_rng(h * 1000 + w)
    binary = rng.random((h, w)) < density
    labels, n = _label_regions(binary)
    want, want_n = ndi.label(binary, structure=np.ones((3, 3), int))
    assert n == want_n and np.array_equal(labels.cpu().numpy(), want)
    if n == 0:
        return
    area = torch.empty((n,), dtype=torch.int32, device='cuda')
    bbox = torch.empty((n, 4), dtype=torch.int32, device='cuda')
    maxd4 = torch.empty((n,), dtype=torch.int32, device='cuda')
    _lib.call('msq_region_props', _dev.ptr(labels), h, w, n, _dev.ptr(area), _dev.ptr(bbox), _dev.ptr(maxd4), _dev.stream())
    assert np.array_equal(area.cpu().numpy(), np.bincount(want.ravel(), minlength=n + 1)[1:])
    rows, cols = np.nonzero(want)
    lab = want[rows, cols] - 1
    d4 = (2 * rows - h) ** 2 + (2 * cols - w) ** 2
    exp_d4 = np.zeros(n, np.int64)
    np.maximum.at(exp_d4, lab, d4)
    assert np.array_equal(maxd4.cpu().numpy(), exp_d4)
    exp_box = np.array([[sl[0].start, sl[1].start, sl[0].stop - 1, sl[1].stop - 1] for sl in ndi.find_objects(want)])
    assert np.array_equal(bbox.cpu().numpy(), exp_box)


@pytest.mark.parametrize('h,w', [(120, 150), (424, 512), (90, 1024), (40, 31)])
def test_region_masks_equal_opencv_and_scipy_for_random_elements(h, w):
    from moseq2_detectron_extract_b200 import _dev, _lib
    rng = np.random.default_rng(w)
    binary = ndi.binary_opening(rng.random((h, w)) < 0.62, iterations=1)             # blobs with holes, some touching the border
    labels, n = _label_regions(binary)
    want_labels = labels.cpu().numpy()
    assert n > 3
    order = rng.permutation(n)[: min(n, 12)].astype(np.int32)
    for trial in range(4):
        se_d = (rng.random((int(rng.integers(1, 9)), int(rng.integers(1, 12)))) < 0.6).astype(np.uint8) if trial != 1 else None
        se_e = (rng.random((int(rng.integers(1, 5)), int(rng.integers(1, 5)))) < 0.7).astype(np.uint8) if trial >= 2 else None
        if se_d is not None and not se_d.any():
            se_d[0, 0] = 1
        if se_e is not None and not se_e.any():
            se_e[-1, -1] = 1
        fill = trial != 3
        masks = torch.empty((len(order), h, w), dtype=torch.uint8, device='cuda')
        boxes = torch.empty((len(order), 4), dtype=torch.int32, device='cuda')
        d_dev = None if se_d is None else torch.from_numpy(se_d).cuda()
        e_dev = None if se_e is None else torch.from_numpy(se_e).cuda()
        _lib.call('msq_region_rois', _dev.ptr(labels), h, w, _dev.ptr(torch.from_numpy(order).cuda()), len(order), _dev.ptr(d_dev),
                  0 if se_d is None else se_d.shape[0], 0 if se_d is None else se_d.shape[1], _dev.ptr(e_dev),
                  0 if se_e is None else se_e.shape[0], 0 if se_e is None else se_e.shape[1], int(fill), _dev.ptr(masks), _dev.ptr(boxes),
                  _dev.stream())
        masks, boxes = masks.cpu().numpy(), boxes.cpu().numpy()
        for i, r in enumerate(order):
            roi = (want_labels == r + 1).astype(np.float32)
            if se_d is not None:
                roi = cv2.dilate(roi, se_d, iterations=1)
            if se_e is not None:
                roi = cv2.erode(roi, se_e, iterations=1)
            roi = ndi.binary_fill_holes(roi) if fill else roi > 0
            assert np.array_equal(masks[i].astype(bool), roi), (trial, i)
            rows, cols = np.nonzero(roi)
            exp = [-1] * 4 if len(rows) == 0 else [rows.min(), cols.min(), rows.max(), cols.max()]
            assert boxes[i].tolist() == exp


@pytest.mark.parametrize('ksize,threshold', [(7, 3000), (5, 400), (3, 60), (1, 8)])
@pytest.mark.parametrize('dtype', ['float32', 'float64', 'uint16'])
def test_gradient_mask_equals_opencv_sobel(ksize, threshold, dtype):
    from moseq2_detectron_extract_b200.proc.roi import _depth_f64, _gradient_mask, sobel_kernels
    deriv, smooth = sobel_kernels(ksize)
    kx, ky = cv2.getDerivKernels(1, 0, ksize, normalize=False, ktype=cv2.CV_64F)
    assert np.array_equal(deriv, kx.ravel()) and np.array_equal(smooth, ky.ravel())
    for h, w in ((200, 260), (33, 17), (424, 512)):
        bg = roi_oracle.synthetic_bground(h=h, w=w, seed=h).astype(dtype)
        if dtype == 'float64':
            bg[::3, ::2] += 0.5                                   # np.median of an even frame count gives half steps
        got = _gradient_mask(_depth_f64(bg), ksize, threshold).cpu().numpy()
        assert np.array_equal(got, roi_oracle.gradient_mask(bg, ksize, threshold))
        assert 0 < got.sum() < got.size

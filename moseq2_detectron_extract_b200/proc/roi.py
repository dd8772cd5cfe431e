"""ROI helpers on the hot path (mirrors reference proc/roi.py:215-254) and the per-session setup: background image
(proc/roi.py:293-307) and ROI detection (proc/roi.py:14-212)."""
import ctypes
from typing import Optional, Union

import numpy as np
import torch

from .. import _dev


def get_bbox(roi) -> Union[np.ndarray, None]:
    """((y_min, x_min), (y_max, x_max)) of the non-zero region or None (ref: proc/roi.py:239-254).
    Callers slice `[y_min:y_max, x_min:x_max]`, i.e. the maxima are treated as exclusive."""
    if isinstance(roi, torch.Tensor):
        roi = roi.detach().cpu().numpy()
    rows = np.flatnonzero((np.asarray(roi) > 0).any(axis=1))
    cols = np.flatnonzero((np.asarray(roi) > 0).any(axis=0))
    if rows.size == 0 or cols.size == 0:
        return None
    return np.array([[rows[0], cols[0]], [rows[-1], cols[-1]]])


def apply_roi(frames, roi):
    """Mask `frames` by `roi` and crop to its bounding box (ref: proc/roi.py:215-236).

    On the extract path this never runs on its own: it is fused into the prep kernel
    (`prep_raw_frames`).  The standalone form is two elementwise device ops."""
    dev = _dev.as_device(frames)
    mask = _dev.as_device(np.asarray(roi.detach().cpu() if isinstance(roi, torch.Tensor) else roi) > 0)
    if dev.dim() == 3:
        dev = dev * mask
    box = get_bbox(roi)
    if box is not None:
        dev = dev[:, int(box[0, 0]):int(box[1, 0]), int(box[0, 1]):int(box[1, 1])]
    return _dev.give_back(dev.contiguous(), frames)


def get_bground_im(frames, med_scale: int = 5):
    """Background image: per-frame `med_scale` x `med_scale` median blur, then the per-pixel median over frames
    (ref: proc/roi.py:293-307; io/session.py:217-218 passes every 500th frame of the session).

    `frames`: (n, H, W) int16 / uint16, numpy or CUDA tensor.  Returns (H, W) float64, bit-identical to
    cv2.medianBlur + np.median.  Unlike the reference the input frames are not overwritten with their blurred
    versions.  16-bit frames allow med_scale 3 or 5 only (the same restriction cv2.medianBlur has)."""
    from .. import _lib
    unsigned = isinstance(frames, np.ndarray) and frames.dtype == np.uint16
    if isinstance(frames, np.ndarray) and frames.dtype not in (np.int16, np.uint16):
        raise TypeError(f'get_bground_im: frames must be int16 or uint16, got {frames.dtype}')
    if isinstance(frames, torch.Tensor) and frames.dtype != torch.int16:
        raise TypeError(f'get_bground_im: tensor frames must be int16, got {frames.dtype}')
    dev = _dev.as_device(frames, torch.int16)
    if dev.dim() != 3:
        raise ValueError(f'get_bground_im: frames must be (nframes, rows, cols), got {tuple(dev.shape)}')
    n, h, w = (int(v) for v in dev.shape)
    out = _dev.empty((h, w), torch.float64)
    scratch = _dev.empty((int(_lib.load().msq_bground_scratch_bytes(n, h, w)) + 16,), torch.uint8)
    _lib.call('msq_get_bground_im', _dev.ptr(dev), n, h, w, int(med_scale), int(unsigned), _dev.ptr(out), _dev.ptr(scratch),
              scratch.numel(), _dev.stream())
    return _dev.give_back(out, frames)


def plane_fit3(points) -> np.ndarray:
    """Plane through 3 points as (a, b, c, d) with unit normal, `a*x + b*y + c*z + d = 0`; all-NaN when the points are
    collinear (ref: proc/roi.py:106-130).  Host helper; the RANSAC kernel evaluates the same formula per candidate."""
    pts = np.asarray(points, dtype=np.float64)
    normal = np.cross(pts[1] - pts[0], pts[2] - pts[0])
    denom = float((normal[0] * normal[0] + normal[1] * normal[1]) + normal[2] * normal[2])
    if denom < np.spacing(1):
        return np.full((4,), np.nan)
    normal = normal / np.sqrt(denom)
    return np.append(normal, ((-pts[0, 0]) * normal[0] + (-pts[0, 1]) * normal[1]) + (-pts[0, 2]) * normal[2])


def _depth_f64(depth_image) -> torch.Tensor:
    if isinstance(depth_image, torch.Tensor) and depth_image.dtype == torch.bool:
        raise TypeError('depth image must be numeric')
    dev = _dev.as_device(depth_image, torch.float64)
    if dev.dim() != 2:
        raise ValueError(f'depth image must be (rows, cols), got {tuple(dev.shape)}')
    return dev


def _plane_distance(depth: torch.Tensor, plane: np.ndarray, tol: float, valid: Optional[torch.Tensor], want_dist: bool, want_bin: bool):
    from .. import _lib
    h, w = (int(v) for v in depth.shape)
    dist = _dev.empty((h * w,), torch.float64) if want_dist else None
    on_plane = _dev.empty((h, w), torch.uint8) if want_bin else None
    host_plane = (ctypes.c_double * 4)(*[float(v) for v in plane])
    _lib.call('msq_plane_distance', _dev.ptr(depth), h, w, host_plane, float(tol), _dev.ptr(valid), _dev.ptr(dist), _dev.ptr(on_plane),
              _dev.stream())
    return dist, on_plane


def _ransac_triples(npoints: int, iters: int) -> np.ndarray:
    """(iters, 3) indices of the candidate triples, from NumPy's global generator.  The reference draws
    `np.random.choice(npoints, 3, replace=True)` once per iteration (proc/roi.py:170); with replacement that is
    `randint(0, npoints, 3)`, and the legacy generator produces the same numbers (and ends in the same state) when all
    iterations are drawn in one call -- so np.random.seed(...) still reproduces the reference's sequence of candidate planes
    (tests/test_host_logic.py checks the equivalence), at 1/150 of the cost of the loop."""
    if iters <= 0:
        return np.zeros((0, 3), np.int64)
    return np.random.randint(0, npoints, size=(iters, 3))


def _plane_ransac_device(depth: torch.Tensor, depth_range, iters: int, noise_tolerance: float, in_ratio: float,
                         mask: Optional[torch.Tensor]) -> np.ndarray:
    from .. import _lib
    h, w = (int(v) for v in depth.shape)
    use = (depth > depth_range[0]) & (depth < depth_range[1])
    if mask is not None:
        use &= mask
    idx = use.flatten().nonzero().squeeze(1).to(torch.int32)
    npoints = int(idx.numel())
    if npoints == 0:
        raise ValueError(f'plane_ransac: no pixel of the image lies inside depth_range={tuple(depth_range)}')
    sel = _ransac_triples(npoints, int(iters))
    sel_dev = _dev.as_device(sel.astype(np.int64))
    planes = _dev.empty((int(iters), 4), torch.float64)
    ninl = _dev.empty((int(iters),), torch.int32)
    sumd = _dev.empty((int(iters),), torch.float64)
    _lib.call('msq_plane_ransac_score', _dev.ptr(idx), npoints, _dev.ptr(depth), h, w, _dev.ptr(sel_dev), int(iters), float(noise_tolerance),
              _dev.ptr(planes), _dev.ptr(ninl), _dev.ptr(sumd), _dev.stream())
    planes, ninl, sumd = planes.cpu().numpy(), ninl.cpu().numpy(), sumd.cpu().numpy()
    best_plane, best_dist, best_num = None, np.inf, 0
    for i in range(int(iters)):                           # the running-best rule is order dependent (proc/roi.py:181-186)
        if np.isnan(planes[i, 0]):
            continue
        mean_dist = sumd[i] / npoints
        if ninl[i] / npoints > in_ratio and ninl[i] > best_num and mean_dist < best_dist:
            best_plane, best_dist, best_num = planes[i].copy(), mean_dist, int(ninl[i])
    if best_plane is None:
        raise RuntimeError(f'plane_ransac: none of the {iters} candidate planes had more than {in_ratio:.0%} of the '
                           f'{npoints} pixels in depth_range={tuple(depth_range)} within {noise_tolerance} of it')
    return best_plane


def plane_ransac(depth_image, depth_range=(650, 750), iters=1000, noise_tolerance=30, in_ratio=0.1, progress_bar=True, mask=None):
    """RANSAC plane fit of a background image (ref: proc/roi.py:133-212).  Returns `(best_plane (4,), dist (rows*cols,))`,
    `dist` being every pixel's distance to the plane (float64; numpy in -> numpy out, CUDA tensor in -> tensor out).

    All `iters` candidates are scored in one kernel launch (inlier count and mean distance over the pixels inside
    `depth_range`), then the reference's running-best rule picks the plane.  Candidate triples are drawn from
    `np.random` exactly like the reference does, so a seeded call fits the same plane.  Where the reference dies with an
    UnboundLocalError because no candidate qualifies, this raises RuntimeError.  `progress_bar` is accepted and ignored."""
    depth = _depth_f64(depth_image)
    valid = None if mask is None else _dev.as_device(mask).to(torch.bool)
    plane = _plane_ransac_device(depth, depth_range, iters, noise_tolerance, in_ratio, valid)
    dist, _ = _plane_distance(depth, plane, noise_tolerance, None, True, False)
    return plane, _dev.give_back(dist, depth_image)


def sobel_kernels(ksize: int):
    """(derivative, smoothing) taps of cv2.getDerivKernels(1, 0, ksize, normalize=False): binomial smoothing of ksize
    taps, and the first difference of the (ksize - 1)-tap binomial; ksize 1 is OpenCV's [-1, 0, 1] x [1]."""
    ksize = int(ksize)
    if ksize < 1 or ksize > 31 or ksize % 2 == 0:
        raise ValueError(f'Sobel kernel size must be odd and in 1..31, got {ksize}')
    if ksize == 1:
        return np.array([-1.0, 0.0, 1.0]), np.array([1.0])
    smooth, deriv = np.ones(1), np.ones(1)
    for _ in range(ksize - 1):
        smooth = np.convolve(smooth, [1.0, 1.0])
    for _ in range(ksize - 2):
        deriv = np.convolve(deriv, [1.0, 1.0])
    return np.convolve(deriv, [1.0, -1.0])[::-1].copy(), smooth


def _gradient_mask(depth: torch.Tensor, ksize: int, threshold: float) -> torch.Tensor:
    from .. import _lib
    h, w = (int(v) for v in depth.shape)
    deriv, smooth = sobel_kernels(ksize)
    mask = _dev.empty((h, w), torch.uint8)
    _lib.call('msq_sobel_gradient_mask', _dev.ptr(depth), h, w, (ctypes.c_double * len(deriv))(*deriv), len(deriv),
              (ctypes.c_double * len(smooth))(*smooth), len(smooth), float(threshold), _dev.ptr(mask), _dev.stream())
    return mask.view(torch.bool)


def _rank_max(values: np.ndarray) -> np.ndarray:
    """scipy.stats.rankdata(values, method='max'): the number of elements <= each value."""
    return np.searchsorted(np.sort(values), values, side='right').astype(np.float64)


_DEFAULT_STREL_DILATE = np.ones((15, 15), np.uint8)       # cv2.getStructuringElement(cv2.MORPH_RECT, (15, 15))


def get_roi(depth_image, strel_dilate=_DEFAULT_STREL_DILATE, strel_erode=None, noise_tolerance=30, weights=(1, .1, 1), overlap_roi=None,
            gradient_filter=False, gradient_kernel=7, gradient_threshold=3000, fill_holes=True, **kwargs):
    """Find the arena floor in a background image (ref: proc/roi.py:14-103; called once per session, io/session.py:234).

    RANSAC plane fit, 8-connected regions of the pixels within `noise_tolerance` of the plane, ranked by area, extent and
    farthest distance from the image centre (weights); every region is dilated / eroded by the given structuring
    elements (any uint8 array as cv2.getStructuringElement returns, or None) and its holes filled.

    Returns the reference's tuple `(rois, roi_plane, bboxes, label_im, ranks, shape_index)`: `rois` / `bboxes` are lists in
    ranked order (best first), masks are bool whatever `fill_holes` is (the reference leaves 0/1 images of the depth dtype
    when `fill_holes=False`).  numpy in -> numpy out; CUDA tensor in -> `rois` and `label_im` stay on the device.
    Labels, region features, ranks and masks are bit-identical to skimage / OpenCV / SciPy for the same plane.
    `gradient_filter=True` masks out pixels whose |cv2.Sobel| response (kernel `gradient_kernel`) reaches
    `gradient_threshold` in x or y before the fit, like the reference."""
    from .. import _lib
    kwargs.pop('progress_bar', None)
    depth = _depth_f64(depth_image)
    h, w = (int(v) for v in depth.shape)
    # gradient_filter: pixels with a steep Sobel response (walls, rims) neither vote for the plane nor belong to it
    valid = _gradient_mask(depth, gradient_kernel, gradient_threshold) if gradient_filter else None
    roi_plane = _plane_ransac_device(depth, kwargs.pop('depth_range', (650, 750)), kwargs.pop('iters', 1000), noise_tolerance,
                                     kwargs.pop('in_ratio', 0.1), valid)
    if kwargs:
        raise TypeError(f'get_roi: unexpected arguments {sorted(kwargs)}')
    _, on_plane = _plane_distance(depth, roi_plane, noise_tolerance, None if valid is None else valid.view(torch.uint8), False, True)

    labels = _dev.empty((h, w), torch.int32)
    count = _dev.empty((1,), torch.int32)
    scratch = _dev.empty((int(_lib.load().msq_label_scratch_bytes(h, w)) + 16,), torch.uint8)
    _lib.call('msq_label_regions', _dev.ptr(on_plane), h, w, _dev.ptr(labels), _dev.ptr(count), _dev.ptr(scratch), scratch.numel(), _dev.stream())
    n_regions = int(count.item())
    area = _dev.empty((n_regions,), torch.int32)
    bbox = _dev.empty((n_regions, 4), torch.int32)
    maxd4 = _dev.empty((n_regions,), torch.int32)
    _lib.call('msq_region_props', _dev.ptr(labels), h, w, n_regions, _dev.ptr(area), _dev.ptr(bbox), _dev.ptr(maxd4), _dev.stream())

    # ranking of a handful of regions: host arithmetic, same expressions as the reference (proc/roi.py:51-72)
    areas = area.cpu().numpy().astype(np.float64)
    box = bbox.cpu().numpy().astype(np.float64)
    extents = areas / ((box[:, 2] - box[:, 0] + 1) * (box[:, 3] - box[:, 1] + 1)) if n_regions else np.zeros((0,))
    dists = np.sqrt(maxd4.cpu().numpy().astype(np.float64) / 4.0)
    ranks = np.vstack((_rank_max(-areas), _rank_max(-extents), _rank_max(dists)))
    weight_array = np.array(weights, 'float32')
    shape_index = np.mean(np.multiply(ranks.astype('float32'), weight_array[:, np.newaxis]), 0).argsort()

    def element(strel):
        if strel is None:
            return None, 0, 0
        arr = np.ascontiguousarray(np.asarray(strel) != 0, dtype=np.uint8)
        if arr.ndim != 2:
            raise ValueError('structuring elements must be 2-D')
        return _dev.as_device(arr), int(arr.shape[0]), int(arr.shape[1])
    se_d, dh, dw = element(strel_dilate)
    se_e, eh, ew = element(strel_erode)
    masks = _dev.empty((n_regions, h, w), torch.uint8)
    boxes = _dev.empty((n_regions, 4), torch.int32)
    order = _dev.as_device(shape_index.astype(np.int32))
    _lib.call('msq_region_rois', _dev.ptr(labels), h, w, _dev.ptr(order), n_regions, _dev.ptr(se_d), dh, dw, _dev.ptr(se_e), eh, ew,
              int(bool(fill_holes)), _dev.ptr(masks), _dev.ptr(boxes), _dev.stream())
    masks = masks.view(torch.bool) if n_regions else masks.to(torch.bool)
    boxes = boxes.cpu().numpy().astype(np.int64)
    keep = list(range(n_regions))
    if overlap_roi is not None and n_regions:
        other = _dev.as_device(overlap_roi).to(torch.bool)
        overlaps = (masks & other).flatten(1).sum(1).cpu().numpy()
        del keep[int(np.argmax(overlaps))]
    on_host = not isinstance(depth_image, torch.Tensor)
    host_masks = masks.cpu().numpy() if on_host else None
    rois = [host_masks[i] if on_host else masks[i] for i in keep]
    bboxes = [None if boxes[i, 0] < 0 else boxes[i].reshape(2, 2) for i in keep]
    label_im = labels.to(torch.int64)
    return rois, roi_plane, bboxes, _dev.give_back(label_im, depth_image), ranks, shape_index

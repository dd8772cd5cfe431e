"""The `proc` operator surface of the extract hot path, served by libmoseq_b200 (sm_100a kernels).

Same function names, argument meaning and error behaviour as the reference's
`moseq2_detectron_extract/proc/proc.py`; each docstring cites the reference lines it replaces.
Inputs may be numpy arrays (copied to the GPU, results come back as numpy) or CUDA torch tensors
(results stay on the device).  There is no CPU implementation: without the built CUDA library and a
CUDA device every function raises.
"""
from __future__ import annotations

import ctypes
import warnings
from typing import List, Optional, Tuple

import numpy as np
import torch

from .. import _dev, _lib
from .roi import get_bbox

_ELLIPSE9 = np.array([[0, 0, 0, 0, 1, 0, 0, 0, 0],
                      [0, 1, 1, 1, 1, 1, 1, 1, 0],
                      [0, 1, 1, 1, 1, 1, 1, 1, 0],
                      [1, 1, 1, 1, 1, 1, 1, 1, 1],
                      [1, 1, 1, 1, 1, 1, 1, 1, 1],
                      [1, 1, 1, 1, 1, 1, 1, 1, 1],
                      [0, 1, 1, 1, 1, 1, 1, 1, 0],
                      [0, 1, 1, 1, 1, 1, 1, 1, 0],
                      [0, 0, 0, 0, 1, 0, 0, 0, 0]], dtype=np.uint8)   # cv2 MORPH_ELLIPSE (9, 9)


class InvalidPixelsError(NotImplementedError):
    """Kept for API compatibility with earlier builds: invalid pixels are now in-painted on the GPU
    (csrc/inpaint.cu), so `prep_raw_frames` no longer raises this."""


# --------------------------------------------------------------------------------------------
# a2  prep_raw_frames
# --------------------------------------------------------------------------------------------
def _bg_code(bground) -> Tuple[int, Optional[torch.Tensor]]:
    if bground is None:
        return _lib.MSQ_BG_NONE, None
    if isinstance(bground, torch.Tensor):
        kind = {torch.float32: 'f4', torch.float64: 'f8', torch.int16: 'u2'}.get(bground.dtype)
        if kind is None:
            bground, kind = bground.to(torch.float64), 'f8'
        return {'f4': _lib.MSQ_BG_F32, 'f8': _lib.MSQ_BG_F64, 'u2': _lib.MSQ_BG_U16}[kind], _dev.as_device(bground)
    arr = np.asarray(bground)
    if arr.dtype == np.float32:
        return _lib.MSQ_BG_F32, _dev.as_device(arr)
    if arr.dtype == np.uint16:
        return _lib.MSQ_BG_U16, _dev.as_device(arr)             # bits travel as int16, the kernel reads uint16
    return _lib.MSQ_BG_F64, _dev.as_device(arr.astype(np.float64))


def _prep_device(frames, bground_im, roi, vmin, vmax, want_invalid: bool, want_bits: bool = False, positive_bits_out: Optional[list] = None):
    """Launch the fused prep kernel; returns (out_u8 (n,h,w), invalid_count (n) int32 or None) and, with
    want_bits, additionally the packed invalid-pixel mask (n, h, ceil(w/8)) uint8.  positive_bits_out: a list that receives
    the positive-pixel bit rows (n, h, ceil(w/32)) int32 the cleaning pass can use instead of re-reading the frames."""
    if isinstance(frames, torch.Tensor) and not frames.is_cuda and frames.dtype == torch.int16 and frames.is_pinned() \
            and frames.is_contiguous():
        _dev.require_cuda()
        raw = frames          # zero-copy: the kernel reads the ROI box straight from page-locked host memory (UVA)
    else:
        raw = _dev.as_device(frames, torch.int16)
    if raw.dim() != 3:
        raise ValueError(f'frames must be (nframes, height, width); got shape {tuple(raw.shape)}')
    n, H, W = (int(v) for v in raw.shape)
    code, bg = _bg_code(bground_im)
    if bg is not None and tuple(bg.shape) != (H, W):
        raise ValueError(f'bground_im shape {tuple(bg.shape)} does not match frames {(H, W)}')
    roi_dev, y0, x0, h, w = None, 0, 0, H, W
    if roi is not None:
        roi_np = roi.detach().cpu().numpy() if isinstance(roi, torch.Tensor) else np.asarray(roi)
        if roi_np.shape != (H, W):
            raise ValueError(f'roi shape {roi_np.shape} does not match frames {(H, W)}')
        roi_dev = _dev.as_device((roi_np > 0).astype(np.uint8))
        box = get_bbox(roi_np)
        if box is not None:      # max-exclusive slicing, ref: proc/roi.py:233-235
            y0, x0, h, w = int(box[0, 0]), int(box[0, 1]), int(box[1, 0] - box[0, 0]), int(box[1, 1] - box[0, 1])
    if h <= 0 or w <= 0:
        empty = _dev.empty((n, max(h, 0), max(w, 0)), torch.uint8)
        return (empty, None, None) if want_bits else (empty, None)
    flags = (_lib.MSQ_PREP_HAS_VMIN if vmin is not None else 0) | (_lib.MSQ_PREP_HAS_VMAX if vmax is not None else 0)
    out = _dev.empty((n, h, w), torch.uint8)
    invalid = _dev.empty((n,), torch.int32) if want_invalid else None
    bits = None
    if want_bits:
        row_bytes = (w + 7) // 8
        bits = torch.zeros(((n * h * row_bytes + 3) // 4 * 4,), dtype=torch.uint8, device='cuda')
    positive = None
    if positive_bits_out is not None:
        positive = _dev.positive_bits_like(out)
        positive_bits_out.append(positive)
    _lib.call('msq_prep_frames_bits', _dev.ptr(raw), n, H, W, _dev.ptr(bg), code, _dev.ptr(roi_dev), y0, x0, h, w,
              float(vmin if vmin is not None else 0.0), float(vmax if vmax is not None else 0.0), flags,
              _dev.ptr(out), _dev.ptr(invalid), _dev.ptr(bits), _dev.ptr(positive), _dev.stream())
    if want_bits:
        return out, invalid, bits
    return out, invalid


def fill_invalid_pixels_device(frames_u8: torch.Tensor, invalid_count: torch.Tensor, invalid_bits: torch.Tensor,
                               radius: int = 3) -> int:
    """In-paint, in place, every frame whose invalid count is non-zero (ref: proc/proc.py:189-210,
    cv2.inpaint(..., 3, cv2.INPAINT_NS) bit-exact; csrc/inpaint.cu).  Returns the number of frames touched."""
    flagged = torch.nonzero(invalid_count, as_tuple=False).flatten().to(torch.int32)
    m = int(flagged.numel())
    if m == 0:
        return 0
    n, h, w = (int(v) for v in frames_u8.shape)
    scratch = _dev.empty((int(_lib.load().msq_inpaint_scratch_bytes(m, h, w)) + 8,), torch.uint8)
    _lib.call('msq_inpaint_frames', _dev.ptr(frames_u8), _dev.ptr(invalid_bits), _dev.ptr(flagged), m, h, w, int(radius),
              _dev.ptr(scratch), scratch.numel(), _dev.stream())
    return m


def prep_raw_frames(frames, bground_im=None, roi=None, vmin: Optional[float] = None, vmax: Optional[float] = None,
                    dtype='uint8', fix_invalid_pixels: bool = True, positive_bits_out: Optional[list] = None):
    """Background-subtract, ROI-mask + crop, clamp and cast raw depth frames (ref: proc/proc.py:129-172).

    One fused kernel (csrc/prep.cu) replaces find_invalid_pixels + `bground - frames` + apply_roi +
    the two fancy-index clamps + astype.  Returns (nframes, roi_height, roi_width) uint8.

    positive_bits_out (not in the reference): a list; the kernel's second output, one bit per prepared pixel that is > 0, is
    appended to it for `ChunkEngine.extract(..., positive_bits=)` -- the cleaning pass then finds the rows the opening can leave
    non-zero from 1/8 of the bytes.  In-painting only rewrites pixels whose bit is set, so the rows stay a valid superset."""
    if np.dtype(dtype) != np.uint8:
        raise NotImplementedError('prep_raw_frames: only dtype=uint8 (the extract path) is implemented')
    if not fix_invalid_pixels:
        out, _ = _prep_device(frames, bground_im, roi, vmin, vmax, want_invalid=False, positive_bits_out=positive_bits_out)
        return _dev.give_back(out, frames)
    out, invalid, bits = _prep_device(frames, bground_im, roi, vmin, vmax, want_invalid=True, want_bits=True,
                                      positive_bits_out=positive_bits_out)
    if invalid is not None and out.numel() > 0:
        fill_invalid_pixels_device(out, invalid, bits)
    return _dev.give_back(out, frames)


def find_invalid_pixels(frames):
    """Mask of Kinect invalid pixels, ones where raw == 0 (ref: proc/proc.py:175-186)."""
    dev = _dev.as_device(frames)
    return _dev.give_back((dev == 0).to(torch.uint8), frames)


# --------------------------------------------------------------------------------------------
# a3  scale_raw_frames
# --------------------------------------------------------------------------------------------
def scale_raw_frames(frames, vmin: float, vmax: float, dtype='uint8'):
    """Linear intensity scale to the uint8 range (ref: proc/proc.py:214-234)."""
    if np.dtype(dtype) != np.uint8:
        raise NotImplementedError('scale_raw_frames: only dtype=uint8 is implemented')
    src = _dev.as_device(frames, torch.uint8)
    out = torch.empty_like(src)
    _lib.call('msq_scale_frames', _dev.ptr(src), _dev.ptr(out), src.numel(), float(vmin), float(vmax),
              int(isinstance(vmin, (int, np.integer)) and not isinstance(vmin, bool)), _dev.stream())
    return _dev.give_back(out, frames)


# --------------------------------------------------------------------------------------------
# a6  clean_frames
# --------------------------------------------------------------------------------------------
def clean_frames(frames, prefilter_space=(3,), prefilter_time=None, strel_tail=_ELLIPSE9, iters_tail=None,
                 frame_dtype='uint8', strel_min=None, iters_min=None, progress_bar=True):
    """3x3 median + one opening with the 9x9 ellipse (ref: proc/proc.py:480-515).

    Only the configuration the extract path uses (`instances_to_features` calls
    `clean_frames(raw, iters_tail=3)`, ref: proc/proc.py:715) is implemented; any other one raises.
    NOTE the reference passes iters_tail into cv2.morphologyEx's `dst` slot, so any positive
    iters_tail means exactly one opening (SURVEY.md trap 3) -- replicated here."""
    if tuple(prefilter_space or ()) != (3,) or prefilter_time is not None or (iters_min is not None and iters_min > 0):
        raise NotImplementedError('clean_frames: only prefilter_space=(3,), no temporal filter, no iters_min')
    if iters_tail is None or iters_tail <= 0:
        raise NotImplementedError('clean_frames: iters_tail must be positive (median + opening)')
    if not np.array_equal(np.asarray(strel_tail), _ELLIPSE9):
        raise NotImplementedError('clean_frames: only the 9x9 MORPH_ELLIPSE structuring element is implemented')
    src = _dev.as_device(frames, torch.uint8)
    n, h, w = (int(v) for v in src.shape)
    out = torch.empty_like(src)
    _dev.clean_frames_ws(src, out)
    return _dev.give_back(out, frames)


# --------------------------------------------------------------------------------------------
# a7  get_frame_features / im_moment_features
# --------------------------------------------------------------------------------------------
def _features_device(cleaned: torch.Tensor, mask: torch.Tensor, frame_threshold: float, want_sums: bool = False):
    n, h, w = (int(v) for v in cleaned.shape)
    centroid = _dev.empty((n, 2), torch.float64)
    orientation = _dev.empty((n,), torch.float64)
    axis = _dev.empty((n, 2), torch.float64)
    sums = _dev.empty((n, 6), torch.int64) if want_sums else None
    flist = _dev.empty((max(n, 1) + 1,), torch.int32)     # scratch: frames the streaming fast path leaves to the general kernel
    _lib.call('msq_frame_features', _dev.ptr(cleaned), _dev.ptr(mask), n, h, w, float(frame_threshold),
              _dev.ptr(centroid), _dev.ptr(orientation), _dev.ptr(axis), _dev.ptr(sums), _dev.ptr(flist), flist.numel() * 4,
              _dev.stream())
    return centroid, orientation, axis, sums


def get_frame_features(frames, frame_threshold: float = 10, mask=np.array([]), mask_threshold: float = -30,
                       use_cc: bool = False, progress_bar: bool = True):
    """Moment features of the largest contour per frame (ref: proc/proc.py:237-302).

    Returns (features, mask) like the reference; `features['contour']` is an empty list (contours are
    never materialised: the kernel integrates the same polygon without tracing it)."""
    cleaned = _dev.as_device(frames, torch.uint8)
    if use_cc and mask_threshold >= 0:
        raise NotImplementedError('get_frame_features: use_cc with mask_threshold >= 0 is not implemented '
                                  '(with the default -30 the largest-CC mask is all-True on uint8 frames)')
    has_mask = (isinstance(mask, torch.Tensor) and mask.numel() > 0) or (isinstance(mask, np.ndarray) and mask.size > 0)
    if has_mask:
        mask_dev = _dev.as_device(mask)
        mask_dev = mask_dev.to(torch.uint8) if mask_dev.dtype != torch.uint8 else mask_dev
        mask_out = mask
    else:
        mask_dev = torch.ones_like(cleaned)
        mask_out = _dev.give_back((cleaned.to(torch.float32) > float(frame_threshold)).to(torch.uint8), frames)
    centroid, orientation, axis, _ = _features_device(cleaned, mask_dev, frame_threshold)
    features = {
        'centroid': _dev.give_back(centroid, frames),
        'orientation': _dev.give_back(orientation, frames),
        'axis_length': _dev.give_back(axis, frames),
        'contour': [],
    }
    return features, mask_out


def im_moment_features(image: np.ndarray) -> dict:
    """Features from the polygon moments of one contour (ref: proc/proc.py:518-549).

    Host helper for callers holding an explicit contour ((K,1,2) or (K,2) integer points); a few dozen
    flops, so it stays on the host.  The extract path never calls it (csrc/features.cu moment_epilogue)."""
    pts = np.asarray(image, dtype=np.float64).reshape(-1, 2)
    x, y = pts[:, 0], pts[:, 1]
    xp, yp = np.roll(x, 1), np.roll(y, 1)                       # previous vertex
    cross = xp * y - x * yp
    a00 = cross.sum()
    a10 = (cross * (xp + x)).sum()
    a01 = (cross * (yp + y)).sum()
    a20 = (cross * (xp * (xp + x) + x * x)).sum()
    a11 = (cross * (xp * ((yp + y) + yp) + x * ((yp + y) + y))).sum()
    a02 = (cross * (yp * (yp + y) + y * y)).sum()
    if not abs(a00) > np.finfo(np.float32).eps:
        return {'orientation': np.nan, 'centroid': np.nan, 'axis_length': [np.nan, np.nan]}
    sgn = 1.0 if a00 > 0 else -1.0
    m00, m10, m01 = a00 * 0.5 * sgn, a10 / 6 * sgn, a01 / 6 * sgn
    m20, m11, m02 = a20 / 12 * sgn, a11 / 24 * sgn, a02 / 12 * sgn
    cx, cy = m10 / m00, m01 / m00
    mu20, mu11, mu02 = m20 - m10 * cx, m11 - m10 * cy, m02 - m01 * cy
    den = mu20 - mu02
    common = np.sqrt(4 * np.square(mu11) + np.square(den))
    return {
        'orientation': -.5 * np.arctan2(2 * mu11, den),
        'centroid': [cx, cy],
        'axis_length': [2 * np.sqrt(2) * np.sqrt((mu20 + mu02 + common) / m00),
                        2 * np.sqrt(2) * np.sqrt((mu20 + mu02 - common) / m00)],
    }


# --------------------------------------------------------------------------------------------
# a13  crop_and_rotate_frame
# --------------------------------------------------------------------------------------------
def crop_and_rotate_frames_batch(frames, centers, angles, crop_size: Tuple[int, int] = (80, 80), frames2=None):
    """Batched form of `crop_and_rotate_frame`: (n,h,w) frames, (n,2) centres, (n,) angles in degrees.
    With `frames2` (e.g. the masks) a second stack is warped by the same transforms in the same launch."""
    src = _dev.as_device(frames, torch.uint8)
    n, h, w = (int(v) for v in src.shape)
    cw, ch = int(crop_size[0]), int(crop_size[1])
    cen = _dev.as_device(centers, torch.float64)
    ang = _dev.as_device(angles, torch.float64)
    out = _dev.empty((n, ch, cw), torch.uint8)
    src2 = out2 = None
    if frames2 is not None:
        src2 = _dev.as_device(frames2, torch.uint8)
        out2 = _dev.empty((n, ch, cw), torch.uint8)
    scratch = _dev.empty((int(_lib.load().msq_crop_scratch_bytes(max(n, 1))) + 16,), torch.uint8)
    _lib.call('msq_crop_rotate', _dev.ptr(src), _dev.ptr(src2), n, h, w, _dev.ptr(cen), _dev.ptr(ang), cw, ch,
              _dev.ptr(out), _dev.ptr(out2), _dev.ptr(scratch), scratch.numel(), _dev.stream())
    if frames2 is None:
        return _dev.give_back(out, frames)
    return _dev.give_back(out, frames), _dev.give_back(out2, frames2)


def crop_and_rotate_frame(frame, center: Tuple[float, float], angle: float, crop_size: Tuple[int, int] = (80, 80)):
    """Rotate one frame about `center` by `angle` degrees and crop (ref: proc/proc.py:305-335).
    Bit-exact with cv2.warpAffine's fixed-point bilinear; NaN or negative centre -> zeros; like the
    reference, any failure yields a zero crop instead of an exception (ref: proc/proc.py:334-335)."""
    try:
        c = np.asarray(center.detach().cpu() if isinstance(center, torch.Tensor) else center, dtype=np.float64).reshape(1, 2)
        if np.any(c < 0):
            warnings.warn(f'Encountered center < 0 ({c[0, 0]}, {c[0, 1]}).')
        src = frame[None] if isinstance(frame, torch.Tensor) else np.asarray(frame)[None]
        out = crop_and_rotate_frames_batch(src, c, np.array([angle], dtype=np.float64), crop_size)
        return out[0]
    except (_lib.MoseqB200Error, _dev.CudaRequiredError):
        raise
    except Exception:       # pylint: disable=broad-except
        return np.zeros((crop_size[0], crop_size[1]), dtype=np.uint8)


# --------------------------------------------------------------------------------------------
# a8-a10  angles / flips / filter
# --------------------------------------------------------------------------------------------
def clamp_angles_deg(angles: np.ndarray) -> np.ndarray:
    """Clamp to [0, 360) (ref: proc/proc.py:688-691)."""
    angles = np.asarray(angles)
    return np.where(angles < 0, 360 + angles, angles) % 360


def flips_from_keypoints(keypoints, centroids, angles, length=80):
    """Estimate flips from keypoint votes (ref: proc/proc.py:851-889).  Returns (flips bool, confidence).
    float64 keypoints (the smoothed ones of the tracking branch) are voted on in float64."""
    f64 = _is_f64(keypoints)
    kp = _dev.as_device(keypoints, torch.float64 if f64 else torch.float32)
    n = int(kp.shape[0])
    cen = _dev.as_device(centroids, torch.float64)
    ang = _dev.as_device(angles, torch.float64)
    if np.isscalar(length):
        length = np.full((n,), float(length))
    lens = _dev.as_device(length, torch.float64)
    flips = _dev.empty((n,), torch.uint8)
    conf = _dev.empty((n,), torch.float64)
    _lib.call('msq_flips_from_keypoints_f64' if f64 else 'msq_flips_from_keypoints', _dev.ptr(kp), _dev.ptr(cen),
              _dev.ptr(ang), _dev.ptr(lens), n, _dev.ptr(flips), _dev.ptr(conf), _dev.stream())
    return _dev.give_back(flips.to(torch.bool), keypoints), _dev.give_back(conf, keypoints)


def iterative_filter_angles(angles, window: int = 3, tolerance: float = 60, max_iters: int = 1000):
    """Iterate the 180-degree flip filter until stable (ref: proc/proc.py:627-654).  Returns (angles, flips)."""
    ang = _dev.as_device(angles, torch.float64)
    n = int(ang.shape[0])
    out = torch.empty_like(ang)
    flips = _dev.empty((n,), torch.uint8)
    _lib.call('msq_iterative_filter_angles', _dev.ptr(ang), n, max(n, 1), int(window), float(tolerance), int(max_iters),
              _dev.ptr(out), _dev.ptr(flips), ctypes.c_void_p(0), _dev.stream())
    return _dev.give_back(out, angles), _dev.give_back(flips.to(torch.bool), angles)


def filter_angles(angles, window: int = 3, tolerance: float = 60):
    """One pass of the flip filter (ref: proc/proc.py:600-624) = the iterative filter with max_iters=0."""
    return iterative_filter_angles(angles, window, tolerance, max_iters=0)[0]


# --------------------------------------------------------------------------------------------
# a5 + glue  instances_to_features
# --------------------------------------------------------------------------------------------
def mask_and_keypoints_from_model_output(model_outputs: List[dict]):
    """First instance of every frame as dense arrays (ref: proc/proc.py:657-685).

    Returns (masks (n,1,h,w) uint8, keypoints (n,1,K,3) float64 NaN-filled, num_instances (n,)) as numpy,
    like the reference.  The device path used by `instances_to_features` avoids this host round trip."""
    masks, kpts, ninst = _gather_instances(model_outputs)
    return (masks.cpu().numpy()[:, None], kpts.to(torch.float64).cpu().numpy()[:, None], ninst)


def _gather_instances(model_outputs: List[dict]):
    """Device tensors: masks (n,h,w) u8, keypoints (n,K,3) f32 (NaN where no instance); host num_instances."""
    _dev.require_cuda()
    first = model_outputs[0]['instances']
    h, w = (int(v) for v in first.pred_masks.shape[1:])
    k = int(first.pred_keypoints.shape[1])
    n = len(model_outputs)
    ninst = np.zeros((n,), dtype=int)
    mask_rows, kp_rows = [], []
    zero_mask = torch.zeros((h, w), dtype=torch.uint8, device='cuda')
    nan_kp = torch.full((k, 3), float('nan'), dtype=torch.float32, device='cuda')
    for i, output in enumerate(model_outputs):
        inst = output['instances']
        ninst[i] = len(inst)
        if ninst[i] > 0:
            mask_rows.append(inst.pred_masks[0].to(device='cuda', dtype=torch.uint8, non_blocking=True))
            kp_rows.append(inst.pred_keypoints[0].to(device='cuda', dtype=torch.float32, non_blocking=True))
        else:
            mask_rows.append(zero_mask)
            kp_rows.append(nan_kp)
    return torch.stack(mask_rows).contiguous(), torch.stack(kp_rows).contiguous(), ninst


def _is_f64(x) -> bool:
    return (isinstance(x, torch.Tensor) and x.dtype == torch.float64) or (isinstance(x, np.ndarray) and x.dtype == np.float64)


def _tracked_angles_and_flips(centroid, orientation_rad, axis, kpts32, point_tracker, angle_tracker):
    """The tracking strategy of the reference (proc/proc.py:730-826) on device tensors:
    1) Kalman-smooth centroids and keypoints (tail tip keeps its raw position, :752);
    2) flips from the smoothed keypoints, angles[flips] = clamp(angles + 180);
    3) per frame: read the tracked angle, defer to it when the keypoints are badly aligned (score < 0.4), flip by 180
       when it disagrees by more than 140 degrees, then update the angle filter with the result.
    `KalmanTracker.sample(1)` is the last filtered state itself (the draw it makes only perturbs a discarded
    observation), so the branch is deterministic; estimate_keypoint_rotation (:765) only feeds the debug TSV."""
    from .kalman import KalmanTracker, KalmanTrackerAngle
    if not isinstance(point_tracker, KalmanTracker) or not isinstance(angle_tracker, KalmanTracker):
        raise TypeError('instances_to_features: point_tracker / angle_tracker must be moseq2_detectron_extract_b200.proc.kalman.KalmanTracker')
    if len(angle_tracker.items) != 1 or not isinstance(angle_tracker.items[0], KalmanTrackerAngle) or not angle_tracker.items[0].degrees:
        raise ValueError('instances_to_features: angle_tracker must track exactly one KalmanTrackerAngle in degrees')
    n = int(centroid.shape[0])
    kp64 = kpts32.to(torch.float64).contiguous()
    xy = kp64[:, :, :2].contiguous()
    if not point_tracker.is_initialized:
        point_tracker.initialize([centroid, xy])
    s_centroid, s_xy = point_tracker.smooth_update([centroid, xy])
    centroid = s_centroid.contiguous()
    kp64[:, :7, :2] = s_xy[:, :7, :]
    e = _dev.empty
    angles, flips, conf, scores = e((n,), torch.float64), e((n,), torch.uint8), e((n,), torch.float64), e((n,), torch.float64)
    st = _dev.stream()
    _lib.call('msq_tracking_prepare', _dev.ptr(orientation_rad), _dev.ptr(axis), _dev.ptr(centroid), _dev.ptr(kp64), n,
              _dev.ptr(angles), _dev.ptr(flips), _dev.ptr(conf), _dev.ptr(scores), st)
    if not angle_tracker.is_initialized:
        angle_tracker.initialize([angles])
    m = angle_tracker.device_model()
    mean, cov = angle_tracker.last_mean.clone(), angle_tracker.last_covar.clone()
    _lib.call('msq_track_angles', _dev.ptr(m['A']), _dev.ptr(m['H']), _dev.ptr(m['Q']), _dev.ptr(m['R']), _dev.ptr(mean),
              _dev.ptr(cov), angle_tracker.n_state, _dev.ptr(angles), _dev.ptr(flips), _dev.ptr(scores), n, st)
    angle_tracker.last_mean, angle_tracker.last_covar = mean, cov
    return centroid, kp64, angles, flips


def instances_to_features(model_outputs: List[dict], raw_frames, point_tracker=None, angle_tracker=None,
                          debug: bool = True) -> dict:
    """Clean frames, moment features, flips and angle post-processing for one chunk (ref: proc/proc.py:700-848).
    Without trackers: keypoint flips + the iterative 180-degree filter (:827-839).  With a point and an angle
    `KalmanTracker` (proc/kalman.py of this package): the Kalman tracking branch (:730-826); `debug` only controlled
    the reference's flip_info.tsv dump and is ignored."""
    from ..engine import ChunkEngine
    masks, kpts, ninst = _gather_instances(model_outputs)
    chunk = _dev.as_device(raw_frames, torch.uint8)
    on_dev = _dev.is_device_tensor(raw_frames)
    conv = (lambda t: t) if on_dev else (lambda t: t.cpu().numpy())
    if point_tracker is not None and angle_tracker is not None:
        res = ChunkEngine.shared().clean_and_features(chunk, masks)
        centroid, kp64, angles, flips = _tracked_angles_and_flips(res['centroid'], res['orientation_rad'], res['axis_length'],
                                                                  kpts, point_tracker, angle_tracker)
        return {
            'cleaned_frames': conv(res['cleaned']),
            'masks': conv(masks),
            'features': {'centroid': conv(centroid), 'orientation': conv(angles),
                         'axis_length': conv(res['axis_length']), 'contour': []},
            'flips': conv(flips.to(torch.bool)),
            'keypoints': conv(kp64),
            'num_instances': ninst,
        }
    res = ChunkEngine.shared().features_only(chunk, masks, kpts)
    return {
        'cleaned_frames': conv(res['cleaned']),
        'masks': conv(masks),
        'features': {'centroid': conv(res['centroid']), 'orientation': conv(res['angle_deg']),
                     'axis_length': conv(res['axis_length']), 'contour': []},
        'flips': conv(res['flips'].to(torch.bool)),
        'keypoints': conv(kpts.to(torch.float64)),
        'num_instances': ninst,
    }

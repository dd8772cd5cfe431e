"""ROI helpers on the hot path (mirrors reference proc/roi.py:215-254)."""
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

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

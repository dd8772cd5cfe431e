"""Host-side helpers of the hot path (mirrors reference proc/util.py)."""
from typing import Tuple

import numpy as np


def convert_pxs_to_mm(coords: np.ndarray, resolution: Tuple[int, int] = (512, 424),
                      field_of_view: Tuple[float, float] = (70.6, 60), true_depth: float = 673.1) -> np.ndarray:
    """Pin-hole pixel -> millimetre conversion (ref: proc/util.py:29-61).

    Tiny (N,2) host arithmetic kept in NumPy for callers that use it directly; inside the extract path the
    same arithmetic runs in the scalars/keypoints kernel (csrc/epilogue.cu px_to_mm)."""
    half_w, half_h = resolution[0] // 2, resolution[1] // 2
    focal_w = resolution[0] / (2 * np.deg2rad(field_of_view[0] / 2))
    focal_h = resolution[1] / (2 * np.deg2rad(field_of_view[1] / 2))
    out = np.zeros_like(coords)
    out[:, 0] = true_depth * (coords[:, 0] - half_w) / focal_w
    out[:, 1] = true_depth * (coords[:, 1] - half_h) / focal_h
    return out


def slice_dict(data: dict, index) -> dict:
    """Apply `index` to every array of a (possibly nested) dict (ref: proc/util.py:80-93)."""
    out = {}
    for key, value in data.items():
        out[key] = slice_dict(value, index) if isinstance(value, dict) else value[index]
    return out

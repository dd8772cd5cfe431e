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


def select_strel(shape: str = 'e', size: Tuple[int, int] = (10, 10)) -> np.ndarray:
    """Structuring element as a uint8 array of `size` = (width, height): 'r...' = rectangle, anything else = ellipse
    (ref: proc/util.py:9-26).  Equal to cv2.getStructuringElement(MORPH_RECT / MORPH_ELLIPSE, size): OpenCV's ellipse
    fills, in row i, the columns within round-half-even(c * sqrt(1 - (i - r)^2 / r^2)) of the centre column c = width // 2,
    r = height // 2."""
    width, height = int(size[0]), int(size[1])
    if width <= 0 or height <= 0:
        raise ValueError(f'select_strel: size must be positive, got {size}')
    if shape[0].lower() == 'r' or (width == 1 and height == 1):
        return np.ones((height, width), np.uint8)
    out = np.zeros((height, width), np.uint8)
    r, c = height // 2, width // 2
    inv_r2 = 1.0 / (float(r) * r) if r else 0.0
    for i in range(height):
        dy = i - r
        if abs(dy) <= r:
            dx = int(np.rint(c * np.sqrt((r * r - dy * dy) * inv_r2)))
            out[i, max(c - dx, 0):min(c + dx + 1, width)] = 1
    return out

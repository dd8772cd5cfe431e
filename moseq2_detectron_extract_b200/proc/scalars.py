"""Per-frame scalars (mirrors reference proc/scalars.py)."""
from typing import Dict

import numpy as np
import torch

from .. import _dev, _lib


def scalar_attributes() -> Dict[str, str]:
    """Scalar names with descriptions (ref: proc/scalars.py:6-33)."""
    return {
        'centroid_x_px': 'X centroid (pixels)',
        'centroid_y_px': 'Y centroid (pixels)',
        'velocity_2d_px': '2D velocity (pixels / frame), note that missing frames are not accounted for',
        'velocity_3d_px': '3D velocity (pixels / frame), note that missing frames are not accounted for, also height is in mm, not pixels for calculation',
        'width_px': 'Mouse width (pixels)',
        'length_px': 'Mouse length (pixels)',
        'area_px': 'Mouse area (pixels)',
        'centroid_x_mm': 'X centroid (mm)',
        'centroid_y_mm': 'Y centroid (mm)',
        'velocity_2d_mm': '2D velocity (mm / frame), note that missing frames are not accounted for',
        'velocity_3d_mm': '3D velocity (mm / frame), note that missing frames are not accounted for',
        'width_mm': 'Mouse width (mm)',
        'length_mm': 'Mouse length (mm)',
        'area_mm': 'Mouse area (mm)',
        'height_ave_mm': 'Mouse average height (mm)',
        'angle': 'Angle (radians, unwrapped)',
        'velocity_theta': 'Angular component of velocity (arctan(vel_x, vel_y))',
    }


def scalars_from_table(table, like=None) -> Dict[str, np.ndarray]:
    """(17, n) float64 table (kernel layout) -> dict with the reference's dtypes
    (area_px integer, height_ave_mm float32, everything else float64)."""
    names = _lib.scalar_names()
    if isinstance(table, torch.Tensor) and not _dev.is_device_tensor(like):
        table = table.cpu().numpy()
    out = {}
    for i, name in enumerate(names):
        col = table[i]
        if name == 'area_px':
            col = col.to(torch.int64) if isinstance(col, torch.Tensor) else col.astype(np.int64)
        elif name == 'height_ave_mm':
            col = col.to(torch.float32) if isinstance(col, torch.Tensor) else col.astype(np.float32)
        out[name] = col
    return out


def compute_scalars(frames, track_features: dict, min_height: float = 10, max_height: float = 100,
                    true_depth: float = 673.1) -> Dict[str, np.ndarray]:
    """17 per-frame scalars (ref: proc/scalars.py:36-120).  `frames` is the masked chunk (chunk * mask);
    velocities are first differences over the whole call (the reference calls this once per chunk)."""
    fr = _dev.as_device(frames, torch.uint8)
    n, h, w = (int(v) for v in fr.shape)
    cen = _dev.as_device(track_features['centroid'], torch.float64)
    ang = _dev.as_device(track_features['orientation'], torch.float64)
    axis = _dev.as_device(track_features['axis_length'], torch.float64)
    table = _dev.empty((_lib.NUM_SCALARS, n), torch.float64)
    kp_dummy = torch.zeros((n, 8, 3), dtype=torch.float32, device='cuda')
    scratch = _dev.empty((int(_lib.load().msq_scalars_scratch_bytes(n)) + 8,), torch.uint8)
    _lib.call('msq_scalars_and_keypoints', _dev.ptr(fr), None, _dev.ptr(fr), _dev.ptr(cen), _dev.ptr(ang), _dev.ptr(axis),
              _dev.ptr(kp_dummy), n, h, w, max(n, 1), float(min_height), float(max_height), float(true_depth),
              _dev.ptr(table), None, _dev.ptr(scratch), scratch.numel(), _dev.stream())
    return scalars_from_table(table, like=frames)

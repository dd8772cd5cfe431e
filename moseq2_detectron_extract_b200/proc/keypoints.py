"""Keypoint helpers (mirrors reference proc/keypoints.py:11-165)."""
from typing import Dict, List, Optional, Tuple, Union

import numpy as np
import torch

from .. import _dev, _lib

default_keypoint_names = ['Nose', 'Left Ear', 'Right Ear', 'Neck', 'Left Hip', 'Right Hip', 'TailBase', 'TailTip']


def rotate_points(points: np.ndarray, center: Tuple[float, float] = (0, 0), angle: float = 0) -> np.ndarray:
    """Rotate (K, 2|3) points about `center` by -`angle` degrees; a third column (scores) is kept
    (ref: proc/keypoints.py:11-39).  Host helper; the extract path rotates inside the epilogue kernels."""
    points = np.asarray(points, dtype=np.float64)
    if points.shape[1] not in (2, 3):
        raise ValueError(f'Expected axis 1 of `points` to have length 2 or 3, but got {points.shape[1]}')
    t = np.deg2rad(-angle)
    c, s = np.cos(t), np.sin(t)
    dx, dy = points[:, 0] - center[0], points[:, 1] - center[1]
    out = points.copy()
    out[:, 0] = (c * dx + (-s) * dy) + center[0]
    out[:, 1] = (s * dx + c * dy) + center[1]
    return out


def rotate_points_batch(points: np.ndarray, centers: np.ndarray, angles: Union[np.ndarray, float]) -> np.ndarray:
    """Per-frame `rotate_points` over (n, K, 2|3) points (ref: proc/keypoints.py:42-64); modifies and returns `points`."""
    if isinstance(angles, (int, float)):
        angles = np.full((points.shape[0],), float(angles))
    elif not isinstance(angles, np.ndarray):
        raise TypeError(f'Expected angles to be of type numpy.ndarray or float, got {type(angles).__name__} instead!')
    for i in range(points.shape[0]):
        points[i] = rotate_points(points[i], centers[i], angles[i])
    return points


def keypoint_attributes(keypoint_names: Optional[List[str]] = None) -> Dict[str, str]:
    """Names + descriptions of the 96 keypoint columns (ref: proc/keypoints.py:67-90)."""
    names = keypoint_names or default_keypoint_names
    out = {}
    for kpn in names:
        for cs in ['reference', 'rotated']:
            out[f'{cs}/{kpn}_x_px'] = f'X position of {kpn} (pixels) in {cs} coordinate system.'
            out[f'{cs}/{kpn}_y_px'] = f'Y position of {kpn} (pixels) in {cs} coordinate system.'
            out[f'{cs}/{kpn}_x_mm'] = f'X position of {kpn} (mm) in {cs} coordinate system.'
            out[f'{cs}/{kpn}_y_mm'] = f'Y position of {kpn} (mm) in {cs} coordinate system.'
            out[f'{cs}/{kpn}_z_mm'] = f'Z position of {kpn} (mm) in {cs} coordinate system.'
            out[f'{cs}/{kpn}_score'] = f'Inference score of {kpn}.'
    return out


def keypoints_from_table(table, like=None) -> Dict[str, np.ndarray]:
    """(96, n) float64 table (kernel layout) -> dict keyed like the reference's keypoints_to_dict."""
    if isinstance(table, torch.Tensor) and not _dev.is_device_tensor(like):
        table = table.cpu().numpy()
    return {name: table[i] for i, name in enumerate(_lib.keypoint_col_names())}


def keypoints_to_dict(keypoints, frames, centers, angles, true_depth: float = 673.1,
                      keypoint_names: Optional[List[str]] = None) -> Dict[str, np.ndarray]:
    """Reference / rotated keypoints in px and mm plus the z lookup (ref: proc/keypoints.py:93-165)."""
    f64 = (isinstance(keypoints, torch.Tensor) and keypoints.dtype == torch.float64) or \
          (isinstance(keypoints, np.ndarray) and keypoints.dtype == np.float64)
    kp = _dev.as_device(keypoints, torch.float64 if f64 else torch.float32)
    fr = _dev.as_device(frames, torch.uint8)
    n, h, w = (int(v) for v in fr.shape)
    if tuple(kp.shape) != (n, 8, 3):
        raise ValueError(f'keypoints must be (nframes, 8, 3); got {tuple(kp.shape)}')
    cen = _dev.as_device(centers, torch.float64)
    ang = _dev.as_device(angles, torch.float64)
    axis = torch.zeros((n, 2), dtype=torch.float64, device='cuda')
    table = _dev.empty((_lib.NUM_KPT_COLS, n), torch.float64)
    scratch = _dev.empty((int(_lib.load().msq_scalars_scratch_bytes(n)) + 8,), torch.uint8)
    _lib.call('msq_scalars_and_keypoints_f64' if f64 else 'msq_scalars_and_keypoints', _dev.ptr(fr), None, _dev.ptr(fr),
              _dev.ptr(cen), _dev.ptr(ang), _dev.ptr(axis), _dev.ptr(kp), n, h, w, max(n, 1), 0.0, 0.0, float(true_depth), None,
              _dev.ptr(table), _dev.ptr(scratch), scratch.numel(), _dev.stream())
    return keypoints_from_table(table, like=keypoints)

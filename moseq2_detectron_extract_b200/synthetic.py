"""Seeded synthetic depth sessions (BASELINE.json `north_star`).

Kinect-v2-shaped 512x424 int16 depth frames (the wire dtype of the reference's raw reader,
reference io/video.py:68): a noisy bucket floor with a procedurally generated mouse-shaped
half-ellipsoid moving along a Lissajous path, plus the per-frame "instance" the R-CNN would
have produced (full-frame bool mask + 8 keypoints).  Everything is generated on the host
from `numpy.random.default_rng(seed)` so that the oracle and the GPU path see identical bytes.

This module replaces, for tests and benchmarks, the reference's `io.session.Session`
(reference io/session.py:24) which needs ffprobe and real recordings.
"""
from __future__ import annotations

import dataclasses
from typing import List, Optional, Tuple

import numpy as np

KEYPOINT_NAMES = [  # reference io/annot.py:51-60
    'Nose', 'Left Ear', 'Right Ear', 'Neck', 'Left Hip', 'Right Hip', 'TailBase', 'TailTip'
]


@dataclasses.dataclass
class SessionGeometry:
    """Sensor + arena geometry of a synthetic session."""
    width: int = 512
    height: int = 424
    bucket_center: Tuple[int, int] = (256, 212)     # (x, y)
    bucket_radius: int = 120
    floor_depth: float = 673.0
    wall_depth: float = 500.0
    mouse_axes: Tuple[float, float] = (35.0, 14.0)  # semi-axes in px (long, short)
    mouse_height: float = 45.0
    path_radius: float = 60.0
    crop_size: Tuple[int, int] = (80, 80)

    @staticmethod
    def kinect_v2() -> "SessionGeometry":
        return SessionGeometry()

    @staticmethod
    def azure() -> "SessionGeometry":
        """Azure-Kinect NFOV-unbinned shaped variant (BASELINE.json configs[4])."""
        return SessionGeometry(width=640, height=576, bucket_center=(320, 288), bucket_radius=200,
                               floor_depth=673.0, wall_depth=500.0, mouse_axes=(56.0, 22.0),
                               mouse_height=45.0, path_radius=100.0, crop_size=(128, 128))


def make_roi(geom: SessionGeometry) -> np.ndarray:
    """Bool disk ROI `(H, W)`; its bounding box is max-exclusive downstream (reference roi.py:235,254)."""
    yy, xx = np.mgrid[0:geom.height, 0:geom.width]
    cx, cy = geom.bucket_center
    return ((xx - cx) ** 2 + (yy - cy) ** 2) <= geom.bucket_radius ** 2


def make_background(geom: SessionGeometry, dtype=np.float32, half_steps: bool = False) -> np.ndarray:
    """Background depth image: floor inside the bucket, wall outside.

    `half_steps=True` adds 0.5 to a checkerboard of pixels, which is what a median over an even
    number of frames can produce (SURVEY.md trap 8) and exercises the float truncation.
    """
    roi = make_roi(geom)
    bg = np.where(roi, geom.floor_depth, geom.wall_depth).astype(np.float64)
    if half_steps:
        yy, xx = np.mgrid[0:geom.height, 0:geom.width]
        bg = bg + 0.5 * ((xx + yy) % 2)
    return bg.astype(dtype)


def roi_bbox(roi: np.ndarray) -> Tuple[int, int, int, int]:
    """(y0, x0, y1, x1) with the reference's max-EXCLUSIVE convention (roi.py:254 + :235)."""
    ys, xs = np.nonzero(roi > 0)
    return int(ys.min()), int(xs.min()), int(ys.max()), int(xs.max())


@dataclasses.dataclass
class SyntheticChunk:
    frames: np.ndarray        # (N, H, W) int16 raw depth
    masks: np.ndarray         # (N, h, w) uint8 {0,1}; instance mask in ROI-bbox space (zeros if no instance)
    keypoints: np.ndarray     # (N, 8, 3) float32 [x, y, score] in ROI-bbox space (NaN if no instance)
    num_instances: np.ndarray  # (N,) int64
    heading_deg: np.ndarray   # (N,) ground-truth heading
    center_xy: np.ndarray     # (N, 2) ground-truth centre in full-frame px


def _body_keypoints(cx, cy, theta, a, b):
    """8 analytic keypoints along/around the body axis; theta = heading (rad, image coords, y down)."""
    ux, uy = np.cos(theta), np.sin(theta)      # along the body, pointing to the nose
    vx, vy = -uy, ux                           # across the body
    spec = [(0.90, 0.0), (0.60, -0.5), (0.60, 0.5), (0.45, 0.0),
            (-0.40, -0.6), (-0.40, 0.6), (-0.95, 0.0), (-1.60, 0.0)]
    pts = np.empty((8, 2), dtype=np.float64)
    for i, (al, ac) in enumerate(spec):
        pts[i, 0] = cx + al * a * ux + ac * b * vx
        pts[i, 1] = cy + al * a * uy + ac * b * vy
    return pts


def generate_chunk(n_frames: int, seed: int = 0, geom: Optional[SessionGeometry] = None, t0: int = 0,
                   invalid_rate: float = 0.0, missing_every: int = 0, noise_sigma: float = 1.0,
                   mask_holes: bool = False, realistic: bool = False) -> SyntheticChunk:
    """Generate `n_frames` consecutive frames starting at session time index `t0`.

    realistic    : a less convenient animal -- the body bends into a banana (curvature swings with time: concave outlines, rows
                   with two runs at many headings), it drags a thin curved tail, and the instance mask is what a Mask R-CNN
                   would hand over: the ground-truth silhouette averaged down to a 28x28 soft mask inside its box, pasted back
                   bilinearly and thresholded at 0.5 (ragged, box-clipped outline; the tail mostly falls below 0.5).

    invalid_rate : probability for a pixel to be a Kinect "invalid" (value 0) pixel.
    missing_every: if >0, every `missing_every`-th frame (offset 7) has no instance (NaN path).
    mask_holes   : punch a small hole and detach a speck in some instance masks (exercises the
                   hole-fill / multi-component logic of the feature kernel).
    """
    geom = geom or SessionGeometry()
    rng = np.random.default_rng([seed, t0, n_frames])
    H, W = geom.height, geom.width
    roi = make_roi(geom)
    y0, x0, y1, x1 = roi_bbox(roi)
    h, w = y1 - y0, x1 - x0
    base = np.where(roi, geom.floor_depth, geom.wall_depth).astype(np.float32)

    frames = np.empty((n_frames, H, W), dtype=np.int16)
    masks = np.zeros((n_frames, h, w), dtype=np.uint8)
    kpts = np.full((n_frames, 8, 3), np.nan, dtype=np.float32)
    ninst = np.ones((n_frames,), dtype=np.int64)
    heading = np.empty((n_frames,), dtype=np.float64)
    centers = np.empty((n_frames, 2), dtype=np.float64)

    a, b = geom.mouse_axes
    reach = int(np.ceil(a * (1.9 if realistic else 1.0))) + 2
    bx, by = geom.bucket_center
    for i in range(n_frames):
        t = t0 + i
        cx = bx + geom.path_radius * np.cos(0.05 * t)
        cy = by + geom.path_radius * np.sin(0.07 * t)
        deg = (3.0 * t) % 360.0
        th = np.deg2rad(deg)
        heading[i] = deg
        centers[i] = (cx, cy)

        depth = base + rng.normal(0.0, noise_sigma, size=(H, W)).astype(np.float32)

        xa, xb = max(int(cx) - reach, 0), min(int(cx) + reach + 1, W)
        ya, yb = max(int(cy) - reach, 0), min(int(cy) + reach + 1, H)
        yy, xx = np.mgrid[ya:yb, xa:xb]
        dx, dy = xx - cx, yy - cy
        along = dx * np.cos(th) + dy * np.sin(th)
        across = -dx * np.sin(th) + dy * np.cos(th)
        kappa = (np.sin(0.031 * t) / 45.0) if realistic else 0.0          # body axis bent into a parabola across = kappa s^2 / 2
        u = along / a
        v = (across - 0.5 * kappa * along * along) / b
        rr = u * u + v * v
        body = np.where(rr < 1.0, geom.mouse_height * np.sqrt(np.clip(1.0 - rr, 0.0, 1.0)), 0.0)
        tail = np.zeros_like(body, dtype=bool)
        if realistic:                                                     # 3 px wide, 8 mm high, curling the other way
            ts = np.linspace(-a - 0.85 * a, -a + 2.0, 60)
            tl = 0.5 * kappa * a * a - 0.02 * np.cos(0.017 * t) * (ts + a) ** 2
            d2 = (along[..., None] - ts) ** 2 + (across[..., None] - tl) ** 2
            tail = (d2.min(axis=-1) <= 1.5 ** 2) & (rr >= 1.0)
            body = np.where(tail, 8.0, body)
        depth[ya:yb, xa:xb] -= body.astype(np.float32)
        frames[i] = np.rint(depth).astype(np.int16)

        if missing_every > 0 and (i % missing_every) == 7 % missing_every:
            ninst[i] = 0
            continue

        # instance mask in ROI-bbox space: slightly generous ellipse support
        blob = (rr < 1.08) | tail
        full = np.zeros((H, W), dtype=np.uint8)
        full[ya:yb, xa:xb] = blob
        m = full[y0:y1, x0:x1].copy()
        if realistic and m.any():
            import cv2
            ys, xs = np.nonzero(m)
            by0, by1, bx0, bx1 = ys.min(), ys.max() + 1, xs.min(), xs.max() + 1
            soft = cv2.resize(m[by0:by1, bx0:bx1].astype(np.float32), (28, 28), interpolation=cv2.INTER_AREA)
            pasted = cv2.resize(soft, (int(bx1 - bx0), int(by1 - by0)), interpolation=cv2.INTER_LINEAR) >= 0.5
            m = np.zeros_like(m)
            m[by0:by1, bx0:bx1] = pasted
        if mask_holes and (i % 5) == 3:
            hx, hy = int(cx) - x0, int(cy) - y0
            if 2 <= hy < h - 2 and 2 <= hx < w - 2:
                m[hy - 1:hy + 1, hx - 1:hx + 2] = 0            # enclosed hole
            sy, sx = min(max(hy + reach - 1, 0), h - 3), min(max(hx + reach - 1, 0), w - 3)
            m[sy:sy + 2, sx:sx + 2] = 1                         # detached speck (usually over floor)
        masks[i] = m
        pts = _body_keypoints(cx - x0, cy - y0, th, a, b)
        pts += rng.normal(0.0, 0.4, size=pts.shape)
        kpts[i, :, :2] = pts.astype(np.float32)
        kpts[i, :, 2] = rng.uniform(0.6, 1.0, size=8).astype(np.float32)

    if invalid_rate > 0:
        bad = rng.random(size=frames.shape) < invalid_rate
        frames[bad] = 0
    return SyntheticChunk(frames, masks, kpts, ninst, heading, centers)


def default_config(geom: Optional[SessionGeometry] = None) -> dict:
    """Config keys the reference's steps read (reference cli.py:333-403)."""
    geom = geom or SessionGeometry()
    return {
        'min_height': 0, 'max_height': 100, 'chunk_size': 1000, 'chunk_overlap': 0, 'batch_size': 10,
        'crop_size': tuple(geom.crop_size), 'true_depth': float(geom.floor_depth), 'use_tracking': False,
        'expected_instances': 1, 'debug_feature_processing': False, 'output_dir': '.', 'nframes': 1000,
        'device': 'cuda', 'model': None,
    }


class SyntheticSession:
    """Duck-type of the slice of `io.session.Session` the producer step uses
    (reference produce_frames_step.py:19-27: `.bground_im`, `.roi`, `.iterate(chunk_size, chunk_overlap)`)."""

    def __init__(self, nframes: int, seed: int = 0, geom: Optional[SessionGeometry] = None, **gen_kwargs):
        self.geom = geom or SessionGeometry()
        self.nframes = int(nframes)
        self.seed = seed
        self.gen_kwargs = gen_kwargs
        self.roi = make_roi(self.geom)
        self.bground_im = make_background(self.geom)
        self.true_depth = float(self.geom.floor_depth)
        self._last: Optional[SyntheticChunk] = None

    def iterate(self, chunk_size: int = 1000, chunk_overlap: int = 0):
        return _SyntheticIterator(self, chunk_size, chunk_overlap)


class _SyntheticIterator:
    """Mirrors `SessionFramesIterator` (reference io/session.py:352-466): yields
    `(frame_idxs, frames)` per chunk and applies an attached depth filter."""

    def __init__(self, session: SyntheticSession, chunk_size: int, chunk_overlap: int):
        self.session = session
        self.filters = []
        from .shard import chunk_ranges
        self.batches: List[range] = chunk_ranges(session.nframes, chunk_size, chunk_overlap)   # reference io/util.py:24-35
        self._pos = 0
        self.instances = []   # ground-truth instances per yielded chunk (tests / bench use them)

    def attach_filter(self, stream=None, filterer=None):
        self.filters.append(filterer)

    def __len__(self):
        return len(self.batches)

    def __iter__(self):
        return self

    def __next__(self):
        if self._pos >= len(self.batches):
            raise StopIteration
        idxs = self.batches[self._pos]
        self._pos += 1
        chunk = generate_chunk(len(idxs), seed=self.session.seed, geom=self.session.geom, t0=idxs[0],
                               **self.session.gen_kwargs)
        self.session._last = chunk
        frames = chunk.frames
        for f in self.filters:
            frames = f(frames)
        return list(idxs), frames

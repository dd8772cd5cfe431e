"""Instance identity tracking across frames (a15): the slice of `norfair` the reference uses, restated.

ref: pipeline/process_features_step.py:35-38 builds `norfair.Tracker(distance_function='euclidean', distance_threshold=50,
initialization_delay=0, hit_counter_max=3)`, :115-129 turns every instance into a one-point `Detection` (mask centre of
mass), :140 calls `tracker.update(detections=...)` once per frame and :145-156 keeps the OLDEST live tracked objects.

**Parity unpinned**: norfair (unpinned in the reference's setup.py) is not installed here; this module follows the library's
published behaviour (norfair 2.x `tracker.py`, `filter.py`) for exactly that configuration:

  * every tracked object runs a constant-velocity Kalman filter per coordinate (OptimizedKalmanFilter: R = 4, Q = 0.1, initial
    position variance 10, velocity variance 1, no position/velocity covariance), `estimate` = its position after `predict()`;
  * `update()`: drop objects whose hit counter fell below 0, step every remaining object (hit counter - 1, age + 1, predict),
    match detections to objects greedily by ascending Euclidean distance below the threshold (smallest distance first, each
    detection and object at most once), `hit()` the matched objects (hit counter + 2, capped at hit_counter_max; point hit
    counter + 2, capped at 4), start a new object from every unmatched detection (hit counter 1; with initialization_delay = 0
    it is initialised at once), return the initialised objects with a non-negative hit counter;
  * `live_points` = point hit counter > 0, i.e. seen in the current frame or recently enough.

Host-side Python like the library itself: a handful of objects per frame, nothing here is on the per-pixel path.
"""
from __future__ import annotations

from typing import Any, List, Optional, Sequence

import numpy as np


class Detection:
    """norfair.Detection: `points` (n_points, dim) -- a 1-D point is promoted to one row -- plus free-form `data`."""

    def __init__(self, points, scores=None, data: Any = None, label=None, embedding=None):
        pts = np.asarray(points, dtype=float)
        self.points = pts[np.newaxis, :] if pts.ndim == 1 else pts
        self.absolute_points = self.points.copy()
        self.scores = scores
        self.data = data
        self.label = label
        self.embedding = embedding
        self.age: Optional[int] = None


class _PointFilter:
    """norfair OptimizedKalmanFilter for state (positions, velocities), one independent 2-state filter per coordinate."""

    def __init__(self, points: np.ndarray, r: float = 4.0, q: float = 0.1, pos_variance: float = 10.0,
                 pos_vel_covariance: float = 0.0, vel_variance: float = 1.0):
        self.dim_z = points.size
        self.x = np.zeros((2 * self.dim_z, 1))
        self.x[:self.dim_z, 0] = points.ravel()
        self.pos_variance = np.full((self.dim_z, 1), float(pos_variance))
        self.pos_vel_covariance = np.full((self.dim_z, 1), float(pos_vel_covariance))
        self.vel_variance = np.full((self.dim_z, 1), float(vel_variance))
        self.q = float(q)
        self.r = np.full((self.dim_z, 1), float(r))

    def predict(self) -> None:
        self.x[:self.dim_z] += self.x[self.dim_z:]

    def update(self, z: np.ndarray, observed: Optional[np.ndarray] = None) -> None:
        z = np.asarray(z, dtype=float).reshape(self.dim_z, 1)
        diag = np.ones((self.dim_z, 1)) if observed is None else np.asarray(observed, dtype=float).reshape(self.dim_z, 1)
        error = (z - self.x[:self.dim_z]) * diag
        vel_plus_cov = self.pos_vel_covariance + self.vel_variance
        added = self.pos_variance + self.pos_vel_covariance + vel_plus_cov + self.q + self.r
        r_over = self.r / added
        v_over = vel_plus_cov / added
        added_or_r = added * (1.0 - diag) + self.r * diag
        self.x[:self.dim_z] += diag * (1.0 - r_over) * error
        self.x[self.dim_z:] += diag * v_over * error
        self.pos_variance = (1.0 - r_over) * added_or_r
        self.pos_vel_covariance = v_over * added_or_r
        self.vel_variance = self.vel_variance + self.q - diag * np.square(v_over) * added


class TrackedObject:
    _next_id = 0

    def __init__(self, initial_detection: Detection, hit_counter_max: int, initialization_delay: int, pointwise_hit_counter_max: int,
                 period: int = 1):
        self.num_points, self.dim_points = initial_detection.points.shape
        self.hit_counter_max = int(hit_counter_max)
        self.pointwise_hit_counter_max = max(int(pointwise_hit_counter_max), period)
        self.initialization_delay = int(initialization_delay)
        self.hit_counter = int(period)
        self.point_hit_counter = np.full((self.num_points,), int(period), dtype=int)
        self.age = 0
        self.last_detection = initial_detection
        self.last_distance: Optional[float] = None
        self.is_initializing = self.hit_counter <= self.initialization_delay
        self.id: Optional[int] = None
        if not self.is_initializing:
            self._acquire_id()
        initial_detection.age = self.age
        self.filter = _PointFilter(initial_detection.absolute_points)

    def _acquire_id(self) -> None:
        self.id = TrackedObject._next_id
        TrackedObject._next_id += 1

    @property
    def hit_counter_is_positive(self) -> bool:
        return self.hit_counter >= 0

    @property
    def estimate(self) -> np.ndarray:
        return self.filter.x[:self.filter.dim_z, 0].reshape(self.num_points, self.dim_points)

    @property
    def live_points(self) -> np.ndarray:
        return self.point_hit_counter > 0

    def tracker_step(self) -> None:
        self.hit_counter -= 1
        self.point_hit_counter -= 1
        self.age += 1
        self.filter.predict()

    def hit(self, detection: Detection, period: int = 1) -> None:
        self.last_detection = detection
        detection.age = self.age
        self.hit_counter = min(self.hit_counter + 2 * period, self.hit_counter_max)
        if self.is_initializing and self.hit_counter > self.initialization_delay:
            self.is_initializing = False
            self._acquire_id()
        self.point_hit_counter = np.minimum(self.point_hit_counter + 2 * period, self.pointwise_hit_counter_max)
        self.filter.update(detection.absolute_points.ravel())


class Tracker:
    """norfair.Tracker for scalar distances ('euclidean' = norm of detection.points - object.estimate, or any callable
    (detection, tracked_object) -> float)."""

    def __init__(self, distance_function='euclidean', distance_threshold: float = 50.0, hit_counter_max: int = 15,
                 initialization_delay: Optional[int] = None, pointwise_hit_counter_max: int = 4, **_unused):
        if distance_function == 'euclidean':
            distance_function = lambda det, obj: float(np.linalg.norm(det.points - obj.estimate))       # noqa: E731
        elif not callable(distance_function):
            raise NotImplementedError(f'Tracker: distance_function {distance_function!r} is not restated (the reference uses "euclidean")')
        self.distance_function = distance_function
        self.distance_threshold = float(distance_threshold)
        self.hit_counter_max = int(hit_counter_max)
        self.initialization_delay = int(self.hit_counter_max / 2) if initialization_delay is None else int(initialization_delay)
        if not 0 <= self.initialization_delay < self.hit_counter_max:
            raise ValueError('initialization_delay must be in [0, hit_counter_max)')
        self.pointwise_hit_counter_max = int(pointwise_hit_counter_max)
        self.tracked_objects: List[TrackedObject] = []

    @staticmethod
    def match_dets_and_objs(distance_matrix: np.ndarray, distance_threshold: float):
        """Greedy minimum-distance matching: smallest entry first, its row and column are then out of the game."""
        dist = distance_matrix.copy()
        det_idxs, obj_idxs = [], []
        if dist.size > 0:
            current_min = dist.min()
            while current_min < distance_threshold:
                flat = int(dist.argmin())
                det_idx, obj_idx = flat // dist.shape[1], flat % dist.shape[1]
                det_idxs.append(det_idx)
                obj_idxs.append(obj_idx)
                dist[det_idx, :] = distance_threshold + 1
                dist[:, obj_idx] = distance_threshold + 1
                current_min = dist.min()
        return det_idxs, obj_idxs

    def _update_objects_in_place(self, objects: Sequence[TrackedObject], detections: Sequence[Detection], period: int):
        if len(detections) == 0 or len(objects) == 0:
            return list(detections), [], list(objects)
        dist = np.array([[self.distance_function(d, o) for o in objects] for d in detections], dtype=float)
        if np.isnan(dist).any():
            raise ValueError('Received nan values from distance function, please check your distance function for errors!')
        det_idxs, obj_idxs = self.match_dets_and_objs(dist, self.distance_threshold)
        for di, oi in zip(det_idxs, obj_idxs):
            objects[oi].hit(detections[di], period=period)
            objects[oi].last_distance = float(dist[di, oi])
        unmatched_dets = [d for i, d in enumerate(detections) if i not in set(det_idxs)]
        matched = [objects[oi] for oi in obj_idxs]
        unmatched_objs = [o for i, o in enumerate(objects) if i not in set(obj_idxs)]
        return unmatched_dets, matched, unmatched_objs

    def update(self, detections: Optional[Sequence[Detection]] = None, period: int = 1) -> List[TrackedObject]:
        detections = list(detections or [])
        self.tracked_objects = [o for o in self.tracked_objects if o.hit_counter_is_positive]
        for obj in self.tracked_objects:
            obj.tracker_step()
        alive = self.tracked_objects
        unmatched, _, _ = self._update_objects_in_place([o for o in alive if not o.is_initializing], detections, period)
        unmatched, _, _ = self._update_objects_in_place([o for o in alive if o.is_initializing], unmatched, period)
        for det in unmatched:
            self.tracked_objects.append(TrackedObject(det, self.hit_counter_max, self.initialization_delay,
                                                      self.pointwise_hit_counter_max, period))
        return self.get_active_objects()

    def get_active_objects(self) -> List[TrackedObject]:
        return [o for o in self.tracked_objects if not o.is_initializing and o.hit_counter_is_positive]

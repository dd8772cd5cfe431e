"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's Kalman tracking branch.

Follows reference proc/kalman.py:281-418 (KalmanTracker over pykalman) and proc/proc.py:730-826 (the tracking strategy
of instances_to_features), on top of oracle/pykalman_standin.py (pykalman itself is absent from this image: "parity
unpinned" for the third-party filter, see that file).  Pinned against the UNMODIFIED reference run on the same stand-in
(tests/golden/kinect_tracking.npz, made by oracle/make_golden.py).
"""
import numpy as np
from numpy import ma
from scipy.linalg import block_diag

import extract_oracle as O
from pykalman_standin import KalmanFilter

EXPECTED_ALIGNMENT = np.array([            # ref proc/proc.py:961-984
    [0, 1, 1, 1, 1, 1, 1], [-1, 0, 0, 1, 1, 1, 1], [-1, 0, 0, 1, 1, 1, 1], [-1, -1, -1, 0, 1, 1, 1],
    [-1, -1, -1, -1, 0, 0, 1], [-1, -1, -1, -1, 0, 0, 1], [-1, -1, -1, -1, -1, -1, 0]])


def coordinate_block(order=3, dt=1.0):
    """ref proc/kalman.py:149-168"""
    d = [1.0, dt, dt ** 2 / 2, dt ** 3 / 6][:order]
    m = np.zeros((order, order))
    for r in range(order):
        for i, c in enumerate(range(r, order)):
            m[r, c] = d[i]
    return m


class Tracker:
    """n_coords scalar coordinates at `order`; obs (T, n_coords) with NaN = missing (ref proc/kalman.py:281-418)."""

    def __init__(self, n_coords, order=3):
        self.n, self.order = n_coords, order
        self.A = block_diag(*[coordinate_block(order)] * n_coords)
        pick = np.zeros((1, order)); pick[0, 0] = 1
        self.H = block_diag(*[pick] * n_coords)
        self.kf = None

    def initialize(self, obs):
        m0 = np.zeros(self.n * self.order)
        if obs.shape[0] > 0:
            m0[::self.order] = obs[0]                     # ref :172-188 (first sample, NaN included)
        self.kf = KalmanFilter(transition_matrices=self.A, observation_matrices=self.H, initial_state_mean=m0,
                               em_vars=['transition_covariance', 'observation_covariance', 'initial_state_covariance'])
        Z = ma.masked_invalid(obs)
        rows = np.isfinite(obs).any(axis=1)
        if np.count_nonzero(rows) > 0:
            self.kf.em(Z[rows], n_iter=10)
        self.last_mean, self.last_cov = self.kf.initial_state_mean, self.kf.initial_state_covariance

    def smooth_update(self, obs):
        if obs.shape[0] == 1:
            return self.filter_update(obs[0])[None]
        means, covs = self.kf.smooth(ma.masked_invalid(obs))
        self.last_mean = self.kf.initial_state_mean = means[-1]
        self.last_cov = self.kf.initial_state_covariance = covs[-1]
        return means[:, ::self.order]

    def filter_update(self, z):
        self.last_mean, self.last_cov = self.kf.filter_update(self.last_mean, self.last_cov, ma.masked_invalid(z))
        return self.last_mean[::self.order]


def tracked_angle(state, order=3):
    """KalmanTrackerAngle.inverse_format_data (ref proc/kalman.py:236-242)"""
    a = np.arctan2(state[0], state[order])
    a = 2 * np.pi + a if a < 0 else a
    return np.rad2deg(a)


def angle_difference(a1, a2):
    d = (a2 - a1) % 360
    return -(360 - d) if d > 180 else d


def alignment_scores(rot_xy):
    """ref proc/proc.py:936-958 on (n, 7, 2) rotated keypoints"""
    x = rot_xy[..., 0]
    with np.errstate(invalid='ignore'):
        signs = np.sign(x[:, :, None] - x[:, None, :])
    signs = np.where(EXPECTED_ALIGNMENT == 0, 0, signs)
    met = np.count_nonzero(signs == EXPECTED_ALIGNMENT, axis=(1, 2)) - np.count_nonzero(EXPECTED_ALIGNMENT == 0)
    return met / np.count_nonzero(EXPECTED_ALIGNMENT)


def track_chunk(centroid, orientation_rad, axis_length, keypoints, point_tracker: Tracker, angle_tracker: Tracker):
    """ref proc/proc.py:720-826.  keypoints (n,8,3) float64, NaN where no instance.  Trackers carry state across chunks.
    Returns centroid, keypoints, angles (deg), flips."""
    kp = keypoints.astype(np.float64).copy()
    n = kp.shape[0]
    with np.errstate(invalid='ignore'):
        lengths = np.max(axis_length, axis=1)
        angles = O.clamp_deg(-np.rad2deg(orientation_rad))
    obs = np.column_stack([centroid, kp[:, :, :2].reshape(n, -1)])
    if point_tracker.kf is None:
        point_tracker.initialize(obs)
    sm = point_tracker.smooth_update(obs)
    centroid = sm[:, :2].copy()
    kp[:, :7, :2] = sm[:, 2:].reshape(n, 8, 2)[:, :7]
    flips, _ = O.keypoint_flips(kp, centroid, angles, lengths)
    angles = angles.copy()
    angles[flips] = O.clamp_deg(angles[flips] + 180)
    rot = O.rotate_about(kp[:, :7, :2], centroid, angles)
    scores = alignment_scores(rot)
    to_obs = lambda a: np.array([np.sin(np.deg2rad(a)), np.cos(np.deg2rad(a))])    # noqa: E731
    if angle_tracker.kf is None:
        angle_tracker.initialize(np.column_stack([np.sin(np.deg2rad(angles)), np.cos(np.deg2rad(angles))]))
    flips = flips.copy()
    for i in range(n):
        predicted = tracked_angle(angle_tracker.last_mean, angle_tracker.order)     # sample(1) == the last state
        with np.errstate(invalid='ignore'):
            rel = angle_difference(predicted, angles[i])
        if scores[i] < 0.4:
            angles[i] = predicted
        elif abs(rel) > 140:
            angles[i] = O.clamp_deg(np.array([angles[i] + 180]))[0]
            flips[i] = not flips[i]
        angle_tracker.filter_update(to_obs(angles[i]))
    return centroid, kp, angles, flips

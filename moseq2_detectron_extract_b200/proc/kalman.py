"""Kalman tracking of centroids, keypoints and angles (a14; mirrors the public surface of reference
proc/kalman.py:101-418 -- same class names, constructor arguments and method meaning -- so that
`ProcessFeaturesStep` and `instances_to_features` can take these trackers in place of the reference's).

The reference wraps `pykalman.KalmanFilter` (NumPy, one Python iteration per time step).  Here the model
matrices, the running state and every observation stay in GPU memory and the filter / RTS smoother / EM
run in `csrc/kalman.cu` (`msq_kalman_smooth`, `msq_kalman_em`, `msq_track_angles`).  Items only describe
the block structure of the model and how user data maps to observations.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np
import torch

from .. import _dev, _lib


def angle_difference(angles1: np.ndarray, angles2: np.ndarray) -> np.ndarray:
    """Signed smaller difference angles2 - angles1 in degrees (ref: proc/kalman.py:93-98)."""
    diff = (np.asarray(angles2, dtype=float) - np.asarray(angles1, dtype=float)) % 360
    return np.where(diff > 180, -(360 - diff), diff)


def timestamps_to_steps(timestamps, step_size=(1 / 30 * 1000)):
    """Discrete number of time steps between observations (ref: proc/kalman.py:10-20)."""
    return np.rint(np.diff(timestamps) / step_size).astype(int)


# ------------------------------------------------------------------------------------------------
# items: what is tracked
# ------------------------------------------------------------------------------------------------
class KalmanTrackerItem:
    """`n_coords` scalar coordinates, each with an `order`-long state (value, velocity, acceleration, jerk).
    Subclasses fix `n_coords` and the mapping between user data and the (T, n_coords) observation matrix."""

    n_coords = 1

    def __init__(self, order: int = 3, delta_t: float = 1.0):
        self.order = int(order)
        self.delta_t = float(delta_t)

    # -- model blocks (host, tiny) -----------------------------------------------------------
    def _coordinate_transition(self) -> np.ndarray:
        dt = self.delta_t
        taylor = [1.0, dt, dt ** 2 / 2, dt ** 3 / 6][:self.order]
        block = np.zeros((self.order, self.order))
        for row in range(self.order):
            block[row, row:] = taylor[:self.order - row]
        return block

    def build_trans_mat(self) -> np.ndarray:
        return np.kron(np.eye(self.n_coords), self._coordinate_transition())

    def build_observ_mat(self) -> np.ndarray:
        pick = np.zeros((1, self.order))
        pick[0, 0] = 1.0
        return np.kron(np.eye(self.n_coords), pick)

    @property
    def state_size(self) -> int:
        return self.n_coords * self.order

    @property
    def obs_size(self) -> int:
        return self.n_coords

    # -- data mapping (device tensors, float64) ------------------------------------------------
    def format_data(self, data: torch.Tensor) -> torch.Tensor:
        return data.reshape(data.shape[0], -1)

    def inverse_format_data(self, states: torch.Tensor) -> torch.Tensor:
        return states[:, ::self.order]

    def build_init_state_means(self, data) -> np.ndarray:
        """[first observation, 0, 0, ...] per coordinate (ref: proc/kalman.py:172-188); host helper."""
        obs = self.format_data(_dev.as_device(data, torch.float64))
        out = np.zeros((self.state_size,))
        if obs.shape[0] > 0:
            out[::self.order] = obs[0].cpu().numpy()
        return out


class KalmanTrackerPoint1D(KalmanTrackerItem):
    n_coords = 1


class KalmanTrackerPoint2D(KalmanTrackerItem):
    n_coords = 2


class KalmanTrackerAngle(KalmanTrackerPoint2D):
    """An angle tracked as the point (sin, cos) on the unit circle (ref: proc/kalman.py:213-242)."""

    def __init__(self, order: int = 3, delta_t: float = 1.0, degrees: bool = True):
        super().__init__(order=order, delta_t=delta_t)
        self.degrees = bool(degrees)

    def format_data(self, data: torch.Tensor) -> torch.Tensor:
        ang = data.reshape(-1)
        if self.degrees:
            ang = torch.deg2rad(ang)
        return torch.stack([torch.sin(ang), torch.cos(ang)], dim=1)

    def inverse_format_data(self, states: torch.Tensor) -> torch.Tensor:
        yx = states[:, ::self.order]
        ang = torch.atan2(yx[:, 0], yx[:, 1])
        ang = torch.where(ang < 0, 2 * np.pi + ang, ang)
        return torch.rad2deg(ang) if self.degrees else ang


class KalmanTrackerNPoints2D(KalmanTrackerItem):
    def __init__(self, n_points: int, order: int = 3, delta_t: float = 1):
        super().__init__(order, delta_t)
        self.n_points = int(n_points)
        self.n_coords = 2 * self.n_points

    def inverse_format_data(self, states: torch.Tensor) -> torch.Tensor:
        return states[:, ::self.order].reshape(states.shape[0], self.n_points, -1)


# ------------------------------------------------------------------------------------------------
# the tracker
# ------------------------------------------------------------------------------------------------
def _block_diag(blocks: List[np.ndarray]) -> np.ndarray:
    rows, cols = sum(b.shape[0] for b in blocks), sum(b.shape[1] for b in blocks)
    out = np.zeros((rows, cols))
    r = c = 0
    for b in blocks:
        out[r:r + b.shape[0], c:c + b.shape[1]] = b
        r, c = r + b.shape[0], c + b.shape[1]
    return out


class KalmanTracker:
    """Device-resident Kalman tracker over a list of items (ref: proc/kalman.py:281-418)."""

    EM_ITERATIONS = 10          # ref proc/kalman.py:334

    def __init__(self, items_to_track: Sequence[KalmanTrackerItem]):
        if items_to_track is None or len(items_to_track) <= 0:
            raise ValueError('You need to supply a list of `KalmanTrackerItem`s to the constructor!')
        steps = [item.delta_t for item in items_to_track]
        if not np.allclose(steps, steps[0]):
            raise ValueError('Timesteps across `KalmanTrackerItem` must be the same! Got: ' + ', '.join(str(t) for t in steps))
        self.items = list(items_to_track)
        self.n_state = sum(it.state_size for it in self.items)
        self.n_obs = sum(it.obs_size for it in self.items)
        self._model: Optional[dict] = None          # A, H, Q, R, m0, P0 on the device once initialised
        self.last_mean: Optional[torch.Tensor] = None
        self.last_covar: Optional[torch.Tensor] = None
        self._workspace: Optional[torch.Tensor] = None

    # ---- plumbing ----------------------------------------------------------------------------
    @property
    def is_initialized(self) -> bool:
        return self._model is not None

    def _check(self, data: Sequence) -> None:
        if len(data) != len(self.items):
            raise ValueError(f'Length of data ({len(data)}) does not equal length of `items_to_track` ({len(self.items)})')

    def _format_data(self, data: Sequence) -> torch.Tensor:
        """(T, n_obs) float64 device matrix; non-finite entries mark a missing observation."""
        self._check(data)
        cols = [it.format_data(_dev.as_device(d, torch.float64)) for it, d in zip(self.items, data)]
        return torch.cat(cols, dim=1).contiguous()

    def _inverse_format_data(self, states: torch.Tensor, like) -> List:
        out, off = [], 0
        for it in self.items:
            out.append(_dev.give_back(it.inverse_format_data(states[:, off:off + it.state_size]).contiguous(), like))
            off += it.state_size
        return out

    def _ws(self, T: int, for_em: bool) -> torch.Tensor:
        need = int(_lib.load().msq_kalman_workspace_bytes(T, self.n_state, self.n_obs, int(for_em))) + 256
        if self._workspace is None or self._workspace.numel() < need:
            self._workspace = _dev.empty((need,), torch.uint8)
        return self._workspace

    @staticmethod
    def _aligned(ws: torch.Tensor):
        off = (-ws.data_ptr()) % 256
        return ws[off:], ws.numel() - off

    def _run(self, obs: torch.Tensor, m0: torch.Tensor, P0: torch.Tensor, predict_first: bool, smooth: bool):
        T = int(obs.shape[0])
        m = self._model
        means = _dev.empty((T, self.n_state), torch.float64)
        last_mean = _dev.empty((self.n_state,), torch.float64)
        last_cov = _dev.empty((self.n_state, self.n_state), torch.float64)
        ws, ws_bytes = self._aligned(self._ws(T, False))
        _lib.call('msq_kalman_smooth', _dev.ptr(m['A']), _dev.ptr(m['H']), _dev.ptr(m['Q']), _dev.ptr(m['R']), _dev.ptr(m0),
                  _dev.ptr(P0), _dev.ptr(obs), T, self.n_state, self.n_obs, int(predict_first), int(smooth), _dev.ptr(means),
                  _dev.ptr(last_mean), _dev.ptr(last_cov), _dev.ptr(ws), ws_bytes, _dev.stream())
        return means, last_mean, last_cov

    # ---- reference API -------------------------------------------------------------------------
    def initialize(self, init_data: Sequence) -> None:
        """Build the model, estimate Q, R and P0 by EM on `init_data` (ref: proc/kalman.py:311-342)."""
        _dev.require_cuda()
        self._check(init_data)
        obs = self._format_data(init_data)
        dev = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).cuda()   # noqa: E731
        m0 = torch.zeros((self.n_state,), dtype=torch.float64, device='cuda')
        if obs.shape[0] > 0:
            # every coordinate starts at its first observation, derivatives at 0 (ref: proc/kalman.py:172-188, 344-347)
            first, off_s, off_o = obs[0], 0, 0
            for it in self.items:
                m0[off_s:off_s + it.state_size:it.order] = first[off_o:off_o + it.obs_size]
                off_s, off_o = off_s + it.state_size, off_o + it.obs_size
        self._model = {
            'A': dev(_block_diag([it.build_trans_mat() for it in self.items])),
            'H': dev(_block_diag([it.build_observ_mat() for it in self.items])),
            'Q': dev(np.eye(self.n_state)), 'R': dev(np.eye(self.n_obs)),
            'm0': m0, 'P0': dev(np.eye(self.n_state)),
        }
        # rows without a single finite value are dropped, the rest (partly missing included) go through EM
        usable = obs[torch.isfinite(obs).any(dim=1)].contiguous()
        if usable.shape[0] >= 2:
            m = self._model
            T = int(usable.shape[0])
            ws, ws_bytes = self._aligned(self._ws(T, True))
            _lib.call('msq_kalman_em', _dev.ptr(m['A']), _dev.ptr(m['H']), _dev.ptr(m['Q']), _dev.ptr(m['R']), _dev.ptr(m['m0']),
                      _dev.ptr(m['P0']), _dev.ptr(usable), T, self.n_state, self.n_obs, self.EM_ITERATIONS, _dev.ptr(ws),
                      ws_bytes, _dev.stream())
        self.last_mean = self._model['m0']
        self.last_covar = self._model['P0']

    def _require_init(self):
        if not self.is_initialized:
            raise RuntimeError('KalmanTracker is not initialized; call initialize(init_data) first')

    def smooth(self, data: Sequence) -> List:
        self._require_init()
        means, _, _ = self._run(self._format_data(data), self._model['m0'], self._model['P0'], False, True)
        return self._inverse_format_data(means, data[0])

    def filter(self, data: Sequence) -> List:
        self._require_init()
        means, _, _ = self._run(self._format_data(data), self._model['m0'], self._model['P0'], False, False)
        return self._inverse_format_data(means, data[0])

    def smooth_update(self, data: Sequence) -> List:
        """Smooth a chunk and keep its last state as the prior of the next one (ref: proc/kalman.py:386-401)."""
        self._require_init()
        obs = self._format_data(data)
        if obs.shape[0] == 1:
            return self.filter_update(data)
        means, last_mean, last_cov = self._run(obs, self._model['m0'], self._model['P0'], False, True)
        self.last_mean = self._model['m0'] = last_mean
        self.last_covar = self._model['P0'] = last_cov
        return self._inverse_format_data(means, data[0])

    def filter_update(self, data: Sequence) -> List:
        """One predict + correct step from the running state (ref: proc/kalman.py:408-418)."""
        self._require_init()
        obs = self._format_data(data)[:1].contiguous()
        means, last_mean, last_cov = self._run(obs, self.last_mean, self.last_covar, True, False)
        self.last_mean, self.last_covar = last_mean, last_cov
        return self._inverse_format_data(means, data[0])

    def sample(self, n_timesteps: int = 1, init_data=None, random_state=None) -> List[np.ndarray]:
        """Look `n_timesteps` ahead (ref: proc/kalman.py:370-377 -> pykalman.KalmanFilter.sample): the first sampled
        state IS the start state, later ones add N(0, Q) noise.  Control-plane helper on the host; the extract path
        reads the tracked angle on the device inside `msq_track_angles`."""
        self._require_init()
        m = {k: v.cpu().numpy() for k, v in self._model.items()}
        if init_data is not None:
            state = np.concatenate([it.build_init_state_means(d) for it, d in zip(self.items, init_data)])
        else:
            state = self.last_mean.cpu().numpy()
        rng = random_state if isinstance(random_state, np.random.RandomState) else (
            np.random.mtrand._rand if random_state is None else np.random.RandomState(random_state))
        states = np.zeros((n_timesteps, self.n_state))
        for t in range(n_timesteps):
            states[t] = state if t == 0 else m['A'] @ states[t - 1] + rng.multivariate_normal(np.zeros(self.n_state), m['Q'])
            rng.multivariate_normal(np.zeros(self.n_obs), m['R'])       # the observation draw pykalman makes (and drops)
        out, off = [], 0
        for it in self.items:
            out.append(it.inverse_format_data(torch.from_numpy(states[:, off:off + it.state_size])).numpy())
            off += it.state_size
        return out

    # ---- device access for the extract path ------------------------------------------------------
    def device_model(self) -> dict:
        self._require_init()
        return self._model

    def __str__(self) -> str:
        if not self.is_initialized:
            return f'KalmanTracker(n_state={self.n_state}, n_obs={self.n_obs}, uninitialized)'
        names = {'A': 'transition_matrices', 'H': 'observation_matrices', 'Q': 'transition_covariance',
                 'R': 'observation_covariance', 'm0': 'initial_state_mean', 'P0': 'initial_state_covariance'}
        parts = [f'{names[k]}:\n{v.cpu().numpy()}\n' for k, v in self._model.items()]
        parts.append(f'n_dim_state:\n{self.n_state}\n')
        parts.append(f'n_dim_obs:\n{self.n_obs}\n')
        return '\n'.join(parts)

    __repr__ = __str__

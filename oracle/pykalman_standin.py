"""TEST INFRASTRUCTURE ONLY -- a stand-in for the third-party `pykalman.KalmanFilter`.

PARITY UNPINNED for this module: pykalman (unpinned in the reference's setup.py:27-45) is not installed in this
image, so the reference's tracking branch (proc/proc.py:730-826, proc/kalman.py:281-418) cannot run against the
real library here.  This file restates the *published* algorithm of pykalman's "standard" linear-Gaussian
Kalman filter -- the subset the reference calls (kalman.py:322-333 constructor + `em`, :376 `sample`, :383/:397
`smooth`, :404 `filter`, :411 `filter_update`) -- with pykalman's documented conventions:

* time-invariant A (transition), H (observation), Q, R, zero offsets; defaults Q = I, R = I, P0 = I;
* the first predicted state is (initial_state_mean, initial_state_covariance) itself (no transition before t = 0);
* an observation with ANY masked component is skipped entirely (gain 0, corrected = predicted);
* gains use the Moore-Penrose pseudo-inverse;
* RTS smoother, lag-one pairwise covariances sigma[t] = P_s[t] J[t-1]^T;
* EM M-step for the variables named in `em_vars` (the reference only ever asks for transition_covariance,
  observation_covariance and initial_state_covariance; initial_state_mean stays as given);
* `sample(n, initial_state)`: states[0] = initial_state, later states add N(0, Q) noise, every observation adds
  N(0, R) noise (so `sample(1, x)` returns x itself and consumes one R-draw).

With it installed as `sys.modules['pykalman']` (oracle/ref_import.py) the UNMODIFIED reference kalman.py and
instances_to_features(tracking) execute, which is how tests/golden/*_tracking.npz were produced.
"""
import numpy as np
from numpy import ma
from scipy import linalg


def _obs_is_masked(z):
    return bool(np.any(ma.getmask(z)))


class KalmanFilter:
    def __init__(self, transition_matrices=None, observation_matrices=None, transition_covariance=None,
                 observation_covariance=None, transition_offsets=None, observation_offsets=None,
                 initial_state_mean=None, initial_state_covariance=None, random_state=None,
                 em_vars=('transition_covariance', 'observation_covariance', 'initial_state_mean',
                          'initial_state_covariance'), n_dim_state=None, n_dim_obs=None):
        self.transition_matrices = None if transition_matrices is None else np.asarray(transition_matrices, float)
        self.observation_matrices = None if observation_matrices is None else np.atleast_2d(np.asarray(observation_matrices, float))
        self.transition_covariance = transition_covariance
        self.observation_covariance = observation_covariance
        self.transition_offsets = transition_offsets
        self.observation_offsets = observation_offsets
        self.initial_state_mean = initial_state_mean
        self.initial_state_covariance = initial_state_covariance
        self.random_state = random_state
        self.em_vars = list(em_vars)
        self.n_dim_state = self.transition_matrices.shape[-1] if n_dim_state is None else n_dim_state
        self.n_dim_obs = self.observation_matrices.shape[0] if n_dim_obs is None else n_dim_obs

    # ---- parameters with pykalman's defaults -------------------------------------------------------------------
    def _params(self):
        s, o = self.n_dim_state, self.n_dim_obs
        A = self.transition_matrices
        H = self.observation_matrices
        Q = np.eye(s) if self.transition_covariance is None else np.asarray(self.transition_covariance, float)
        R = np.eye(o) if self.observation_covariance is None else np.asarray(self.observation_covariance, float)
        b = np.zeros(s) if self.transition_offsets is None else np.asarray(self.transition_offsets, float)
        d = np.zeros(o) if self.observation_offsets is None else np.asarray(self.observation_offsets, float)
        m0 = np.zeros(s) if self.initial_state_mean is None else np.asarray(self.initial_state_mean, float)
        P0 = np.eye(s) if self.initial_state_covariance is None else np.asarray(self.initial_state_covariance, float)
        return A, H, Q, R, b, d, m0, P0

    @staticmethod
    def _parse(X):
        Z = ma.asarray(X)
        if Z.ndim == 1:
            Z = Z[:, None] if Z.shape[0] != 1 else Z[None, :]
        return ma.atleast_2d(Z)

    # ---- one predict / one correct -----------------------------------------------------------------------------
    @staticmethod
    def _predict(A, Q, b, m, P):
        return A @ m + b, A @ (P @ A.T) + Q

    @staticmethod
    def _correct(H, R, d, m_pred, P_pred, z):
        if not _obs_is_masked(z):
            z_pred = H @ m_pred + d
            S = H @ (P_pred @ H.T) + R
            K = P_pred @ (H.T @ linalg.pinv(S))
            m = m_pred + K @ (np.asarray(z) - z_pred)
            P = P_pred - K @ (H @ P_pred)
        else:
            K = np.zeros((P_pred.shape[0], H.shape[0]))
            m, P = m_pred, P_pred
        return K, m, P

    def _filter(self, Z):
        A, H, Q, R, b, d, m0, P0 = self._params()
        T, s = Z.shape[0], self.n_dim_state
        mp, Pp = np.zeros((T, s)), np.zeros((T, s, s))
        mf, Pf = np.zeros((T, s)), np.zeros((T, s, s))
        for t in range(T):
            if t == 0:
                mp[t], Pp[t] = m0, P0
            else:
                mp[t], Pp[t] = self._predict(A, Q, b, mf[t - 1], Pf[t - 1])
            _, mf[t], Pf[t] = self._correct(H, R, d, mp[t], Pp[t], Z[t])
        return mp, Pp, mf, Pf

    def _smooth(self, mp, Pp, mf, Pf):
        A = self.transition_matrices
        T, s = mf.shape
        ms, Ps, J = np.zeros((T, s)), np.zeros((T, s, s)), np.zeros((max(T - 1, 0), s, s))
        ms[-1], Ps[-1] = mf[-1], Pf[-1]
        for t in reversed(range(T - 1)):
            J[t] = Pf[t] @ (A.T @ linalg.pinv(Pp[t + 1]))
            ms[t] = mf[t] + J[t] @ (ms[t + 1] - mp[t + 1])
            Ps[t] = Pf[t] + J[t] @ ((Ps[t + 1] - Pp[t + 1]) @ J[t].T)
        return ms, Ps, J

    # ---- public API used by the reference ----------------------------------------------------------------------
    def filter(self, X):
        _, _, mf, Pf = self._filter(self._parse(X))
        return mf, Pf

    def smooth(self, X):
        mp, Pp, mf, Pf = self._filter(self._parse(X))
        ms, Ps, _ = self._smooth(mp, Pp, mf, Pf)
        return ms, Ps

    def filter_update(self, filtered_state_mean, filtered_state_covariance, observation=None, **_):
        A, H, Q, R, b, d, _, _ = self._params()
        if observation is None:
            z = ma.array(np.zeros(self.n_dim_obs), mask=True)
        else:
            z = ma.asarray(observation)
        m_pred, P_pred = self._predict(A, Q, b, np.asarray(filtered_state_mean, float), np.asarray(filtered_state_covariance, float))
        _, m, P = self._correct(H, R, d, m_pred, P_pred, z)
        return m, P

    def sample(self, n_timesteps, initial_state=None, random_state=None):
        A, H, Q, R, b, d, m0, P0 = self._params()
        rs = self.random_state if random_state is None else random_state
        rng = rs if isinstance(rs, np.random.RandomState) else (np.random.mtrand._rand if rs is None else np.random.RandomState(rs))
        states = np.zeros((n_timesteps, self.n_dim_state))
        obs = np.zeros((n_timesteps, self.n_dim_obs))
        if initial_state is None:
            initial_state = rng.multivariate_normal(m0, P0)
        for t in range(n_timesteps):
            if t == 0:
                states[t] = initial_state
            else:
                states[t] = A @ states[t - 1] + b + rng.multivariate_normal(np.zeros(self.n_dim_state), Q)
            obs[t] = H @ states[t] + d + rng.multivariate_normal(np.zeros(self.n_dim_obs), R)
        return states, ma.array(obs)

    def em(self, X, y=None, n_iter=10, em_vars=None):
        Z = self._parse(X)
        A, H, Q, R, b, d, m0, P0 = self._params()
        # unspecified parameters take their defaults before the first E-step
        self.transition_covariance, self.observation_covariance = Q, R
        self.transition_offsets, self.observation_offsets = b, d
        self.initial_state_mean, self.initial_state_covariance = m0, P0
        em_vars = self.em_vars if em_vars is None else list(em_vars)
        T = Z.shape[0]
        for _ in range(n_iter):
            mp, Pp, mf, Pf = self._filter(Z)
            ms, Ps, J = self._smooth(mp, Pp, mf, Pf)
            pair = np.zeros_like(Ps)
            for t in range(1, T):
                pair[t] = Ps[t] @ J[t - 1].T
            if 'transition_covariance' in em_vars:
                acc = np.zeros_like(Q)
                for t in range(T - 1):
                    err = ms[t + 1] - A @ ms[t] - b
                    VA = pair[t + 1] @ A.T
                    acc += np.outer(err, err) + A @ (Ps[t] @ A.T) + Ps[t + 1] - VA - VA.T
                self.transition_covariance = acc / (T - 1)
            if 'observation_covariance' in em_vars:
                acc, n_obs = np.zeros_like(R), 0
                for t in range(T):
                    if not _obs_is_masked(Z[t]):
                        err = np.asarray(Z[t]) - H @ ms[t] - d
                        acc += np.outer(err, err) + H @ (Ps[t] @ H.T)
                        n_obs += 1
                self.observation_covariance = acc / n_obs if n_obs > 0 else acc
            if 'initial_state_mean' in em_vars:
                self.initial_state_mean = ms[0].copy()
            if 'initial_state_covariance' in em_vars:
                x0, mu = ms[0], np.asarray(self.initial_state_mean, float)
                self.initial_state_covariance = (Ps[0] + np.outer(x0, x0) - np.outer(mu, x0) - np.outer(x0, mu)
                                                 + np.outer(mu, mu))
        return self

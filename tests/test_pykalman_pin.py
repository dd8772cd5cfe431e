"""CPU: pin of oracle/pykalman_standin.py (f1).  pykalman itself is not installable here, so the stand-in the reference's
tracking branch runs on (oracle/ref_import.py) is checked against an INDEPENDENT computation that shares no code and no
recursion with it: the linear-Gaussian state-space model

    x_0 ~ N(m0, P0),   x_t = A x_{t-1} + b + w_t,  w ~ N(0, Q),   z_t = H x_t + d + v_t,  v ~ N(0, R)

is written out as ONE joint Gaussian over (x_0 .. x_{T-1}, z_0 .. z_{T-1}); filtered / smoothed means and covariances and the
lag-one cross-covariances are then plain Gaussian conditioning on the observed z (np.linalg.solve on the big covariance).
That settles, by first principles rather than by reading pykalman:
  * the first prediction is the prior itself (z_0 observes x_0, no transition before t = 0)           [convention (i)]
  * a missing observation leaves the prediction untouched (its row is simply not conditioned on)      [convention (ii), whole-vector masks]
  * Kalman filter, RTS smoother and the pairwise covariances P_s[t] J[t-1]^T                           [algorithms]
  * the EM M-step for Q, R, P0 (Shumway & Stoffer; pykalman `_em_*`): expected complete-data moments from the exact posterior,
    and every EM iteration raises the exact marginal log-likelihood                                     [em()]
  * filter_update == one step of filter();  sample(1, x) returns x                                     [kalman.py:376,411]
What stays a documented pykalman convention without an independent derivation: an observation with only SOME components masked
is skipped as a whole (pykalman tests `np.any(np.ma.getmask(z))`), and gains use a pseudo-inverse (identical to the inverse for
the positive-definite innovation covariances of this model)."""
import numpy as np
import pytest
from numpy import ma

import pykalman_standin as PK


def random_system(rng, s, o, T, offsets=True):
    A = np.eye(s) + 0.3 * rng.normal(size=(s, s))
    H = rng.normal(size=(o, s))

    def spd(k, scale):
        M = rng.normal(size=(k, k))
        return scale * (M @ M.T + k * np.eye(k))
    Q, R, P0 = spd(s, 0.05), spd(o, 0.2), spd(s, 0.5)
    b = 0.1 * rng.normal(size=s) if offsets else np.zeros(s)
    d = 0.1 * rng.normal(size=o) if offsets else np.zeros(o)
    m0 = rng.normal(size=s)
    return A, H, Q, R, b, d, m0, P0


def joint_gaussian(A, H, Q, R, b, d, m0, P0, T):
    """Mean and covariance of (x_0..x_{T-1}, z_0..z_{T-1}) by propagating the linear model symbolically: x = M e + c with
    e = (x_0 - m0, w_1 .. w_{T-1}) independent, so Cov(x) = M blockdiag(P0, Q, .., Q) M^T."""
    s, o = A.shape[0], H.shape[0]
    M = np.zeros((T * s, T * s))
    c = np.zeros(T * s)
    M[:s, :s] = np.eye(s)
    c[:s] = m0
    for t in range(1, T):
        M[t * s:(t + 1) * s] = A @ M[(t - 1) * s:t * s]
        M[t * s:(t + 1) * s, t * s:(t + 1) * s] = np.eye(s)
        c[t * s:(t + 1) * s] = A @ c[(t - 1) * s:t * s] + b
    D = np.zeros((T * s, T * s))
    D[:s, :s] = P0
    for t in range(1, T):
        D[t * s:(t + 1) * s, t * s:(t + 1) * s] = Q
    Sx = M @ D @ M.T
    Hb = np.kron(np.eye(T), H)
    mean = np.concatenate([c, Hb @ c + np.tile(d, T)])
    cov = np.block([[Sx, Sx @ Hb.T], [Hb @ Sx, Hb @ Sx @ Hb.T + np.kron(np.eye(T), R)]])
    return mean, cov


def condition(mean, cov, T, s, o, Z, observed_steps):
    """Posterior of all states given z_t for t in observed_steps (whole vectors)."""
    nx = T * s
    idx = np.concatenate([nx + t * o + np.arange(o) for t in observed_steps]) if len(observed_steps) else np.zeros(0, int)
    if len(idx) == 0:
        return mean[:nx], cov[:nx, :nx], 0.0
    z = np.concatenate([Z[t] for t in observed_steps])
    Szz, Sxz = cov[np.ix_(idx, idx)], cov[:nx, idx]
    G = np.linalg.solve(Szz, Sxz.T).T
    r = z - mean[idx]
    loglik = -0.5 * (r @ np.linalg.solve(Szz, r) + np.linalg.slogdet(Szz)[1] + len(idx) * np.log(2 * np.pi))
    return mean[:nx] + G @ r, cov[:nx, :nx] - G @ Sxz.T, loglik


def make_filter(A, H, Q, R, b, d, m0, P0, **kw):
    return PK.KalmanFilter(transition_matrices=A, observation_matrices=H, transition_covariance=Q, observation_covariance=R,
                           transition_offsets=b, observation_offsets=d, initial_state_mean=m0, initial_state_covariance=P0, **kw)


@pytest.mark.parametrize('seed,s,o,T,missing', [(0, 2, 1, 6, ()), (1, 3, 2, 7, (2, 3)), (2, 3, 3, 5, (0,)), (3, 4, 2, 8, (1, 6, 7))])
def test_filter_and_smoother_equal_joint_gaussian_conditioning(seed, s, o, T, missing):
    rng = np.random.default_rng(seed)
    sysm = random_system(rng, s, o, T)
    A, H, Q, R, b, d, m0, P0 = sysm
    mean, cov = joint_gaussian(*sysm, T)
    Z = rng.multivariate_normal(mean, cov)[T * s:].reshape(T, o)
    Zm = ma.asarray(Z.copy())
    for t in missing:
        Zm[t] = ma.masked
    seen = [t for t in range(T) if t not in missing]
    kf = make_filter(*sysm)
    mf, Pf = kf.filter(Zm)
    ms, Ps = kf.smooth(Zm)
    for t in range(T):                                             # filter: condition on z_0 .. z_t
        pm, pc, _ = condition(mean, cov, T, s, o, Z, [u for u in seen if u <= t])
        assert np.allclose(mf[t], pm[t * s:(t + 1) * s], rtol=1e-9, atol=1e-10)
        assert np.allclose(Pf[t], pc[t * s:(t + 1) * s, t * s:(t + 1) * s], rtol=1e-8, atol=1e-10)
    pm, pc, _ = condition(mean, cov, T, s, o, Z, seen)             # smoother: condition on everything observed
    assert np.allclose(ms.ravel(), pm, rtol=1e-8, atol=1e-9)
    for t in range(T):
        assert np.allclose(Ps[t], pc[t * s:(t + 1) * s, t * s:(t + 1) * s], rtol=1e-7, atol=1e-9)
    # lag-one cross covariances Cov(x_t, x_{t-1} | Z) = P_s[t] J[t-1]^T, the quantity the M-step uses
    mp_, Pp_, mf_, Pf_ = kf._filter(kf._parse(Zm))
    _, _, J = kf._smooth(mp_, Pp_, mf_, Pf_)
    for t in range(1, T):
        assert np.allclose(Ps[t] @ J[t - 1].T, pc[t * s:(t + 1) * s, (t - 1) * s:t * s], rtol=1e-7, atol=1e-9)
    # convention (i): with nothing observed at t = 0 the filtered state IS the prior
    if 0 in missing:
        assert np.allclose(mf[0], m0) and np.allclose(Pf[0], P0)
    # filter_update (kalman.py:411) is one filter step; a None observation is a prediction
    m1, P1 = kf.filter_update(mf[1], Pf[1], observation=Zm[2])
    assert np.allclose(m1, mf[2]) and np.allclose(P1, Pf[2])
    mN, PN = kf.filter_update(mf[1], Pf[1], observation=None)
    assert np.allclose(mN, A @ mf[1] + b) and np.allclose(PN, A @ Pf[1] @ A.T + Q)


@pytest.mark.parametrize('seed,s,o,T,missing', [(5, 2, 2, 9, (3,)), (6, 3, 1, 8, ()), (7, 3, 2, 10, (0, 4, 5))])
def test_em_m_step_equals_expected_moments_and_raises_likelihood(seed, s, o, T, missing):
    rng = np.random.default_rng(seed)
    sysm = random_system(rng, s, o, T, offsets=False)              # the reference's trackers have zero offsets
    A, H, Q, R, b, d, m0, P0 = sysm
    mean, cov = joint_gaussian(*sysm, T)
    Z = rng.multivariate_normal(mean, cov)[T * s:].reshape(T, o)
    Zm = ma.asarray(Z.copy())
    for t in missing:
        Zm[t] = ma.masked
    seen = [t for t in range(T) if t not in missing]
    em_vars = ['transition_covariance', 'observation_covariance', 'initial_state_covariance']      # kalman.py:329-333
    kf = make_filter(*sysm, em_vars=em_vars).em(Zm, n_iter=1)
    pm, pc, ll0 = condition(mean, cov, T, s, o, Z, seen)
    X = pm.reshape(T, s)

    def blk(i, j):
        return pc[i * s:(i + 1) * s, j * s:(j + 1) * s]
    # E[(x_{t+1} - A x_t)(x_{t+1} - A x_t)^T | Z] summed over t, from the exact posterior
    Qn = np.zeros((s, s))
    for t in range(T - 1):
        e = X[t + 1] - A @ X[t]
        Qn += np.outer(e, e) + blk(t + 1, t + 1) - A @ blk(t, t + 1) - blk(t + 1, t) @ A.T + A @ blk(t, t) @ A.T
    Qn /= T - 1
    Rn = np.zeros((o, o))
    for t in seen:
        e = Z[t] - H @ X[t]
        Rn += np.outer(e, e) + H @ blk(t, t) @ H.T
    Rn /= len(seen)
    P0n = blk(0, 0) + np.outer(X[0] - m0, X[0] - m0)               # the initial mean is NOT in em_vars: it stays m0
    assert np.allclose(kf.transition_covariance, Qn, rtol=1e-7, atol=1e-9)
    assert np.allclose(kf.observation_covariance, Rn, rtol=1e-7, atol=1e-9)
    assert np.allclose(kf.initial_state_covariance, P0n, rtol=1e-7, atol=1e-9)
    assert np.allclose(kf.initial_state_mean, m0)
    # EM never lowers the exact marginal likelihood of the observed data
    lls = [ll0]
    kf2 = make_filter(*sysm, em_vars=em_vars)
    for _ in range(4):
        kf2.em(Zm, n_iter=1)
        mean2, cov2 = joint_gaussian(A, H, np.asarray(kf2.transition_covariance), np.asarray(kf2.observation_covariance), b, d, m0,
                                     np.asarray(kf2.initial_state_covariance), T)
        lls.append(condition(mean2, cov2, T, s, o, Z, seen)[2])
    assert all(b2 >= a2 - 1e-9 for a2, b2 in zip(lls, lls[1:])), lls


def test_defaults_and_sample_conventions():
    """Q = I, R = I, P0 = I, m0 = 0 when unspecified (kalman.py:322-333 passes no covariances before em);
    sample(1, x) returns x itself (kalman.py:376)."""
    rng = np.random.default_rng(9)
    A, H = np.array([[1.0, 1.0], [0.0, 1.0]]), np.array([[1.0, 0.0]])
    kf = PK.KalmanFilter(transition_matrices=A, observation_matrices=H)
    Z = rng.normal(size=(5, 1))
    mean, cov = joint_gaussian(A, H, np.eye(2), np.eye(1), np.zeros(2), np.zeros(1), np.zeros(2), np.eye(2), 5)
    pm, _, _ = condition(mean, cov, 5, 2, 1, Z, list(range(5)))
    assert np.allclose(kf.smooth(Z)[0].ravel(), pm, rtol=1e-9, atol=1e-10)
    x = np.array([3.0, -1.0])
    states, _ = kf.sample(1, initial_state=x, random_state=0)
    assert np.array_equal(states[0], x)
    # a partially masked observation is skipped as a whole (pykalman convention, not derivable: documented above)
    kf2 = PK.KalmanFilter(transition_matrices=np.eye(2), observation_matrices=np.eye(2))
    z = ma.array([[1.0, 2.0]], mask=[[False, True]])
    mf, Pf = kf2.filter(z)
    assert np.allclose(mf[0], 0) and np.allclose(Pf[0], np.eye(2))

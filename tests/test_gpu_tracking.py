"""GPU parity of the Kalman tracking branch.

* msq_kalman_smooth / msq_kalman_em vs oracle/pykalman_standin.py (the restated pykalman filter) on random models;
* instances_to_features(tracking) and ProcessFeaturesStep(use_tracking=True) vs tests/golden/kinect_tracking.npz, the
  outputs of the UNMODIFIED reference (proc/proc.py:730-826 over proc/kalman.py) run on that stand-in.
Tolerances: float64 throughout; the GPU sums in a different order and inverts with Gauss-Jordan where pykalman
pseudo-inverts by SVD, so smoothed positions agree to <= 1e-7 px and angles to <= 1e-6 degrees (BASELINE: 0.5 px, 1e-3 rad).
"""
import numpy as np
import pytest
from numpy import ma

import make_golden
import tracking_oracle as TO
from cases import assert_close, golden
from pykalman_standin import KalmanFilter

pytestmark = pytest.mark.gpu
torch = pytest.importorskip('torch')


def _random_model(rng, n_coords, order=3):
    tr = TO.Tracker(n_coords, order)
    S, Ob = tr.A.shape[0], tr.H.shape[0]

    def spd(n, scale):
        m = rng.normal(size=(n, n))
        return scale * (m @ m.T / n + np.eye(n))
    return tr.A, tr.H, spd(S, 0.05), spd(Ob, 0.5), rng.normal(size=S), spd(S, 2.0)


def _obs(rng, T, H, missing):
    walk = np.cumsum(rng.normal(size=(T, H.shape[0])), axis=0) + rng.normal(scale=0.3, size=(T, H.shape[0]))
    for t in missing:
        walk[t, rng.integers(0, H.shape[0])] = np.nan          # ONE missing component skips the whole row
    return walk


def _device_smooth(A, H, Q, R, m0, P0, obs, predict_first=False, smooth=True):
    from moseq2_detectron_extract_b200 import _dev, _lib
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).cuda()    # noqa: E731
    T, S, Ob = obs.shape[0], A.shape[0], H.shape[0]
    means, lm, lc = _dev.empty((T, S), torch.float64), _dev.empty((S,), torch.float64), _dev.empty((S, S), torch.float64)
    nbytes = int(_lib.load().msq_kalman_workspace_bytes(T, S, Ob, 0))
    ws = _dev.empty((nbytes + 256,), torch.uint8)
    off = (-ws.data_ptr()) % 256
    t = [d(x) for x in (A, H, Q, R, m0, P0, obs)]
    _lib.call('msq_kalman_smooth', *[_dev.ptr(x) for x in t], T, S, Ob, int(predict_first), int(smooth), _dev.ptr(means),
              _dev.ptr(lm), _dev.ptr(lc), _dev.ptr(ws[off:]), nbytes, _dev.stream())
    torch.cuda.synchronize()
    return means.cpu().numpy(), lm.cpu().numpy(), lc.cpu().numpy()


@pytest.mark.parametrize('n_coords,T', [(18, 150), (2, 64), (5, 2), (1, 1)])
def test_kalman_smooth_matches_restated_pykalman(n_coords, T):
    rng = np.random.default_rng(n_coords * 100 + T)
    A, H, Q, R, m0, P0 = _random_model(rng, n_coords)
    obs = _obs(rng, T, H, missing=[t for t in (3, 4, 17, T - 1) if 0 < t < T])
    kf = KalmanFilter(transition_matrices=A, observation_matrices=H, transition_covariance=Q, observation_covariance=R,
                      initial_state_mean=m0, initial_state_covariance=P0)
    ref_f, ref_fc = kf.filter(ma.masked_invalid(obs))
    got_f, lm, lc = _device_smooth(A, H, Q, R, m0, P0, obs, smooth=False)
    assert_close(got_f, ref_f, 1e-9, 1e-9, what='filtered means')
    assert_close(lm, ref_f[-1], 1e-9, 1e-9, what='last mean')
    assert_close(lc, ref_fc[-1], 1e-9, 1e-10, what='last covariance')
    ref_s, _ = kf.smooth(ma.masked_invalid(obs))
    got_s, lm, lc = _device_smooth(A, H, Q, R, m0, P0, obs, smooth=True)
    assert_close(got_s, ref_s, 1e-8, 1e-8, what='smoothed means')
    assert_close(lc, ref_fc[-1], 1e-9, 1e-10, what='last covariance (smoother)')


def test_kalman_full_chunk_steady_state_and_linearity():
    """BASELINE-size chunk (1000 steps, 54 states).  The filter kernel stops updating covariances once they repeat and restarts
    when an observation is missing: bursts of missing rows at several places exercise entering and leaving that mode, and
    the result must still match the restated pykalman filter / smoother.  Size-independent property: with the covariances
    fixed, filtered and smoothed means are linear in (observations, prior mean)."""
    rng = np.random.default_rng(2024)
    A, H, Q, R, m0, P0 = _random_model(rng, 18)
    T = 1000
    missing = [1, 2, 40, 41, 42, 300, 555, 556, 998, 999]
    obs = _obs(rng, T, H, missing)
    kf = KalmanFilter(transition_matrices=A, observation_matrices=H, transition_covariance=Q, observation_covariance=R,
                      initial_state_mean=m0, initial_state_covariance=P0)
    ref_f, ref_fc = kf.filter(ma.masked_invalid(obs))
    ref_s, _ = kf.smooth(ma.masked_invalid(obs))
    got_f, lm, lc = _device_smooth(A, H, Q, R, m0, P0, obs, smooth=False)
    got_s, _, _ = _device_smooth(A, H, Q, R, m0, P0, obs, smooth=True)
    assert_close(got_f, ref_f, 1e-9, 1e-9, what='filtered means, 1000 steps')
    assert_close(got_s, ref_s, 1e-8, 1e-8, what='smoothed means, 1000 steps')
    assert_close(lc, ref_fc[-1], 1e-9, 1e-11, what='last covariance, 1000 steps')
    obs2 = _obs(rng, T, H, missing)                              # same missing pattern, other values
    m02 = rng.normal(size=m0.shape)
    s2, _, _ = _device_smooth(A, H, Q, R, m02, P0, obs2, smooth=True)
    mix, _, _ = _device_smooth(A, H, Q, R, 0.25 * m0 - 3.0 * m02, P0, 0.25 * obs - 3.0 * obs2, smooth=True)
    assert_close(mix, 0.25 * got_s - 3.0 * s2, 1e-9, 1e-8, what='linearity of the smoother')


def test_kalman_filter_update_and_all_missing():
    rng = np.random.default_rng(5)
    A, H, Q, R, m0, P0 = _random_model(rng, 4)
    kf = KalmanFilter(transition_matrices=A, observation_matrices=H, transition_covariance=Q, observation_covariance=R,
                      initial_state_mean=m0, initial_state_covariance=P0)
    z = rng.normal(size=(1, 4))
    ref_m, ref_c = kf.filter_update(m0, P0, z[0])
    got, lm, lc = _device_smooth(A, H, Q, R, m0, P0, z, predict_first=True, smooth=False)
    assert_close(lm, ref_m, 1e-11, 1e-12, what='filter_update mean')
    assert_close(lc, ref_c, 1e-11, 1e-12, what='filter_update covariance')
    z[0, 2] = np.inf                                             # masked_invalid also masks infinities
    ref_m, ref_c = kf.filter_update(m0, P0, ma.masked_invalid(z[0]))
    got, lm, lc = _device_smooth(A, H, Q, R, m0, P0, z, predict_first=True, smooth=False)
    assert_close(lm, ref_m, 1e-12, 1e-13, what='prediction only mean')
    assert_close(lc, ref_c, 1e-12, 1e-13, what='prediction only covariance')


@pytest.mark.parametrize('n_coords,T,iters', [(18, 120, 3), (2, 90, 10)])
def test_kalman_em_matches_restated_pykalman(n_coords, T, iters):
    from moseq2_detectron_extract_b200 import _dev, _lib
    rng = np.random.default_rng(7 + n_coords)
    A, H, _, _, m0, _ = _random_model(rng, n_coords)
    obs = _obs(rng, T, H, missing=[5, 6, 40])
    m0[::3] = obs[0]
    kf = KalmanFilter(transition_matrices=A, observation_matrices=H, initial_state_mean=m0,
                      em_vars=['transition_covariance', 'observation_covariance', 'initial_state_covariance'])
    kf.em(ma.masked_invalid(obs), n_iter=iters)
    S, Ob = A.shape[0], H.shape[0]
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).cuda()    # noqa: E731
    Ad, Hd, Qd, Rd, md, Pd, od = d(A), d(H), d(np.eye(S)), d(np.eye(Ob)), d(m0), d(np.eye(S)), d(obs)
    nbytes = int(_lib.load().msq_kalman_workspace_bytes(T, S, Ob, 1))
    ws = _dev.empty((nbytes + 256,), torch.uint8)
    off = (-ws.data_ptr()) % 256
    _lib.call('msq_kalman_em', _dev.ptr(Ad), _dev.ptr(Hd), _dev.ptr(Qd), _dev.ptr(Rd), _dev.ptr(md), _dev.ptr(Pd), _dev.ptr(od),
              T, S, Ob, iters, _dev.ptr(ws[off:]), nbytes, _dev.stream())
    torch.cuda.synchronize()
    assert_close(Qd.cpu().numpy(), kf.transition_covariance, 1e-7, 1e-9, what='transition_covariance')
    assert_close(Rd.cpu().numpy(), kf.observation_covariance, 1e-7, 1e-9, what='observation_covariance')
    assert_close(Pd.cpu().numpy(), kf.initial_state_covariance, 1e-7, 1e-9, what='initial_state_covariance')


def test_kalman_rejects_bad_arguments():
    from moseq2_detectron_extract_b200 import _lib
    lib = _lib.load()
    assert lib.msq_kalman_workspace_bytes(0, 6, 2, 0) == 0
    with pytest.raises(_lib.MoseqB200Error, match='n_state'):
        _lib.call('msq_kalman_smooth', None, None, None, None, None, None, None, 4, 80, 2, 0, 1, None, None, None, None, 0, None)
    with pytest.raises(_lib.MoseqB200Error, match='null'):
        _lib.call('msq_kalman_smooth', None, None, None, None, None, None, None, 4, 6, 2, 0, 1, None, None, None, None, 0, None)


def _tracking_inputs():
    from moseq2_detectron_extract_b200 import synthetic
    import extract_oracle as O
    geom = synthetic.SessionGeometry.kinect_v2()
    kw = {k: v for k, v in make_golden.TRACKING_CASE.items() if k != 'chunks'}
    chunk = synthetic.generate_chunk(geom=geom, **kw)
    roi, bg = synthetic.make_roi(geom), synthetic.make_background(geom)
    prep = O.prep_frames(chunk.frames, bg, roi, 0, 100)
    return geom, chunk, prep, synthetic.default_config(geom)


def _fake_outputs(chunk, sl):
    return [{'instances': make_golden.FakeInstances(chunk.masks[i], chunk.keypoints[i], chunk.num_instances[i] > 0)}
            for i in range(sl.start, sl.stop)]


def test_tracking_branch_matches_reference():
    import moseq2_detectron_extract_b200.proc as P
    from moseq2_detectron_extract_b200.proc.kalman import (KalmanTracker, KalmanTrackerAngle, KalmanTrackerNPoints2D,
                                                           KalmanTrackerPoint2D)
    geom, chunk, prep, cfg = _tracking_inputs()
    g = golden('kinect_tracking')
    pt = KalmanTracker([KalmanTrackerPoint2D(order=3, delta_t=1.0), KalmanTrackerNPoints2D(8, order=3, delta_t=1.0)])
    at = KalmanTracker([KalmanTrackerAngle(order=3, delta_t=1.0, degrees=True)])
    assert not pt.is_initialized and pt.n_state == 54 and pt.n_obs == 18 and at.n_state == 6
    start = 0
    for c, n in enumerate(make_golden.TRACKING_CASE['chunks']):
        sl = slice(start, start + n)
        feats = P.instances_to_features(_fake_outputs(chunk, sl), prep[sl], pt, at, debug=False)
        assert np.isnan(g[f'c{c}/keypoints']).any()                                   # frames without an instance are in
        assert_close(feats['features']['centroid'], g[f'c{c}/centroid'], 0, 1e-7, what=f'chunk {c} smoothed centroid')
        assert_close(feats['keypoints'], g[f'c{c}/keypoints'], 0, 1e-7, what=f'chunk {c} smoothed keypoints')
        assert_close(feats['features']['orientation'], g[f'c{c}/orientation'], 0, 1e-6, what=f'chunk {c} tracked angle')
        assert np.array_equal(feats['flips'], g[f'c{c}/flips'])
        assert_close(pt.last_mean.cpu().numpy(), g[f'c{c}/point_last_mean'], 0, 1e-7, what='point tracker state')
        assert_close(at.last_mean.cpu().numpy(), g[f'c{c}/angle_last_mean'], 0, 1e-7, what='angle tracker state')
        start += n
    m = pt.device_model()
    assert_close(m['Q'].cpu().numpy(), g['point_transition_covariance'], 1e-6, 1e-9, what='EM transition covariance')
    assert_close(m['R'].cpu().numpy(), g['point_observation_covariance'], 1e-6, 1e-9, what='EM observation covariance')
    assert_close(at.device_model()['Q'].cpu().numpy(), g['angle_transition_covariance'], 1e-6, 1e-10, what='angle EM Q')
    # the same chunks through the restated oracle (independent of the stored file)
    pt2, at2 = TO.Tracker(18), TO.Tracker(2)
    import extract_oracle as O
    cleaned = O.clean_frames_cv2(prep[:100])
    f = O.frame_features_cv2(cleaned, chunk.masks[:100])
    cen, kp, ang, fl = TO.track_chunk(f['centroid'], f['orientation'], f['axis_length'], chunk.keypoints[:100], pt2, at2)
    assert_close(cen, g['c0/centroid'], 0, 1e-12, what='oracle centroid')
    assert_close(ang, g['c0/orientation'], 0, 1e-9, what='oracle angle')


def test_tracker_api_round_trips():
    """KalmanTracker surface (ref proc/kalman.py:281-418): filter / smooth leave the state alone, smooth_update and
    filter_update advance it, sample(1) returns the running state, numpy in -> numpy out."""
    from moseq2_detectron_extract_b200.proc.kalman import KalmanTracker, KalmanTrackerAngle, KalmanTrackerPoint2D, angle_difference
    rng = np.random.default_rng(3)
    pts = np.cumsum(rng.normal(size=(80, 2)), axis=0)
    pts[7] = np.nan
    tr = KalmanTracker([KalmanTrackerPoint2D(order=3, delta_t=1.0)])
    with pytest.raises(RuntimeError):
        tr.smooth([pts])
    with pytest.raises(ValueError):
        tr.initialize([pts, pts])
    tr.initialize([pts])
    ref = TO.Tracker(2)
    ref.initialize(pts)
    before = tr.last_mean.clone()
    sm, = tr.smooth([pts])
    fl, = tr.filter([pts])
    assert isinstance(sm, np.ndarray) and sm.shape == (80, 2) and torch.equal(before, tr.last_mean)
    ref_s, _ = ref.kf.smooth(ma.masked_invalid(pts))
    ref_f, _ = ref.kf.filter(ma.masked_invalid(pts))
    assert_close(sm, ref_s[:, ::3], 0, 1e-7, what='smooth')
    assert_close(fl, ref_f[:, ::3], 0, 1e-7, what='filter')
    up, = tr.smooth_update([pts])
    assert_close(up, ref.smooth_update(pts), 0, 1e-7, what='smooth_update')
    one, = tr.filter_update([pts[-1:] + 1.0])
    assert_close(one[0], ref.filter_update(pts[-1] + 1.0), 0, 1e-7, what='filter_update')
    s, = tr.sample(1)
    assert_close(s[0], tr.last_mean.cpu().numpy()[::3], 0, 0, what='sample(1) is the running state')
    ang = KalmanTracker([KalmanTrackerAngle(order=3, delta_t=1.0, degrees=True)])
    deg = (np.arange(60) * 7.0) % 360
    ang.initialize([deg])
    out, = ang.smooth([deg])
    assert np.all(np.abs(angle_difference(out, deg)) < 20)
    assert np.allclose(angle_difference(np.array([350.0, 10.0]), np.array([10.0, 350.0])), [20.0, -20.0])


def test_tracking_step_matches_reference():
    """ProcessFeaturesStep(use_tracking=True) over the two chunks: scalars, keypoint table and crops of the reference."""
    from moseq2_detectron_extract_b200 import _dev
    from moseq2_detectron_extract_b200.pipeline import ProcessFeaturesStep
    geom, chunk, prep, cfg = _tracking_inputs()
    g = golden('kinect_tracking')
    cfg = dict(cfg, use_tracking=True, expected_instances=1)
    step = ProcessFeaturesStep(cfg, 'features')
    step.initialize()
    start = 0
    for c, n in enumerate(make_golden.TRACKING_CASE['chunks']):
        sl = slice(start, start + n)
        data = {'batch': c, 'chunk': prep[sl], 'frame_idxs': list(range(start, start + n)), 'offset': 0,
                '_dense_instances': (_dev.as_device(chunk.masks[sl]), _dev.as_device(chunk.keypoints[sl], torch.float32),
                                     chunk.num_instances[sl])}
        out = step.process(data)
        assert_close(out['features']['features']['orientation'], g[f'c{c}/orientation'], 0, 1e-6, what='orientation')
        assert np.array_equal(out['features']['flips'], g[f'c{c}/flips'])
        for k in g.files:
            if k.startswith(f'c{c}/scalars/'):
                name = k.split('/', 2)[2]
                tol = 1e-4 if 'velocity' in name else 1e-6
                assert_close(out['scalars'][name], g[k], tol, 1e-6, what=k)
            elif k.startswith(f'c{c}/keypoints/'):
                name = k.split('/', 2)[2]
                assert_close(out['keypoints'][name], g[k], 1e-6, 1e-5, what=k)
        # crops: the transform differs from the reference's by ~1e-7 px / 1e-6 degrees, so a tap can round the other way
        diff = np.abs(out['depth_frames'].astype(int) - g[f'c{c}/depth_frames'].astype(int))
        assert diff.max() <= 1 and np.count_nonzero(diff) <= 1e-3 * diff.size, (diff.max(), np.count_nonzero(diff))
        start += n


def test_tracking_edge_chunks_match_oracle():
    """Chunk shapes the reference handles specially: a single-frame chunk (smooth_update -> filter_update, proc/kalman.py:392-395),
    a chunk in which every frame lacks an instance (prediction only; the tracked angle replaces the NaN one), then normal
    frames again -- trackers carry their state across all of them.  Checked against oracle/tracking_oracle.py."""
    import extract_oracle as O
    import moseq2_detectron_extract_b200.proc as P
    from moseq2_detectron_extract_b200 import synthetic
    from moseq2_detectron_extract_b200.proc.kalman import (KalmanTracker, KalmanTrackerAngle, KalmanTrackerNPoints2D,
                                                           KalmanTrackerPoint2D)
    geom = synthetic.SessionGeometry.kinect_v2()
    chunk = synthetic.generate_chunk(n_frames=111, seed=21, geom=geom, t0=5)
    roi, bg = synthetic.make_roi(geom), synthetic.make_background(geom)
    prep = O.prep_frames(chunk.frames, bg, roi, 0, 100)
    masks, kpts, present = chunk.masks.copy(), chunk.keypoints.copy(), (chunk.num_instances > 0).copy()
    gone = slice(81, 91)
    masks[gone] = 0
    kpts[gone] = np.nan
    present[gone] = False
    pt = KalmanTracker([KalmanTrackerPoint2D(order=3, delta_t=1.0), KalmanTrackerNPoints2D(8, order=3, delta_t=1.0)])
    at = KalmanTracker([KalmanTrackerAngle(order=3, delta_t=1.0, degrees=True)])
    pt_ref, at_ref = TO.Tracker(18), TO.Tracker(2)
    for sl in (slice(0, 80), slice(80, 81), gone, slice(91, 111)):
        outs = [{'instances': make_golden.FakeInstances(masks[i], kpts[i], bool(present[i]))} for i in range(sl.start, sl.stop)]
        feats = P.instances_to_features(outs, prep[sl], pt, at, debug=False)
        f = O.frame_features_cv2(O.clean_frames_cv2(prep[sl]), masks[sl])
        cen, kp, ang, fl = TO.track_chunk(f['centroid'], f['orientation'], f['axis_length'], kpts[sl], pt_ref, at_ref)
        what = f'frames {sl.start}..{sl.stop}'
        assert_close(feats['features']['centroid'], cen, 0, 1e-6, what=what + ' centroid')
        assert_close(feats['keypoints'], kp, 0, 1e-6, what=what + ' keypoints')
        assert_close(feats['features']['orientation'], ang, 0, 1e-5, what=what + ' angle')
        assert np.array_equal(feats['flips'], fl), what
        if sl == gone:
            assert np.isfinite(feats['features']['centroid']).all() and np.isfinite(feats['features']['orientation']).all()
            assert np.isnan(feats['keypoints'][:, 7, :2]).all()           # the tail tip keeps its raw (missing) position

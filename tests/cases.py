"""Shared helpers: rebuild the seeded inputs of the golden cases and load the stored reference outputs."""
import os

import numpy as np

import make_golden  # oracle/make_golden.py (on sys.path via conftest); importing it does not touch /root/reference

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
CASE_NAMES = list(make_golden.CASES)
_cache = {}


def case_inputs(name):
    """-> (geom, chunk, roi, bground, cfg) regenerated bit-identically from the seeds."""
    if name not in _cache:
        from moseq2_detectron_extract_b200 import synthetic
        geom, chunk, roi, bg = make_golden.build_case(name)
        _cache[name] = (geom, chunk, roi, bg, synthetic.default_config(geom))
    return _cache[name]


def golden(name):
    return np.load(os.path.join(GOLDEN_DIR, name + '.npz'))


def assert_close(a, b, rtol, atol=0.0, what=''):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, f'{what}: shape {a.shape} vs {b.shape}'
    assert np.array_equal(np.isnan(a), np.isnan(b)), f'{what}: NaN pattern differs'
    ok = ~np.isnan(a)
    err = np.abs(a[ok] - b[ok])
    tol = atol + rtol * np.abs(b[ok])
    assert np.all(err <= tol), f'{what}: max err {err.max() if err.size else 0} (tol rtol={rtol}, atol={atol})'

"""TEST INFRASTRUCTURE ONLY -- import the *unmodified* reference from /root/reference.

Used (a) by oracle/make_golden.py to generate the committed fixtures under tests/golden/ and
(b) by tests that cross-check oracle/extract_oracle.py against the real reference when
/root/reference is present (this container only; the GPU box does not have it).

The reference drags in third-party modules that are absent here and never executed on the
extract hot path.  They are replaced by inert stubs *before* import.
`skimage.measure.label/regionprops` ARE executed by get_roi (session setup) and get the stand-in of
oracle/skimage_standin.py.  `pykalman.KalmanFilter` IS executed on the tracking branch (reference proc/kalman.py) and gets the restated
stand-in of oracle/pykalman_standin.py.  `bottleneck.move_median` IS executed (reference proc/proc.py:618) so it gets a functional
stand-in built on pandas' trailing rolling median, which has the same semantics
(trailing window, NaN-skipping, min_count == min_periods).
"""
import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("MOSEQ_REFERENCE_ROOT", "/root/reference")

_STUBS = [
    "matplotlib", "matplotlib.pyplot", "matplotlib.cm", "matplotlib.colors", "matplotlib.figure",
    "matplotlib.backends", "matplotlib.backends.backend_agg",
    "h5py", "ruamel", "ruamel.yaml", "tifffile", "imageio", "skimage", "skimage.measure", "skimage.draw",
    "skimage.morphology", "skimage.filters",
    "pycocotools", "pycocotools.mask", "detectron2", "detectron2.data", "detectron2.structures",
    "detectron2.utils", "detectron2.utils.visualizer", "detectron2.utils.colormap", "detectron2.data.catalog",
    "norfair", "click", "statsmodels", "statsmodels.api",
]


class _Anything(types.ModuleType):
    """Module whose every attribute is a harmless placeholder class."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        placeholder = type(name, (), {"__init__": lambda self, *a, **k: None,
                                      "__call__": lambda self, *a, **k: None})
        setattr(self, name, placeholder)
        return placeholder


def _move_median(a, window, min_count=None, axis=-1, ddof=0):
    import numpy as np
    import pandas as pd
    if min_count is None:
        min_count = window
    return pd.Series(np.asarray(a, dtype=float)).rolling(window, min_periods=min_count).median().to_numpy()


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "moseq2_detectron_extract"))


def load():
    """Return a namespace with the reference's hot-path modules (proc, scalars, keypoints, roi, util)."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    for name in _STUBS:
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:  # absent -> stub
                sys.modules[name] = _Anything(name)
    if isinstance(sys.modules.get("skimage.measure"), _Anything):    # absent -> restated stand-in for get_roi
        import skimage_standin
        sys.modules["skimage.measure"].label = skimage_standin.label
        sys.modules["skimage.measure"].regionprops = skimage_standin.regionprops
        sys.modules["skimage"].measure = sys.modules["skimage.measure"]
    import pandas  # noqa: F401  (must be imported BEFORE the bottleneck stand-in exists: pandas probes it)
    if "bottleneck" not in sys.modules:
        try:
            importlib.import_module("bottleneck")
        except Exception:
            bn = types.ModuleType("bottleneck")
            bn.move_median = _move_median
            sys.modules["bottleneck"] = bn
    if "pykalman" not in sys.modules:
        try:
            importlib.import_module("pykalman")
        except Exception:       # absent -> restated stand-in (oracle/pykalman_standin.py, "parity unpinned")
            import pykalman_standin
            pk = types.ModuleType("pykalman")
            pk.KalmanFilter = pykalman_standin.KalmanFilter
            sys.modules["pykalman"] = pk
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    ns = types.SimpleNamespace()
    ns.proc = importlib.import_module("moseq2_detectron_extract.proc.proc")
    ns.roi = importlib.import_module("moseq2_detectron_extract.proc.roi")
    ns.scalars = importlib.import_module("moseq2_detectron_extract.proc.scalars")
    ns.keypoints = importlib.import_module("moseq2_detectron_extract.proc.keypoints")
    ns.util = importlib.import_module("moseq2_detectron_extract.proc.util")
    ns.kalman = importlib.import_module("moseq2_detectron_extract.proc.kalman")
    return ns

"""
TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Imports the *unmodified* reference (pearsonlab/autoencoded-vocal-analysis, mounted
read-only at /root/reference) inside THIS container so that golden vectors can be
generated from the reference's own code and the oracle restatements under oracle/
can be pinned against it.  /root/reference does not exist on the GPU box: nothing
that runs there (pytest -m gpu, smoke(), bench.py) may import this module.

Why stubs are needed:
  * ava/models/vae.py:29-30 imports vae_dataset (-> h5py) and plotting.grid_plot
    (-> matplotlib) at module top; ava/models/utils.py:8-11 imports affinewarp.
    None of these are installed here, and none is touched by the hot path.
  * ava/preprocessing/utils.py:11 imports scipy.interpolate.interp2d, removed in
    SciPy >= 1.14.  The shim below follows SciPy's own transition guide:
    RectBivariateSpline(kx=1, ky=1) == plain bilinear interpolation, plus the
    fill_value rule of interp2d(bounds_error=False).
"""
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("AVA_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "ava"))


class _Interp2dShim:
    """interp2d(x, y, z, bounds_error=False, fill_value=v) for kind='linear'."""

    def __init__(self, x, y, z, kind='linear', copy=True, bounds_error=False, \
        fill_value=None):
        from scipy.interpolate import RectBivariateSpline
        assert kind == 'linear' and not bounds_error
        self.x = np.asarray(x, dtype=np.float64)
        self.y = np.asarray(y, dtype=np.float64)
        self.fill_value = fill_value
        # z has shape [len(y), len(x)] in interp2d convention.
        self.spline = RectBivariateSpline(self.x, self.y, \
                np.asarray(z, dtype=np.float64).T, kx=1, ky=1, s=0)

    def __call__(self, x, y, assume_sorted=False):
        x = np.atleast_1d(np.asarray(x, dtype=np.float64))
        y = np.atleast_1d(np.asarray(y, dtype=np.float64))
        out = self.spline(x, y).T
        if self.fill_value is not None:
            bad_x = (x < self.x[0]) | (x > self.x[-1])
            bad_y = (y < self.y[0]) | (y > self.y[-1])
            out[:, bad_x] = self.fill_value
            out[bad_y, :] = self.fill_value
        return out


def install_stubs():
    """Register the empty stand-in modules the reference needs to import."""
    def _mod(name, **attrs):
        if name in sys.modules:
            return sys.modules[name]
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m
    try:
        import h5py  # noqa: F401
    except ImportError:
        _mod('h5py', File=None)
    try:
        import affinewarp  # noqa: F401
    except ImportError:
        aw = _mod('affinewarp', PiecewiseWarping=None)
        cv = _mod('affinewarp.crossval', paramsearch=None)
        aw.crossval = cv
    try:
        import matplotlib  # noqa: F401
    except ImportError:
        mpl = _mod('matplotlib', use=lambda *a, **k: None)
        plt = _mod('matplotlib.pyplot', switch_backend=lambda *a, **k: None)
        mpl.pyplot = plt
    import scipy.interpolate
    if not hasattr(scipy.interpolate, '_ava_interp2d_shimmed'):
        try:
            # Present-but-raising stub in SciPy >= 1.14.
            scipy.interpolate.interp2d([0, 1], [0, 1], [[0, 1], [1, 2]])
        except Exception:
            scipy.interpolate.interp2d = _Interp2dShim
        scipy.interpolate._ava_interp2d_shimmed = True


def import_reference():
    """Return the reference's (vae, preprocessing.utils, window_vae_dataset,
    vae_dataset) modules, imported unmodified."""
    if not reference_available():
        raise RuntimeError("reference not present at " + REFERENCE_ROOT)
    install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import ava.models.vae as ref_vae
    import ava.preprocessing.utils as ref_pre
    import ava.models.window_vae_dataset as ref_win
    import ava.models.vae_dataset as ref_ds
    # The module-level name bound at import time must be the shim too.
    import scipy.interpolate
    ref_pre.interp2d = scipy.interpolate.interp2d
    return ref_vae, ref_pre, ref_win, ref_ds

"""Shared helpers for the parity tests (golden digests, tolerances)."""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
N_DIGEST = 4096


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)


def digest_indices(numel, salt=0):
    rng = np.random.default_rng(77 + salt)
    return rng.integers(0, numel, size=N_DIGEST)


def digest(a, salt=0):
    a = np.asarray(a, dtype=np.float64).ravel()
    idx = digest_indices(a.size, salt)
    return np.concatenate([[a.sum(), np.sqrt((a * a).sum())], a[idx]])


def rel_err(a, b):
    """max |a-b| / max |b| -- the scale-aware error used for 'rtol' on whole
    tensors (an elementwise rtol is meaningless for entries near zero)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.abs(b).max()
    if den == 0:
        return np.abs(a).max()
    return np.abs(a - b).max() / den


def l2_err(a, b):
    a = np.asarray(a, dtype=np.float64).ravel()
    b = np.asarray(b, dtype=np.float64).ravel()
    den = np.sqrt((b * b).sum())
    return np.sqrt(((a - b) ** 2).sum()) / den if den > 0 else np.sqrt((a * a).sum())


def check_against_golden(g, prefix, name, value, tol, kink_factor=10.0):
    """Compare `value` with golden entry prefix+name (full or digest).

    Pass if the max-norm relative error is <= tol.  Gradients of a ReLU network are
    discontinuous where a pre-activation crosses zero: a 1e-7 forward rounding difference
    can flip one unit's mask (observed: 1 of 917,504 units of convt6 at B=7), which puts an
    isolated spike into the few gradient entries that unit feeds.  Such entries are allowed
    up to kink_factor*tol as long as the tensor as a whole (relative L2 error) is within tol."""
    key = prefix + name
    if key in g.files:
        ref = g[key]
        got = np.asarray(value).reshape(ref.shape)
        err, err2 = rel_err(got, ref), l2_err(got, ref)
    else:
        ref = g[key + "__digest"]
        got = digest(value)
        # sum is ill-conditioned; compare L2 norm + samples.
        nrm = abs(got[1] - ref[1]) / max(ref[1], 1e-30)
        err = max(nrm, rel_err(got[2:], ref[2:]))
        err2 = max(nrm, l2_err(got[2:], ref[2:]))
    ok = err <= tol or (err2 <= tol and err <= kink_factor * tol)
    assert ok, "%s: max-norm rel err %.3e, L2 rel err %.3e > %.1e" % (key, err, err2, tol)
    return err

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


def check_against_golden(g, prefix, name, value, tol):
    """Compare `value` with golden entry prefix+name (full or digest)."""
    key = prefix + name
    if key in g.files:
        ref = g[key]
        err = rel_err(np.asarray(value).reshape(ref.shape), ref)
    else:
        ref = g[key + "__digest"]
        got = digest(value)
        # sum is ill-conditioned; compare L2 norm + samples.
        err = max(abs(got[1] - ref[1]) / max(ref[1], 1e-30),
                  rel_err(got[2:], ref[2:]))
    assert err <= tol, "%s: rel err %.3e > %.1e" % (key, err, tol)
    return err

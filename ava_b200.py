"""Importable alias for the hyphenated package directory `autoencoded-vocal-analysis_b200`.

    import ava_b200
    from ava_b200.models.vae import VAE
"""
import importlib
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
_pkg = importlib.import_module("autoencoded-vocal-analysis_b200")
sys.modules[__name__] = _pkg
for _sub in ("models", "preprocessing", "plotting", "data"):
    try:
        sys.modules[__name__ + "." + _sub] = importlib.import_module(_pkg.__name__ + "." + _sub)
    except ImportError:
        pass

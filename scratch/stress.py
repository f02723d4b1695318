"""Repeat the per-layer backward checks at an awkward batch size to catch races."""
import sys, importlib
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import tests.test_gpu_kernels as T
L = importlib.import_module("autoencoded-vocal-analysis_b200._lib"); L.lib()
bad = 0
for rep in range(int(sys.argv[1])):
    for l in range(14):
        for nb in (True, False):
            try:
                T.test_bnconv_bwd(L, l, nb)
            except AssertionError as e:
                bad += 1; print("FAIL rep", rep, "layer", l, nb, str(e)[:80])
        for tr in (True, False):
            try:
                T.test_bnconv_fwd(L, l, tr)
            except AssertionError as e:
                bad += 1; print("FAIL fwd rep", rep, "layer", l, tr, str(e)[:80])
print("failures:", bad)

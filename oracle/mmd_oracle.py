"""
TEST INFRASTRUCTURE ONLY -- never imported by the product path.

numpy restatement of the reference's MMD^2 estimators (ava/plotting/mmd_plots.py), pinned
against outputs of the reference itself (tests/golden/mmd_cases.npz, oracle/make_golden.py).
"""
import numpy as np

EPSILON = 1e-8   # mmd_plots.py:34


def estimate_median_sigma(latent, n=10000, seed=42):
    """mmd_plots.py:450-476."""
    np.random.seed(seed)
    arr = np.zeros(n)
    for i in range(n):
        i1, i2 = np.random.randint(len(latent)), np.random.randint(len(latent))
        arr[i] = np.sum(np.power(latent[i1] - latent[i2], 2))
    np.random.seed(None)
    return np.sqrt(np.median(arr) + EPSILON)


def _gram_sum(x, y, A):
    d2 = ((x[:, None, :] - y[None, :, :]) ** 2).sum(-1)
    return np.exp(A * d2)


def estimate_mmd2(latent, i1, i2, sigma, max_n=None, seed=None):
    """mmd_plots.py:255-295 (the three double loops, vectorised)."""
    A = -0.5 / (sigma ** 2)
    n1, n2 = len(i1), len(i2)
    if max_n is not None:
        np.random.seed(seed)
        n1, n2 = min(max_n, n1), min(max_n, n2)
        if n1 < len(i1):
            np.random.shuffle(i1)
            i1 = i1[:n1]
        if n2 < len(i2):
            np.random.shuffle(i2)
            i2 = i2[:n2]
        np.random.seed(None)
    x, y = latent[i1], latent[i2]
    k11, k22, k12 = _gram_sum(x, x, A), _gram_sum(y, y, A), _gram_sum(x, y, A)
    term_1 = np.triu(k11, 1).sum() * 2 / (n1 * (n1 - 1))
    term_2 = np.triu(k22, 1).sum() * 2 / (n2 * (n2 - 1))
    term_3 = k12.sum() * 2 / (n1 * n2)
    return term_1 + term_2 - term_3


def estimate_mmd2_linear_time(latent, i1, i2, sigma):
    """mmd_plots.py:298-312."""
    A = -0.5 / (sigma ** 2)
    n = min(len(i1), len(i2))
    m = n // 2
    k = lambda x, y: np.exp(A * np.sum(np.power(x - y, 2)))          # noqa: E731
    h = lambda x1, y1, x2, y2: k(x1, x2) + k(y1, y2) - k(x1, y2) - k(x2, y1)   # noqa: E731
    term = 0.0
    for i in range(m):
        term += h(latent[i1[2 * i]], latent[i2[2 * i]], latent[i1[2 * i + 1]], latent[i2[2 * i + 1]])
    return term / m

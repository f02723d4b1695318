"""
TEST INFRASTRUCTURE ONLY -- never imported by the product path.

CPU restatement (numpy, float64) of what the reference computes in
``DataContainer._make_latent_mean_pca_projection`` (ava/data/data_container.py:538-551):

    PCA(n_components=2, copy=False, random_state=42).fit_transform(latent_means)

The arithmetic lives in a third-party dependency that is not vendored in the reference:
**scikit-learn** (``setup.py:30`` ``scikit-learn``, unpinned; installed here: 1.9.0),
``sklearn/decomposition/_pca.py``: ``_fit`` picks the solver (``'full'`` when
max(N, D) <= 500, ``'covariance_eigh'`` when D <= 1000 and N >= 10*D, :509-523),
``_fit_full`` (:548-640) centres, decomposes, orders the eigenvalues descending, clips
negative ones to 0 and fixes the signs with ``svd_flip(U, Vt, u_based_decision=False)``
(``sklearn/utils/extmath.py``: the entry of largest magnitude in each row of Vt is made
positive); ``fit_transform`` returns ``(X - mean) @ components_.T``.  Both solver routes
compute the same quantities; this restatement follows the covariance route.

Pinned against scikit-learn itself run in this container: ``tests/golden/pca_cases.npz``
(``oracle/make_golden.py pca``), see ``tests/test_oracle_golden.py``.
"""
import numpy as np


def pca_fit(x, n_components=2):
    """x: [N,D].  Returns dict(mean, components [K,D], explained_variance [K],
    explained_variance_ratio [K], covariance [D,D])."""
    x = np.asarray(x, dtype=np.float64)
    n = x.shape[0]
    mean = x.mean(axis=0)
    xc = x - mean
    cov = xc.T @ xc / (n - 1)
    evals, evecs = np.linalg.eigh(cov)
    evals = np.clip(evals[::-1], 0.0, None)
    vt = evecs[:, ::-1].T.copy()
    # svd_flip(u_based_decision=False)
    idx = np.argmax(np.abs(vt), axis=1)
    signs = np.sign(vt[np.arange(vt.shape[0]), idx])
    vt *= signs[:, None]
    k = n_components
    return {
        "mean": mean,
        "covariance": cov,
        "components": vt[:k],
        "explained_variance": evals[:k],
        "explained_variance_ratio": evals[:k] / evals.sum(),
    }


def pca_fit_transform(x, n_components=2):
    """[N,D] -> ([N,K] embedding, fit dict)."""
    fit = pca_fit(x, n_components)
    x = np.asarray(x, dtype=np.float64)
    return (x - fit["mean"]) @ fit["components"].T, fit


def synth_latents(seed, n, d=32):
    """Anisotropic, correlated, off-centre Gaussian cloud (distinct leading eigenvalues)."""
    rng = np.random.default_rng(seed)
    scales = np.linspace(3.0, 0.2, d) * (1.0 + 0.05 * rng.standard_normal(d))
    q, _ = np.linalg.qr(rng.standard_normal((d, d)))
    return (rng.standard_normal((n, d)) * scales) @ q.T + rng.standard_normal(d) * 0.7

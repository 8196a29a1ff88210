"""PCA oracle.  TEST INFRASTRUCTURE.

The reference fits ``sklearn.decomposition.IncrementalPCA(n_components=2, batch_size=500)`` on the
latent matrix and projects with ``transform`` (/root/reference/scratchpad/IEEE_ICIP_paper/
latent_vis.py:94-143).  scikit-learn is a third-party dependency of the reference (unpinned in its
setup.py); scikit-learn 1.9.0 is importable in this image, so ``incremental_pca_reference`` runs the
literal call.  IncrementalPCA keeping only 2 components between batches is a lossy streaming
approximation of PCA; BASELINE.json.north_star specifies covariance accumulation + eigensolve, i.e.
exact PCA, which ``exact_pca`` restates (float64 covariance, ``eigh``, sklearn's deterministic sign
rule ``svd_flip(u_based_decision=False)``: the largest-|v| entry of each component is positive).
"""
import numpy as np


def exact_pca(h, n_components=2):
    """Returns (mean (d,), components (k,d), explained_variance (k,), z (n,k)) in float64."""
    h = np.asarray(h, dtype=np.float64)
    n = h.shape[0]
    mean = h.mean(axis=0)
    hc = h - mean
    cov = hc.T @ hc / (n - 1)
    evals, evecs = np.linalg.eigh(cov)
    order = np.argsort(evals)[::-1][:n_components]
    comps = evecs[:, order].T.copy()
    for i in range(comps.shape[0]):
        j = np.argmax(np.abs(comps[i]))
        if comps[i, j] < 0:
            comps[i] = -comps[i]
    return mean, comps, evals[order], hc @ comps.T


def pca_from_moments(count, s1, s2, n_components=2):
    """Same PCA from the streaming moments the GPU path accumulates: count, sum h, sum h h^T."""
    s1 = np.asarray(s1, dtype=np.float64)
    s2 = np.asarray(s2, dtype=np.float64)
    mean = s1 / count
    cov = (s2 - count * np.outer(mean, mean)) / (count - 1)
    evals, evecs = np.linalg.eigh(cov)
    order = np.argsort(evals)[::-1][:n_components]
    comps = evecs[:, order].T.copy()
    for i in range(comps.shape[0]):
        j = np.argmax(np.abs(comps[i]))
        if comps[i, j] < 0:
            comps[i] = -comps[i]
    return mean, comps, evals[order]


def incremental_pca_reference(h, n_components=2, batch_size=500):
    """latent_vis.py:107-109,133 verbatim: fit then transform."""
    from sklearn.decomposition import IncrementalPCA
    pca = IncrementalPCA(n_components=n_components, batch_size=batch_size)
    pca.fit(h)
    return pca, pca.transform(h)


def rescale_z(z):
    """latent_vis.py:146-154: per-component min-max to [-5, 5]."""
    z = np.array(z, dtype=np.float64, copy=True)
    for i in range(z.shape[1]):
        z[:, i] = ((z[:, i] - z[:, i].min()) / (z[:, i].max() - z[:, i].min()) - 0.5) * 10.0
    return z

"""Latent-space PCA on the GPU.

Replaces ``fit_PCA`` / ``transform_PCA`` / ``rescale_z`` of
/root/reference/scratchpad/IEEE_ICIP_paper/latent_vis.py:94-154 (sklearn ``IncrementalPCA(2,
batch_size=500).fit`` + ``.transform``).  ``LatentPCA`` keeps sklearn's attribute names
(``mean_``, ``components_``, ``explained_variance_``, ``n_components``) and its sign convention, but
computes the EXACT principal components from streamed float64 moments (count, sum h, sum h h^T)
accumulated by a GPU kernel, so that the fit is a single pass, order independent and, with several
ranks, one all-reduce of d*d + d + 1 doubles (BASELINE.json north_star).  IncrementalPCA with 2
retained components is itself a lossy streaming approximation of this.
"""
import numpy as np
import torch

from .. import _device, _lib, sharding


class LatentPCA:
    def __init__(self, n_components=2, batch_size=None, process_group=None):
        self.n_components = int(n_components)
        self.batch_size = batch_size          # accepted for signature compatibility; unused
        self.process_group = process_group    # torch.distributed group (None = default group if initialised)
        self._moments = None
        self._d = None
        self.mean_ = self.components_ = self.explained_variance_ = None
        self._d_mean = self._d_comps = None

    # ---- streaming fit
    def partial_fit(self, h):
        """Accumulates the moments of a batch of latent vectors h (n,d) (numpy or torch CUDA)."""
        t = _device.to_device(h)
        if t.dtype != torch.float32:
            t = t.float()
        t = t.contiguous()
        n, d = t.shape
        if self._moments is None:
            if d > 256:
                raise ValueError("latent dimension must be <= 256")
            self._d = d
            self._moments = torch.zeros(1 + d + d * d, dtype=torch.float64, device=t.device)
        elif d != self._d:
            raise ValueError("latent dimension changed between batches")
        _lib.check(_lib.lib().te_pca_accumulate(t.data_ptr(), n, d, self._moments.data_ptr(), _device.stream_ptr()),
                   "te_pca_accumulate")
        return self

    def finalize(self, distributed=None):
        """All-reduces the moments across ranks (when torch.distributed is initialised) and solves
        the d x d eigenproblem on the GPU.  Every rank ends with identical components."""
        if self._moments is None:
            raise ValueError("no data: call partial_fit first")
        if distributed is None:
            distributed = sharding.world(self.process_group)[1] > 1
        m = self._moments
        if distributed:
            m = sharding.allreduce_sum_(m.clone(), self.process_group)
        d, k = self._d, self.n_components
        dev = m.device
        self._d_mean = torch.empty(d, dtype=torch.float64, device=dev)
        self._d_comps = torch.empty((k, d), dtype=torch.float64, device=dev)
        evar = torch.empty(k, dtype=torch.float64, device=dev)
        work = torch.empty(2 * d * d, dtype=torch.float64, device=dev)
        _lib.check(_lib.lib().te_pca_solve(m.data_ptr(), d, k, self._d_mean.data_ptr(), self._d_comps.data_ptr(),
                                           evar.data_ptr(), work.data_ptr(), _device.stream_ptr()), "te_pca_solve")
        self.n_samples_seen_ = int(round(float(m[0].item())))
        self.mean_ = self._d_mean.cpu().numpy()
        self.components_ = self._d_comps.cpu().numpy()
        self.explained_variance_ = evar.cpu().numpy()
        return self

    def fit(self, h, y=None):
        self._moments = None
        self.partial_fit(h)
        return self.finalize()

    def transform(self, h):
        """(h - mean_) @ components_.T -> (n, n_components) float32, where h lives."""
        if self._d_comps is None:
            raise ValueError("PCA is not fitted")
        t = _device.to_device(h)
        if t.dtype != torch.float32:
            t = t.float()
        t = t.contiguous()
        n, d = t.shape
        if d != self._d:
            raise ValueError("latent dimension mismatch")
        z = torch.empty((n, self.n_components), dtype=torch.float32, device=t.device)
        _lib.check(_lib.lib().te_pca_project(t.data_ptr(), n, d, self._d_mean.data_ptr(), self._d_comps.data_ptr(),
                                             self.n_components, z.data_ptr(), _device.stream_ptr()), "te_pca_project")
        return _device.to_host_like(z, h)

    def fit_transform(self, h, y=None):
        return self.fit(h).transform(h)

    def transform_all_gather(self, h):
        """Projects this rank's latents and all-gathers the (n_local, k) projections (equal n per rank)."""
        z = self.transform(_device.to_device(h))
        return sharding.allgather_rows(z, self.process_group)


def _latent_columns(N):
    return ["$h_%i$" % i for i in range(N)]


def fit_PCA(dfN, N, ncomps=2, transform=False):
    """latent_vis.py:94-116: dfN is a DataFrame with columns $h_0$..$h_{N-1}$."""
    pca = LatentPCA(n_components=ncomps, batch_size=500)
    h = np.asarray(dfN[_latent_columns(N)], dtype=np.float32)
    pca.fit(h)
    if transform:
        return pca, transform_PCA(dfN, N, pca, ncomps=ncomps)
    return pca


def transform_PCA(dfN, N, pca, ncomps=2):
    """latent_vis.py:118-143."""
    import pandas as pd
    h = np.asarray(dfN[_latent_columns(N)], dtype=np.float32)
    z = pca.transform(h)
    df = pd.DataFrame(data=z, columns=["$z_%i$" % ii for ii in range(ncomps)])
    if "label" in dfN.columns:
        for col in ("label", "measurement", "shape", "param"):
            if col in dfN.columns:
                df[col] = dfN[col].values
    elif "label_idx" in dfN.columns:
        df["label_idx"] = dfN["label_idx"].values
    return df


def rescale_z(df):
    """latent_vis.py:146-154: per-component min-max to [-5, 5]."""
    z = np.asarray(df[["$z_0$", "$z_1$"]].copy(), dtype=np.float64)
    for ii in range(2):
        z[:, ii] = ((z[:, ii] - z[:, ii].min()) / (z[:, ii].max() - z[:, ii].min()) - 0.5) * 10.0
    df[["$z_0$", "$z_1$"]] = z
    return df

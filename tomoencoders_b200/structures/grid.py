"""``Grid``: a table of equal-width cubic patches on a regular lattice, with GPU gather / scatter.

Drop-in for ``tomo_encoders.structures.grid.Grid`` (/root/reference/tomo_encoders/structures/
grid.py:26-357): same constructor, dict keys (``points``, ``width``, ``source``), attributes
(``points``, ``wd``, ``vol_shape``) and methods.  ``extract`` / ``fill_patches_in_volume`` run on the
GPU (libtomoenc.so); the triple Python loop that builds the lattice is vectorised.  ``copy`` works
(the reference's passes a wrong keyword, grid.py:222-225).
"""
import numpy as np
from numpy.random import default_rng

from . import _table


class Grid(dict):

    def __init__(self, vol_shape, initialize_by="grid-params", xp=np, **kwargs):
        self.vol_shape = tuple(int(v) for v in vol_shape)
        self.xp = np
        initializers = {"data": self._set_from_data,
                        "grid-params": self._set_regular_grid,
                        "file": self._load_from_disk}
        if initialize_by not in initializers:
            raise NotImplementedError("unknown initialize_by = %s" % str(initialize_by))
        self["source"] = initialize_by
        self.points, self.wd = initializers[initialize_by](**kwargs)
        self["points"] = self.points
        self["width"] = self.wd

    def __len__(self):
        return len(self.points)

    def dump(self, fpath):
        """grid.py:55-62; datasets vol_shape, points, width."""
        data = {"vol_shape": np.asarray(self.vol_shape), "points": self.points, "width": np.asarray(self.wd)}
        if str(fpath).endswith(".npz"):
            np.savez(fpath, **data)
            return
        from .hdf5_lite import h5_module           # h5py when installed, else the format subset in hdf5_lite.py
        with h5_module().File(fpath, "w") as hf:
            for k, v in data.items():
                hf.create_dataset(k, data=v)

    def _set_regular_grid(self, width=None, n_points=None):
        """grid.py:64-105: corners of the non-overlapping lattice, z outermost, x innermost."""
        nd = len(self.vol_shape)
        assert not any(self.vol_shape[i] % width > 0 for i in range(nd)), "vol_shape must be multiple of patch width"
        axes = [np.arange(self.vol_shape[i] // width, dtype=np.int64) * int(width) for i in range(nd)]
        mesh = np.meshgrid(*axes, indexing="ij")
        points = np.stack([m.reshape(-1) for m in mesh], axis=1)
        if n_points is not None:
            n_points = min(n_points, len(points))
            idxs = np.sort(default_rng().choice(points.shape[0], n_points, replace=False))
            points = points[idxs].copy()
        return points.astype(np.uint32), int(width)

    def _load_from_disk(self, fpath=None):
        if str(fpath).endswith(".npz"):
            hf = dict(np.load(fpath))
        else:
            from .hdf5_lite import h5_module
            with h5_module().File(fpath, "r") as f:
                hf = {k: np.asarray(f[k]) for k in f.keys()}
        self.vol_shape = tuple(int(v) for v in hf["vol_shape"])
        return np.asarray(hf["points"]).astype(np.uint32), int(hf["width"])

    def _set_from_data(self, points=None, width=None):
        points = np.asarray(points)
        p = points.astype(np.int64)
        if points.ndim != 2 or points.shape[1] != len(self.vol_shape):
            raise ValueError("points must have shape (n, %d)" % len(self.vol_shape))
        if np.any(p < 0) or np.any(p + int(width) > np.asarray(self.vol_shape, dtype=np.int64)[None, :]):
            raise ValueError("Some points are invalid")
        return points.astype(np.uint32), int(width)

    def append(self, more_patches):
        if self.vol_shape != more_patches.vol_shape:
            raise ValueError("patches data is not compatible. Ensure that big volume shapes match")
        assert self.wd == more_patches.wd, "this is a Grid data structure. All widths must be equal"
        self.points = np.concatenate([self.points, more_patches.points], axis=0)
        self["points"] = self.points

    def slices(self):
        n, nd = self.points.shape
        a = self.points.astype(np.int64)
        out = np.empty((n, nd), dtype=object)
        for ii in range(n):
            for jj in range(nd):
                out[ii, jj] = slice(int(a[ii, jj]), int(a[ii, jj]) + self.wd)
        return out

    def centers(self):
        return self.points.astype(np.int64) + self.wd // 2

    def _is_within_cylindrical_crop(self, mask_ratio, height_ratio):
        """grid.py:171-189."""
        assert self.vol_shape[1] == self.vol_shape[2], "must be tomographic CT volume (ny = nx = n)"
        nz, n = self.vol_shape[:2]
        centers = self.centers()
        radii = np.sqrt(np.power(centers[:, 1] - n / 2.0, 2) + np.power(centers[:, 2] - n / 2.0, 2))
        heights = np.abs(centers[:, 0] - nz / 2.0)
        return (radii < mask_ratio * n / 2.0) & (heights < height_ratio * nz / 2.0)

    def filter_by_cylindrical_mask(self, mask_ratio=0.9, height_ratio=1.0):
        return self.filter_by_condition(self._is_within_cylindrical_crop(mask_ratio, height_ratio))

    def _derive(self, sel):
        return Grid(self.vol_shape, initialize_by="data", points=self.points[sel].copy(), width=self.wd)

    def filter_by_condition(self, cond_list):
        """grid.py:200-220."""
        cond_list = np.asarray(cond_list)
        if cond_list.shape[0] != len(self.points):
            raise ValueError("length of condition list must same as the current number of stored points")
        if cond_list.ndim == 2:
            cond_list = np.prod(cond_list, axis=1)
        elif cond_list.ndim > 2:
            raise ValueError("condition list must have 1 or 2 dimensions like so (n_pts,) or (n_pts, n_conditions)")
        return self._derive(np.nonzero(cond_list)[0])

    def copy(self):
        return self._derive(slice(None))

    def select_by_range(self, s_sel):
        return self._derive(slice(s_sel[0], s_sel[1], None))

    def pop(self, n_pop):
        if n_pop > 0:
            return self._derive(slice(n_pop, None, None))
        if n_pop < 0:
            return self._derive(slice(None, n_pop, None))
        return self.copy()

    def rescale(self, fac):
        """grid.py:260-268: scales vol_shape, points and width."""
        fac = int(fac)
        new_vol_shape = tuple(int(v * fac) for v in self.vol_shape)
        return Grid(new_vol_shape, initialize_by="data", points=self.points.astype(np.int64) * fac,
                    width=int(self.wd * fac))

    def select_by_indices(self, idxs):
        return self._derive(np.asarray(idxs))

    def select_random_sample(self, n_points):
        idxs = np.sort(default_rng().choice(self.points.shape[0], n_points, replace=False))
        return self.select_by_indices(idxs)

    def sort_by(self, feature):
        feature = np.asarray(feature)
        assert feature.ndim == 1, "feature must be 1D array"
        assert len(feature) == len(self.points), "length mismatch"
        return self.select_by_indices(np.argsort(feature))

    def _widths(self):
        return np.full(self.points.shape, self.wd, dtype=np.uint32)

    def extract(self, vol):
        """grid.py:313-336: (n, wd, wd, wd) with vol's dtype, gathered on the GPU."""
        assert tuple(vol.shape) == tuple(self.vol_shape), \
            "Shape of big volume does not match vol_shape attribute of patches data"
        nd = len(self.vol_shape)
        return _table.gather(vol, self.vol_shape, self.points, self._widths(), (self.wd,) * nd)

    def fill_patches_in_volume(self, sub_vols, vol_out):
        """grid.py:338-348: vol_out[slices_i] = sub_vols[i] in place, on the GPU."""
        _table.scatter(sub_vols, vol_out, self.vol_shape, self.points, self._widths())

"""``Patches``: a table of cuboid patches inside a big volume, with GPU gather / scatter.

Drop-in for ``tomo_encoders.structures.patches.Patches`` (/root/reference/tomo_encoders/structures/
patches.py:28-855): same constructor, same dict keys / attributes, same initialisers, same method
names and error behaviour.  Differences, all deliberate:

* ``extract`` / ``fill_patches_in_volume`` run on the GPU through libtomoenc.so (no CPU fallback);
  a torch CUDA tensor plays the role the reference gives to a cupy array (device in -> device out).
* the O(n) Python loops of the reference (grid construction, validity check, slices) are vectorised
  so that n = 10^6 patches is practical.
* coordinates are validated in int64, so negative inputs raise instead of wrapping through uint32.
* ``dump`` / ``initialize_by="file"`` write / read the reference's HDF5 schema through h5py when it is importable and
  through ``structures/hdf5_lite.py`` (a dependency-free subset of the file format) otherwise; ``.npz`` paths also work.
"""
import numpy as np
from numpy.random import default_rng

from . import _table


class Patches(dict):

    def __init__(self, vol_shape, initialize_by="data", features=None, names=[], xp=np, **kwargs):
        self.vol_shape = tuple(int(v) for v in vol_shape)
        self.xp = np  # coordinate tables always live on the host
        initializers = {"data": self._check_data,
                        "slices": self._from_slices,
                        "grid": self._set_grid,
                        "regular-grid": self._set_regular_grid,
                        "octree-grid": self._set_regular_grid,
                        "multiple-grids": self._set_multiple_grids,
                        "random-fixed-width": self._get_random_fixed_width,
                        "random": self._get_random}
        self["source"] = initialize_by
        self.features = None
        self.feature_names = []
        if initialize_by == "file":
            self._load_from_disk(**kwargs)
            self.check_valid = None
        elif initialize_by in initializers:
            self.points, self.widths, self.check_valid = initializers[initialize_by](**kwargs)
        else:
            raise NotImplementedError("unknown initialize_by = %s" % str(initialize_by))
        self._set_tables(self.points, self.widths)
        if initialize_by == "file":
            feats, fnames = self.features, self.feature_names
            self.features, self.feature_names = None, []
            self.add_features(feats, fnames)
        else:
            self.add_features(features, names)

    # ------------------------------------------------------------------ construction helpers
    def _set_tables(self, points, widths):
        p = np.asarray(points)
        w = np.asarray(widths)
        if p.ndim != 2 or p.shape != w.shape or p.shape[1] != len(self.vol_shape):
            raise ValueError("dimension mismatch for points and corresponding widths")
        self._validate(p.astype(np.int64), w.astype(np.int64))
        self.points = p.astype(np.uint32)
        self.widths = w.astype(np.uint32)
        self["points"] = self.points
        self["widths"] = self.widths

    def _validate(self, p, w):
        """patches.py:416-428 (_check_valid_points), vectorised and evaluated in int64."""
        bad = (p < 0) | (p + w > np.asarray(self.vol_shape, dtype=np.int64)[None, :])
        if bad.any():
            ii = int(np.nonzero(bad.any(axis=1))[0][0])
            print("Patch %i, %s, %s is invalid" % (ii, str(p[ii]), str(w[ii])))
            raise ValueError("Some points are invalid")

    def _check_valid_points(self):
        self._validate(self.points.astype(np.int64), self.widths.astype(np.int64))

    def __len__(self):
        return len(self.points)

    # ------------------------------------------------------------------ disk
    def dump(self, fpath):
        """patches.py:81-91: HDF5 file with the datasets vol_shape, points, widths, features, feature_names (h5py when it is
        installed, else the format subset in structures/hdf5_lite.py); a path ending in .npz writes a numpy archive of the
        same arrays instead."""
        data = {"vol_shape": np.asarray(self.vol_shape), "points": self.points, "widths": self.widths}
        if self.features is not None:
            data["features"] = self.features
        if any(self.feature_names):
            data["feature_names"] = np.asarray(self.feature_names, dtype="S")
        if str(fpath).endswith(".npz"):
            np.savez(fpath, **data)
            return
        from .hdf5_lite import h5_module
        with h5_module().File(fpath, "w") as hf:
            for k, v in data.items():
                hf.create_dataset(k, data=v)

    def _load_from_disk(self, fpath=None):
        """patches.py:94-110 (initialize_by = "file")."""
        if str(fpath).endswith(".npz"):
            hf = dict(np.load(fpath))
        else:
            from .hdf5_lite import h5_module
            with h5_module().File(fpath, "r") as f:
                hf = {k: np.asarray(f[k]) for k in f.keys()}
        self.vol_shape = tuple(int(v) for v in hf["vol_shape"])
        self.points = np.asarray(hf["points"])
        self.widths = np.asarray(hf["widths"])
        self.features = np.asarray(hf["features"]) if "features" in hf else None
        self.feature_names = [s.decode("UTF-8") for s in hf["feature_names"]] if "feature_names" in hf else []

    # ------------------------------------------------------------------ features
    def add_features(self, features, names=[]):
        """patches.py:112-151."""
        if features is None:
            return
        features = np.asarray(features)
        if features.ndim == 1:
            features = features.reshape(-1, 1)
        if len(self.points) != len(features):
            raise ValueError("numbers of anchor points and corresponding features must match")
        if features.ndim != 2:
            raise ValueError("shape of features array not valid")
        if self.features is not None:
            self.features = np.concatenate([self.features, features], axis=1)
        else:
            self.features = features
        names = list(names) if names is not None else []
        if not any(names):
            names = ["Unnamed_f%i" % i for i in range(features.shape[-1])]
        else:
            assert len(names) == features.shape[-1], "names of features must match array shape"
        self.feature_names = self.feature_names + names
        for ii, key in enumerate(self.feature_names):
            self[key] = self.features[:, ii]

    def append(self, more_patches):
        """patches.py:153-186 (in place)."""
        if self.vol_shape != more_patches.vol_shape:
            raise ValueError("patches data is not compatible. Ensure that big volume shapes match")
        self.points = np.concatenate([self.points, more_patches.points], axis=0)
        self.widths = np.concatenate([self.widths, more_patches.widths], axis=0)
        self["points"], self["widths"] = self.points, self.widths
        if self.features is None:
            return
        self.features = np.concatenate([self.features, more_patches.features], axis=0)
        if self.feature_names != more_patches.feature_names:
            raise ValueError("feature names in self do not match input")
        for ii, key in enumerate(self.feature_names):
            self[key] = self.features[:, ii]

    # ------------------------------------------------------------------ initialisers
    def _check_data(self, points=None, widths=None, check_valid=None):
        if len(points) != len(widths):
            raise ValueError("number of anchor points and corresponding widths must match")
        if np.shape(points)[-1] != np.shape(widths)[-1]:
            raise ValueError("dimension mismatch for points and corresponding widths")
        return np.asarray(points), np.asarray(widths), check_valid

    def _from_slices(self, s=None, check_valid=None):
        nd = len(s[0])
        points = np.asarray([[s[i][a].start for a in range(nd)] for i in range(len(s))]).astype(int)
        widths = np.asarray([[s[i][a].stop - s[i][a].start for a in range(nd)] for i in range(len(s))]).astype(int)
        return points, widths, check_valid

    def _check_stride(self, patch_size, stride):
        """patches.py:216-232: widths = patch_size * stride, limited by the volume."""
        nd = len(self.vol_shape)
        if stride is None:
            return patch_size
        max_possible_stride = min(self.vol_shape[i] // patch_size[i] for i in range(nd))
        if stride > max_possible_stride:
            raise ValueError("Cannot preserve aspect ratio with given combination of patch_size and stride. "
                             "Try lower values.")
        return tuple(int(patch_size[i] * stride) for i in range(nd))

    @staticmethod
    def _lattice(nsteps, stepsize):
        axes = [np.arange(int(n), dtype=np.int64) * int(s) for n, s in zip(nsteps, stepsize)]
        mesh = np.meshgrid(*axes, indexing="ij")          # first axis outermost, last innermost
        return np.stack([m.reshape(-1) for m in mesh], axis=1)

    @staticmethod
    def _subsample(points, widths, n_points):
        if n_points is None:
            return points, widths
        n_points = min(n_points, len(points))
        idxs = np.sort(default_rng().choice(points.shape[0], n_points, replace=False))
        return points[idxs].copy(), widths[idxs].copy()

    def _set_grid(self, patch_size=None, stride=1, n_points=None):
        """patches.py:261-313: overlapping grid that covers the full volume."""
        patch_size = self._check_stride(patch_size, stride)
        m, p = list(self.vol_shape), list(patch_size)
        nsteps = [int(np.ceil(m[i] / p[i])) for i in range(len(m))]
        stepsize = [((m[i] - p[i]) // (nsteps[i] - 1)) if m[i] != p[i] else 0 for i in range(len(m))]
        points = self._lattice(nsteps, stepsize)
        widths = np.tile(np.asarray(p, dtype=np.int64), (len(points), 1))
        points, widths = self._subsample(points, widths, n_points)
        return points, widths, False

    def _set_regular_grid(self, patch_size=None, n_points=None):
        """patches.py:315-360: non-overlapping grid, volume cropped to a multiple of the patch."""
        patch_size = self._check_stride(patch_size, 1)
        m, p = list(self.vol_shape), list(patch_size)
        nsteps = [int(m[i] // p[i]) for i in range(len(m))]
        points = self._lattice(nsteps, p)
        widths = np.tile(np.asarray(p, dtype=np.int64), (len(points), 1))
        points, widths = self._subsample(points, widths, n_points)
        return points, widths, False

    def _set_multiple_grids(self, min_patch_size=None, max_stride=None, n_points=None):
        """patches.py:234-259: grids for stride 1..max_stride concatenated."""
        points, widths = [], []
        for stride in range(1, max_stride + 1):
            p, w, _ = self._set_grid(patch_size=min_patch_size, stride=stride)
            points.append(p)
            widths.append(w)
        points = np.concatenate(points, axis=0)
        widths = np.concatenate(widths, axis=0)
        points, widths = self._subsample(points, widths, n_points)
        return points, widths, False

    def _get_random_fixed_width(self, patch_size=None, n_points=None):
        """patches.py:363-382: corners ~ randint(0, vol - patch) per axis (upper bound exclusive),
        drawn from the legacy global numpy RNG in the same order as the reference."""
        nd = len(self.vol_shape)
        patch_size = self._check_stride(patch_size, 1)
        points = np.asarray([np.random.randint(0, self.vol_shape[i] - patch_size[i], n_points)
                             for i in range(nd)]).T
        widths = np.tile(np.asarray(patch_size, dtype=np.int64), (n_points, 1))
        return points, widths, False

    def _get_random(self, min_patch_size=None, max_stride=None, n_points=None):
        """patches.py:384-413: stride ~ randint(1, max_stride) per patch, width = stride * min size."""
        nd = len(self.vol_shape)
        self._check_stride(min_patch_size, max_stride)
        random_strides = np.random.randint(1, max_stride, n_points)
        points = np.empty((n_points, nd), dtype=np.int64)
        widths = np.empty((n_points, nd), dtype=np.int64)
        for k, stride in enumerate(random_strides):
            cur = self._check_stride(min_patch_size, int(stride))
            points[k] = [np.random.randint(0, self.vol_shape[i] - cur[i]) for i in range(nd)]
            widths[k] = cur
        return points, widths, False

    # ------------------------------------------------------------------ coordinate views
    def slices(self, binning=None):
        """patches.py:436-453: (n, ndim) object array of python slices."""
        n, nd = self.points.shape
        if binning is None:
            binning = np.ones(n, dtype=np.int64)
        elif isinstance(binning, (int, np.integer)):
            binning = np.full(n, int(binning), dtype=np.int64)
        a = self.points.astype(np.int64)
        w = self.widths.astype(np.int64)
        out = np.empty((n, nd), dtype=object)
        for ii in range(n):
            for jj in range(nd):
                out[ii, jj] = slice(int(a[ii, jj]), int(a[ii, jj] + w[ii, jj]), int(binning[ii]))
        return out

    def centers(self):
        """patches.py:455-466."""
        return (self.points.astype(np.int64) + self.widths.astype(np.int64) // 2)

    def features_to_numpy(self, names):
        if not any(self.feature_names):
            raise ValueError("feature names must be defined first.")
        return np.asarray([self.features[:, self.feature_names.index(name)] for name in names]).T

    def _is_within_cylindrical_crop(self, mask_ratio, height_ratio):
        """patches.py:481-498."""
        assert self.vol_shape[1] == self.vol_shape[2], "must be tomographic CT volume (ny = nx = n)"
        nz, n = self.vol_shape[:2]
        centers = self.centers()
        radii = np.sqrt(np.power(centers[:, 1] - n / 2.0, 2) + np.power(centers[:, 2] - n / 2.0, 2))
        heights = np.abs(centers[:, 0] - nz / 2.0)
        return (radii < mask_ratio * n / 2.0) & (heights < height_ratio * nz / 2.0)

    def filter_by_cylindrical_mask(self, mask_ratio=0.9, height_ratio=1.0):
        return self.filter_by_condition(self._is_within_cylindrical_crop(mask_ratio, height_ratio))

    def _derive(self, sel, vol_shape=None, points=None, widths=None):
        pts = self.points[sel].copy() if points is None else points
        wds = self.widths[sel].copy() if widths is None else widths
        return Patches(self.vol_shape if vol_shape is None else vol_shape, initialize_by="data",
                       points=pts, widths=wds,
                       features=None if self.features is None else self.features[sel].copy(),
                       names=self.feature_names.copy() if any(self.feature_names) else [])

    def filter_by_condition(self, cond_list):
        """patches.py:510-531: keep rows where ALL conditions hold."""
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
        """patches.py:562-581: n_pop > 0 drops from the beginning, < 0 from the end."""
        if n_pop > 0:
            return self._derive(slice(n_pop, None, None))
        if n_pop < 0:
            return self._derive(slice(None, n_pop, None))
        return self.copy()

    def perturb(self, perturb_size_vox):
        """patches.py:583-603: jitter corners by U[-s, s), clamped into the volume (int64 inside,
        so no uint32 wrap-around)."""
        nd = len(self.vol_shape)
        perturbations = np.random.randint(0, perturb_size_vox * 2, (len(self), nd)) - perturb_size_vox
        wd = self.widths[0, :].astype(np.int64)
        points = self.points.astype(np.int64) + perturbations
        points[points < 0] = 0
        for ia in range(nd):
            hi = self.vol_shape[ia] - wd[ia]
            points[points[:, ia] > hi, ia] = hi
        return self._derive(slice(None), points=points, widths=self.widths.copy())

    def rescale(self, fac, new_vol_shape):
        """patches.py:605-628."""
        fac = int(fac)
        px_max = np.max(self.points.astype(np.int64), axis=0) * fac
        if np.any(np.asarray(new_vol_shape) - px_max < 0):
            raise ValueError("new volume shape is inappropriate")
        return self._derive(slice(None), vol_shape=new_vol_shape, points=self.points.astype(np.int64) * fac,
                            widths=self.widths.astype(np.int64) * fac)

    def select_by_indices(self, idxs):
        return self._derive(np.asarray(idxs))

    def select_random_sample(self, n_points):
        idxs = np.sort(default_rng().choice(self.points.shape[0], n_points, replace=False))
        return self.select_by_indices(idxs)

    def _feature_column(self, feature, ife):
        if feature is None:
            return self.features[:, ife]
        feature = np.asarray(feature)
        if feature.ndim != 1:
            raise ValueError("feature must be 1D array")
        if len(feature) != len(self.points):
            raise ValueError("length of feature array must match number of patch points")
        return feature

    def sort_by_feature(self, feature=None, ife=None):
        return self.select_by_indices(np.argsort(self._feature_column(feature, ife)))

    def select_by_plane(self, plane_axis, plane_idx):
        lo = self.points[:, plane_axis].astype(np.int64)
        hi = lo + self.widths[:, plane_axis].astype(np.int64)
        return self.filter_by_condition(np.stack([plane_idx > lo, plane_idx < hi], axis=1))

    def select_by_feature(self, n_selections, feature=None, ife=None, selection_by="highest"):
        """patches.py:700-732."""
        idxs = np.argsort(self._feature_column(feature, ife))
        n_selections = min(n_selections, len(self.points))
        if selection_by == "highest":
            idxs = idxs[::-1]
        return self.select_by_indices(idxs[:n_selections])

    # ------------------------------------------------------------------ data plane
    def _calc_binning(self, patch_size):
        """patches.py:734-746: isotropic integer binning, >= 1."""
        bin_vals = self.widths // np.asarray(patch_size, dtype=np.uint32)
        if np.sum(np.max(bin_vals, axis=1) != np.min(bin_vals, axis=1)) > 0:
            raise ValueError("aspect ratios of some patches don't match!! Cannot bin to patch_size")
        if np.any(bin_vals == 0):
            raise ValueError("patch_size cannot be larger than any given patch in the list")
        if np.any(self.widths % np.asarray(patch_size, dtype=np.uint32)):
            # the reference computes this condition (patches.py:739) but never acts on it; the
            # resulting ragged rows cannot be stacked, so it is an error here.
            raise ValueError("patch widths must be integer multiples of patch_size")
        return bin_vals[:, 0]

    def extract(self, vol, patch_size):
        """patches.py:748-781.  Returns (n,) + patch_size with vol's dtype, gathered on the GPU:
        out[i] = vol[z:z+wz:b, y:y+wy:b, x:x+wx:b] with b = widths // patch_size.  With
        patch_size=None returns a list of unbinned sub-volumes (one gather per distinct width)."""
        assert tuple(vol.shape) == tuple(self.vol_shape), \
            "Shape of big volume does not match vol_shape attribute of patches data"
        if patch_size is not None:
            self._calc_binning(patch_size)
            return _table.gather(vol, self.vol_shape, self.points, self.widths, tuple(patch_size))
        out = [None] * len(self)
        uniq, inv = np.unique(self.widths, axis=0, return_inverse=True)
        inv = np.asarray(inv).reshape(-1)
        for g in range(len(uniq)):
            idx = np.nonzero(inv == g)[0]
            sub = _table.gather(vol, self.vol_shape, self.points[idx], self.widths[idx], tuple(int(v) for v in uniq[g]))
            for k, i in enumerate(idx):
                out[i] = sub[k]
        return out

    def fill_patches_in_volume(self, sub_vols, vol_out):
        """patches.py:811-821: vol_out[slices_i] = sub_vols[i] in index order (last index wins on
        overlap), in place, on the GPU."""
        _table.scatter(sub_vols, vol_out, self.vol_shape, self.points, self.widths)

    def upsample_patches(self, sub_vols, upsampling_fac):
        raise NotImplementedError("not implemented in the reference either (patches.py:841)")

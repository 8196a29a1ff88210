"""``DataFile``: slice-wise reading and writing of one 3-D dataset in an HDF5 file or a TIFF series.

Drop-in for ``tomo_encoders.structures.datafile.DataFile`` (/root/reference/tomo_encoders/structures/datafile.py:83-508;
SURVEY.md 8f row 5): same constructor arguments, attributes (``d_shape``, ``d_type``, ``chunk_shape``, ``slice_size``,
``d_size_GB``) and methods (``create_new``, ``get_stats``, ``show_stats``, ``est_chunking``, ``read_slice``, ``read_data``,
``read_sequence``, ``read_chunk``, ``read_full``, ``write_full``, ``write_data``, ``write_chunk``), same error behaviour.
It sits on either side of the hot path (the volume that is segmented comes out of one, the mask goes into one); it is
host-side I/O and has no GPU part.

HDF5 goes through h5py when it is installed (the reference's dependency) and through the dependency-free subset in
``hdf5_lite.py`` otherwise; TIFF series through Pillow (the reference uses skimage / tifffile, which are not installed):
one single-page file per z slice, named ``<folder>/<folder name><number>.tif`` as ``write_tiffseq`` (datafile.py:543-583)
names them.
"""
import glob
import os
import shutil

import numpy as np

from .hdf5_lite import h5_module

_TIFF_MODES = {np.dtype(np.uint8): "L", np.dtype(np.uint16): "I;16", np.dtype(np.float32): "F"}


def _message(text, flag):
    if flag:
        print(text)


def read_tiffseq(userfilepath="", procs=None, s=None):
    """datafile.py:510-540: the sorted *.tif / *.tiff files of a folder (optionally a python slice of them) as a 3-D array."""
    from PIL import Image
    if not userfilepath:
        raise ValueError("File path is required.")
    paths = sorted(glob.glob(os.path.join(userfilepath, "*.tif")))
    if not paths:
        paths = sorted(glob.glob(os.path.join(userfilepath, "*.tiff")))
    if s is not None:
        paths = paths[s] if isinstance(s, slice) else [paths[i] for i in s]
    imgs = []
    for p in paths:
        with Image.open(p) as im:
            imgs.append(np.array(im))
    return np.asarray(imgs)


def write_tiffseq(S, SaveDir="", increment_flag=False, suffix_len=None):
    """datafile.py:543-583: one TIFF per slice of S along axis 0 named <folder name><number>.tif, numbers starting at 1 and
    zero-padded to suffix_len; increment_flag continues after the highest number already in the folder, otherwise an
    existing folder is replaced."""
    from PIL import Image
    S = np.asarray(S)
    if S.dtype not in _TIFF_MODES:
        raise ValueError("Data type %s is not supported." % S.dtype)
    if not suffix_len:
        if increment_flag:
            raise ValueError("suffix_len required if increment_flag is True.")
        suffix_len = len(str(S.shape[0]))
    last_num = 0
    if not os.path.exists(SaveDir):
        os.makedirs(SaveDir)
    elif not increment_flag:
        shutil.rmtree(SaveDir)
        os.makedirs(SaveDir)
    else:
        paths = sorted(glob.glob(os.path.join(SaveDir, "*.tif"))) or sorted(glob.glob(os.path.join(SaveDir, "*.tiff")))
        if paths:
            last_num = int(os.path.splitext(paths[-1])[0][-suffix_len:])
    base = os.path.basename(os.path.normpath(SaveDir))
    for i in range(S.shape[0]):
        Image.fromarray(S[i]).save(os.path.join(SaveDir, base + str(i + 1 + last_num).zfill(suffix_len) + ".tif"), format="TIFF")


class DataFile:
    """Points at a 3-D dataset in a TIFF series or an HDF5 file (datafile.py:83-172)."""

    def __init__(self, fname, data_tag=None, tiff=False, chunk_shape=None, chunk_size=None, chunked_slice_size=None,
                 d_shape=None, d_type=None, VERBOSITY=1, autodetect_format=False):
        self.fname = fname
        self.data_tag = data_tag
        self.VERBOSITY = VERBOSITY
        if autodetect_format:
            self.get_format()
        else:
            self.tiff_mode = tiff
        self.exists = os.path.exists(self.fname)
        self.chunked_slice_size = chunked_slice_size
        self.chunk_size = chunk_size
        self.chunk_shape = chunk_shape
        if (d_type is not None) and (d_shape is not None):
            self.d_type = np.dtype(d_type)
            self.d_shape = tuple(int(v) for v in d_shape)
            self.get_slice_sizes()
            self.est_chunking()
            _message("\n" + "#" * 50 + "\nNew dataset will be created in %s: " % ("tiff folder" if self.tiff_mode else "hdf5 file")
                     + os.path.basename(self.fname), self.VERBOSITY > 1)
        elif self.exists:
            self.get_stats()
            self.get_slice_sizes()
        else:
            raise ValueError("Inputs d_shape and d_type are required for new dataset as file does not already exist.\n"
                             "Path does not exist:\n%s" % fname)

    # ------------------------------------------------------------------ format / stats
    def get_format(self):
        """datafile.py:174-188 (the reference compares the extension WITH its dot against names without one, so it never
        recognises an hdf5 file; here '.hdf5' / '.h5' are recognised)."""
        if not os.path.exists(self.fname):
            raise ValueError("Trying to detect file format but path does not exist:\n%s" % self.fname)
        if os.path.isdir(self.fname):
            self.tiff_mode = True
        elif os.path.splitext(self.fname)[-1] in (".hdf5", ".h5", "hdf5", "h5"):
            self.tiff_mode = False
            if self.data_tag is None:
                raise ValueError("dataset-name required for hdf5")
        else:
            raise ValueError("input file type not recognized. must be tiff folder or hdf5 file")

    def set_verbosity(self, VERBOSITY):
        self.VERBOSITY = VERBOSITY

    def create_new(self, overwrite=False):
        """datafile.py:196-228: an empty (zero-filled) dataset of d_shape / d_type / chunk_shape in the hdf5 file; for a TIFF
        folder only the overwrite check."""
        if self.tiff_mode:
            if os.path.exists(self.fname):
                if overwrite:
                    shutil.rmtree(self.fname)
                    _message("Removed old contents in tiff folder %s" % os.path.basename(self.fname), self.VERBOSITY > 0)
                else:
                    raise ValueError("tiff folder already exists and overwrite is not allowed.")
            return
        if self.data_tag is None:
            raise ValueError("data_tag: argument missing for creating new hdf5 file.")
        h5 = h5_module()
        if os.path.exists(self.fname):
            with h5.File(self.fname, "r") as hf:
                present = self.data_tag in hf
            if present:
                if overwrite:
                    os.remove(self.fname)
                    _message("Removed old hdf5 file %s" % os.path.basename(self.fname), self.VERBOSITY > 0)
                else:
                    raise ValueError("hdf5 dataset already exists and overwrite is not allowed.")
        self.est_chunking()
        with h5.File(self.fname, "r+" if os.path.exists(self.fname) else "w") as hf:
            hf.create_dataset(self.data_tag, shape=self.d_shape, dtype=self.d_type, chunks=self.chunk_shape)
        _message("New hdf5 dataset %s created in file %s" % (self.data_tag, os.path.basename(self.fname)), self.VERBOSITY > 0)

    def get_stats(self, return_output=False):
        """datafile.py:230-258."""
        if not self.tiff_mode:
            with h5_module().File(self.fname, "r") as hf:
                ds = hf[self.data_tag]
                self.d_shape = tuple(int(v) for v in ds.shape)
                self.d_type = np.dtype(ds.dtype)
                self.chunk_shape = ds.chunks
        else:
            paths = glob.glob(os.path.join(self.fname, "*.tif")) or glob.glob(os.path.join(self.fname, "*.tiff"))
            if not paths:
                raise ValueError("no tiff files in %s" % self.fname)
            img = read_tiffseq(self.fname, s=slice(0, 1))[0]
            self.d_shape = (len(paths),) + tuple(img.shape)
            self.d_type = img.dtype
            self.chunk_shape = None
        out = "\n" + "#" * 50 + "\nFound existing %s: " % ("tiff folder" if self.tiff_mode else "hdf5 file") \
            + os.path.basename(self.fname) + "\nDataset shape: %s" % (str(self.d_shape))
        if return_output:
            return out
        _message(out, self.VERBOSITY > 0)

    def get_slice_sizes(self):
        """datafile.py:260-275: float32, uint8 and uint16 voxels, as in the reference."""
        fac = {np.dtype(np.float32): 4.0, np.dtype(np.uint8): 1.0, np.dtype(np.uint16): 2.0}.get(np.dtype(self.d_type))
        if fac is None:
            raise ValueError("Data type %s is not supported." % self.d_type)
        s = self.d_shape
        self.slice_size = (np.prod(s[1:]) * fac / 1e9, np.prod(s[::2]) * fac / 1e9, np.prod(s[:-1]) * fac / 1e9)
        self._bytes_per_voxel = fac
        self.d_size_GB = fac * np.prod(s) / 1e9

    def show_stats(self, return_output=False):
        out = "\nDataset shape: %s\nDataset size: %.2f GB" % (str(self.d_shape), self.d_size_GB)
        if not self.tiff_mode:
            out += "\nChunk shape: %s" % (str(self.chunk_shape))
        for i, size in enumerate(self.slice_size):
            out += "\nSlice size along %i: %4.2f MB" % (i, 1000.0 * size)
        if return_output:
            return out
        _message(out, self.VERBOSITY > -1)

    def est_chunking(self):
        """datafile.py:298-320: chunk_shape > chunk_size > chunked_slice_size."""
        if self.tiff_mode:
            self.chunked_shape = None
            return
        if self.chunk_shape is not None:
            return
        if self.chunk_size is not None:
            fac = np.cbrt(self.chunk_size / self.d_size_GB)
            self.chunk_shape = tuple(int(self.d_shape[i] * fac) for i in range(3))
            return
        if self.chunked_slice_size is not None:
            if self.chunked_slice_size > (self.slice_size[0] * self.d_shape[0]):
                raise ValueError("chunked_slice_size cannot be larger than dataset size")
            self.chunk_shape = tuple(int(np.ceil(self.chunked_slice_size / one)) for one in self.slice_size)
            _message("Estimated Chunk shape: %s" % str(self.chunk_shape), self.VERBOSITY > 1)
            return
        self.chunk_shape = None

    # ------------------------------------------------------------------ reading
    def read_slice(self, axis=None, slice_idx=None):
        img, _ = self.read_chunk(axis=axis, slice_start=slice_idx, slice_end=slice_idx + 1)
        return img[0]

    def read_data(self, slice_3D=(slice(None, None),) * 3):
        """datafile.py:333-348 (the reference falls through to h5py after reading a TIFF series; here the TIFF branch returns)."""
        if self.tiff_mode:
            ch = np.asarray(read_tiffseq(self.fname, s=slice_3D[0]))
            return ch[:, slice_3D[1], slice_3D[2]]
        with h5_module().File(self.fname, "r") as hf:
            return np.asarray(hf[self.data_tag][slice_3D[0], slice_3D[1], slice_3D[2]])

    def read_sequence(self, idxs):
        if self.tiff_mode:
            return np.asarray(read_tiffseq(self.fname, s=list(idxs)))
        with h5_module().File(self.fname, "r") as hf:
            return np.asarray(hf[self.data_tag][list(idxs), ...])

    def read_chunk(self, axis=None, slice_start=None, chunk_shape=None, max_GB=10.0, slice_end="", skip_fac=None,
                   return_slice=True):
        """datafile.py:361-430: slices [slice_start, slice_end) along `axis` (slice_end from max_GB when not given), moved to
        the front."""
        if slice_end == "":
            if (chunk_shape is None) and (not self.tiff_mode):
                chunk_shape = self.chunk_shape
            chunk_len = chunk_shape[axis] if chunk_shape is not None else 1
            slice_len = max(1, int(np.round(max_GB / self.slice_size[axis]) // chunk_len)) * chunk_len
            slice_end = min(slice_len + slice_start, self.d_shape[axis])
        elif slice_end is None:
            slice_end = self.d_shape[axis]
        s = slice(slice_start, slice_end, skip_fac)
        if not self.tiff_mode:
            with h5_module().File(self.fname, "r") as hf:
                key = [slice(None)] * 3
                key[axis] = s
                ch = np.moveaxis(np.asarray(hf[self.data_tag][tuple(key)]), axis, 0)
        else:
            if axis != 0:
                raise ValueError("TIFF data format does not support multi-axial slicing.")
            ch = np.asarray(read_tiffseq(self.fname, s=s))
        return (ch, s) if return_slice else ch

    def read_full(self, skip_fac=None):
        ch, _ = self.read_chunk(axis=0, slice_start=0, slice_end=self.d_shape[0], skip_fac=skip_fac)
        return ch

    # ------------------------------------------------------------------ writing
    def write_full(self, ch):
        self.write_chunk(ch, axis=0, s=slice(0, self.d_shape[0]))

    def write_data(self, ch, slice_3D=None):
        if self.tiff_mode:
            raise ValueError("Not supported for tiff data")
        with h5_module().File(self.fname, "r+") as hf:
            hf[self.data_tag][slice_3D[0], slice_3D[1], slice_3D[2]] = ch

    def write_chunk(self, ch, axis=None, s=None):
        """datafile.py:469-505."""
        if not self.tiff_mode:
            if not os.path.exists(self.fname):
                raise ValueError("hdf5 file needs to exist before writing chunks.")
            with h5_module().File(self.fname, "r+") as hf:
                key = [slice(None)] * 3
                key[axis] = s
                hf[self.data_tag][tuple(key)] = np.moveaxis(np.asarray(ch), 0, axis)
        else:
            if axis != 0:
                raise ValueError("TIFF data format does not support multi-axial slicing.")
            if os.path.exists(self.fname) and (s.start == 0):
                raise FileExistsError("TIFF folder to be written already exists. Please delete first.")
            write_tiffseq(np.asarray(ch).astype(self.d_type), SaveDir=self.fname, increment_flag=True,
                          suffix_len=len(str(self.d_shape[axis])))

"""Import the reference's own modules, unmodified, in a container without cupy/h5py/tensorflow.

Test infrastructure only.  ``/root/reference`` does not exist on the GPU box, so callers must
guard with ``available()``; the fixtures under ``tests/golden`` are generated with this loader by
``tests/golden/make_golden.py`` and committed.

The stubs only satisfy module-level imports: on the extract path the reference uses nothing but
``cp.get_array_module`` (``structures/patches.py:759``, ``structures/grid.py:324``).
"""
import importlib.util
import os
import sys
import types

import numpy as np

REF_ROOT = os.environ.get("TOMOENC_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REF_ROOT, "tomo_encoders", "structures", "patches.py"))


def _stub(name, **attrs):
    if name in sys.modules and not getattr(sys.modules[name], "__tomoenc_stub__", False):
        return sys.modules[name]
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__tomoenc_stub__ = True
    sys.modules[name] = m
    return m


def _install_stubs():
    _stub("cupy", get_array_module=lambda *a: np)
    _stub("h5py")
    tf = _stub("tensorflow")
    k = _stub("tensorflow.keras")
    lay = _stub("tensorflow.keras.layers", UpSampling3D=object)
    tf.keras = k
    k.layers = lay


_VOID_STUBS = ("cupyx", "cupyx.scipy", "cupyx.scipy.ndimage", "tifffile", "skimage", "skimage.measure", "tomo_encoders",
               "tomo_encoders.misc", "tomo_encoders.misc.voxel_processing", "tomo_encoders.misc.ellipse_rad_max")


def _remove_stubs():
    for name in ("cupy", "h5py", "tensorflow", "tensorflow.keras", "tensorflow.keras.layers") + _VOID_STUBS:
        m = sys.modules.get(name)
        if m is not None and getattr(m, "__tomoenc_stub__", False):
            del sys.modules[name]


def _load(relpath, name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF_ROOT, relpath))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


_cache = {}


def load():
    """Returns a namespace with RefPatches, RefGrid, ref_voxel (module)."""
    if "ns" in _cache:
        return _cache["ns"]
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    _install_stubs()
    try:
        ns = types.SimpleNamespace()
        ns.RefPatches = _load("tomo_encoders/structures/patches.py", "_ref_patches").Patches
        ns.RefGrid = _load("tomo_encoders/structures/grid.py", "_ref_grid").Grid
        ns.ref_voxel = _load("tomo_encoders/misc/voxel_processing.py", "_ref_voxel")
        # structures/voids.py (count_voids / import_from_grid / export_grid): its module-level imports of cupyx, tifffile,
        # skimage and of the tomo_encoders package itself are satisfied by stubs that hand it the reference's own Grid
        from scipy.ndimage import label as _label_np
        _stub("cupyx"); _stub("cupyx.scipy"); _stub("cupyx.scipy.ndimage", label=_label_np)
        _stub("tifffile", imsave=None, imread=None)
        _stub("skimage"); _stub("skimage.measure", marching_cubes=None)
        _stub("tomo_encoders", Grid=ns.RefGrid)
        _stub("tomo_encoders.misc")
        _stub("tomo_encoders.misc.voxel_processing", TimerGPU=getattr(ns.ref_voxel, "TimerGPU", None),
              TimerCPU=getattr(ns.ref_voxel, "TimerCPU", None))
        _stub("tomo_encoders.misc.ellipse_rad_max", get_all_max_ellip_rad=None)
        ns.RefVoids = _load("tomo_encoders/structures/voids.py", "_ref_voids").Voids
    finally:
        _remove_stubs()
    _cache["ns"] = ns
    return ns

"""ctypes binding of libtomoenc.so (C ABI declared in include/tomoenc.h) and its in-tree build.

There is no CPU fallback: ``lib()`` raises if the shared library is missing or cannot be loaded, and
every data-plane call in this package goes through it.
"""
import ctypes
import glob
import os
import subprocess
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(_HERE, "libtomoenc.so")
HEADER = os.path.join(_ROOT, "include", "tomoenc.h")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]

TE_U8, TE_U16, TE_F16, TE_F32, TE_BF16 = 0, 1, 2, 3, 4
TE_MODE_BF16, TE_MODE_FP16, TE_MODE_FP32_CHECK = 0, 1, 2
TE_HINT_ALIGNED = 1
TE_HINT_UNBINNED = 2
MODES = {"bf16": TE_MODE_BF16, "fp16": TE_MODE_FP16, "fp32": TE_MODE_FP32_CHECK,
         "fp32check": TE_MODE_FP32_CHECK}


class TomoEncError(RuntimeError):
    pass


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _stale():
    if not os.path.isfile(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + [HEADER]
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force=False, verbose=False):
    """Compiles csrc/*.cu for sm_100a into tomoencoders_b200/libtomoenc.so (nvcc cross-compiles
    without a GPU): one object per source (only stale ones are recompiled, in parallel) under
    tomoencoders_b200/_build/, then one link.  Returns the library path."""
    if not force and not _stale():
        return LIB_PATH
    from concurrent.futures import ThreadPoolExecutor
    nvcc = os.environ.get("NVCC", "nvcc")
    obj_dir = os.path.join(_HERE, "_build")
    os.makedirs(obj_dir, exist_ok=True)
    headers = glob.glob(os.path.join(CSRC, "*.cuh")) + [HEADER]
    t_hdr = max(os.path.getmtime(h) for h in headers)
    flags = [f for f in NVCC_FLAGS if f != "-shared"]
    jobs, objs = [], []
    for src in sources():
        obj = os.path.join(obj_dir, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        if force or not os.path.isfile(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), t_hdr):
            jobs.append([nvcc] + flags + ["-c", "-o", obj, src])

    def run(cmd):
        if verbose:
            print(" ".join(cmd))
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise TomoEncError("nvcc failed:\n%s\n%s" % (r.stdout, r.stderr))

    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        list(ex.map(run, jobs))
    tmp = LIB_PATH + ".tmp.%d" % os.getpid()
    # cuFFT (the FBP ramp filter, csrc/recon_ops.cu) comes from the toolkit next to nvcc; same path on the GPU box
    cuda_lib = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "lib64")
    run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", tmp] + objs +
        ["-L" + cuda_lib, "-lcufft", "-Xlinker", "-rpath=" + cuda_lib])
    os.replace(tmp, LIB_PATH)
    return LIB_PATH


class NetDesc(ctypes.Structure):
    """struct te_net_desc (include/tomoenc.h)."""
    _fields_ = [
        ("kind", ctypes.c_int32),
        ("n_blocks", ctypes.c_int32),
        ("n_filters", ctypes.c_int32 * 8),
        ("pool_size", ctypes.c_int32 * 8),
        ("isconcat", ctypes.c_int32 * 8),
        ("batch_norm", ctypes.c_int32),
        ("kern_size", ctypes.c_int32),
        ("kern_size_upconv", ctypes.c_int32),
        ("n_hidden", ctypes.c_int32),
        ("hidden_units", ctypes.c_int32 * 8),
        ("pool_flag", ctypes.c_int32),
        ("patch", ctypes.c_int32 * 3),
        ("max_batch", ctypes.c_int32),
        ("lrelu_alpha", ctypes.c_float),
        ("bn_eps", ctypes.c_float),
        ("mode", ctypes.c_int32),
    ]


class TeTensor(ctypes.Structure):
    """struct te_tensor (include/tomoenc.h): an NDHWC channel slice handed to the training operators."""
    _fields_ = [("ptr", ctypes.c_void_p), ("dtype", ctypes.c_int32), ("ctot", ctypes.c_int32), ("coff", ctypes.c_int32),
                ("C", ctypes.c_int32), ("N", ctypes.c_int32), ("D", ctypes.c_int32), ("H", ctypes.c_int32),
                ("W", ctypes.c_int32)]


_P = ctypes.c_void_p
_I64 = ctypes.c_int64
_I32 = ctypes.c_int32
_F = ctypes.c_float
_D = ctypes.c_double
_TP = ctypes.POINTER(TeTensor)
_SIGNATURES = {
    # name: (restype, argtypes)
    "te_version": (ctypes.c_int, []),
    "te_last_error": (ctypes.c_char_p, []),
    "te_launch_count": (_I64, []),
    "te_launch_count_add": (_I64, [_I64]),
    "te_gather": (ctypes.c_int, [_P, ctypes.c_int, ctypes.POINTER(_I64), _P, _P, _I64,
                                 ctypes.POINTER(_I32), _P, ctypes.c_int, _P]),
    "te_gather_perm": (ctypes.c_int, [_P, ctypes.c_int, ctypes.POINTER(_I64), _P, _P, _I64,
                                      ctypes.POINTER(_I32), _P, ctypes.c_int, _P, _P]),
    "te_gather_norm_perm": (ctypes.c_int, [_P, ctypes.c_int, ctypes.POINTER(_I64), _P, _P, _I64,
                                           ctypes.POINTER(_I32), ctypes.c_int, ctypes.c_int, ctypes.c_float,
                                           ctypes.c_float, _P, ctypes.c_int, _P, _P]),
    "te_gather_norm": (ctypes.c_int, [_P, ctypes.c_int, ctypes.POINTER(_I64), _P, _P, _I64,
                                      ctypes.POINTER(_I32), ctypes.c_int, ctypes.c_int, ctypes.c_float,
                                      ctypes.c_float, _P, ctypes.c_int, _P]),
    "te_gather_standardize": (ctypes.c_int, [_P, ctypes.c_int, ctypes.POINTER(_I64), _P, _P, _I64,
                                             ctypes.POINTER(_I32), ctypes.c_int, ctypes.c_float, ctypes.c_float, _P, _P,
                                             ctypes.c_int, _P]),
    "te_patch_std": (ctypes.c_int, [_P, ctypes.c_int, ctypes.POINTER(_I64), _P, _P, _I64, ctypes.POINTER(_I32), _P, _P]),
    "te_training_pairs": (ctypes.c_int, [_P, ctypes.c_int, _P, ctypes.c_int, ctypes.POINTER(_I64), _P, _P, _I64,
                                         ctypes.POINTER(_I32), _P, _P, _P, ctypes.c_uint64, _P, _P, _P]),
    "te_scatter": (ctypes.c_int, [_P, ctypes.c_int, _P, _I64, ctypes.POINTER(_I32), _P,
                                  ctypes.POINTER(_I64), _P, ctypes.c_int, _P]),
    "te_minmax_strided": (ctypes.c_int, [_P, ctypes.c_int, ctypes.POINTER(_I64), ctypes.c_int, _P, _P, _P]),
    "te_upload_sampled": (ctypes.c_int, [_P, _I32, ctypes.POINTER(_I64), _I32, _P, _P]),
    "te_minmax_nd": (ctypes.c_int, [_P, ctypes.c_int, ctypes.POINTER(_I64), ctypes.POINTER(_I64), _P, _P, _P]),
    "te_histogram_nd": (ctypes.c_int, [_P, ctypes.c_int, ctypes.POINTER(_I64), ctypes.POINTER(_I64), _P, _I32, _P, _P]),
    "te_rescale": (ctypes.c_int, [_P, ctypes.c_int, _P, ctypes.c_int, _I64, _F, _F, _P]),
    "te_edge_map": (ctypes.c_int, [_P, ctypes.c_int, ctypes.POINTER(_I64), _P, _P]),
    "te_cylindrical_mask": (ctypes.c_int, [_P, ctypes.c_int, ctypes.POINTER(_I64), _I64, ctypes.c_uint32, _P]),
    "te_cyl_values": (ctypes.c_int, [_P, ctypes.c_int, ctypes.POINTER(_I64), ctypes.POINTER(_I64), _P, _P, _P, _I64, _P, _P]),
    "te_net_weight_count": (_I64, [ctypes.POINTER(NetDesc)]),
    "te_net_create": (ctypes.c_int, [ctypes.POINTER(NetDesc), ctypes.POINTER(_P)]),
    "te_net_set_weights": (ctypes.c_int, [_P, _P, _I64]),
    "te_net_destroy": (ctypes.c_int, [_P]),
    "te_net_workspace_bytes": (_I64, [_P]),
    "te_unet_forward": (ctypes.c_int, [_P, _P, _I64, _P, ctypes.c_int, _P, _P]),
    "te_encoder_forward": (ctypes.c_int, [_P, _P, _I64, _P, _P]),
    "te_decoder_forward": (ctypes.c_int, [_P, _P, _I64, _P, _P]),
    "te_net_layer_output": (_I64, [_P, ctypes.c_int, _P, _I64, _P]),
    "te_net_num_layers": (ctypes.c_int, [_P]),
    "te_net_flops_per_patch": (ctypes.c_double, [_P]),
    "te_net_layer_info": (ctypes.c_int, [_P, ctypes.c_int, ctypes.POINTER(ctypes.c_char_p),
                                         ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int)]),
    "te_net_set_profiling": (ctypes.c_int, [_P, ctypes.c_int]),
    "te_net_overflow": (ctypes.c_int, [_P, ctypes.c_int, ctypes.POINTER(ctypes.c_int), _P]),
    "te_net_layer_time": (ctypes.c_int, [_P, ctypes.c_int, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(_I64)]),
    "te_net_layer_kernel": (ctypes.c_int, [_P, ctypes.c_int, ctypes.c_char_p, ctypes.c_int]),
    "te_pca_accumulate": (ctypes.c_int, [_P, _I64, _I32, _P, _P]),
    "te_pca_solve": (ctypes.c_int, [_P, _I32, _I32, _P, _P, _P, _P, _P]),
    "te_pca_project": (ctypes.c_int, [_P, _I64, _I32, _P, _P, _I32, _P, _P]),
    # training operators
    "te_tr_conv3_wpack_bytes": (_I64, [_I32, _I32]),
    "te_tr_conv3": (ctypes.c_int, [_TP, _TP, _P, _P, _TP, _P, _P]),
    "te_conv1_kernel_kind": (ctypes.c_int, [_TP, _TP]),
    "te_tr_flip_weights": (ctypes.c_int, [_P, _I32, _I32, _I32, _I32, _P, _P]),
    "te_tr_conv3_wgrad": (ctypes.c_int, [_TP, _TP, _P, _I32, _I32, _P, _P]),
    "te_tr_bn_stats": (ctypes.c_int, [_TP, _P, _P]),
    "te_tr_bn_act_fwd": (ctypes.c_int, [_TP, _TP, _P, _P, _F, _P]),
    "te_tr_bn_act_bwd_reduce": (ctypes.c_int, [_TP, _TP, _P, _P, _P, _P, _F, _P, _P]),
    "te_tr_bn_act_bwd_apply": (ctypes.c_int, [_TP, _TP, _TP, _P, _P, _P, _P, _P, _P, _F, _P, _P]),
    "te_tr_maxpool_fwd": (ctypes.c_int, [_TP, _TP, _I32, _P]),
    "te_tr_maxpool_bwd": (ctypes.c_int, [_TP, _TP, _TP, _I32, _I32, _P]),
    "te_tr_upconv_fwd": (ctypes.c_int, [_TP, _P, _P, _TP, _I32, _F, _P, _P]),
    "te_tr_upconv_bwd": (ctypes.c_int, [_TP, _TP, _TP, _P, _I32, _F, _P, _P, _TP, _TP, _P, _P]),
    "te_tr_head_loss": (ctypes.c_int, [_TP, _P, _P, _P, _TP, _P, _P, _D, _P, _P, _P]),
    "te_tr_dense_fwd": (ctypes.c_int, [_P, _I64, _I32, _I32, _P, _P, _P, _P]),
    "te_tr_dense_bwd": (ctypes.c_int, [_P, _P, _I64, _I32, _I32, _P, _P, _P, _P, _P]),
    "te_tr_head_ae_loss": (ctypes.c_int, [_TP, _P, _P, _P, _TP, _P, _P, _D, _F, _P, _P, _P]),
    "te_tr_bn_finalize": (ctypes.c_int, [_P, _I32, _D, _P, _P, _F, _P, _P, _P, _P, _P, _P]),
    "te_tr_bn_bwd_finalize": (ctypes.c_int, [_P, _P, _I32, _D, _P, _P, _P, _P, _P]),
    "te_tr_take_stats": (ctypes.c_int, [_P, _I32, _I32, _P, _P]),
    "te_label_workspace_bytes": (_I64, [_I64]),
    "te_label": (ctypes.c_int, [_P, ctypes.c_int, ctypes.POINTER(_I64), ctypes.c_int, _P, _P, _P, _P]),
    "te_label_stats": (ctypes.c_int, [_P, ctypes.POINTER(_I64), _I64, _P, _P, _P]),
    "te_label_crop_masks": (ctypes.c_int, [_P, ctypes.POINTER(_I64), _P, _P, _I64, _P, _I64, _P, _P]),
    "te_paint_boxes": (ctypes.c_int, [_P, _P, _P, _I64, _I64, _P, ctypes.POINTER(_I64), _P, _P]),
    "te_tile_any": (ctypes.c_int, [_P, ctypes.c_int, ctypes.POINTER(_I64), _I32, _P, _P]),
    # reconstruction
    "te_fbp_padding": (None, [_I32, ctypes.POINTER(_I32), ctypes.POINTER(_I32)]),
    "te_fbp_workspace_bytes": (_I64, [_I64, _I32]),
    "te_fbp_filter": (ctypes.c_int, [_P, _I64, _I32, _P, _I64, _P]),
    "te_fbp_release_plans": (ctypes.c_int, []),
    "te_preprocess": (ctypes.c_int, [_P, _P, _P, _I32, _I32, _I32, _P]),
    "te_make_mask": (ctypes.c_int, [_P, ctypes.POINTER(_I64), _P, _I64, _I32, _P]),
    "te_bp_workspace_bytes": (_I64, [_I32, _I32, _I32]),
    "te_bp_prepare": (ctypes.c_int, [_P, _P, _I32, _I32, _I32, _P, _P]),
    "te_bp_box": (ctypes.c_int, [_P, _I32, _I32, _I32, _F, _I32, _I32, _I32, _I32, _I32, _I32, _P, _I32, _I32, _P]),
    "te_bp_patches": (ctypes.c_int, [_P, _I32, _I32, _I32, _F, _P, _I64, _I32, _P, _I32, _I32, _F, _F, _I32, _P]),
    "te_bp_points": (ctypes.c_int, [_P, _I32, _I32, _I32, _F, _P, _I32, _I32, _P, _I32, _P]),
    "te_peer_create": (ctypes.c_int, [_I32, _I32, _I32, _I32, ctypes.POINTER(_P)]),
    "te_peer_export": (ctypes.c_int, [_P, _P]),
    "te_peer_import": (ctypes.c_int, [_P, _P]),
    "te_peer_allreduce_f64": (ctypes.c_int, [_P, _P, _I32, _P]),
    "te_peer_destroy": (ctypes.c_int, [_P]),
    "te_tr_adam": (ctypes.c_int, [_P, _P, _P, _P, _I64, _F, _F, _F, _F, _I64, _F, _P]),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lock = threading.Lock()
_lib = None


def lib():
    """Loads libtomoenc.so (once).  Raises TomoEncError if it has not been built: the product path
    never falls back to a CPU implementation."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.isfile(LIB_PATH):
            raise TomoEncError(
                "libtomoenc.so is missing (%s): build it with `python -c 'import __graft_entry__ as g; "
                "g.build()'`; there is no CPU fallback" % LIB_PATH)
        try:
            handle = ctypes.CDLL(LIB_PATH)
        except OSError as e:  # pragma: no cover
            raise TomoEncError("cannot load %s: %s" % (LIB_PATH, e))
        for name, (res, args) in _SIGNATURES.items():
            try:
                fn = getattr(handle, name)
            except AttributeError:
                raise TomoEncError("libtomoenc.so does not export %s (stale build?)" % name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = lib().te_last_error().decode("utf-8", "replace")
        raise TomoEncError("%s failed (%d): %s" % (what or "libtomoenc call", rc, msg))


def i64x3(v):
    return (_I64 * 3)(int(v[0]), int(v[1]), int(v[2]))


def i32x3(v):
    return (_I32 * 3)(int(v[0]), int(v[1]), int(v[2]))

"""Processor base classes and the model handle they own.

Drop-in for /root/reference/tomo_encoders/neural_nets/keras_processor.py:25-253
(``GenericKerasProcessor`` -> ``Vox2VoxProcessor_fCNN`` / ``EmbeddingLearner``): same constructor
(``model_initialization``, ``descriptor_tag``, ``**model_params``), ``self.models`` dict,
``predict_patches(model_key, x, chunk_size, out_arr, min_max=None, TIMEIT=False)``,
``calc_voxel_min_max``, ``rescale_data``, ``test_speeds``.  ``keras.Model`` is replaced by ``Model``
below, which keeps the few Keras methods callers use (``predict``, ``get_weights``,
``set_weights``, ``count_params``, ``input_shape`` / ``output_shape``) and runs on libtomoenc.so.

Differences: the per-chunk clip / rescale is evaluated in fp32 on the
GPU, fused into the kernel that stages the chunk, instead of numpy arithmetic in the input dtype;
the last chunk is not edge-padded (patches are independent, so the result is identical); the
reference's quirks Q4/Q5/Q7 (NameError with min_max, autoencoder branch on output_type, rounding of
embeddings) are not reproduced.
"""
import os
import time
import warnings

import numpy as np
import torch

from .. import _device, _lib
from ..misc.voxel_processing import _find_min_max
from . import _spec
from ._net import Net

DEFAULT_MODE = os.environ.get("TOMOENC_MODE", "fp16")


class Model:
    """One network of the reference (kind: "unet", "encoder" or "decoder") with its weights.

    The U-net is fully convolutional (Input((None,None,None,1)), Unet3D.py:242): execution plans
    (``Net`` handles: workspace + TMA descriptors) are created per (patch size, batch) on first use
    and cached."""

    def __init__(self, kind, model_params, model_size=None, mode=None, name="model", seed=None):
        self.kind = kind
        self.name = name
        self.mode = mode or DEFAULT_MODE
        self.model_params = dict(model_params)
        self.model_size = None if model_size is None else tuple(int(s) for s in model_size)
        p = self.model_params
        if "n_blocks" not in p:
            p["n_blocks"] = len(p["n_filters"])
        if kind == "unet":
            self.spec = _spec.unet_spec(**p)
        elif kind == "encoder":
            self.spec = _spec.encoder_spec(self.model_size, **p)
        elif kind == "decoder":
            self.spec = _spec.decoder_spec(self.model_size, **p)
        else:
            raise ValueError(kind)
        self._weights = _spec.keras_default_init(self.spec, seed)
        self._nets = {}

    # ---- Keras-like surface
    def get_weights(self):
        return [w.copy() for w in self._weights]

    def set_weights(self, weights):
        weights = [np.asarray(w, dtype=np.float32) for w in weights]
        if len(weights) != len(self.spec):
            raise ValueError("expected %d weight arrays, got %d" % (len(self.spec), len(weights)))
        for w, (kind, shape) in zip(weights, self.spec):
            if tuple(w.shape) != tuple(shape):
                raise ValueError("weight %s: expected shape %s, got %s" % (kind, shape, w.shape))
        self._weights = weights
        for net in self._nets.values():
            net.set_weights(self._weights)

    def count_params(self):
        return int(sum(int(np.prod(s)) for _, s in self.spec))

    @property
    def input_shape(self):
        if self.kind == "decoder":
            return (None, self.latent)
        return (None,) + (self.model_size or (None, None, None)) + (1,)

    @property
    def output_shape(self):
        if self.kind == "encoder":
            return (None, self.latent)
        return (None,) + (self.model_size or (None, None, None)) + (1,)

    @property
    def latent(self):
        hu = self.model_params.get("hidden_units")
        return int(hu[-1]) if hu else 0

    def save(self, filepath, include_optimizer=False):
        np.savez(filepath, **{"w%04d" % i: w for i, w in enumerate(self._weights)})

    def load_weights(self, filepath):
        with np.load(filepath) as f:
            self.set_weights([f["w%04d" % i] for i in range(len(f.files))])

    # ---- fp16 range guard
    def fp16_overflowed(self):
        """True (and the flags are cleared) when any fp16 plan of this model saw an activation leave the finite fp16 range
        since the last check."""
        hit = False
        for net in self._nets.values():
            hit = net.overflowed(reset=True) or hit
        return hit

    def set_mode(self, mode):
        """Switches the arithmetic mode ("fp16" / "bf16" / "fp32"); execution plans are rebuilt on next use."""
        if mode != self.mode:
            self.mode = mode
            self._nets = {}

    # ---- execution
    def net(self, patch, max_batch):
        patch = tuple(int(p) for p in patch)
        if self.model_size is not None and patch != self.model_size:
            raise ValueError("model was built for input size %s, got %s" % (self.model_size, patch))
        key = (patch, int(max_batch))
        if key not in self._nets:
            # keep the cache small: plans for smaller batches of the same patch size are superseded
            for k in [k for k in self._nets if k[0] == patch and k[1] < max_batch]:
                del self._nets[k]
            net = Net(self.kind, patch, max_batch, mode=self.mode, **self.model_params)
            net.set_weights(self._weights)
            self._nets[key] = net
        return self._nets[key]

    def net_for(self, patch, n):
        patch = tuple(int(p) for p in patch)
        best = None
        for (pp, mb), net in self._nets.items():
            if pp == patch and mb >= n and (best is None or mb < best.max_batch):
                best = net
        return best if best is not None else self.net(patch, n)

    def predict(self, x, batch_size=32, verbose=0):
        """keras.Model.predict: x (nb,pz,py,px,1) -> same shape float32 (unet / decoder input is
        (nb, latent)); encoder -> (nb, latent) float32.  Host arrays in, host arrays out."""
        x = np.asarray(x)
        dev = _device.device()
        if self.kind == "decoder":
            z = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).to(dev)
            out = []
            for a in range(0, len(z), batch_size):
                zc = z[a:a + batch_size].contiguous()
                out.append(self.net_for(self.model_size, len(zc)).decoder_forward(zc))
            return torch.cat(out).cpu().numpy()[..., np.newaxis]
        assert x.ndim == 5 and x.shape[-1] == 1, "x must be 5-dimensional (batch_size, nz, ny, nx, 1)."
        patch = x.shape[1:4]
        outs = []
        for a in range(0, len(x), batch_size):
            net = self.net_for(patch, min(batch_size, len(x) - a))
            xc = torch.from_numpy(np.ascontiguousarray(x[a:a + batch_size, ..., 0], dtype=np.float32)).to(dev)
            xc = xc.to(net.act_dtype).contiguous()
            if self.kind == "unet":
                outs.append(net.unet_forward(xc, out_dtype=torch.float32))
            else:
                outs.append(net.encoder_forward(xc))
        out = torch.cat(outs).cpu().numpy()
        return out[..., np.newaxis] if self.kind == "unet" else out


def _standardized(vol_dev, vol_shape, d_points, d_widths, n, patch, min_max, net):
    """Gather + optional rescale + per-patch min-max standardisation (build_CAE_3D(stdinput=True),
    scratchpad/IEEE_ICIP_paper/porosity_encoders.py:89-98) into network-input layout."""
    out = torch.empty((n,) + tuple(patch), dtype=net.act_dtype, device=vol_dev.device)
    mm = torch.empty((max(n, 1), 2), dtype=torch.float32, device=vol_dev.device)
    lo, hi = (0.0, 1.0) if min_max is None else (float(min_max[0]), float(min_max[1]))
    _lib.check(_lib.lib().te_gather_standardize(vol_dev.data_ptr(), _device.te_dtype(vol_dev), _lib.i64x3(vol_shape),
                                                d_points.data_ptr(), d_widths.data_ptr(), n, _lib.i32x3(patch),
                                                0 if min_max is None else 1, lo, hi, mm.data_ptr(), out.data_ptr(), net.act_te,
                                                _device.stream_ptr()), "te_gather_standardize")
    return out


def normalize_chunk(x_dev, min_max, clip, net, standardize=False):
    """(n,pz,py,px) device tensor of any supported dtype -> activation dtype of ``net``, with the
    clip / rescale of predict_patches fused (te_gather_norm on the chunk viewed as a volume)."""
    x_dev = _device.to_supported(x_dev)
    n, pz, py, px = x_dev.shape
    out = torch.empty((n, pz, py, px), dtype=net.act_dtype, device=x_dev.device)
    pts = torch.zeros((n, 3), dtype=torch.int32, device=x_dev.device)
    pts[:, 0] = torch.arange(n, device=x_dev.device, dtype=torch.int32) * pz
    wds = torch.tensor([pz, py, px], dtype=torch.int32, device=x_dev.device).repeat(n, 1).contiguous()
    if standardize:
        return _standardized(x_dev, (n * pz, py, px), pts, wds, n, (pz, py, px), min_max, net)
    lo, hi = (0.0, 1.0) if min_max is None else (float(min_max[0]), float(min_max[1]))
    L = _lib.lib()
    _lib.check(L.te_gather_norm(x_dev.data_ptr(), _device.te_dtype(x_dev), _lib.i64x3((n * pz, py, px)),
                                pts.data_ptr(), wds.data_ptr(), n, _lib.i32x3((pz, py, px)),
                                0 if min_max is None else 1, 1 if clip else 0, lo, hi, out.data_ptr(), net.act_te,
                                _device.stream_ptr()), "te_gather_norm")
    return out


def gather_normalized(vol_dev, vol_shape, d_points, d_widths, n, patch, min_max, clip, net, standardize=False):
    """Fused gather + clip + rescale straight from the volume into network-input layout."""
    if standardize:
        return _standardized(vol_dev, vol_shape, d_points, d_widths, n, patch, min_max, net)
    out = torch.empty((n,) + tuple(patch), dtype=net.act_dtype, device=vol_dev.device)
    lo, hi = (0.0, 1.0) if min_max is None else (float(min_max[0]), float(min_max[1]))
    L = _lib.lib()
    _lib.check(L.te_gather_norm(vol_dev.data_ptr(), _device.te_dtype(vol_dev), _lib.i64x3(vol_shape),
                                d_points.data_ptr(), d_widths.data_ptr(), n, _lib.i32x3(patch),
                                0 if min_max is None else 1, 1 if clip else 0, lo, hi, out.data_ptr(), net.act_te,
                                _device.stream_ptr()), "te_gather_norm")
    return out


_SLAB_BYTES = 16 << 20        # z-slab size of the streamed host <-> device copies in process_volume


class GenericKerasProcessor:
    """keras_processor.py:25-196."""

    def __init__(self, model_initialization="define-new", descriptor_tag="misc", **kwargs):
        self.input_type = "data"
        self.output_type = getattr(self, "output_type", "data")
        self.mode = kwargs.pop("mode", None) or DEFAULT_MODE
        if model_initialization == "define-new":
            self._build_models(descriptor_tag=descriptor_tag, **kwargs)
        elif model_initialization == "load-model":
            self._load_models(**kwargs)
        else:
            raise NotImplementedError("method is not implemented")

    # ---- fp16 range guard: Keras computes in float32 and cannot overflow where fp16 activations can (|x| > 65504).  The conv
    #      epilogues flag it (te_net_overflow); the call is then repeated with bf16 activations (same exponent range as
    #      float32) and the model stays in bf16 from there on.
    def _model_keys(self, model_key):
        return ["encoder", "decoder"] if model_key == "autoencoder" else [model_key]

    def _guard_fp16(self, model_key, run):
        """run() executes the forward work; if an fp16 plan overflowed, switches the models to bf16 and runs again."""
        out = run()
        models = [self.models.get(k) for k in self._model_keys(model_key)]
        models = [m for m in models if isinstance(m, Model) and m.mode == "fp16"]
        if any([m.fp16_overflowed() for m in models]):
            warnings.warn("fp16 activations overflowed (|x| > 65504) in model %r: re-running with bf16 activations; the model "
                          "stays in bf16 mode (construct the processor with mode='bf16' to avoid the repeated pass)" % model_key,
                          RuntimeWarning, stacklevel=3)
            for m in models:
                m.set_mode("bf16")
            out = run()
        return out

    # ---- CUDA-graph replay of the per-chunk kernel sequence (process_volume, embed_volume)
    use_graphs = os.environ.get("TOMOENC_NO_GRAPH", "0") != "1"
    _GRAPH_CACHE_MAX = 8
    _STAGING_MAX_BYTES = 48 << 30

    def _staging(self, tag, shape, dtype, dev):
        """Device staging buffer for a HOST volume / result, kept between calls (one per tag): the same host-to-host call
        repeated on volumes of one shape then runs on stable device addresses, so its captured chunk graphs are reused."""
        cache = self.__dict__.setdefault("_staging_bufs", {})
        t = cache.get(tag)
        if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype or t.device != dev:
            cache.pop(tag, None)
            t = torch.empty(shape, dtype=dtype, device=dev)
            if t.numel() * t.element_size() <= self._STAGING_MAX_BYTES:
                cache[tag] = t
        return t

    def _chunk_graph(self, net, key, build):
        """Captured chunk graphs live on the plan (``Net``) whose workspace they address, so they die with it."""
        cache = net.__dict__.setdefault("_chunk_graphs", {})
        g = cache.pop(key, None)
        if g is None:
            g = build()
            while len(cache) >= self._GRAPH_CACHE_MAX:
                cache.pop(next(iter(cache)))
        cache[key] = g                                   # most recently used last
        return g

    # ---- helpers shared by the subclasses
    _clip_input = False      # SurfaceSegmenter clips to min_max before rescaling (surface_segmenter.py:272-275)

    def print_layers(self, modelkey):
        model = self.models[modelkey]
        print("\n".join("%s    ::    %s" % (str(s), k) for k, s in model.spec))

    def _run_model(self, model_key, x_act, net_cache):
        raise NotImplementedError

    def predict_patches(self, model_key, x, chunk_size, out_arr, min_max=None, TIMEIT=False):
        """keras_processor.py:78-146.  x: (nb,pz,py,px,1) numpy / torch (host or CUDA); returns an
        array of the same shape: uint8 labels if output_type == "labels", otherwise float."""
        assert x.ndim == 5, "x must be 5-dimensional (batch_size, nz, ny, nx, 1)."
        t0 = time.time()
        nb = len(x)
        on_device = _device.is_device_array(x)
        labels = self.output_type == "labels"
        if out_arr is not None:
            assert tuple(out_arr.shape) == tuple(x.shape), \
                "x and out_arr shapes must be equal and 4-dimensional (batch_size, nz, ny, nx, 1)"
        patch = tuple(int(s) for s in x.shape[1:4])
        dev = _device.device()
        out_t_dtype = torch.uint8 if labels else torch.float32
        res = torch.empty((nb,) + patch, dtype=out_t_dtype, device=dev)

        def run():
            for a in range(0, nb, chunk_size):
                b = min(a + chunk_size, nb)
                xc = x[a:b, ..., 0]
                xc = _device.to_device(xc if on_device else np.ascontiguousarray(xc))
                self._predict_chunk(model_key, xc, res[a:b], min_max, chunk_size)
        self._guard_fp16(model_key, run)
        if on_device:
            out = res.unsqueeze(-1)
            if out_arr is not None:
                out_arr.copy_(out)
                out = out_arr
        else:
            host = res.cpu().numpy()[..., np.newaxis]
            if labels:
                out = host                                      # np.round(...).astype(np.uint8) (:135)
            else:
                tgt = out_arr.dtype if out_arr is not None else (x.dtype if np.issubdtype(np.asarray(x[:0]).dtype, np.floating) else np.float32)
                out = host.astype(tgt, copy=False)
            if out_arr is not None and not labels:
                out_arr[...] = out
                out = out_arr
        t_unit = (time.time() - t0) * 1000.0 / max(nb, 1)
        if TIMEIT:
            print("inf. time p. input patch size %s = %.2f ms, nb = %i" % (str(patch), t_unit, nb))
            print("\n")
            return out, t_unit
        return out

    def _predict_chunk(self, model_key, xc, out, min_max, chunk_size):
        """xc: (n,pz,py,px) device tensor (raw values); out: (n,pz,py,px) uint8 / float32 view."""
        patch = tuple(xc.shape[1:])
        if model_key == "autoencoder":
            enc = self.models["encoder"].net_for(patch, max(chunk_size, len(xc)))
            dec = self.models["decoder"].net_for(patch, max(chunk_size, len(xc)))
            xa = normalize_chunk(xc, min_max, self._clip_input, enc)
            y = dec.decoder_forward(enc.encoder_forward(xa))
            out.copy_(y if out.dtype == torch.float32 else (y > 0.5).to(torch.uint8))
            return
        model = self.models[model_key]
        net = model.net_for(patch, max(chunk_size, len(xc)))
        xa = normalize_chunk(xc, min_max, self._clip_input, net)
        if model.kind == "unet":
            net.unet_forward_into(xa, out)
        elif model.kind == "decoder":
            raise ValueError("the decoder takes latent vectors; use model_key='autoencoder'")
        else:
            raise ValueError("model %s does not produce patches; use predict_embeddings" % model_key)

    def calc_voxel_min_max(self, vol, sampling_factor, TIMEIT=False):
        return _find_min_max(vol, sampling_factor, TIMEIT=TIMEIT)

    def rescale_data(self, data, min_val, max_val):
        eps = 1e-12
        return (data - min_val) / (max_val - min_val + eps)

    def _msg_exec_time(self, func, t_exec):
        print("TIME: %s: %.2f seconds" % (func.__name__, t_exec))

    # ---- persistence (npz instead of Keras HDF5: h5py / TF are not dependencies of this package)
    def save_models(self, model_path):
        for model_key, model in self.models.items():
            if model is None:
                continue
            model.save(os.path.join(model_path, "%s_%s.npz" % (model_key, self.model_tag)))

    def train(self, *args, **kwargs):
        """Training lives in the subclasses that the reference trains (SurfaceSegmenter.train,
        SelfSupervisedCAE.train); the generic base has no data generator (keras_processor.py:25-196)."""
        raise NotImplementedError("%s does not define train(); use SurfaceSegmenter.train or SelfSupervisedCAE.train"
                                  % type(self).__name__)


class Vox2VoxProcessor_fCNN(GenericKerasProcessor):
    """keras_processor.py:199-243."""

    def test_speeds(self, chunk_size, n_reps=3, input_size=None, model_key="segmenter"):
        if input_size is None:
            input_size = (64, 64, 64)
        for _ in range(n_reps):
            x = np.random.uniform(0, 1, tuple([chunk_size] + list(input_size) + [1])).astype(np.float32)
            t0 = time.time()
            self.predict_patches(model_key, x, chunk_size, None, min_max=(-1, 1))
            t_unit = (time.time() - t0) * 1000.0 / len(x)
            print(f"inf. time per patch {input_size} = {t_unit:.2f} ms, nb = {len(x)}")
            print(f"inf. time per voxel {(t_unit / (np.prod(input_size)) * 1.0e6):.2f} ns")
            print("\n")

    def random_data_generator(self, batch_size, input_size=(64, 64, 64)):
        while True:
            x_shape = tuple([batch_size] + list(input_size) + [1])
            x = np.random.uniform(0, 1, x_shape)
            y = np.random.uniform(0, 1, x_shape)
            x[x == 0] = 1.0e-12
            y[y == 0] = 1.0e-12
            yield x, y

    # ---- fused volume path (gather -> net -> stitch on the device; the patches never exist in HBM
    #      as an (n,P,P,P) array of the volume dtype)
    def process_volume(self, vol, patches, chunk_size, min_max=None, model_key="segmenter", out=None):
        """Equivalent to
            x = patches.extract(vol[, patch_size]); y = self.predict_patches(model_key, x[...,None], chunk_size, None, min_max)
            patches.fill_patches_in_volume(y[...,0], out)
        (/root/reference/scratchpad/void_metrology_v2/bin/vis_voids.py:64-73) for fixed-width
        patches, executed chunk by chunk on the device.  vol: numpy / torch (host or CUDA).  Returns
        ``out`` (allocated as zeros of vol's shape if None): uint8 for labels, float32 otherwise.

        Launch structure: gather+clip+rescale -> 17 fused layers -> stitch of ONE chunk is captured once as a CUDA graph
        (per plan, chunk length, volume / result address) and replayed for every chunk; only the chunk's corner table
        (a 12 n byte device copy) changes between replays."""
        return self._guard_fp16(model_key, lambda: self._process_volume(vol, patches, chunk_size, min_max, model_key, out))

    def _process_volume(self, vol, patches, chunk_size, min_max, model_key, out):
        from ..structures import _table
        vs = tuple(int(v) for v in patches.vol_shape)
        assert tuple(vol.shape) == vs, "Shape of big volume does not match vol_shape attribute of patches data"
        widths = patches.widths if hasattr(patches, "widths") else patches._widths()
        if len(np.unique(np.asarray(widths), axis=0)) > 1:
            raise ValueError("process_volume needs fixed-width patches")
        n = len(patches)
        labels = self.output_type == "labels"
        host = not _device.is_device_array(vol)
        w3 = tuple(int(v) for v in widths[0]) if n else None
        # Host volumes are uploaded in z slabs on a side stream; a chunk only waits for the slabs it reads, so the
        # upload overlaps the network (and, with a pinned `out`, finished slabs of the result go back the same way).
        src = _device.as_host_tensor(vol) if host and n else None
        o_dtype = torch.uint8 if labels else torch.float32
        if src is not None:
            dev = _device.device()
            t_vol = self._staging("vol", vs, src.dtype, dev)
            plane_bytes = vs[1] * vs[2] * src.element_size()
            zs = max(1, min(vs[0], -(-_SLAB_BYTES // max(1, plane_bytes))))
            n_slabs = -(-vs[0] // zs)
            side = _device.side_stream()
            side.wait_stream(torch.cuda.current_stream())
            ev_up = [torch.cuda.Event() for _ in range(n_slabs)]
            issued = [0]
            pinned_in = src.is_pinned()

            def upload_to(s_last):
                s_last = min(s_last, n_slabs - 1)
                with torch.cuda.stream(side):
                    while issued[0] <= s_last:
                        k = issued[0]
                        t_vol[k * zs:(k + 1) * zs].copy_(src[k * zs:(k + 1) * zs], non_blocking=True)
                        ev_up[k].record(side)
                        issued[0] += 1
        else:
            # float64 / signed integer volumes (numpy defaults): one-shot upload, cast to float32 on the device
            t_vol = _device.to_supported(_device.to_device(vol))
        dev = t_vol.device
        o_es = 1 if labels else 4
        disjoint = n > 0 and _table._disjoint(patches.points, w3, vs) and (w3[2] * o_es) % 16 == 0
        # disjoint patches that fill the volume overwrite every voxel: the previous contents of `out` are not needed
        covers = disjoint and n * w3[0] * w3[1] * w3[2] == vs[0] * vs[1] * vs[2]
        if out is None:
            t_out = torch.zeros(vs, dtype=o_dtype, device=dev)
        elif covers and not _device.is_device_array(out) and _device.as_host_tensor(out) is not None:
            if tuple(out.shape) != vs or _device.as_host_tensor(out).dtype != o_dtype:
                raise TypeError("out must be a %s array of shape %s" % (o_dtype, vs))
            t_out = self._staging("out", vs, o_dtype, dev)
        else:
            t_out = _device.to_device(out)
            if t_out.dtype != o_dtype:
                raise TypeError("out must be %s" % o_dtype)
        if n == 0:
            return _device.to_host_like(t_out, vol) if out is None else out
        model = self.models[model_key]
        net = model.net_for(w3, min(chunk_size, n))
        d_pts = _device.points_to_device(patches.points)
        d_wds = _device.points_to_device(widths)
        if not disjoint:
            # overlapping patches: stitch once at the end through the ordered path
            all_out = torch.empty((n,) + w3, dtype=o_dtype, device=dev)
        L = _lib.lib()
        hint = _table._aligned_hint(np.asarray(patches.points), None, w3, t_out.element_size()) if disjoint else 0
        pz = np.asarray(patches.points)[:, 0].astype(np.int64)
        starts = list(range(0, n, chunk_size))
        # pipelined download: slabs below every later chunk are final once the current chunk is done
        out_host = _device.as_host_tensor(out) if (src is not None and out is not None and disjoint) else None
        stream_out = out_host is not None and out_host.is_pinned() and out_host.dtype == o_dtype
        if stream_out:
            zmin_after = [vs[0]] * len(starts)
            run = vs[0]
            for ci in range(len(starts) - 1, -1, -1):
                zmin_after[ci] = run
                run = min(run, int(pz[starts[ci]:starts[ci] + chunk_size].min()))
            down = 0                                     # slabs [0, down) have been sent back
        lo, hi = (0.0, 1.0) if min_max is None else (float(min_max[0]), float(min_max[1]))
        vs64, w32 = _lib.i64x3(vs), _lib.i32x3(w3)

        def chunk_body(pts_t, wds_t, nn, xa, y):
            """gather + clip + rescale -> network -> stitch of one chunk, all on the current stream."""
            _lib.check(L.te_gather_norm(t_vol.data_ptr(), _device.te_dtype(t_vol), vs64, pts_t.data_ptr(), wds_t.data_ptr(), nn,
                                        w32, 0 if min_max is None else 1, 1 if self._clip_input else 0, lo, hi,
                                        xa.data_ptr(), net.act_te, _device.stream_ptr()), "te_gather_norm")
            net.unet_forward_into(xa, y)
            if disjoint:
                _lib.check(L.te_scatter(y.data_ptr(), y.element_size(), pts_t.data_ptr(), nn, w32, t_out.data_ptr(), vs64,
                                        None, hint, _device.stream_ptr()), "te_scatter")

        graphs = self.use_graphs and disjoint and not net.profiling and len(starts) > 1
        for ci, a in enumerate(starts):
            b = min(a + chunk_size, n)
            if src is not None:
                need = min(n_slabs - 1, (int(pz[a:b].max()) + int(np.asarray(widths)[a:b, 0].max()) - 1) // zs)
                upload_to(n_slabs - 1 if pinned_in else need + 1)
                torch.cuda.current_stream().wait_event(ev_up[need])
            if graphs:
                key = (b - a, t_vol.data_ptr(), t_out.data_ptr(), str(t_vol.dtype), vs, w3, min_max is None, lo, hi,
                       self._clip_input, hint)
                g = self._chunk_graph(net, key, lambda: _ChunkGraph(b - a, w3, net.act_dtype, o_dtype, dev))
                g.run(d_pts[a:b], d_wds[a:b], chunk_body)
            else:
                xa = torch.empty((b - a,) + w3, dtype=net.act_dtype, device=dev)
                y = all_out[a:b] if not disjoint else torch.empty((b - a,) + w3, dtype=o_dtype, device=dev)
                chunk_body(d_pts[a:b], d_wds[a:b], b - a, xa, y)
            if stream_out:
                final = min(n_slabs, zmin_after[ci] // zs) if zmin_after[ci] < vs[0] else n_slabs
                if final > down:
                    ev = torch.cuda.Event()
                    ev.record(torch.cuda.current_stream())
                    with torch.cuda.stream(side):
                        side.wait_event(ev)
                        out_host[down * zs:final * zs].copy_(t_out[down * zs:final * zs], non_blocking=True)
                    down = final
        if src is not None:
            upload_to(n_slabs - 1)
            torch.cuda.current_stream().wait_stream(side)
        if not disjoint:
            _table.scatter(all_out, t_out, vs, patches.points, widths)
        if out is None:
            return _device.to_host_like(t_out, vol) if host else t_out
        if stream_out:
            side.synchronize()                           # every slab has been written into the caller's array
            return out
        if host:
            res = t_out.cpu()
            if isinstance(out, torch.Tensor):
                out.copy_(res)
            else:
                out[...] = res.numpy()
        elif t_out.data_ptr() != out.data_ptr():
            out.copy_(t_out)
        return out


class _ChunkGraph:
    """One chunk of the fused volume path (gather+clip+rescale -> network -> stitch) as a CUDA graph.

    The kernels read the chunk's corner table from a static device buffer; ``run`` copies the current chunk's corners
    there and replays.  The first ``run`` executes the body eagerly (it is that chunk's real work and doubles as the
    warm-up: shared-memory attributes, TMA descriptors), counts its kernel launches and then captures."""

    def __init__(self, n, w3, act_dtype, o_dtype, dev, out_shape=None):
        self.n = n
        self.pts = torch.empty((n, 3), dtype=torch.int32, device=dev)
        self.wds = torch.empty((n, 3), dtype=torch.int32, device=dev)
        self.xa = torch.empty((n,) + tuple(w3), dtype=act_dtype, device=dev)
        self.y = torch.empty((n,) + tuple(w3) if out_shape is None else tuple(out_shape), dtype=o_dtype, device=dev)
        self.graph = None
        self.kernels = 0
        self.varying_widths = False       # multi-scale patch lists (binned gathers): the width table changes per chunk too

    def run(self, pts_chunk, wds_chunk, body):
        self.pts.copy_(pts_chunk)
        if self.graph is not None and self.varying_widths:
            self.wds.copy_(wds_chunk)
        if self.graph is not None:
            self.graph.replay()
            _lib.lib().te_launch_count_add(self.kernels)
            return
        self.wds.copy_(wds_chunk)
        L = _lib.lib()
        c0 = L.te_launch_count()
        body(self.pts, self.wds, self.n, self.xa, self.y)
        self.kernels = int(L.te_launch_count() - c0)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            body(self.pts, self.wds, self.n, self.xa, self.y)
        L.te_launch_count_add(-self.kernels)             # the capture pass went through the counting entry points but enqueued nothing
        self.graph = g


class EmbeddingLearner(GenericKerasProcessor):
    """keras_processor.py:245-253."""

    def predict_embeddings(self, x, chunk_size, min_max=None, TIMEIT=False):
        raise NotImplementedError

"""Training step of the U-net segmenter on the GPU.

Replaces the Keras ``fit`` step behind ``SurfaceSegmenter.train``
(/root/reference/tomo_encoders/neural_nets/surface_segmenter.py:43-66; model compiled with ``Adam()``
and ``BinaryCrossentropy()`` at :322-323; graph ``build_Unet_3D`` neural_nets/Unet3D.py:199-291 with
``BatchNormalization(momentum=0.9, epsilon=1e-5)`` in training mode, :78).  The graph is walked here;
every operator is a C-ABI call into libtomoenc.so (include/tomoenc.h, "training operators"):
forward convolutions and their data gradients on the tcgen05 implicit-GEMM kernel, weight gradients
on mma.sync tiles, BN / LeakyReLU / pooling / loss as HBM-bound element-wise kernels.  Master
weights, Adam moments and gradients are fp32 in ONE flat buffer in Keras ``get_weights()`` order, so
that the data-parallel exchange is a single all-reduce; BN batch statistics are all-reduced as well
(``sync_bn``), which reproduces the reference's single-device batch semantics on any number of GPUs.
"""
import ctypes
import os

import numpy as np
import torch

from .. import _device, _lib, sharding
from . import _spec

BN_MOMENTUM = 0.9      # Unet3D.py:78
BN_EPS = 1e-5
LRELU_ALPHA = 0.2      # Unet3D.py:38


TeTensor = _lib.TeTensor


class _Act:
    """One NDHWC activation tensor (torch buffer) plus, lazily, the buffer of its gradient."""

    def __init__(self, n, size, c, dtype, device, te_dtype, need_grad=True):
        self.n, self.size, self.c = n, tuple(size), c
        self.t = torch.empty((n,) + self.size + (c,), dtype=dtype, device=device)
        self.g = torch.empty_like(self.t) if need_grad else None
        self.te_dtype = te_dtype

    def _desc(self, t):
        d = TeTensor()
        d.ptr, d.dtype, d.ctot, d.coff, d.C = t.data_ptr(), self.te_dtype, self.c, 0, self.c
        d.N, (d.D, d.H, d.W) = self.n, self.size
        return d

    @property
    def x(self):
        return self._desc(self.t)

    @property
    def dx(self):
        return self._desc(self.g)

    @property
    def nvox(self):
        return self.n * int(np.prod(self.size))


class _Vec:
    """A (N, n) fp32 activation of the dense part of the autoencoder (+ its gradient), described to the
    operators as an (N, 1, 1, 1, n) tensor."""

    def __init__(self, n, c, device, need_grad=True):
        self.n, self.c, self.size = n, c, (1, 1, 1)
        self.t = torch.empty((n, c), dtype=torch.float32, device=device)
        self.g = torch.empty_like(self.t) if need_grad else None

    def _desc(self, t):
        d = TeTensor()
        d.ptr, d.dtype, d.ctot, d.coff, d.C = t.data_ptr(), _lib.TE_F32, self.c, 0, self.c
        d.N, d.D, d.H, d.W = self.n, 1, 1, 1
        return d

    @property
    def x(self):
        return self._desc(self.t)

    @property
    def dx(self):
        return self._desc(self.g)

    @property
    def nvox(self):
        return self.n


class UnetTrainer:
    """Owns the weights of one U-net and runs fit steps on batches of (x, y) patch pairs.

    mode: "fp32" (CUDA-core check mode), "bf16" or "fp16" (tensor-core path; fp32 master weights)."""

    def __init__(self, model_params, patch, batch, mode="bf16", lr=1e-3, beta_1=0.9, beta_2=0.999, epsilon=1e-7,
                 sync_bn=True, process_group=None, weights=None, seed=None):
        _device.require_cuda()
        p = dict(model_params)
        p.setdefault("n_blocks", len(p["n_filters"]))
        for k, v in (("activation", "lrelu"), ("batch_norm", True), ("kern_size", 3), ("kern_size_upconv", 2)):
            if p.setdefault(k, v) != v:
                raise NotImplementedError("training is implemented for %s=%r (the value of every shipped model)" % (k, v))
        nb = p["n_blocks"]
        self.params = p
        self.pools = _spec._pools(p.get("pool_size", 2), nb)
        self.isconcat = list(p.get("isconcat") or [False] * nb)
        if any(q not in (1, 2) for q in self.pools):
            raise NotImplementedError("training supports pool sizes 1 and 2")
        self.patch = tuple(int(s) for s in patch)
        self.batch = int(batch)
        if mode not in ("fp32", "bf16"):
            raise NotImplementedError("training runs in 'bf16' (tensor cores, fp32 master weights) or 'fp32' (check mode); "
                                      "fp16 activations would need loss scaling (dL/dlogit ~ 1/n_voxels underflows)")
        self.mode = mode
        self.dev = _device.device()
        self.tdtype = {"fp32": torch.float32, "bf16": torch.bfloat16, "fp16": torch.float16}[mode]
        self.te_dtype = {"fp32": _lib.TE_F32, "bf16": _lib.TE_BF16, "fp16": _lib.TE_F16}[mode]
        self.lr, self.b1, self.b2, self.eps = lr, beta_1, beta_2, epsilon
        self.sync_bn, self.group = sync_bn, process_group
        self.step_count = 0
        self.graph_launches = 0
        self.L = _lib.lib()
        self.spec = self._make_spec(p)
        sizes = [int(np.prod(s)) for _, s in self.spec]
        self.offsets = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
        n_par = int(self.offsets[-1])
        self.theta = torch.zeros(n_par, dtype=torch.float32, device=self.dev)
        self.grad = torch.zeros_like(self.theta)
        self.m = torch.zeros_like(self.theta)
        self.v = torch.zeros_like(self.theta)
        self.set_weights(weights if weights is not None else _spec.keras_default_init(self.spec, seed))
        self._plan()

    def _make_spec(self, p):
        return _spec.unet_spec(**{k: p[k] for k in ("n_filters", "n_blocks", "activation", "batch_norm", "kern_size",
                                                    "kern_size_upconv", "isconcat", "pool_size") if k in p})

    # ------------------------------------------------------------------ weights
    def set_weights(self, weights):
        if len(weights) != len(self.spec):
            raise ValueError("expected %d weight arrays, got %d" % (len(self.spec), len(weights)))
        flat = np.concatenate([np.asarray(w, dtype=np.float32).reshape(-1) for w in weights])
        if flat.size != self.theta.numel():
            raise ValueError("weight sizes do not match the model")
        self.theta.copy_(torch.from_numpy(flat))

    def get_weights(self):
        flat = self.theta.cpu().numpy()
        return [flat[self.offsets[i]:self.offsets[i + 1]].reshape(s).copy() for i, (_, s) in enumerate(self.spec)]

    def get_gradients(self):
        flat = self.grad.cpu().numpy()
        return [flat[self.offsets[i]:self.offsets[i + 1]].reshape(s).copy() if k not in ("bn_mean", "bn_var") else None
                for i, (k, s) in enumerate(self.spec)]

    def _w(self, i):
        return self.theta[self.offsets[i]:self.offsets[i + 1]]

    def _g(self, i):
        return self.grad[self.offsets[i]:self.offsets[i + 1]]

    # ------------------------------------------------------------------ plan
    def _plan(self):
        """Layer list in Keras creation order (= spec order); every entry knows its weight indices,
        its input activation(s) and owns its output buffers."""
        p, nb, n = self.params, self.params["n_blocks"], self.batch
        mk = lambda size, c, grad=True: _Act(n, size, c, self.tdtype, self.dev, self.te_dtype, grad)  # noqa: E731
        self.inp = mk(self.patch, 1, grad=False)
        layers, wi = [], 0
        size, cur = list(self.patch), self.inp
        skips = []

        def conv(src, src2, cout):
            nonlocal wi
            lay = dict(op="conv", w=wi, src=src, src2=src2, cin=src.c + (src2.c if src2 else 0), cout=cout,
                       y=mk(src.size, cout, grad=False), a=mk(src.size, cout))
            wi += 6
            layers.append(lay)
            return lay["a"]

        for i in range(nb):
            f = p["n_filters"][i]
            cur = conv(cur, None, f)
            cur = conv(cur, None, 2 * f)
            skips.append(cur)
            if self.pools[i] > 1:
                if any(s % 2 for s in size):
                    raise ValueError("pool size must divide the spatial extent")
                size = [s // 2 for s in size]
                out = mk(size, cur.c)
                layers.append(dict(op="pool", src=cur, a=out))
                cur = out
        nf = cur.c
        cur = conv(cur, None, nf)
        cur = conv(cur, None, 2 * nf)
        for i in range(nb - 1, -1, -1):
            f = p["n_filters"][i]
            skip = None
            if self.pools[i] > 1:
                size = [s * 2 for s in size]
                up = mk(size, cur.c)
                layers.append(dict(op="upconv", w=wi, src=cur, a=up))
                wi += 2
                cur = up
                if self.isconcat[i]:
                    skip = skips[i]
            cur = conv(cur, skip, 2 * f)
            cur = conv(cur, None, 2 * f)
        layers.append(dict(op="head", w=wi, src=cur))
        wi += 2
        assert wi == len(self.spec), (wi, len(self.spec))
        self.layers = layers
        self._alloc_scratch()

    def _alloc_scratch(self):
        layers = self.layers
        maxc = max(max(l.get("cout", 1), l.get("cin", 1)) for l in layers if l["op"] != "dense")
        maxc = max([maxc] + [l["cout"] for l in layers if l["op"] == "dense"])
        # scratch
        self.bstat = torch.zeros_like(self.theta)
        self.mov_mask = torch.zeros_like(self.theta)
        self.stats = torch.zeros(2 * max(maxc, 1024), dtype=torch.float64, device=self.dev)
        wpack = max(int(self.L.te_tr_conv3_wpack_bytes(l["cin"], l["cout"])) for l in layers if l["op"] == "conv")
        wpack = max(wpack, max(int(self.L.te_tr_conv3_wpack_bytes(l["cout"], l["cin"])) for l in layers if l["op"] == "conv"))
        self.wflip = torch.empty(27 * max(l["cin"] * l["cout"] for l in layers if l["op"] == "conv"), dtype=torch.float32,
                                 device=self.dev)
        self.zero_bias = torch.zeros(maxc, dtype=torch.float32, device=self.dev)
        ups = [l for l in layers if l["op"] == "upconv"]
        self.dz2 = torch.empty(max([l["a"].t.numel() for l in ups] + [1]), dtype=self.tdtype, device=self.dev)
        for l in ups:      # packed-weight scratch must also hold the transposed-conv (8 sub-positions) and pointwise forms
            wpack = max(wpack, 8 * l["src"].c * l["a"].c * 2)
        self.wpack = torch.empty(max(wpack, 16), dtype=torch.uint8, device=self.dev)
        self.loss = torch.zeros(1, dtype=torch.float64, device=self.dev)
        self.dl = torch.empty(self.batch * int(np.prod(self.patch)), dtype=torch.float32, device=self.dev)
        self.labels = torch.empty((self.batch,) + self.patch, dtype=torch.uint8, device=self.dev)
        self.loss3 = torch.zeros(3, dtype=torch.float64, device=self.dev)
        mk = lambda size, c, grad=True: _Act(self.batch, size, c, self.tdtype, self.dev, self.te_dtype, grad)  # noqa: E731
        for lay in layers:
            if lay["op"] in ("conv", "dense"):
                c = lay["cout"]
                if lay["op"] == "conv":
                    lay["cat"] = mk(lay["src"].size, lay["cin"]) if (lay["src2"] is not None and self.mode == "fp32") else None
                lay["bn"] = {k: torch.empty(c, dtype=torch.float32, device=self.dev)
                             for k in ("rstd", "scale", "shift", "dbeta_n", "dgamma_n")}
                wi = lay["w"]
                lay["bn"]["mean"] = self.bstat[self.offsets[wi + 4]:self.offsets[wi + 5]]
                lay["bn"]["var"] = self.bstat[self.offsets[wi + 5]:self.offsets[wi + 6]]
                self.mov_mask[self.offsets[wi + 4]:self.offsets[wi + 6]] = 1.0

    def workspace_bytes(self):
        tot = 0
        for lay in self.layers:
            for k in ("y", "a"):
                if k in lay:
                    tot += lay[k].t.numel() * lay[k].t.element_size() * (2 if lay[k].g is not None else 1)
        return tot

    # ------------------------------------------------------------------ helpers
    def _call(self, name, *args):
        _lib.check(getattr(self.L, name)(*args), name)

    def _reduce_stats(self, c2):
        s = self.stats[:c2]
        if self.sync_bn:
            sharding.allreduce_sum_(s, self.group)
        return s

    # ------------------------------------------------------------------ BN + LeakyReLU helpers (conv and dense layers)
    def _bn_forward(self, y_desc, a_desc, c, nvox, wi, bn, world, st):
        """self.stats is all zero on entry and on exit (every consumer kernel re-zeroes what it read)."""
        br = ctypes.byref
        self._call("te_tr_bn_stats", br(y_desc), self.stats.data_ptr(), st)
        self._reduce_stats(2 * c)
        self._call("te_tr_bn_finalize", self.stats.data_ptr(), c, float(nvox * world), self._w(wi + 2).data_ptr(),
                   self._w(wi + 3).data_ptr(), BN_EPS, bn["mean"].data_ptr(), bn["var"].data_ptr(), bn["rstd"].data_ptr(),
                   bn["scale"].data_ptr(), bn["shift"].data_ptr(), st)
        self._call("te_tr_bn_act_fwd", br(y_desc), br(a_desc), bn["scale"].data_ptr(), bn["shift"].data_ptr(), LRELU_ALPHA, st)

    def _bn_backward(self, y_desc, da_desc, c, nvox, wi, bn, world, st):
        """da -> dy in place; parameter gradients use the LOCAL sums (the gradient all-reduce adds the
        ranks up), the normalisation backward the GLOBAL batch sums."""
        br = ctypes.byref
        self._call("te_tr_bn_act_bwd_reduce", br(y_desc), br(da_desc), bn["scale"].data_ptr(), bn["shift"].data_ptr(),
                   bn["mean"].data_ptr(), bn["rstd"].data_ptr(), LRELU_ALPHA, self.stats.data_ptr(), st)
        local = None
        if self.sync_bn and world > 1:
            local = self.stats[:2 * c].clone()
            self._reduce_stats(2 * c)
        self._call("te_tr_bn_bwd_finalize", self.stats.data_ptr(), None if local is None else local.data_ptr(), c,
                   float(nvox * world), self._g(wi + 3).data_ptr(), self._g(wi + 2).data_ptr(), bn["dbeta_n"].data_ptr(),
                   bn["dgamma_n"].data_ptr(), st)
        self._call("te_tr_bn_act_bwd_apply", br(y_desc), br(da_desc), br(da_desc), bn["scale"].data_ptr(), bn["shift"].data_ptr(),
                   bn["mean"].data_ptr(), bn["rstd"].data_ptr(), bn["dbeta_n"].data_ptr(), bn["dgamma_n"].data_ptr(),
                   LRELU_ALPHA, self.stats.data_ptr(), st)
        self._take(c, 2 * c, self._g(wi + 1), st)       # bias gradient of the conv / dense layer in front of the BN

    def _take(self, n, n_zero, dst, st):
        """dst[:n] = stats[:n] (fp64 -> fp32), stats[:n_zero] = 0."""
        self._call("te_tr_take_stats", self.stats.data_ptr(), n, n_zero, dst.data_ptr(), st)

    # ------------------------------------------------------------------ step
    def forward_backward(self, x, y=None):
        """x: (n,pz,py,px[,1]) float (already rescaled to the network's input range), y: same shape
        {0,1} (segmenter; the autoencoder reconstructs x).  Leaves the global-batch mean loss in
        self.loss and the LOCAL gradient contribution in self.grad (sum over ranks = gradient of the
        global-batch loss).  Returns the loss tensor."""
        self._load_batch(x, y)
        return self._run()

    def _load_batch(self, x, y=None):
        tx = _device.to_device(x)
        if tx.dim() == 5:
            tx = tx[..., 0]
        if tuple(tx.shape) != (self.batch,) + self.patch:
            raise ValueError("expected a batch of shape %s, got %s" % ((self.batch,) + self.patch, tuple(tx.shape)))
        self.inp.t.copy_(tx.reshape(self.inp.t.shape))
        if y is not None:
            ty = _device.to_device(y)
            if ty.dim() == 5:
                ty = ty[..., 0]
            self.labels.copy_(ty.reshape(self.labels.shape))

    def _run(self):
        """Forward + backward on the batch held in self.inp / self.labels: only stream-ordered work on static buffers
        (C-ABI launches, a few torch element-wise ops, the SyncBN all-reduces), so it can be captured in a CUDA graph."""
        st = _device.stream_ptr()
        world = sharding.world(self.group)[1] if self.sync_bn else 1
        self.grad.zero_()
        self.loss.zero_()
        br = ctypes.byref
        # ---------------- forward
        for lay in self.layers:
            op = lay["op"]
            if op == "conv":
                wi, c = lay["w"], lay["cout"]
                if lay.get("cat") is not None:          # fp32 check mode: the concatenation is materialised
                    lay["cat"].t.copy_(torch.cat([lay["src"].t, lay["src2"].t], dim=-1))
                    xin, x2 = lay["cat"], None
                else:
                    xin, x2 = lay["src"], (br(lay["src2"].x) if lay["src2"] is not None else None)
                self._call("te_tr_conv3", br(xin.x), x2, self._w(wi).data_ptr(), self._w(wi + 1).data_ptr(),
                           br(lay["y"].x), self.wpack.data_ptr(), st)
                self._bn_forward(lay["y"].x, lay["a"].x, c, lay["y"].nvox, wi, lay["bn"], world, st)
            elif op == "pool":
                self._call("te_tr_maxpool_fwd", br(lay["src"].x), br(lay["a"].x), 2, st)
            elif op == "upconv":
                wi = lay["w"]
                self._call("te_tr_upconv_fwd", br(lay["src"].x), self._w(wi).data_ptr(), self._w(wi + 1).data_ptr(),
                           br(lay["a"].x), 2, LRELU_ALPHA, self.wpack.data_ptr(), st)
            elif op == "flatten":                      # Keras Flatten of NDHWC = the memory order; dense layers run in fp32
                lay["a"].t.copy_(lay["src"].t.reshape(self.batch, -1))
            elif op == "dense":
                wi, c, src = lay["w"], lay["cout"], lay["src"]
                self._call("te_tr_dense_fwd", src.t.data_ptr(), self.batch, lay["cin"], c, self._w(wi).data_ptr(),
                           self._w(wi + 1).data_ptr(), lay["y"].t.data_ptr(), st)
                self._bn_forward(lay["y"].x, lay["a"].x, c, self.batch, wi, lay["bn"], world, st)
            elif op == "reshape":
                lay["a"].t.copy_(lay["src"].t.reshape(lay["a"].t.shape))
            elif op == "head":
                wi, src = lay["w"], lay["src"]
                n_tot = float(src.nvox * sharding.world(self.group)[1])
                self._call("te_tr_head_loss", br(src.x), self._w(wi).data_ptr(), self._w(wi + 1).data_ptr(),
                           self.labels.data_ptr(), br(src.dx), self.loss.data_ptr(), self.stats.data_ptr(), n_tot,
                           self.dl.data_ptr(), None, st)
                self._take(src.c + 1, 2 * src.c, self._g(wi), st)       # head kernel (C) and bias (1) are adjacent in the flat buffer
            elif op == "head_ae":
                wi, src = lay["w"], lay["src"]
                self.loss3.zero_()
                n_tot = float(src.nvox * sharding.world(self.group)[1])
                self._call("te_tr_head_ae_loss", br(src.x), self._w(wi).data_ptr(), self._w(wi + 1).data_ptr(),
                           self.inp.t.data_ptr(), br(src.dx), self.loss3.data_ptr(), self.stats.data_ptr(), n_tot,
                           float(self.reg_weight), self.dl.data_ptr(), None, st)
                self.loss.copy_(self.loss3[:1])
                self._take(src.c + 1, 2 * src.c, self._g(wi), st)
            else:
                raise ValueError(op)
        # ---------------- backward
        pending_pool = {}                     # id(skip activation) -> True once the decoder wrote its gradient
        # the flat gradient buffer is in forward (Keras) order and the backward pass walks the layers in reverse, so the
        # gradients are final from the tail of the buffer downwards: all-reduce them in buckets on a side stream while the
        # remaining layers' backward kernels run (data-parallel steps only)
        self._pend_hi = int(self.offsets[-1])
        prev_w = None
        for lay in reversed(self.layers):
            if prev_w is not None:
                self._grad_ready(int(self.offsets[prev_w]))
            prev_w = lay.get("w", prev_w)
            op = lay["op"]
            if op in ("head", "head_ae"):
                continue
            if op == "conv":
                wi, c, a, yb, bn = lay["w"], lay["cout"], lay["a"], lay["y"], lay["bn"]
                self._bn_backward(yb.x, a.dx, c, yb.nvox, wi, bn, world, st)
                dy = a.dx
                cin = lay["cin"]
                if lay.get("cat") is not None:
                    cat, s1, s2 = lay["cat"], lay["src"], lay["src2"]
                    self._call("te_tr_conv3_wgrad", br(cat.x), br(dy), self._g(wi).data_ptr(), cin, 0, None, st)
                    self._call("te_tr_flip_weights", self._w(wi).data_ptr(), cin, c, 0, cin, self.wflip.data_ptr(), st)
                    self._call("te_tr_conv3", br(dy), None, self.wflip.data_ptr(), self.zero_bias.data_ptr(), br(cat.dx),
                               self.wpack.data_ptr(), st)
                    s1.g.copy_(cat.g[..., :s1.c])
                    s2.g.copy_(cat.g[..., s1.c:])
                    pending_pool[id(s2)] = True
                    continue
                # weight gradients (+ bias gradient through the first call)
                off = 0
                for k, src in enumerate([lay["src"], lay["src2"]]):
                    if src is None:
                        continue
                    self._call("te_tr_conv3_wgrad", br(src.x), br(dy), self._g(wi).data_ptr(), cin, off, None, st)
                    off += src.c
                # data gradients
                off = 0
                for src in (lay["src"], lay["src2"]):
                    if src is None:
                        continue
                    if src.g is not None:
                        self._call("te_tr_flip_weights", self._w(wi).data_ptr(), cin, c, off, src.c, self.wflip.data_ptr(), st)
                        self._call("te_tr_conv3", br(dy), None, self.wflip.data_ptr(), self.zero_bias.data_ptr(), br(src.dx),
                                   self.wpack.data_ptr(), st)
                        pending_pool[id(src)] = True
                    off += src.c
            elif op == "pool":
                src = lay["src"]
                acc = 1 if pending_pool.get(id(src)) else 0
                self._call("te_tr_maxpool_bwd", br(src.x), br(lay["a"].dx), br(src.dx), 2, acc, st)
            elif op == "upconv":
                wi, src, a = lay["w"], lay["src"], lay["a"]
                dz2 = None
                if self.mode != "fp32":
                    d = TeTensor()
                    d.ptr, d.dtype, d.ctot, d.coff, d.C = self.dz2.data_ptr(), self.te_dtype, 8 * a.c, 0, 8 * a.c
                    d.N, (d.D, d.H, d.W) = src.n, src.size
                    dz2 = br(d)
                self._call("te_tr_upconv_bwd", br(src.x), br(a.x), br(a.dx), self._w(wi).data_ptr(), 2, LRELU_ALPHA,
                           self._g(wi).data_ptr(), self.stats.data_ptr(), br(src.dx), dz2, self.wpack.data_ptr(), st)
                self._take(a.c, 2 * a.c, self._g(wi + 1), st)
            elif op == "flatten":
                lay["src"].g.copy_(lay["a"].g.reshape(lay["src"].g.shape))
            elif op == "dense":
                wi, c, src, a, yb = lay["w"], lay["cout"], lay["src"], lay["a"], lay["y"]
                self._bn_backward(yb.x, a.dx, c, self.batch, wi, lay["bn"], world, st)
                self._call("te_tr_dense_bwd", src.t.data_ptr(), a.g.data_ptr(), self.batch, lay["cin"], c, self._w(wi).data_ptr(),
                           self._g(wi).data_ptr(), self.stats.data_ptr(), src.g.data_ptr() if src.g is not None else None, st)
                self._take(c, 2 * c, self._g(wi + 1), st)
            elif op == "reshape":
                lay["src"].g.copy_(lay["a"].g.reshape(lay["src"].g.shape))
        self._grad_ready(0, final=True)
        return self.loss

    GRAD_BUCKET = 1 << 19                     # elements (2 MB of fp32): M_a01's 4.77 M gradients go out in ~5 buckets

    def _grad_ready(self, lo, final=False):
        """grad[lo:] is final.  Data-parallel steps: once a bucket's worth is pending (or at the end of the backward pass)
        its all-reduce is issued on the communication stream behind an event, so it overlaps the backward kernels of the
        layers still to come; _post_backward joins the stream.  OPT-IN (TOMOENC_GRAD_BUCKETS=1); the default is one flat
        all-reduce at the end: validated at 2 ranks only (tests/test_zz_late_additions_gpu.py), and a run at 8 ranks with this AND the
        peer-memory SyncBN reducer enabled hung -- two spin-waiting mechanisms on concurrent streams."""
        if sharding.world(self.group)[1] == 1 or os.environ.get("TOMOENC_GRAD_BUCKETS", "0") != "1":
            return
        hi = self._pend_hi
        if hi <= lo or not (final or hi - lo >= self.GRAD_BUCKET):
            return
        if getattr(self, "_comm", None) is None:
            self._comm = torch.cuda.Stream()
        ev = torch.cuda.Event()
        ev.record()
        with torch.cuda.stream(self._comm):
            self._comm.wait_event(ev)
            torch.distributed.all_reduce(self.grad[lo:hi], op=torch.distributed.ReduceOp.SUM, group=self.group)
        self._pend_hi = lo
        self._bucketed = True

    def _post_backward(self):
        """Gradient all-reduce over the ranks (joined here when it went out in buckets during the backward pass) + BN
        moving statistics from the (global) batch statistics of the last forward pass (graph-capturable: no host-side
        state)."""
        if getattr(self, "_bucketed", False):
            torch.cuda.current_stream().wait_stream(self._comm)
            self._bucketed = False
        else:
            sharding.allreduce_sum_(self.grad, self.group)
        # moving = moving * momentum + batch * (1 - momentum) for every BN layer at once: the batch statistics live in
        # a buffer laid out like theta (bn["mean"] / bn["var"] are views into it), mov_mask selects those entries
        self.theta.addcmul_(self.mov_mask, self.bstat - self.theta, value=1.0 - BN_MOMENTUM)

    def _adam(self):
        self.step_count += 1
        self._call("te_tr_adam", self.theta.data_ptr(), self.grad.data_ptr(), self.m.data_ptr(), self.v.data_ptr(),
                   self.theta.numel(), self.lr, self.b1, self.b2, self.eps, self.step_count, 1.0, _device.stream_ptr())

    def apply_gradients(self):
        """All-reduces the flat gradient buffer over the ranks, updates the BN moving statistics and applies Keras
        Adam (whose update of the moving-statistics entries is exactly zero: their gradient is)."""
        self._post_backward()
        self._adam()

    # ------------------------------------------------------------------ CUDA-graph replay of the fit step
    def _graph_wanted(self):
        """TOMOENC_TRAIN_GRAPH=0 disables the replay.  With several ranks the SyncBN and gradient all-reduces (NCCL)
        are captured with the kernels (measured at 2 ranks: same loss to 5 digits as the eager step)."""
        if os.environ.get("TOMOENC_TRAIN_GRAPH", "1") == "0" or self.mode == "fp32":
            return False
        if sharding.world(self.group)[1] > 1 and torch.distributed.get_backend(self.group) != "nccl":
            return False
        return True

    def _capture(self):
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        n0 = int(self.L.te_launch_count())
        with torch.cuda.graph(g):
            self._run()
            self._post_backward()
        self.graph_launches = int(self.L.te_launch_count()) - n0      # library kernels replayed per step
        self._graph = g

    def train_step(self, x, y=None):
        """One fit step on a batch; returns the (global-batch mean) loss as a python float.  After two eager steps
        (allocations, kernel attributes, communicators) the ~250 launches of forward + backward + all-reduces are
        replayed as ONE CUDA graph; Adam (host-side step count) stays outside."""
        self._load_batch(x, y)
        if not self._graph_wanted():
            self._run()
            self.apply_gradients()
        else:
            if getattr(self, "_graph", None) is None and getattr(self, "_eager_steps", 0) >= 2:
                self._capture()
            if getattr(self, "_graph", None) is not None:
                self._graph.replay()
                self._adam()
            else:
                self._eager_steps = getattr(self, "_eager_steps", 0) + 1
                self._run()
                self.apply_gradients()
        loss = sharding.allreduce_sum_(self.loss.clone(), self.group)
        return float(loss.item())


class CAETrainer(UnetTrainer):
    """Fit steps of the convolutional autoencoder (``SelfSupervisedCAE.train``,
    /root/reference/tomo_encoders/neural_nets/autoencoders.py:674-718; custom step
    ``RegularizedAutoencoder.train_step`` :498-526): encoder ``build_encoder_r`` (:246-347) -> decoder
    ``build_decoder_r`` (:350-447), loss = mean squared error + weight * KL(x || decoded) (weight 1/250),
    Adam.  The flat weight buffer holds the encoder's arrays followed by the decoder's, both in Keras
    order.  Conv layers as in UnetTrainer; the dense layers (a few KB per sample) run in fp32."""

    def __init__(self, model_params, model_size, batch, reg_weight=1.0 / 250.0, regularization_type="kl", **kwargs):
        if regularization_type != "kl":
            raise NotImplementedError("only the 'kl' regularisation (the reference's default) is implemented")
        self.reg_weight = float(reg_weight)
        super().__init__(model_params, model_size, batch, **kwargs)

    def _spec_params(self, p):
        keys = ("n_filters", "n_blocks", "activation", "batch_norm", "kern_size", "kern_size_upconv", "hidden_units",
                "pool_size", "POOL_FLAG")
        return {k: p[k] for k in keys if k in p}

    def _make_spec(self, p):
        if "hidden_units" not in p:
            raise ValueError("hidden_units is required")
        if any(q != 2 for q in self.pools):
            raise NotImplementedError("autoencoder training supports pool size 2")
        enc = _spec.encoder_spec(self.patch, **self._spec_params(p))
        dec = _spec.decoder_spec(self.patch, **self._spec_params(p))
        self.n_enc = len(enc)
        return enc + dec

    def _plan(self):
        p, nb, n = self.params, self.params["n_blocks"], self.batch
        mk = lambda size, c, grad=True: _Act(n, size, c, self.tdtype, self.dev, self.te_dtype, grad)  # noqa: E731
        hidden = [int(h) for h in p["hidden_units"]]
        pool_flag = bool(p.get("POOL_FLAG", True))
        self.inp = mk(self.patch, 1, grad=False)
        layers, wi = [], 0
        size, cur = list(self.patch), self.inp

        def conv(src, cout):
            nonlocal wi
            lay = dict(op="conv", w=wi, src=src, src2=None, cin=src.c, cout=cout, y=mk(src.size, cout, grad=False),
                       a=mk(src.size, cout))
            wi += 6
            layers.append(lay)
            return lay["a"]

        def pool(src):
            nonlocal size
            if any(s % 2 for s in size):
                raise ValueError("pool size must divide the spatial extent")
            size = [s // 2 for s in size]
            out = mk(size, src.c)
            layers.append(dict(op="pool", src=src, a=out))
            return out

        def dense(src, cout):
            nonlocal wi
            lay = dict(op="dense", w=wi, src=src, cin=src.c, cout=cout, y=_Vec(n, cout, self.dev, need_grad=False),
                       a=_Vec(n, cout, self.dev))
            wi += 6
            layers.append(lay)
            return lay["a"]

        def upconv(src):
            nonlocal wi, size
            size = [s * 2 for s in size]
            up = mk(size, src.c)
            layers.append(dict(op="upconv", w=wi, src=src, a=up))
            wi += 2
            return up

        # ---- encoder (autoencoders.py:297-345)
        for i in range(nb):
            f = p["n_filters"][i]
            cur = conv(cur, f)
            cur = conv(cur, 2 * f)
            cur = pool(cur)
        if pool_flag:
            cur = pool(cur)
        pre_size, pre_c = list(size), cur.c
        flat = _Vec(n, pre_c * int(np.prod(size)), self.dev)
        layers.append(dict(op="flatten", src=cur, a=flat))
        vec = flat
        for h in hidden:
            vec = dense(vec, h)
        self.latent_layer = layers[-1]
        assert wi == self.n_enc, (wi, self.n_enc)
        # ---- decoder (autoencoders.py:407-447)
        for h in hidden[::-1][1:]:
            vec = dense(vec, h)
        vec = dense(vec, flat.c)
        cur = mk(pre_size, pre_c)
        layers.append(dict(op="reshape", src=vec, a=cur))
        size = list(pre_size)
        if pool_flag:
            cur = upconv(cur)
        for i in range(nb - 1, -1, -1):
            f = p["n_filters"][i]
            cur = upconv(cur)
            cur = conv(cur, f)
            cur = conv(cur, f)
        layers.append(dict(op="head_ae", w=wi, src=cur))
        wi += 2
        assert wi == len(self.spec), (wi, len(self.spec))
        self.layers = layers
        self._alloc_scratch()

    def get_model_weights(self):
        """(encoder weights, decoder weights), each in Keras get_weights() order."""
        w = self.get_weights()
        return w[:self.n_enc], w[self.n_enc:]

    def losses(self):
        """(total, pixel mse, KL) of the last step, local share of the global-batch means."""
        return tuple(float(v) for v in self.loss3.cpu())

"""``Voids`` and ``label``: the labelling step that follows the stitched segmentation.

Drop-in for the data-structure part of /root/reference/tomo_encoders/structures/voids.py:76-373 (``count_voids``,
``import_from_grid``, ``export_grid`` and the selection helpers) and for ``cupyx.scipy.ndimage.label`` as called at
/root/reference/tomo_encoders/tasks/digital_zoom.py:41.  The voxel work -- connected components, bounding boxes and
sizes (``scipy.ndimage.find_objects``), component masks, painting the masks back and selecting the grid tiles that
contain a void -- runs in libtomoenc.so (csrc/label_ops.cu) through the C ABI; the per-void bookkeeping (lists of a few
thousand boxes) stays in numpy, as in the reference.  Meshing / plotting / disk formats of the reference class
(marching cubes, open3d, tifffile, h5py) are outside the hot path and not provided.
"""
import numpy as np
import torch

from .. import _device, _lib
from .grid import Grid


def _conn_of(structure):
    if structure is None:
        return 6                                           # scipy's default: faces only
    s = np.asarray(structure.cpu() if isinstance(structure, torch.Tensor) else structure)
    if s.shape == (3, 3, 3):
        if np.all(s != 0):
            return 26
        face = np.zeros((3, 3, 3), bool)
        face[1, 1, :] = face[1, :, 1] = face[:, 1, 1] = True
        if np.array_equal(s != 0, face):
            return 6
    raise NotImplementedError("label: structure must be None / the 6-neighbour cross or np.ones((3,3,3))")


def _label_device(t_vol, conn):
    """t_vol: contiguous CUDA tensor (Z,Y,X) -> (labels int32 CUDA tensor, n python int)."""
    shape = tuple(int(s) for s in t_vol.shape)
    nvox = int(np.prod(shape))
    labels = torch.empty(shape, dtype=torch.int32, device=t_vol.device)
    n_dev = torch.zeros(1, dtype=torch.int64, device=t_vol.device)
    L = _lib.lib()
    ws = torch.empty(max(int(L.te_label_workspace_bytes(nvox)), 16), dtype=torch.uint8, device=t_vol.device)
    _lib.check(L.te_label(t_vol.data_ptr(), _device.te_dtype(t_vol), _lib.i64x3(shape), conn, labels.data_ptr(), n_dev.data_ptr(),
                          ws.data_ptr(), _device.stream_ptr()), "te_label")
    return labels, int(n_dev.item())


def label(input, structure=None, output=None):
    """scipy.ndimage.label / cupyx.scipy.ndimage.label for 3-D arrays: (labels int32, number of components).
    Numpy in -> numpy out, CUDA tensor in -> CUDA tensor out."""
    if input.ndim != 3:
        raise NotImplementedError("label: 3-D volumes only")
    conn = _conn_of(structure)
    t = _device.to_device(input)
    if t.dtype == torch.bool:
        t = t.to(torch.uint8)
    elif t.dtype in (torch.int32, torch.int64, torch.float64):
        t = (t != 0).to(torch.uint8)
    labels, n = _label_device(t, conn)
    res = _device.to_host_like(labels, input)
    if output is not None:
        output[...] = res
        return n
    return res, n


def _stats(t_lab, n):
    """(bbox (n,6) int32 numpy: z0,y0,x0,z1,y1,x1; sizes (n,) int64 numpy) of labels 1..n."""
    shape = tuple(int(s) for s in t_lab.shape)
    bbox = torch.empty((max(n, 1), 6), dtype=torch.int32, device=t_lab.device)
    sizes = torch.empty(max(n, 1), dtype=torch.int64, device=t_lab.device)
    _lib.check(_lib.lib().te_label_stats(t_lab.data_ptr(), _lib.i64x3(shape), n, bbox.data_ptr(), sizes.data_ptr(),
                                         _device.stream_ptr()), "te_label_stats")
    return bbox[:n].cpu().numpy(), sizes[:n].cpu().numpy()


def _crop_masks(t_lab, boxes, ids):
    """boxes (m,6) int, ids (m,) int -> list of m numpy bool arrays (labels[box] == id), one device pass + one D2H."""
    m = len(boxes)
    if m == 0:
        return []
    boxes = np.ascontiguousarray(boxes, dtype=np.int32)
    ext = (boxes[:, 3:] - boxes[:, :3]).astype(np.int64)
    nv = ext.prod(axis=1)
    offs = np.concatenate([[0], np.cumsum(nv)]).astype(np.int64)
    dev = t_lab.device
    out = torch.empty(max(int(offs[-1]), 1), dtype=torch.uint8, device=dev)
    d_boxes = torch.from_numpy(boxes).to(dev)
    d_ids = torch.from_numpy(np.ascontiguousarray(ids, dtype=np.int32)).to(dev)
    d_offs = torch.from_numpy(offs).to(dev)
    shape = tuple(int(s) for s in t_lab.shape)
    L = _lib.lib()
    for a in range(0, m, 65535):
        b = min(m, a + 65535)
        _lib.check(L.te_label_crop_masks(t_lab.data_ptr(), _lib.i64x3(shape), d_boxes[a:b].data_ptr(), d_ids[a:b].data_ptr(), b - a,
                                         d_offs[a:b].data_ptr(), int(nv[a:b].max()), out.data_ptr(), _device.stream_ptr()),
                   "te_label_crop_masks")
    host = out.cpu().numpy().view(np.bool_)
    return [host[offs[i]:offs[i + 1]].reshape(tuple(ext[i])) for i in range(m)]


class Voids(dict):
    """voids.py:76-373."""

    def __init__(self):
        self["sizes"] = []
        self["cents"] = []
        self["cpts"] = []
        self["s_voids"] = []
        self["x_voids"] = []
        self["x_boundary"] = []
        self["surfaces"] = []
        self.vol_shape = (None, None, None)
        self.b = 1
        self.pad_bb = 0
        self.marching_cubes_algo = "skimage"

    def __len__(self):
        return len(self["x_voids"])

    def transform_linear_shift(self, shift):
        """voids.py:95-104."""
        shift = np.asarray(shift)
        self["cents"] = np.asarray(self["cents"]) + shift
        self["cpts"] = np.asarray(self["cpts"]) + shift
        self["s_voids"] = [tuple(slice(s[i].start + shift[i], s[i].stop + shift[i]) for i in range(3)) for s in self["s_voids"]]

    # ------------------------------------------------------------------ labelled map -> list of voids
    def count_voids(self, V_lab, b, dust_thresh, boundary_loc=(0, 0, 0)):
        """voids.py:263-303.  V_lab: label image (numpy or CUDA tensor, any integer dtype) as returned by ``label``."""
        self.dust_thresh = dust_thresh
        self.vol_shape = tuple(int(s) for s in V_lab.shape)
        t_lab = _device.to_device(V_lab)
        if t_lab.dtype != torch.int32:
            t_lab = t_lab.to(torch.int32)
        n = int(t_lab.max().item()) if t_lab.numel() else 0
        boundary_id = int(t_lab[tuple(int(v) for v in boundary_loc)].item()) if t_lab.numel() else 0
        bbox, sizes = _stats(t_lab, n)
        cpt_all, ept_all = bbox[:, :3].astype(np.int64), bbox[:, 3:].astype(np.int64)
        present = sizes > 0                              # find_objects yields None for missing labels; the reference never meets one
        keep = present & np.all(np.clip(ept_all - cpt_all - dust_thresh, 0, None) != 0, axis=1)
        vs = np.asarray(self.vol_shape, dtype=np.int64)
        lo = np.maximum(cpt_all - self.pad_bb, 0)
        hi = np.minimum(ept_all + self.pad_bb, vs)
        idx = np.nonzero(keep)[0]
        masks = _crop_masks(t_lab, np.concatenate([lo[idx], hi[idx]], axis=1), idx + 1)
        self["sizes"], self["cents"], self["cpts"], self["x_voids"], self["s_voids"] = [], [], [], [], []
        for k, i in enumerate(idx):
            if i + 1 == boundary_id:
                self["x_boundary"] = masks[k].copy()
                continue
            self["sizes"].append(int(sizes[i]) if self.pad_bb == 0 else int(masks[k].sum()))
            self["cents"].append(list((cpt_all[i] + ept_all[i]) // 2))
            self["cpts"].append(list(cpt_all[i]))
            self["x_voids"].append(masks[k])
            self["s_voids"].append(tuple(slice(int(lo[i, a]), int(hi[i, a])) for a in range(3)))
        self["sizes"] = np.asarray(self["sizes"])
        self["cents"] = np.asarray(self["cents"])
        self["cpts"] = np.asarray(self["cpts"])
        self.b = b
        self.n_voids = len(self["sizes"])
        print(f"\tSTAT: voids found - {self.n_voids}")
        return self

    # ------------------------------------------------------------------ refined masks from a segmented grid
    def import_from_grid(self, voids_b, x_grid, p_grid):
        """voids.py:188-260: stitch the segmented tiles, cut every (up-scaled) bounding box out of the stitched volume and
        keep the largest 26-connected component inside it."""
        dev = _device.device()
        x_dev = _device.to_device(x_grid)
        Vp = torch.zeros(tuple(int(s) for s in p_grid.vol_shape), dtype=torch.uint8, device=dev)
        p_grid.fill_patches_in_volume(x_dev, Vp)
        self["sizes"], self["cents"], self["cpts"], self["s_voids"], self["x_voids"] = [], [], [], [], []
        self.vol_shape = p_grid.vol_shape
        self.b = 1
        b = voids_b.b
        for s_b in voids_b["s_voids"]:
            s = tuple(slice(s_b[i].start * b, s_b[i].stop * b) for i in range(3))
            crop = Vp[s].contiguous()
            self["s_voids"].append(s)
            self["cents"].append([int((s[i].start + s[i].stop) // 2) for i in range(3)])
            self["cpts"].append([int(s[i].start) for i in range(3)])
            lab, n_objs = _label_device(crop, 26)
            if n_objs > 1:
                _, counts = _stats(lab, n_objs)
                i_main = int(np.argmax(counts))
                void = _crop_masks(lab, np.asarray([[0, 0, 0] + list(lab.shape)]), [i_main + 1])[0].astype(np.uint8)
            else:
                void = (crop > 0).to(torch.uint8).cpu().numpy()
            self["sizes"].append(np.sum(void))
            self["x_voids"].append(void)
        self["sizes"] = np.asarray(self["sizes"])
        self["cents"] = np.asarray(self["cents"])
        self["cpts"] = np.asarray(self["cpts"])
        for key in voids_b.keys():
            if key not in self.keys():
                self[key] = voids_b[key]
        return self

    # ------------------------------------------------------------------ selections (host bookkeeping, voids.py:305-354)
    def select_by_indices(self, idxs):
        self["x_voids"] = [self["x_voids"][ii] for ii in idxs]
        self["sizes"] = self["sizes"][idxs]
        self["cents"] = self["cents"][idxs]
        self["cpts"] = self["cpts"][idxs]
        self["s_voids"] = [self["s_voids"][ii] for ii in idxs]

    def select_by_size(self, size_thresh_um, pixel_size_um=1.0, sel_type="geq"):
        if size_thresh_um <= 0:
            return
        size_thresh = size_thresh_um / (self.b * pixel_size_um)
        idxs = np.arange(len(self["sizes"]))
        cond_list = np.asarray([1 if self["sizes"][idx] >= size_thresh ** 3 else 0 for idx in idxs])
        if sel_type == "leq":
            cond_list = cond_list ^ 1
        idxs = idxs[cond_list == 1]
        print(f"\tSTAT: size thres: {size_thresh:.2f} pixel length for {size_thresh_um:.2f} um threshold")
        self.select_by_indices(idxs)

    def sort_by_size(self, reverse=False):
        idxs = np.argsort(self["sizes"])
        if reverse:
            idxs = idxs[::-1]
        self.select_by_indices(idxs)

    def select_around_void(self, void_id, radius_um, pixel_size_um=1.0):
        radius_pix = radius_um / (self.b * pixel_size_um)
        idxs = np.arange(len(self["sizes"]))
        dist = np.linalg.norm((self["cents"] - self["cents"][void_id]), ord=2, axis=1)
        self["distance_%i" % void_id] = dist
        self.select_by_indices(idxs[dist < radius_pix])

    # ------------------------------------------------------------------ voids -> sparse grid for the next pass
    def paint(self):
        """The uint8 volume ``export_grid`` builds (voids.py:358-361): every void's mask written into its bounding box, in
        order (later boxes overwrite earlier ones).  CUDA tensor."""
        dev = _device.device()
        shape = tuple(int(s) for s in self.vol_shape)
        V = torch.zeros(shape, dtype=torch.uint8, device=dev)
        m = len(self["s_voids"])
        if m == 0:
            return V
        boxes = np.asarray([[s[0].start, s[1].start, s[2].start, s[0].stop, s[1].stop, s[2].stop] for s in self["s_voids"]], dtype=np.int32)
        nv = (boxes[:, 3:] - boxes[:, :3]).astype(np.int64).prod(axis=1)
        offs = np.concatenate([[0], np.cumsum(nv)]).astype(np.int64)
        flat = np.concatenate([np.ascontiguousarray(x, dtype=np.uint8).reshape(-1) for x in self["x_voids"]])
        d_masks, d_offs, d_boxes = torch.from_numpy(flat).to(dev), torch.from_numpy(offs).to(dev), torch.from_numpy(boxes).to(dev)
        owner = torch.empty(shape, dtype=torch.int32, device=dev)
        L = _lib.lib()
        for a in range(0, m, 65535):       # chunks keep their order: a later chunk's boxes overwrite the earlier chunks'
            b = min(m, a + 65535)
            _lib.check(L.te_paint_boxes(d_masks.data_ptr(), d_offs[a:b].data_ptr(), d_boxes[a:b].data_ptr(), b - a, int(nv[a:b].max()),
                                        V.data_ptr(), _lib.i64x3(shape), owner.data_ptr(), _device.stream_ptr()), "te_paint_boxes")
        return V

    def export_grid(self, wd):
        """voids.py:356-373: the tiles of Grid(vol_shape, wd // b) that contain void voxels, rescaled to full resolution."""
        wd = wd // self.b
        V_bin = self.paint()
        p3d = Grid(tuple(V_bin.shape), width=wd)
        flags = torch.zeros(max(len(p3d), 1), dtype=torch.uint8, device=V_bin.device)
        _lib.check(_lib.lib().te_tile_any(V_bin.data_ptr(), 1, _lib.i64x3(tuple(V_bin.shape)), wd, flags.data_ptr(),
                                          _device.stream_ptr()), "te_tile_any")
        is_sel = flags[:len(p3d)].cpu().numpy() > 0
        p3d_sel = p3d.filter_by_condition(is_sel)
        if self.b > 1:
            p3d_sel = p3d_sel.rescale(self.b)
        r_fac = len(p3d_sel) * (wd ** 3) / np.prod(p3d_sel.vol_shape)
        print(f"\tSTAT: 1/r value: {1 / r_fac:.4g}")
        return p3d_sel, r_fac

"""Builds oracle/_ref/librec_ref.so: the reference's OWN back-projection kernels compiled for the host.  TEST
INFRASTRUCTURE ONLY (checker for oracle/recon_np.py and generator of tests/golden/recon_golden.npz).

The kernel source is the string literal `source` of /root/reference/tomo_encoders/reconstruction/cuda_kernels.py:21-150
(CUDA C handed to cupy.RawModule).  It is read from where it lies at build time, wrapped between a thread-index shim
(blockDim / blockIdx / threadIdx as thread-local variables, `__global__` empty, `__cosf/__sinf` -> cosf/sinf) and host
launchers that walk the reference's launch shapes (cuda_kernels.py:176-177, 205-210, 233-236, 260-263, 289-292), and piped
to g++ on stdin: nothing of the reference is written into the repository, only the .so under oracle/_ref/ (git-ignored).
-ffp-contract=off keeps one rounding per operation, which is what oracle/recon_np.py restates.
"""
import ctypes
import os
import re
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("TOMOENC_REFERENCE_ROOT", "/root/reference")
REF_FILE = os.path.join(REF_ROOT, "tomo_encoders", "reconstruction", "cuda_kernels.py")
OUT_DIR = os.path.join(HERE, "_ref")
LIB = os.path.join(OUT_DIR, "librec_ref.so")

_PRE = r"""
#include <math.h>
struct te_dim3 { int x, y, z; };
static thread_local te_dim3 blockDim, blockIdx, threadIdx;
#define __global__
#define __cosf cosf
#define __sinf sinf
"""

_POST = r"""
#define TE_FOR_GRID(gx, gy, gz, bx, by, bz, CALL)                                   \
  blockDim.x = bx; blockDim.y = by; blockDim.z = bz;                                \
  for (int Bz = 0; Bz < (gz); ++Bz) for (int By = 0; By < (gy); ++By) for (int Bx = 0; Bx < (gx); ++Bx)   \
  for (int Tz = 0; Tz < (bz); ++Tz) for (int Ty = 0; Ty < (by); ++Ty) for (int Tx = 0; Tx < (bx); ++Tx) {  \
    blockIdx.x = Bx; blockIdx.y = By; blockIdx.z = Bz; threadIdx.x = Tx; threadIdx.y = Ty; threadIdx.z = Tz; \
    CALL;                                                                           \
  }
static int cdiv(int a, int b) { return (a + b - 1) / b; }
extern "C" {
void launch_rec_pts(float* f, float* g, float* theta, int* pts, float center, int ntheta, int nz, int n, int npts) {
  TE_FOR_GRID(cdiv(npts, 1024), 1, 1, 1024, 1, 1, rec_pts(f, g, theta, pts, center, ntheta, nz, n, npts))
}
void launch_rec_pts_xy(float* f, float* g, float* theta, int* pts, float center, int ntheta, int nz, int n, int npts) {
  TE_FOR_GRID(cdiv(npts, 256), cdiv(nz, 4), 1, 256, 4, 1, rec_pts_xy(f, g, theta, pts, center, ntheta, nz, n, npts))
}
void launch_rec(float* f, float* g, float* theta, float center, int ntheta, int nz, int n, int stx, int px, int sty, int py,
                int stz, int pz) {
  TE_FOR_GRID(cdiv(px, 16), cdiv(py, 16), cdiv(pz, 4), 16, 16, 4,
              rec(f, g, theta, center, ntheta, nz, n, stx, px, sty, py, stz, pz))
}
void launch_rec_mask(float* f, float* g, float* theta, float center, int ntheta, int nz, int n) {
  TE_FOR_GRID(cdiv(n, 16), cdiv(n, 16), cdiv(nz, 4), 16, 16, 4, rec_mask(f, g, theta, center, ntheta, nz, n))
}
void launch_rec_all(float* f, float* g, float* theta, float center, int ntheta, int nz, int n) {
  TE_FOR_GRID(cdiv(n, 16), cdiv(n, 16), cdiv(nz, 4), 16, 16, 4, rec_all(f, g, theta, center, ntheta, nz, n))
}
}
"""


def reference_present():
    return os.path.isfile(REF_FILE)


def built():
    return os.path.isfile(LIB)


def build(force=False):
    """Compiles the library when /root/reference is mounted (this container); on the GPU box the prebuilt file is used."""
    if built() and not force:
        return LIB
    if not reference_present():
        raise FileNotFoundError("reference not mounted and %s not prebuilt" % LIB)
    text = open(REF_FILE).read()
    m = re.search(r'source\s*=\s*"""(.*?)"""', text, re.S)
    if m is None:
        raise RuntimeError("kernel source string not found in %s" % REF_FILE)
    os.makedirs(OUT_DIR, exist_ok=True)
    unit = _PRE + m.group(1) + _POST
    cmd = ["g++", "-x", "c++", "-", "-O2", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", "-o", LIB]
    r = subprocess.run(cmd, input=unit, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("g++ failed on the reference kernels:\n" + r.stderr)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
    return _lib


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def rec_all(data, theta, center):
    data, theta = _f(data), _f(theta)
    ntheta, nz, n = data.shape
    obj = np.zeros((nz, n, n), np.float32)
    lib().launch_rec_all(_p(obj), _p(data), _p(theta), ctypes.c_float(center), ntheta, nz, n)
    return obj


def rec_mask(obj, data, theta, center):
    data, theta = _f(data), _f(theta)
    ntheta, nz, n = data.shape
    obj = _f(obj).copy()
    lib().launch_rec_mask(_p(obj), _p(data), _p(theta), ctypes.c_float(center), ntheta, nz, n)
    return obj


def rec_patch(data, theta, center, stx, px, sty, py, stz, pz):
    data, theta = _f(data), _f(theta)
    ntheta, nz, n = data.shape
    obj = np.zeros((pz, py, px), np.float32)
    lib().launch_rec(_p(obj), _p(data), _p(theta), ctypes.c_float(center), ntheta, nz, n, int(stx), int(px), int(sty), int(py),
                     int(stz), int(pz))
    return obj


def rec_pts(data, theta, center, pts):
    data, theta = _f(data), _f(theta)
    ntheta, nz, n = data.shape
    pts = np.ascontiguousarray(pts, dtype=np.int32)
    obj = np.zeros(len(pts), np.float32)
    lib().launch_rec_pts(_p(obj), _p(data), _p(theta), _p(pts), ctypes.c_float(center), ntheta, nz, n, len(pts))
    return obj


def rec_pts_xy(data, theta, center, pts):
    data, theta = _f(data), _f(theta)
    ntheta, nz, n = data.shape
    pts = np.ascontiguousarray(pts, dtype=np.int32)
    obj = np.zeros(len(pts) * nz, np.float32)
    lib().launch_rec_pts_xy(_p(obj), _p(data), _p(theta), _p(pts), ctypes.c_float(center), ntheta, nz, n, len(pts))
    return obj


if __name__ == "__main__":
    print(build(force=True))

// Filtered back-projection next to the hot path: B200 kernels behind the reference's
// reconstruction API -- /root/reference/tomo_encoders/reconstruction/cuda_kernels.py:21-160 (rec_pts, rec_pts_xy, rec,
// rec_mask, rec_all), prep.py:21-77 (calc_padding, fbp_filter, preprocess), recon.py:135-306 (recon_patches_3d,
// extract_segmented, make_mask).
//
// Per voxel (x, y, z) every reference kernel evaluates, in fp32 and in this order, for k = 0 .. ntheta-1:
//     sp = (x - n/2) * cos(theta[k]) - (y - n/2) * sin(theta[k]) + center
//     s0 = roundf(sp)                                   (half away from zero)
//     if (0 <= s0 < n-1)  f += g[k][z][s0] + (g[k][z][s0+1] - g[k][z][s0]) * (sp - s0) / n
// (the "/ n" and the round-instead-of-floor are the reference's; they are kept).  sp, s0 and the weight do not depend on
// z, so the design here is column-wise:
//   * te_bp_prepare re-lays the filtered projections out as gT[zb][k][q][s][4] (z = 16 zb + 4 q + j): a voxel column
//     reads eight 128-bit words per angle (4 z quads x {s0, s0 + 1}) instead of 32 scalars, and the trigonometry is
//     tabulated once per angle (cos / sin evaluated in fp64, rounded to fp32);
//   * one thread owns CPT (x, y) columns x 16 z accumulators; a warp covers 8 x 4 columns, whose detector samples for
//     one angle span <= ~10 consecutive s: with s next to the quad in memory one warp-wide load touches 1-2 cache
//     lines.  (The first layout, [s][16 z], touched ~5 lines per load and ran at the L1 wavefront rate: 2 cycles per
//     line and load, 5.6 updates / clk / SM measured; see profiles/r01_notes.md.)  Per column and angle the address /
//     weight arithmetic is amortised over 16 voxels;
//   * the same kernel body serves the whole volume, a mask (rec_mask), one box (rec), a list of patches (the fused
//     recon -> clip -> rescale -> network-input path of recon_patches_3d + extract_segmented, which never materialises
//     the tomogram) and point lists.
// EXACT = true evaluates the expression with one IEEE rounding per reference operation (no FMA contraction, true
// division): bit-exact against oracle/recon_np.py and the reference kernels compiled for the host (oracle/_ref).
// EXACT = false (the product default) contracts to FMAs and multiplies by 1/n: what nvcc/NVRTC do to the reference
// source, up to their choice of contraction.  The bound of this kernel is the fp32 / issue pipe (3 lane-ops per
// voxel-angle), not HBM: the projections of one z block are read once per wave of CTAs and live in L2.
#include <cufft.h>

#include <map>
#include <mutex>

#include "te_common.cuh"

namespace te {
namespace {

constexpr int kZT = 16;          // z values per column block
constexpr int kTrigChunk = 1024;  // angles staged in shared memory at a time

struct BpParams {
  const float* gT;     // [nzb][ntheta + 1][4][n][4]
  const float2* trig;  // [ntheta] (cos, sin)
  int ntheta, nz, n, nzb;
  float center;
  // volume / box output: out[(tz*py + ty)*px + tx] for voxel (stx+tx, sty+ty, stz+tz)
  float* out;
  int stx, px, sty, py, stz, pz;
  int masked;  // rec_mask: only voxels with out > 0 are reconstructed, the others become 0
  // patch list
  const uint32_t* points;  // (npatch, 3) z, y, x
  long long npatch;
  int wd, tiles_x, tiles_y, zblocks;  // per patch
  void* pout;
  int out_kind, normalize;  // out_kind: 0 float32, 1 float16, 2 bfloat16
  float lo, hi;
  // point lists
  const int* pts;
  int npts;
};

// One angle step of CPT columns: all 8 * CPT 128-bit loads are issued before the arithmetic (16 independent loads in
// flight per thread at CPT = 2).  Layout: float4 index ((k * 4 + q) * n + s) inside the z block's slice, q = z quad; the
// slice ends with one all-zero "angle" (k = ntheta) that out-of-range samples are redirected to, so the loop has no branch:
// g0 = g1 = 0 contributes exactly 0.
template <int CPT, bool EXACT>
__device__ __forceinline__ void bp_columns(const BpParams& p, const float2* __restrict__ trig_s, int k0, int kn,
                                           const float4* __restrict__ gz4, const float (&fx)[CPT], const float (&fy)[CPT],
                                           const bool (&active)[CPT], float (&acc)[CPT][kZT]) {
  const float nf = static_cast<float>(p.n);
  const float inv_n = 1.0f / nf;
  const int n = p.n, zero_off = p.ntheta * 4 * n;
  int kofs = k0 * 4 * n;
#pragma unroll 1
  for (int kk = 0; kk < kn; ++kk, kofs += 4 * n) {
    const float2 cs = trig_s[kk];
    int off[CPT];
    float w[CPT];
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
      float sp;
      if (EXACT) sp = __fadd_rn(__fsub_rn(__fmul_rn(fx[c], cs.x), __fmul_rn(fy[c], cs.y)), p.center);
      else sp = fmaf(fx[c], cs.x, -(fy[c] * cs.y)) + p.center;
      const float s0f = roundf(sp);
      const int s0 = static_cast<int>(s0f);
      const bool ok = active[c] && s0 >= 0 && s0 < n - 1;
      off[c] = ok ? kofs + s0 : zero_off;
      w[c] = __fsub_rn(sp, s0f);
      if (!EXACT) w[c] *= inv_n;
    }
    float4 a[CPT][4], b[CPT][4];
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        a[c][q] = __ldg(gz4 + off[c] + q * n);
        b[c][q] = __ldg(gz4 + off[c] + q * n + 1);
      }
    }
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
      const float* g0 = reinterpret_cast<const float*>(a[c]);
      const float* g1 = reinterpret_cast<const float*>(b[c]);
      if (EXACT) {
#pragma unroll
        for (int i = 0; i < kZT; ++i)
          acc[c][i] = __fadd_rn(acc[c][i], __fadd_rn(g0[i], __fdiv_rn(__fmul_rn(__fsub_rn(g1[i], g0[i]), w[c]), nf)));
      } else {
#pragma unroll
        for (int i = 0; i < kZT; ++i) acc[c][i] += fmaf(g1[i] - g0[i], w[c], g0[i]);
      }
    }
  }
}

enum { MODE_VOLUME = 0, MODE_PATCHES = 1, MODE_PTS_XY = 2 };

// 256 threads: warp w covers an 8 x 4 block of columns, the CTA 32 x 8, CPT columns per thread stacked in y (32 x 8*CPT).
template <int MODE, int CPT, bool EXACT>
__global__ void __launch_bounds__(256, 2) bp_kernel(const BpParams p) {
  __shared__ float2 trig_s[kTrigChunk];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int lx = (warp & 3) * 8 + (lane & 7), ly = (warp >> 2) * 4 + (lane >> 3);

  int x[CPT], y[CPT];     // voxel coordinates of the columns
  bool valid[CPT];
  int zb;                 // z block of gT
  int zlo, zhi;           // voxel z range this CTA may store: [zlo, zhi)
  long long obase[CPT];   // output index of (column, z = 0 of the output box)
  long long zstride;      // output stride between z planes
  int zorg;               // voxel z of output plane 0

  if (MODE == MODE_VOLUME) {
    zb = p.stz / kZT + blockIdx.z;
    zlo = max(p.stz, zb * kZT);
    zhi = min(p.stz + p.pz, zb * kZT + kZT);
    zstride = static_cast<long long>(p.px) * p.py;
    zorg = p.stz;
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
      const int tx = blockIdx.x * 32 + lx, ty = (blockIdx.y * CPT + c) * 8 + ly;
      valid[c] = tx < p.px && ty < p.py;
      x[c] = p.stx + tx;
      y[c] = p.sty + ty;
      obase[c] = static_cast<long long>(ty) * p.px + tx;
    }
  } else if (MODE == MODE_PATCHES) {
    const int per = p.tiles_x * p.tiles_y * p.zblocks;
    const long long ip = blockIdx.x / per;
    int r = static_cast<int>(blockIdx.x - ip * per);
    const int tzb = r / (p.tiles_x * p.tiles_y);
    r -= tzb * p.tiles_x * p.tiles_y;
    const int tyb = r / p.tiles_x, txb = r - tyb * p.tiles_x;
    const int pz0 = static_cast<int>(p.points[ip * 3 + 0]), py0 = static_cast<int>(p.points[ip * 3 + 1]),
              px0 = static_cast<int>(p.points[ip * 3 + 2]);
    zb = pz0 / kZT + tzb;
    zlo = max(pz0, zb * kZT);
    zhi = min(pz0 + p.wd, zb * kZT + kZT);
    if (zlo >= zhi) return;  // uniform per CTA
    zstride = static_cast<long long>(p.wd) * p.wd;
    zorg = pz0;
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
      const int tx = txb * 32 + lx, ty = (tyb * CPT + c) * 8 + ly;
      valid[c] = tx < p.wd && ty < p.wd;
      x[c] = px0 + tx;
      y[c] = py0 + ty;
      obase[c] = ip * zstride * p.wd + static_cast<long long>(ty) * p.wd + tx;
    }
  } else {  // MODE_PTS_XY: columns are the (y, x) points; out[z * npts + i]
    zb = blockIdx.y;
    zlo = zb * kZT;
    zhi = min(p.nz, zlo + kZT);
    zstride = p.npts;
    zorg = 0;
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
      const int i = (blockIdx.x * CPT + c) * 256 + threadIdx.x;
      valid[c] = i < p.npts;
      y[c] = valid[c] ? p.pts[2 * i] : 0;
      x[c] = valid[c] ? p.pts[2 * i + 1] : 0;
      obase[c] = i;
    }
  }

  float acc[CPT][kZT];
  unsigned mbits[CPT];
#pragma unroll
  for (int c = 0; c < CPT; ++c) {
#pragma unroll
    for (int i = 0; i < kZT; ++i) acc[c][i] = 0.f;
    mbits[c] = 0xffffu;
    if (MODE == MODE_VOLUME && p.masked) {
      mbits[c] = 0u;
      if (valid[c])
        for (int z = zlo; z < zhi; ++z)
          if (p.out[obase[c] + (z - zorg) * zstride] > 0.f) mbits[c] |= 1u << (z - zb * kZT);
    }
  }

  const float4* gz4 = reinterpret_cast<const float4*>(p.gT) + static_cast<size_t>(zb) * (p.ntheta + 1) * 4 * p.n;
  const int half = p.n / 2;
  float fx[CPT], fy[CPT];
  bool active[CPT];
  bool any_active = false;
#pragma unroll
  for (int c = 0; c < CPT; ++c) {
    fx[c] = static_cast<float>(x[c] - half);
    fy[c] = static_cast<float>(y[c] - half);
    active[c] = valid[c] && mbits[c] != 0u;
    any_active |= active[c];
  }
  any_active = __any_sync(0xffffffffu, any_active);
  for (int k0 = 0; k0 < p.ntheta; k0 += kTrigChunk) {
    const int kn = min(kTrigChunk, p.ntheta - k0);
    __syncthreads();
    for (int i = threadIdx.x; i < kn; i += blockDim.x) trig_s[i] = p.trig[k0 + i];
    __syncthreads();
    if (any_active) bp_columns<CPT, EXACT>(p, trig_s, k0, kn, gz4, fx, fy, active, acc);
  }

#pragma unroll
  for (int c = 0; c < CPT; ++c) {
    if (!valid[c]) continue;
#pragma unroll
    for (int i = 0; i < kZT; ++i) {
      const int z = zb * kZT + i;
      if (z < zlo || z >= zhi) continue;
      float v = (mbits[c] >> i) & 1u ? acc[c][i] : 0.f;
      const long long o = obase[c] + (z - zorg) * zstride;
      if (MODE == MODE_PATCHES) {
        if (p.normalize) {  // recon.py:283-285: clip, then (v - min) / (max - min)
          v = fminf(fmaxf(v, p.lo), p.hi);
          v = __fdiv_rn(__fsub_rn(v, p.lo), __fsub_rn(p.hi, p.lo));
        }
        if (p.out_kind == 1) reinterpret_cast<__half*>(p.pout)[o] = __float2half_rn(v);
        else if (p.out_kind == 2) reinterpret_cast<__nv_bfloat16*>(p.pout)[o] = __float2bfloat16_rn(v);
        else reinterpret_cast<float*>(p.pout)[o] = v;
      } else {
        p.out[o] = v;
      }
    }
  }
}

// rec_pts: one (z, y, x) point per thread, scalar loads from the blocked layout
template <bool EXACT>
__global__ void __launch_bounds__(256) bp_pts_kernel(const BpParams p) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.npts) return;
  const int z = p.pts[3 * i], y = p.pts[3 * i + 1], x = p.pts[3 * i + 2];
  const float fx = static_cast<float>(x - p.n / 2), fy = static_cast<float>(y - p.n / 2);
  const float nf = static_cast<float>(p.n), inv_n = 1.0f / nf;
  const float* gz = p.gT + (static_cast<size_t>(z / kZT) * (p.ntheta + 1) * 4 + ((z % kZT) >> 2)) * p.n * 4 + (z & 3);
  float f = 0.f;
  for (int k = 0; k < p.ntheta; ++k) {
    const float2 cs = __ldg(p.trig + k);
    float sp;
    if (EXACT) sp = __fadd_rn(__fsub_rn(__fmul_rn(fx, cs.x), __fmul_rn(fy, cs.y)), p.center);
    else sp = fmaf(fx, cs.x, -(fy * cs.y)) + p.center;
    const float s0f = roundf(sp);
    const int s0 = static_cast<int>(s0f);
    if (s0 >= 0 && s0 < p.n - 1) {
      const float w = __fsub_rn(sp, s0f);
      const float* q = gz + (static_cast<size_t>(k) * 4 * p.n + s0) * 4;
      const float g0 = __ldg(q), g1 = __ldg(q + 4);
      if (EXACT) f = __fadd_rn(f, __fadd_rn(g0, __fdiv_rn(__fmul_rn(__fsub_rn(g1, g0), w), nf)));
      else f += fmaf(g1 - g0, w * inv_n, g0);
    }
  }
  p.out[i] = f;
}

// (ntheta, nz, n) -> gT[zb][k][q][s][4] (z = zb * 16 + q * 4 + j), z beyond nz zero-filled.  One CTA: one angle, one z
// block, 64 detector columns; every z quad is written as 1 KB of contiguous memory.
__global__ void __launch_bounds__(256) bp_relayout_kernel(const float* __restrict__ g, float* __restrict__ gT, int ntheta,
                                                          int nz, int n) {
  __shared__ float tile[kZT][65];
  const int k = blockIdx.y, zb = blockIdx.z, s_base = blockIdx.x * 64;
  for (int i = threadIdx.x; i < kZT * 64; i += blockDim.x) {
    const int zl = i >> 6, s = i & 63;
    const int z = zb * kZT + zl;
    tile[zl][s] = (z < nz && s_base + s < n) ? g[(static_cast<size_t>(k) * nz + z) * n + s_base + s] : 0.f;
  }
  __syncthreads();
  float* dst = gT + (static_cast<size_t>(zb) * (ntheta + 1) + k) * 4 * n * 4;
  for (int i = threadIdx.x; i < kZT * 64; i += blockDim.x) {
    const int q = i >> 8, s = (i >> 2) & 63, j = i & 3;
    if (s_base + s < n) dst[(static_cast<size_t>(q) * n + s_base + s) * 4 + j] = tile[q * 4 + j][s];
  }
}

__global__ void bp_trig_kernel(const float* __restrict__ theta, float2* __restrict__ trig, int ntheta) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= ntheta) return;
  const double t = static_cast<double>(theta[k]);
  trig[k] = make_float2(static_cast<float>(cos(t)), static_cast<float>(sin(t)));
}

inline size_t gT_floats(int ntheta, int nz, int n) {
  const size_t nzb = (nz + kZT - 1) / kZT;
  return nzb * (ntheta + 1) * static_cast<size_t>(n) * kZT;   // + one all-zero angle per z block
}

int fill_params(BpParams& p, const void* ws, int ntheta, int nz, int n, float center) {
  TE_CHECK_ARG(ws != nullptr, "back-projection: null workspace (call te_bp_prepare first)");
  TE_CHECK_ARG(ntheta > 0 && nz > 0 && n > 1, "back-projection: bad projection shape (%d, %d, %d)", ntheta, nz, n);
  memset(&p, 0, sizeof(p));
  p.gT = reinterpret_cast<const float*>(ws);
  p.trig = reinterpret_cast<const float2*>(p.gT + gT_floats(ntheta, nz, n));
  p.ntheta = ntheta;
  p.nz = nz;
  p.n = n;
  p.nzb = (nz + kZT - 1) / kZT;
  p.center = center;
  return TE_OK;
}

// ------------------------------------------------------------------------------------------- FBP filter / preprocess
// padded[r][j] = data[r][clamp(j - pad_left, 0, n-1)] for j < n_pad (np.pad mode='edge'); row pitch n_pad + 2 floats
// (in-place R2C layout).  Grid: (column blocks, rows) -- no per-element division.
__global__ void __launch_bounds__(256) fbp_pad_kernel(const float* __restrict__ data, float* __restrict__ padded, int n, int n_pad,
                                                      int pad_left) {
  const int pitch = n_pad + 2;
  const int j = blockIdx.x * 256 + threadIdx.x;
  if (j >= pitch) return;
  const long long r = blockIdx.y;
  float v = 0.f;
  if (j < n_pad) {
    int s = j - pad_left;
    s = s < 0 ? 0 : (s > n - 1 ? n - 1 : s);
    v = __ldg(data + r * n + s);
  }
  padded[r * pitch + j] = v;
}

// spectrum *= rfftfreq(n_pad)[j] / n_pad (the 1/n_pad is numpy's irfft normalisation, which cuFFT leaves out)
__global__ void __launch_bounds__(256) fbp_ramp_kernel(float2* __restrict__ spec, int n_pad) {
  const int nc = n_pad / 2 + 1;
  const int j = blockIdx.x * 256 + threadIdx.x;
  if (j >= nc) return;
  const float inv = 1.0f / static_cast<float>(n_pad);
  const float w = (static_cast<float>(j) * inv) * inv;
  float2* p = spec + static_cast<long long>(blockIdx.y) * nc + j;
  float2 v = *p;
  v.x *= w;
  v.y *= w;
  *p = v;
}

__global__ void __launch_bounds__(256) fbp_crop_kernel(const float* __restrict__ padded, float* __restrict__ data, int n, int n_pad,
                                                       int pad_left) {
  const int pitch = n_pad + 2;
  const int s = blockIdx.x * 256 + threadIdx.x;
  if (s >= n) return;
  const long long r = blockIdx.y;
  data[r * n + s] = padded[r * pitch + pad_left + s];
}

// prep.py:63-77 with the identity median filter removed: data = -log(max((data - dark) / max(flat - dark, 1e-6), 1e-6)),
// dark / flat (nz, n) broadcast over the angles
__global__ void __launch_bounds__(256) preprocess_kernel(float* __restrict__ data, const float* __restrict__ dark,
                                                         const float* __restrict__ flat, long long plane) {
  float* row = data + static_cast<long long>(blockIdx.y) * plane;   // one angle per blockIdx.y
  for (long long j = blockIdx.x * 256LL + threadIdx.x; j < plane; j += gridDim.x * 256LL) {
    const float d = __ldg(dark + j);
    const float v = __fdiv_rn(__fsub_rn(row[j], d), fmaxf(__fsub_rn(__ldg(flat + j), d), 1.0e-6f));
    row[j] = -logf(fmaxf(v, 1.0e-6f));
  }
}

// box painting for make_mask (recon.py:369-381): mask = 0, then 1 inside every wd^3 box
__global__ void __launch_bounds__(256) make_mask_kernel(float* __restrict__ mask, const uint32_t* __restrict__ points,
                                                        long long npatch, int wd, int nz, int ny, int nx) {
  const long long per = static_cast<long long>(wd) * wd * wd;
  const long long total = npatch * per;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += gridDim.x * 256LL) {
    const long long ip = i / per;
    int r = static_cast<int>(i - ip * per);
    const int tz = r / (wd * wd);
    r -= tz * wd * wd;
    const int ty = r / wd, tx = r - ty * wd;
    const long long z = points[ip * 3] + tz, y = points[ip * 3 + 1] + ty, x = points[ip * 3 + 2] + tx;
    if (z < nz && y < ny && x < nx) mask[(z * ny + y) * nx + x] = 1.0f;
  }
}

struct PlanKey {
  int n_pad;
  long long batch;
  bool operator<(const PlanKey& o) const { return n_pad != o.n_pad ? n_pad < o.n_pad : batch < o.batch; }
};
struct Plans {
  cufftHandle fwd, inv;
};
std::mutex g_plan_mutex;
std::map<PlanKey, Plans> g_plans;

int get_plans(int n_pad, long long batch, Plans& out) {
  std::lock_guard<std::mutex> lock(g_plan_mutex);
  auto it = g_plans.find(PlanKey{n_pad, batch});
  if (it != g_plans.end()) {
    out = it->second;
    return TE_OK;
  }
  Plans pl;
  int nn[1] = {n_pad};
  int inembed[1] = {n_pad + 2}, onembed[1] = {n_pad / 2 + 1};
  cufftResult r1 = cufftPlanMany(&pl.fwd, 1, nn, inembed, 1, n_pad + 2, onembed, 1, n_pad / 2 + 1, CUFFT_R2C,
                                 static_cast<int>(batch));
  cufftResult r2 = cufftPlanMany(&pl.inv, 1, nn, onembed, 1, n_pad / 2 + 1, inembed, 1, n_pad + 2, CUFFT_C2R,
                                 static_cast<int>(batch));
  if (r1 != CUFFT_SUCCESS || r2 != CUFFT_SUCCESS) {
    set_error("cufftPlanMany(n = %d, batch = %lld) failed: %d / %d", n_pad, batch, static_cast<int>(r1), static_cast<int>(r2));
    return TE_ERR_CUDA;
  }
  g_plans[PlanKey{n_pad, batch}] = pl;
  out = pl;
  return TE_OK;
}

inline int grid_for(long long total) {
  const long long b = (total + 255) / 256;
  return static_cast<int>(b < 1 ? 1 : (b > kNumSMs * 16LL ? kNumSMs * 16LL : b));
}

}  // namespace
}  // namespace te

using namespace te;

extern "C" void te_fbp_padding(int n, int* pad_left, int* pad_right) {
  // prep.py:21-31: n_pad = ceil(1.5 n / 8) * 8
  const double np = ceil(n * 1.5 / 8.0) * 8.0;
  const int n_pad = static_cast<int>(np);
  const int l = (n_pad - n) / 2;
  if (pad_left) *pad_left = l;
  if (pad_right) *pad_right = n_pad - n - l;
}

extern "C" int64_t te_fbp_workspace_bytes(int64_t rows, int n) {
  int l, r;
  te_fbp_padding(n, &l, &r);
  const int64_t pitch = n + l + r + 2;
  // rows are filtered in batches that keep the padded block inside L2
  int64_t batch = (48LL << 20) / (pitch * 4);
  batch = batch < 1 ? 1 : (batch > rows ? rows : batch);
  batch = batch > 65535 ? 65535 : batch;   // rows ride on gridDim.y
  return batch * pitch * 4;
}

extern "C" int te_fbp_filter(float* data, int64_t rows, int n, void* workspace, int64_t workspace_bytes, void* stream) {
  TE_CHECK_ARG(data && workspace, "te_fbp_filter: null pointer");
  TE_CHECK_ARG(rows > 0 && n > 1, "te_fbp_filter: bad shape (%lld, %d)", static_cast<long long>(rows), n);
  int pl, pr;
  te_fbp_padding(n, &pl, &pr);
  const int n_pad = n + pl + pr;
  const int64_t pitch = n_pad + 2;
  int64_t batch = workspace_bytes / (pitch * 4) < rows ? workspace_bytes / (pitch * 4) : rows;
  batch = batch > 65535 ? 65535 : batch;
  TE_CHECK_ARG(batch >= 1, "te_fbp_filter: workspace too small (%lld bytes)", static_cast<long long>(workspace_bytes));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* pad = reinterpret_cast<float*>(workspace);
  for (int64_t r0 = 0; r0 < rows; r0 += batch) {
    const int64_t nb = rows - r0 < batch ? rows - r0 : batch;
    Plans plans;
    const int rc = get_plans(n_pad, nb, plans);
    if (rc != TE_OK) return rc;
    float* d = data + r0 * n;
    fbp_pad_kernel<<<dim3((pitch + 255) / 256, nb), 256, 0, st>>>(d, pad, n, n_pad, pl);
    TE_LAUNCH_CHECK();
    if (cufftSetStream(plans.fwd, st) != CUFFT_SUCCESS || cufftSetStream(plans.inv, st) != CUFFT_SUCCESS) {
      set_error("cufftSetStream failed");
      return TE_ERR_CUDA;
    }
    if (cufftExecR2C(plans.fwd, pad, reinterpret_cast<cufftComplex*>(pad)) != CUFFT_SUCCESS) {
      set_error("cufftExecR2C failed");
      return TE_ERR_CUDA;
    }
    fbp_ramp_kernel<<<dim3((n_pad / 2 + 256) / 256, nb), 256, 0, st>>>(reinterpret_cast<float2*>(pad), n_pad);
    TE_LAUNCH_CHECK();
    if (cufftExecC2R(plans.inv, reinterpret_cast<cufftComplex*>(pad), pad) != CUFFT_SUCCESS) {
      set_error("cufftExecC2R failed");
      return TE_ERR_CUDA;
    }
    fbp_crop_kernel<<<dim3((n + 255) / 256, nb), 256, 0, st>>>(pad, d, n, n_pad, pl);
    TE_LAUNCH_CHECK();
  }
  return TE_OK;
}

extern "C" int te_fbp_release_plans(void) {
  std::lock_guard<std::mutex> lock(g_plan_mutex);
  for (auto& kv : g_plans) {
    cufftDestroy(kv.second.fwd);
    cufftDestroy(kv.second.inv);
  }
  g_plans.clear();
  return TE_OK;
}

extern "C" int te_preprocess(float* data, const float* dark, const float* flat, int ntheta, int nz, int n, void* stream) {
  TE_CHECK_ARG(data && dark && flat, "te_preprocess: null pointer");
  TE_CHECK_ARG(ntheta > 0 && nz > 0 && n > 0, "te_preprocess: bad shape");
  TE_CHECK_ARG(ntheta <= 65535, "te_preprocess: at most 65535 angles");
  const long long plane = static_cast<long long>(nz) * n;
  const int bx = static_cast<int>((plane + 255) / 256 > 64 ? 64 : (plane + 255) / 256);
  preprocess_kernel<<<dim3(bx, ntheta), 256, 0, static_cast<cudaStream_t>(stream)>>>(data, dark, flat, plane);
  TE_LAUNCH_CHECK();
  return TE_OK;
}

extern "C" int te_make_mask(float* mask, const int64_t shape[3], const uint32_t* points, int64_t n, int32_t wd, void* stream) {
  TE_CHECK_ARG(mask && shape, "te_make_mask: null pointer");
  TE_CHECK_ARG(n == 0 || points, "te_make_mask: null points");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TE_CUDA(cudaMemsetAsync(mask, 0, sizeof(float) * shape[0] * shape[1] * shape[2], st));
  if (n == 0) return TE_OK;
  const long long total = static_cast<long long>(n) * wd * wd * wd;
  make_mask_kernel<<<grid_for(total), 256, 0, st>>>(mask, points, n, wd, static_cast<int>(shape[0]), static_cast<int>(shape[1]),
                                                   static_cast<int>(shape[2]));
  TE_LAUNCH_CHECK();
  return TE_OK;
}

extern "C" int64_t te_bp_workspace_bytes(int ntheta, int nz, int n) {
  if (ntheta <= 0 || nz <= 0 || n <= 0) return 0;
  return static_cast<int64_t>(gT_floats(ntheta, nz, n) * sizeof(float) + static_cast<size_t>(ntheta) * sizeof(float2));
}

extern "C" int te_bp_prepare(const float* data, const float* theta, int ntheta, int nz, int n, void* workspace, void* stream) {
  TE_CHECK_ARG(data && theta && workspace, "te_bp_prepare: null pointer");
  TE_CHECK_ARG(ntheta > 0 && nz > 0 && n > 1, "te_bp_prepare: bad projection shape (%d, %d, %d)", ntheta, nz, n);
  TE_CHECK_ARG(ntheta <= 65535, "te_bp_prepare: at most 65535 angles");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* gT = reinterpret_cast<float*>(workspace);
  float2* trig = reinterpret_cast<float2*>(gT + gT_floats(ntheta, nz, n));
  const int nzb = (nz + kZT - 1) / kZT;
  TE_CHECK_ARG(nzb <= 65535, "te_bp_prepare: nz too large");
  TE_CHECK_ARG(static_cast<long long>(ntheta + 1) * 4 * n < (1LL << 31), "te_bp_prepare: ntheta * n too large");
  for (int zb = 0; zb < nzb; ++zb)   // the all-zero angle that out-of-range samples read
    TE_CUDA(cudaMemsetAsync(gT + (static_cast<size_t>(zb) * (ntheta + 1) + ntheta) * 4 * n * 4, 0, sizeof(float) * 16 * n, st));
  bp_relayout_kernel<<<dim3((n + 63) / 64, ntheta, nzb), 256, 0, st>>>(data, gT, ntheta, nz, n);
  TE_LAUNCH_CHECK();
  bp_trig_kernel<<<(ntheta + 255) / 256, 256, 0, st>>>(theta, trig, ntheta);
  TE_LAUNCH_CHECK();
  return TE_OK;
}

extern "C" int te_bp_box(const void* workspace, int ntheta, int nz, int n, float center, int stx, int px, int sty, int py,
                         int stz, int pz, float* out, int masked, int exact, void* stream) {
  BpParams p;
  const int rc = fill_params(p, workspace, ntheta, nz, n, center);
  if (rc != TE_OK) return rc;
  TE_CHECK_ARG(out != nullptr, "te_bp_box: null output");
  TE_CHECK_ARG(px > 0 && py > 0 && pz > 0 && stz >= 0 && stz + pz <= nz, "te_bp_box: z range [%d, %d) outside [0, %d)", stz,
               stz + pz, nz);
  p.out = out;
  p.stx = stx; p.px = px; p.sty = sty; p.py = py; p.stz = stz; p.pz = pz;
  p.masked = masked;
  constexpr int CPT = 2;
  const int zblocks = (stz + pz - 1) / kZT - stz / kZT + 1;
  dim3 grid((px + 31) / 32, (py + 8 * CPT - 1) / (8 * CPT), zblocks);
  TE_CHECK_ARG(grid.y <= 65535 && grid.z <= 65535, "te_bp_box: box too large");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (exact) bp_kernel<MODE_VOLUME, CPT, true><<<grid, 256, 0, st>>>(p);
  else bp_kernel<MODE_VOLUME, CPT, false><<<grid, 256, 0, st>>>(p);
  TE_LAUNCH_CHECK();
  return TE_OK;
}

extern "C" int te_bp_patches(const void* workspace, int ntheta, int nz, int n, float center, const uint32_t* points, int64_t npatch,
                             int32_t wd, void* out, int out_dtype, int normalize, float clip_lo, float clip_hi, int exact,
                             void* stream) {
  BpParams p;
  const int rc = fill_params(p, workspace, ntheta, nz, n, center);
  if (rc != TE_OK) return rc;
  if (npatch == 0) return TE_OK;
  TE_CHECK_ARG(points && out, "te_bp_patches: null pointer");
  TE_CHECK_ARG(wd > 0 && wd <= nz, "te_bp_patches: patch width %d against nz = %d", wd, nz);
  TE_CHECK_ARG(out_dtype == TE_F32 || out_dtype == TE_F16 || out_dtype == TE_BF16,
               "te_bp_patches: output must be float32, float16 or bfloat16");
  constexpr int CPT = 2;
  p.points = points;
  p.npatch = npatch;
  p.wd = wd;
  p.tiles_x = (wd + 31) / 32;
  p.tiles_y = (wd + 8 * CPT - 1) / (8 * CPT);
  p.zblocks = (wd + kZT - 2) / kZT + 1;  // an unaligned corner straddles one more block
  p.pout = out;
  p.out_kind = out_dtype == TE_F16 ? 1 : (out_dtype == TE_BF16 ? 2 : 0);
  p.normalize = normalize;
  p.lo = clip_lo;
  p.hi = clip_hi;
  const long long blocks = npatch * p.tiles_x * p.tiles_y * p.zblocks;
  TE_CHECK_ARG(blocks < (1LL << 31), "te_bp_patches: too many patches for one launch");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (exact) bp_kernel<MODE_PATCHES, CPT, true><<<static_cast<unsigned>(blocks), 256, 0, st>>>(p);
  else bp_kernel<MODE_PATCHES, CPT, false><<<static_cast<unsigned>(blocks), 256, 0, st>>>(p);
  TE_LAUNCH_CHECK();
  return TE_OK;
}

extern "C" int te_bp_points(const void* workspace, int ntheta, int nz, int n, float center, const int32_t* pts, int32_t npts,
                            int xy_mode, float* out, int exact, void* stream) {
  BpParams p;
  const int rc = fill_params(p, workspace, ntheta, nz, n, center);
  if (rc != TE_OK) return rc;
  if (npts == 0) return TE_OK;
  TE_CHECK_ARG(pts && out, "te_bp_points: null pointer");
  p.pts = pts;
  p.npts = npts;
  p.out = out;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (xy_mode) {
    dim3 grid((npts + 255) / 256, p.nzb);
    if (exact) bp_kernel<MODE_PTS_XY, 1, true><<<grid, 256, 0, st>>>(p);
    else bp_kernel<MODE_PTS_XY, 1, false><<<grid, 256, 0, st>>>(p);
  } else {
    if (exact) bp_pts_kernel<true><<<(npts + 255) / 256, 256, 0, st>>>(p);
    else bp_pts_kernel<false><<<(npts + 255) / 256, 256, 0, st>>>(p);
  }
  TE_LAUNCH_CHECK();
  return TE_OK;
}

// 3x3x3 conv weight gradient on tcgen05 (row a11 of SURVEY.md section 8: the backward pass of the
// Conv3D layers of /root/reference/tomo_encoders/neural_nets/Unet3D.py:20-82 under Keras `fit`).
//
//   dW[tz][ty][tx][ci][co] = sum_v X[v + (tz-1, ty-1, tx-1)][ci] * dY[v][co]          (NDHWC tensors)
//
// GEMM view: the contraction (K) runs over voxels, the SLOW dimension of both operands, so both are
// MN-major UMMA operands: a shared-memory row is one voxel (32- or 64-byte row of channels, TMA
// swizzle), 8 consecutive rows are one K group, and a second stride places further channel "atoms".
// That second stride is used for the TAPS (validated on hardware by tools/wgrad_probe.cu):
//   A = dY plane box (16 y x 11 x', 32 output channels, SWIZZLE_64B); M = 128 = 4 atoms of 32
//       channels whose stride (LBO) is ONE row: atom j is the box shifted by j voxels in x, i.e. the
//       tap tx = 2 - j (the fourth atom is junk and is dropped in the epilogue);
//   B = X plane box (18 y x 8 x, CB input channels); N = 3 atoms of CB channels whose stride is one
//       y row of the box: atom i is the tap ty = i;
//   the tap tz selects which of three resident X planes is multiplied and which accumulator
//   (TMEM columns [tz * 3 CB, (tz + 1) * 3 CB)) receives the product.
// One k-step = 16 voxels = 8 x of two y rows (SBO = the y pitch of the box), 8 k-steps per plane,
// 24 UMMAs (M 128, N 3 CB, K 16) per dY plane.
//
// Work unit = (patch, z segment of up to 16 planes, 16 x 8 column); a CTA owns one (CB input
// channels) x (32 output channels) block of all 27 taps for the whole launch: 9 CB TMEM columns of
// fp32 accumulators, flushed ONCE with coalesced atomics.  Warp 0 = TMA producer (X planes and dY
// planes in two rings; TMA zero fill supplies the per-patch "same" padding), warps 1-3 = UMMA issuers
// (one tz tap each), warps 4-7 = epilogue.
#include <cuda.h>

#include "sm100_ptx.cuh"
#include "te_common.cuh"

namespace te {
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

namespace {

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    cudaDriverEntryPointQueryResult q;
    void* p = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess) fn = (EncodeTiledFn)p;
  }
  return fn;
}

constexpr int kZS = 16;              // planes per z segment
constexpr int kXE = 11;              // x extent of the dY box: 8 + 3 shifts
constexpr int kNXS = 10, kNYS = 6;   // ring depths (X planes / dY planes): ~5 planes of prefetch
constexpr uint32_t kYBytes = 16 * kXE * 64;      // 11264: multiple of the 512-byte swizzle repeat
constexpr int kThreads = 256;             // warp 0 producer, warps 1-3 UMMA issuers (one tz tap each), warps 4-7 epilogue

struct WuParams {
  float* dw; int dw_cin, ci_off, Cout;
  int x_coff, dy_coff;
  int D, H, W;
  int tiles_x, tiles_y, zsegs, zs;
  int nci, pairs, ngroups;
  long long units;
  uint32_t fmt;                      // 0 fp16, 1 bf16
};

template <int CB>
__global__ void __launch_bounds__(kThreads, 1)
wgrad_umma_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_dy, const WuParams p) {
  constexpr uint32_t ROWB = CB * 2;
  constexpr uint32_t XBYTES = 18 * 8 * ROWB;                       // 9216 / 4608
  constexpr uint32_t XSLOT = (XBYTES + 1023) / 1024 * 1024;
  constexpr uint32_t BLAYOUT = CB == 32 ? kLayoutSW64 : kLayoutSW32;
  constexpr uint32_t NCOLS = 9 * CB, TCOLS = CB == 32 ? 512 : 256;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* xs = smem;
  uint8_t* ys = smem + kNXS * XSLOT;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ys + kNYS * kYBytes);
  uint64_t *x_full = bars, *x_empty = bars + kNXS, *y_full = bars + 2 * kNXS, *y_empty = y_full + kNYS, *acc_full = y_empty + kNYS;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int pair = blockIdx.x % p.pairs, group = blockIdx.x / p.pairs;
  const int cic = pair % p.nci, coc = pair / p.nci;

  if (tid == 0) {
    for (int i = 0; i < kNXS; ++i) { mbar_init(x_full + i, 1); mbar_init(x_empty + i, 3); }
    for (int i = 0; i < kNYS; ++i) { mbar_init(y_full + i, 1); mbar_init(y_empty + i, 3); }
    mbar_init(acc_full, 3);
    fence_barrier_init();
    prefetch_tmap(&tmap_x); prefetch_tmap(&tmap_dy);
  }
  if (warp == 1) { tmem_alloc(tmem_slot, TCOLS); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const long long first = group, step = p.ngroups;
  const bool has_work = group < p.ngroups && first < p.units;

  if (warp == 0) {
    // ------------------------------------------------------------------ producer
    if (elect_one()) {
      uint32_t xc = 0, yc = 0;
      for (long long u = first; u < p.units && group < p.ngroups; u += step) {
        unsigned t = static_cast<unsigned>(u);             // 32-bit divisions (unit counts are far below 2^31)
        const unsigned ntx = p.tiles_x, nty = p.tiles_y, nzs = p.zsegs;
        const int x0 = static_cast<int>(t % ntx) * 8; t /= ntx;
        const int y0 = static_cast<int>(t % nty) * 16; t /= nty;
        const int z0 = static_cast<int>(t % nzs) * p.zs; t /= nzs;
        const int n = static_cast<int>(t);
        for (int pi = 0; pi < p.zs + 2; ++pi) {
          const uint32_t s = xc % kNXS;
          mbar_wait(x_empty + s, ((xc / kNXS) & 1u) ^ 1u);
          mbar_expect_tx(x_full + s, XBYTES);
          tma_load_5d(xs + s * XSLOT, &tmap_x, x_full + s, p.x_coff + cic * CB, x0, y0 - 1, z0 - 1 + pi, n);
          ++xc;
          if (pi >= 2) {
            const uint32_t q = yc % kNYS;
            mbar_wait(y_empty + q, ((yc / kNYS) & 1u) ^ 1u);
            mbar_expect_tx(y_full + q, kYBytes);
            tma_load_5d(ys + q * kYBytes, &tmap_dy, y_full + q, p.dy_coff + coc * 32, x0 - 1, y0, z0 + pi - 2, n);
            ++yc;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp <= 3) {
    // ------------------------------------------------------------------ UMMA issuers: warp 1 + tz owns the tap plane tz
    // (its own accumulator, TMEM columns [tz * 3 CB, (tz + 1) * 3 CB)).  One issuer would stall at every plane boundary:
    // the descriptors live in uniform registers, and rewriting them for the next plane waits until the queued UMMAs have
    // read them, i.e. the issue queue drains (measured: tensor pipe busy 66 % of the time).  Three issuers keep the queue
    // fed from two warps while the third turns around.  A ring slot is free once all three have committed (count 3).
    const int dz = warp - 1;
    const uint32_t idesc = make_idesc_f16(128, 3 * CB, p.fmt & 1u) | (1u << 15) | (1u << 16);     // A and B MN-major
    const uint64_t da0 = make_smem_desc(smem_u32(ys), 64, kXE * 64, kLayoutSW64, 0);
    const uint64_t db0 = make_smem_desc(smem_u32(xs), 8 * ROWB, 8 * ROWB, BLAYOUT, 0);
    const uint32_t acc = tmem + dz * 3 * CB;
    uint32_t xc = 0, yc = 0;             // xc = ring count of plane pi = 0 of the current unit
    bool fresh = true;
    for (long long u = first; u < p.units && group < p.ngroups; u += step) {
      for (int zi = 0; zi < p.zs; ++zi) {
        const uint32_t c = xc + zi + dz;                 // the X plane this warp multiplies
        mbar_wait(x_full + c % kNXS, (c / kNXS) & 1u);
        const uint32_t q = yc % kNYS;
        mbar_wait(y_full + q, (yc / kNYS) & 1u);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t da = da0 + static_cast<uint64_t>((q * kYBytes) >> 4);
          const uint64_t db = db0 + static_cast<uint64_t>(((c % kNXS) * XSLOT) >> 4);
#pragma unroll
          for (int s = 0; s < 8; ++s)
            umma_ss(acc, da + static_cast<uint64_t>((s * 2 * kXE * 64) >> 4), db + static_cast<uint64_t>((s * 16 * ROWB) >> 4), idesc,
                    (fresh && s == 0) ? 0u : 1u);
          umma_commit(y_empty + q);
          // plane pi is used by the steps zi = pi - tz: this warp is done with plane c; it also stands in for the uses that
          // never happen at the ends of the unit (plane 0 is only used by tz 0, plane 1 by tz 0 and 1, ...)
          umma_commit(x_empty + c % kNXS);
          if (zi == 0) {                                   // planes below this warp's first one
            for (int k = 0; k < dz; ++k) umma_commit(x_empty + (xc + k) % kNXS);
          }
          if (zi == p.zs - 1) {                            // planes above this warp's last one
            for (int k = dz + 1; k < 3; ++k) umma_commit(x_empty + (xc + zi + k) % kNXS);
          }
        }
        __syncwarp();
        fresh = false;
        ++yc;
      }
      xc += p.zs + 2;
    }
    if (has_work && elect_one()) umma_commit(acc_full);
    __syncwarp();
  } else if (has_work) {
    // ------------------------------------------------------------------ epilogue (once per CTA)
    mbar_wait(acc_full, 0);
    tc_fence_after();
    const int lg = warp & 3;             // TMEM lane group of this warp = x-shift atom j
    if (lg < 3 && p.fmt < 2) {
      const int tx = 2 - lg;
      const int co = coc * 32 + lane;
#pragma unroll 1
      for (int dz = 0; dz < 3; ++dz)
#pragma unroll 1
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int c0 = 0; c0 < CB; c0 += 16) {
            uint32_t r[16];
            tmem_ld16(tmem + (static_cast<uint32_t>(lg * 32) << 16) + dz * 3 * CB + i * CB + c0, r);
            tmem_ld_wait();
            float* dst = p.dw + (static_cast<long long>((dz * 3 + i) * 3 + tx) * p.dw_cin + p.ci_off + cic * CB + c0) * p.Cout + co;
#pragma unroll
            for (int e = 0; e < 16; ++e) atomicAdd(dst + static_cast<long long>(e) * p.Cout, __uint_as_float(r[e]));
          }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, TCOLS);
}

int make_tmap_plane(CUtensorMap* tm, const te_tensor& v, int cb, int bx, int by) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled is not available"); return TE_ERR_CUDA; }
  cuuint64_t gdim[5] = {(cuuint64_t)v.ctot, (cuuint64_t)v.W, (cuuint64_t)v.H, (cuuint64_t)v.D, (cuuint64_t)v.N};
  cuuint64_t gstr[4] = {(cuuint64_t)v.ctot * 2, (cuuint64_t)v.W * v.ctot * 2, (cuuint64_t)v.H * v.W * v.ctot * 2,
                        (cuuint64_t)v.D * v.H * v.W * v.ctot * 2};
  cuuint32_t box[5] = {(cuuint32_t)cb, (cuuint32_t)bx, (cuuint32_t)by, 1, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(tm, v.dtype == TE_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, v.ptr, gdim,
                   gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, cb == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return TE_ERR_CUDA; }
  return TE_OK;
}

template <int CB>
cudaError_t launch(const CUtensorMap& tx, const CUtensorMap& td, const WuParams& p, int grid, cudaStream_t st) {
  constexpr int XSLOT = (18 * 8 * CB * 2 + 1023) / 1024 * 1024;
  constexpr int smem = 1024 + kNXS * XSLOT + kNYS * kYBytes + 256;
  static unsigned long long attr = 0;
  cudaError_t e = smem_attr_once(wgrad_umma_kernel<CB>, smem, &attr);
  if (e != cudaSuccess) return e;
  wgrad_umma_kernel<CB><<<grid, kThreads, smem, st>>>(tx, td, p);
  return cudaGetLastError();
}
}  // namespace

bool wgrad_umma_ok(const te_tensor& x, const te_tensor& dy) {
  static const bool off = getenv("TOMOENC_NO_UMMA_WGRAD") != nullptr;
  return !off && x.dtype != TE_F32 && x.dtype == dy.dtype && x.C % 16 == 0 && dy.C % 32 == 0 && x.ctot % 8 == 0 && dy.ctot % 8 == 0 &&
         x.coff % 8 == 0 && dy.coff % 8 == 0 && x.W >= 8 && (reinterpret_cast<uintptr_t>(x.ptr) & 15) == 0 &&
         (reinterpret_cast<uintptr_t>(dy.ptr) & 15) == 0;
}

// dw[27][dw_cin][dy.C] += wgrad(x, dy), input channels [ci_off, ci_off + x.C)
int wgrad_umma(const te_tensor& x, const te_tensor& dy, float* dw, int dw_cin, int ci_off, cudaStream_t st) {
  const int cb = x.C % 32 == 0 ? 32 : 16;
  WuParams p;
  p.dw = dw; p.dw_cin = dw_cin; p.ci_off = ci_off; p.Cout = dy.C;
  p.x_coff = x.coff; p.dy_coff = dy.coff;
  p.D = x.D; p.H = x.H; p.W = x.W;
  p.tiles_x = (x.W + 7) / 8; p.tiles_y = (x.H + 15) / 16;
  p.zs = x.D < kZS ? x.D : kZS;
  p.zsegs = (x.D + p.zs - 1) / p.zs;
  p.units = static_cast<long long>(x.N) * p.zsegs * p.tiles_y * p.tiles_x;
  p.nci = x.C / cb;
  p.pairs = p.nci * (dy.C / 32);
  long long ng = kNumSMs / p.pairs;
  if (ng < 1) ng = 1;
  if (ng > p.units) ng = p.units;
  p.ngroups = static_cast<int>(ng);
  p.fmt = x.dtype == TE_F16 ? 0u : 1u;
  if (getenv("TOMOENC_WGRAD_NOEPI")) p.fmt |= 2u;   // experiment
  CUtensorMap tx, td;
  int rc = make_tmap_plane(&tx, x, cb, 8, 18);
  if (rc != TE_OK) return rc;
  rc = make_tmap_plane(&td, dy, 32, kXE, 16);
  if (rc != TE_OK) return rc;
  const int grid = p.pairs * p.ngroups;
  cudaError_t e = cb == 32 ? launch<32>(tx, td, p, grid, st) : launch<16>(tx, td, p, grid, st);
  count_launch();
  if (e != cudaSuccess) { set_error("wgrad_umma launch failed: %s", cudaGetErrorString(e)); return TE_ERR_CUDA; }
  return TE_OK;
}

}  // namespace te

// TMA-staged patch gather / scatter (sm_100a): the B200 fast path behind te_gather, te_gather_norm and
// te_scatter (include/tomoenc.h) for unbinned patches of a volume whose row pitch is a multiple of
// 16 bytes.  Reference semantics: Patches.extract / Grid.extract and fill_patches_in_volume
// (/root/reference/tomo_encoders/structures/patches.py:748-781, 811-821; grid.py:313-348).
//
// One 3-D tensor map (X, Y, Z) describes the volume.  A work unit is PZC consecutive z planes of one
// patch: one TMA box load (px, py, PZC) at the patch corner -- any corner, the TMA unit does the
// re-alignment -- lands the planes densely in shared memory, which is exactly the layout of the
// patch array, so they leave again as ONE contiguous bulk store (gather) or, for the fused
// clip / rescale variant, are converted by the CTA's threads and written with 128-bit stores.
// Scatter is the mirror image: contiguous bulk load, TMA box store into the volume.
// A ring of stages keeps ~190 KB of loads in flight per SM with a single issuing thread.
#include "gather_tma.cuh"
#include "sm100_ptx.cuh"

#include <cuda.h>
#include <type_traits>

namespace te {
namespace {

constexpr int kStageBytes = 36 * 1024;
constexpr int kStages = 6;
constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    cudaDriverEntryPointQueryResult q;
    void* p = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess) fn = (EncodeTiledFn)p;
  }
  return fn;
}

__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      :
      : "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, const void* src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               :
               : "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void bulk_store(void* dst, const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
               :
               : "l"(reinterpret_cast<uint64_t>(dst)), "r"(smem_u32(src)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

struct TmaGatherParams {
  const uint32_t* points;
  const uint32_t* widths;     // (n,3) or nullptr; a patch with widths[i][0] != pz is binned -> element-wise path inside the kernel
  uint8_t* patches;           // (n, pz, py, px) array: destination of the gather, source of the scatter
  uint8_t* vol;               // raw volume pointer (binned / unaligned fallbacks)
  long long n;
  long long Y, X;
  int pz, py, px;
  int pzc;                    // z planes per unit
  int units_per_patch;        // pz / pzc
  uint32_t dense_bytes;       // px * py * pzc * ES          (box of the dense map)
  uint32_t padded_bytes;      // (px + 16/ES) * py * pzc * ES (box of the padded map)
  int es;
  // fused clip / rescale (convert kernel)
  int rescale, do_clip;
  float lo, hi;
};

struct Unit { long long i; int zc; uint32_t z0, y0, x0; bool binned; };

__device__ __forceinline__ Unit load_unit(const TmaGatherParams& p, long long u) {
  Unit t;
  t.i = u / p.units_per_patch;
  t.zc = static_cast<int>(u - t.i * p.units_per_patch);
  t.z0 = __ldg(p.points + 3 * t.i);
  t.y0 = __ldg(p.points + 3 * t.i + 1);
  t.x0 = __ldg(p.points + 3 * t.i + 2);
  t.binned = p.widths != nullptr && __ldg(p.widths + 3 * t.i) != static_cast<uint32_t>(p.pz);
  return t;
}

template <typename TIn> __device__ __forceinline__ float cvt_in(TIn v);
template <> __device__ __forceinline__ float cvt_in<uint8_t>(uint8_t v) { return static_cast<float>(v); }
template <> __device__ __forceinline__ float cvt_in<uint16_t>(uint16_t v) { return static_cast<float>(v); }
template <> __device__ __forceinline__ float cvt_in<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float cvt_in<float>(float v) { return v; }

template <typename TOut> __device__ __forceinline__ TOut cvt_out(float v);
template <> __device__ __forceinline__ float cvt_out<float>(float v) { return v; }
template <> __device__ __forceinline__ __half cvt_out<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 cvt_out<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

constexpr int kThreadsG = 256;

// NW consecutive 32-bit words starting at an arbitrarily aligned shared-memory byte address.
template <int NW>
__device__ __forceinline__ void lds_unaligned(const uint8_t* src, uint32_t (&w)[NW]) {
  const uint32_t a = smem_u32(src);
  const uint32_t sh = (a & 3u) * 8u;
  const uint32_t* q = reinterpret_cast<const uint32_t*>(src - (a & 3u));
  if (sh == 0) {
#pragma unroll
    for (int j = 0; j < NW; ++j) w[j] = q[j];
  } else {
    uint32_t prev = q[0];
#pragma unroll
    for (int j = 0; j < NW; ++j) {
      const uint32_t next = q[j + 1];
      w[j] = __funnelshift_r(prev, next, sh);
      prev = next;
    }
  }
}

// ------------------------------------------------------------------------------------------ gather
// RAW = true : bit-exact copy (TOut = TIn is the unsigned integer type of the element size).
// RAW = false: clip / rescale in fp32, TOut in {bf16, f16, f32}.
// A TMA box must start on a 16-byte boundary of the innermost dimension (an unaligned start
// coordinate faults with "illegal instruction": tools/tma_copy_probe.cu).  Units whose corner is
// aligned (every Grid / regular-grid tile) are copied smem -> global by ONE bulk store in RAW mode;
// the others are loaded through the PADDED map (box 16 bytes wider, start rounded down) and
// re-aligned by the CTA's threads on the way out.
template <typename TIn, typename TOut, bool RAW>
__global__ void __launch_bounds__(kThreadsG, 1)
gather_tma_kernel(const __grid_constant__ CUtensorMap tmap_dense, const __grid_constant__ CUtensorMap tmap_pad,
                  const TmaGatherParams p) {
  constexpr int ES = sizeof(TIn);
  constexpr int EPI = RAW ? 16 / ES : 8;                 // elements per work item of a thread
  constexpr int PADE = 16 / ES;                          // elements of padding per row of the padded box
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) mbar_init(full + s, 1);
    fence_barrier_init();
    prefetch_tmap(&tmap_dense);
    prefetch_tmap(&tmap_pad);
  }
  __syncthreads();
  const long long units = p.n * p.units_per_patch;
  const long long first = blockIdx.x, step = gridDim.x;
  const long long mine = first < units ? (units - first + step - 1) / step : 0;
  const float denom = p.hi - p.lo + 1e-12f;
  const int items_per_row = p.px / EPI;
  const int items = p.pzc * p.py * items_per_row;
  const int elems = p.pzc * p.py * p.px;
  const uint32_t pitch_pad = static_cast<uint32_t>(p.px + PADE) * ES;
  long long issued = 0;
  auto issue = [&](long long k) {              // warp 0, warp-uniform; one elected lane issues
    if (tid < 32) {
      const Unit t = load_unit(p, first + k * step);
      const int s = static_cast<int>(k % kStages);
      const bool aligned = (t.x0 * ES) % 16 == 0;
      if (elect_one()) {
        if (t.binned) {
          mbar_arrive(full + s);               // converted straight from global memory
        } else if (RAW && aligned) {
          mbar_expect_tx(full + s, p.dense_bytes);
          tma_load_3d(smem + s * kStageBytes, &tmap_dense, full + s, static_cast<int>(t.x0), static_cast<int>(t.y0),
                      static_cast<int>(t.z0) + t.zc * p.pzc);
        } else {
          mbar_expect_tx(full + s, p.padded_bytes);
          tma_load_3d(smem + s * kStageBytes, &tmap_pad, full + s, static_cast<int>(t.x0 & ~static_cast<uint32_t>(PADE - 1)),
                      static_cast<int>(t.y0), static_cast<int>(t.z0) + t.zc * p.pzc);
        }
      }
      __syncwarp();
    }
  };
  for (; issued < mine && issued < kStages - 1; ++issued) issue(issued);
  for (long long k = 0; k < mine; ++k) {
    const int s = static_cast<int>(k % kStages);
    mbar_wait(full + s, static_cast<uint32_t>((k / kStages) & 1));
    const long long u = first + k * step;
    const Unit t = load_unit(p, u);
    const bool aligned = (t.x0 * ES) % 16 == 0;
    TOut* dst = reinterpret_cast<TOut*>(p.patches) + u * elems;
    const uint8_t* stage = smem + s * kStageBytes;
    if (t.binned) {
      // decimating gather (binning b > 1): element-wise from global memory
      const uint32_t b = __ldg(p.widths + 3 * t.i) / static_cast<uint32_t>(p.pz);
      const TIn* vol = reinterpret_cast<const TIn*>(p.vol);
      for (int e = tid; e < elems; e += kThreadsG) {
        const int ix = e % p.px, iy = (e / p.px) % p.py, iz = e / (p.px * p.py) + t.zc * p.pzc;
        const long long so = ((static_cast<long long>(t.z0) + static_cast<long long>(iz) * b) * p.Y +
                              (static_cast<long long>(t.y0) + static_cast<long long>(iy) * b)) * p.X +
                             t.x0 + static_cast<long long>(ix) * b;
        if constexpr (RAW) {
          dst[e] = vol[so];
        } else {
          float v = cvt_in<TIn>(vol[so]);
          if (p.rescale) {
            if (p.do_clip) v = fminf(fmaxf(v, p.lo), p.hi);
            v = (v - p.lo) / denom;
          }
          dst[e] = cvt_out<TOut>(v);
        }
      }
    } else if (RAW && aligned) {
      if (tid < 32) {
        if (elect_one()) bulk_store(dst, stage, p.dense_bytes);
        __syncwarp();
      }
    } else {
      const uint32_t shift_b = (t.x0 & static_cast<uint32_t>(PADE - 1)) * ES;
      for (int it = tid; it < items; it += kThreadsG) {
        const int row = it / items_per_row, col = (it - row * items_per_row) * EPI;
        const uint8_t* src = stage + row * pitch_pad + shift_b + col * ES;
        constexpr int NW = EPI * ES / 4;
        uint32_t w[NW];
        lds_unaligned<NW>(src, w);
        if constexpr (RAW) {
          *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(dst) + (static_cast<long long>(row) * p.px + col) * ES) =
              make_uint4(w[0], w[1], w[2], w[3]);
        } else {
          float f[8];
          if constexpr (ES == 1) {
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = static_cast<float>((w[j >> 2] >> ((j & 3) * 8)) & 0xffu);
          } else if constexpr (ES == 2) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const uint16_t h = static_cast<uint16_t>(w[j >> 1] >> ((j & 1) * 16));
              if constexpr (std::is_same<TIn, __half>::value) f[j] = __half2float(__ushort_as_half(h));
              else f[j] = static_cast<float>(h);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = __uint_as_float(w[j]);
          }
          if (p.rescale) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float v = f[j];
              if (p.do_clip) v = fminf(fmaxf(v, p.lo), p.hi);
              f[j] = (v - p.lo) / denom;
            }
          }
          TOut* o = dst + static_cast<long long>(row) * p.px + col;
          if constexpr (sizeof(TOut) == 4) {
            reinterpret_cast<float4*>(o)[0] = make_float4(f[0], f[1], f[2], f[3]);
            reinterpret_cast<float4*>(o)[1] = make_float4(f[4], f[5], f[6], f[7]);
          } else {
            uint32_t ow[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if constexpr (std::is_same<TOut, __nv_bfloat16>::value) {
                __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
                ow[j] = *reinterpret_cast<uint32_t*>(&h);
              } else {
                __half2 h = __floats2half2_rn(f[2 * j], f[2 * j + 1]);
                ow[j] = *reinterpret_cast<uint32_t*>(&h);
              }
            }
            *reinterpret_cast<uint4*>(o) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
          }
        }
      }
    }
    // every thread is done reading stage s; the bulk stores (RAW) track their own shared-memory
    // reads through the async group of warp 0's elected lane (always lane 0)
    if (RAW && tid < 32) {
      if (elect_one()) {
        bulk_commit();
        // the stage refilled below was consumed in the PREVIOUS iteration: at most the group just
        // committed may still be reading shared memory
        if (issued < mine) bulk_wait_read<1>();
      }
      __syncwarp();
    }
    __syncthreads();
    if (issued < mine) { issue(issued); ++issued; }
  }
  if (RAW && tid < 32) {
    if (elect_one()) bulk_wait_all();
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------ random-corner row gather
// Short patch rows (16..240 bytes) at arbitrary byte corners -- the access pattern of BASELINE configs[0] / configs[2].
// The SIMT row-group kernel (gather_scatter.cu) spends ~45 warp instructions per 256 bytes on addresses, shuffles and
// predicates and is issue-bound at ~3 TB/s (ncu: 53 % of the issue slots at 32 % occupancy).  Here the TMA unit does the
// address generation: a unit (pzc planes of one patch) is ONE box load whose x start is rounded DOWN to 16 bytes and
// whose rows are 16 bytes wider than the patch row (a box must start on a 16-byte boundary, see above), so a stage holds
// rows of (row_bytes + 16) bytes with the wanted bytes at offset sh = x0 * es % 16.  Eight consumer warps re-align on
// the way out: two aligned 128-bit shared-memory loads, four funnel shifts and one 128-bit coalesced store per 16
// output bytes (the unit's output is contiguous: item i lands at dst + 16 i).  One producer warp runs ahead by the depth
// of the ring (~200 KB of loads in flight per SM); corners are prefetched one unit, the processing order two units ahead.
constexpr int kRowsConsumers = 256;
constexpr int kRowsThreads = 32 + kRowsConsumers;
constexpr int kRowsStageMax = 50 * 1024;
constexpr int kRowsMaxStages = 8;
constexpr int kRowsSmemData = 200 * 1024;

struct RowsParams {
  const uint32_t* points;
  const uint32_t* perm;       // optional processing order (work slot -> patch index)
  uint8_t* out;
  long long n;
  int es, pzc, upp;           // element size, planes per unit, units per patch
  int rows_unit;              // py * pzc
  int cpr;                    // 16-byte chunks per patch row
  uint32_t pitch;             // bytes per staged row = row_bytes + 16
  uint32_t box_bytes;         // pitch * rows_unit
  uint32_t stage_bytes;       // box_bytes rounded up to 128
  int nstages;
  uint32_t unit_out_bytes;    // rows_unit * row_bytes
};

struct RowsMeta { uint8_t* dst; uint32_t sh; uint32_t pad; };

__device__ __forceinline__ void tma_load_3d_hint(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5}], [%2], %6;"
      :
      : "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "l"(pol)
      : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void stg_cs128(void* p, uint4 v) {
  asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// low 128 bits of (hi:lo) >> (32 Q + r) bits; Q words are a compile-time selection, r < 32 bits a funnel shift
template <int Q>
__device__ __forceinline__ uint4 realign(const uint4& lo, const uint4& hi, uint32_t r) {
  const uint32_t w0 = Q == 0 ? lo.x : Q == 1 ? lo.y : Q == 2 ? lo.z : lo.w;
  const uint32_t w1 = Q == 0 ? lo.y : Q == 1 ? lo.z : Q == 2 ? lo.w : hi.x;
  const uint32_t w2 = Q == 0 ? lo.z : Q == 1 ? lo.w : Q == 2 ? hi.x : hi.y;
  const uint32_t w3 = Q == 0 ? lo.w : Q == 1 ? hi.x : Q == 2 ? hi.y : hi.z;
  const uint32_t w4 = Q == 0 ? hi.x : Q == 1 ? hi.y : Q == 2 ? hi.z : hi.w;
  return make_uint4(__funnelshift_r(w0, w1, r), __funnelshift_r(w1, w2, r), __funnelshift_r(w2, w3, r), __funnelshift_r(w3, w4, r));
}

// items [ctid, items) of one staged unit, 256 consumer threads, four independent items per thread in flight
template <int CPR, int Q>
__device__ __forceinline__ void rows_copy_unit(uint32_t stage, uint8_t* __restrict__ dst, int items, uint32_t pitch, int cpr_rt,
                                               uint32_t r, int ctid) {
  constexpr int U = 4;
  for (int it0 = ctid; it0 < items; it0 += U * kRowsConsumers) {
    uint4 lo[U], hi[U];
#pragma unroll
    for (int k = 0; k < U; ++k) {
      const int it = it0 + k * kRowsConsumers;
      if (it < items) {
        const uint32_t row = CPR > 0 ? static_cast<uint32_t>(it) / CPR : static_cast<uint32_t>(it) / static_cast<uint32_t>(cpr_rt);
        const uint32_t c = static_cast<uint32_t>(it) - row * (CPR > 0 ? CPR : cpr_rt);
        const uint32_t a = stage + row * pitch + 16u * c;
        lo[k] = lds128(a);
        hi[k] = lds128(a + 16u);
      }
    }
#pragma unroll
    for (int k = 0; k < U; ++k) {
      const int it = it0 + k * kRowsConsumers;
      if (it < items) stg_cs128(dst + 16ll * it, realign<Q>(lo[k], hi[k], r));
    }
  }
}

template <int CPR>
__global__ void __launch_bounds__(kRowsThreads, 1)
gather_rows_tma_kernel(const __grid_constant__ CUtensorMap tmap, const RowsParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + kRowsSmemData);
  uint64_t* empty = full + kRowsMaxStages;
  RowsMeta* meta = reinterpret_cast<RowsMeta*>(empty + kRowsMaxStages);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int NS = p.nstages;
  if (tid == 0) {
    for (int s = 0; s < NS; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, kRowsConsumers / 32); }
    fence_barrier_init();
    prefetch_tmap(&tmap);
  }
  __syncthreads();
  const long long units = p.n * p.upp;
  const long long first = blockIdx.x, step = gridDim.x;
  const long long mine = first < units ? (units - first + step - 1) / step : 0;
  if (warp == 0) {
    // ============================== producer ==============================
    auto patch_of = [&](long long u) -> long long {
      const long long slot = p.upp == 1 ? u : u / p.upp;
      return p.perm ? static_cast<long long>(__ldg(p.perm + slot)) : slot;
    };
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    if (mine > 0) {
      long long i_cur = patch_of(first);
      long long i_nxt = mine > 1 ? patch_of(first + step) : 0;
      uint32_t z0 = __ldg(p.points + 3 * i_cur), y0 = __ldg(p.points + 3 * i_cur + 1), x0 = __ldg(p.points + 3 * i_cur + 2);
      int s = 0;
      uint32_t ph = 0;
      for (long long k = 0; k < mine; ++k) {
        const long long u = first + k * step;
        const long long i = i_cur;
        const uint32_t cz = z0, cy = y0, cx = x0;
        if (k + 1 < mine) {
          i_cur = i_nxt;
          z0 = __ldg(p.points + 3 * i_cur); y0 = __ldg(p.points + 3 * i_cur + 1); x0 = __ldg(p.points + 3 * i_cur + 2);
          if (k + 2 < mine) i_nxt = patch_of(u + 2 * step);
        }
        const int sub = p.upp == 1 ? 0 : static_cast<int>(u % p.upp);
        mbar_wait(empty + s, ph ^ 1);
        if (lane == 0) {
          const uint32_t xb = cx * static_cast<uint32_t>(p.es), sh = xb & 15u;
          meta[s].dst = p.out + (i * p.upp + sub) * static_cast<long long>(p.unit_out_bytes);
          meta[s].sh = sh;
          mbar_expect_tx(full + s, p.box_bytes);
          tma_load_3d_hint(smem + static_cast<size_t>(s) * p.stage_bytes, &tmap, full + s, static_cast<int>(xb - sh), static_cast<int>(cy),
                           static_cast<int>(cz) + sub * p.pzc, pol);
        }
        __syncwarp();
        if (++s == NS) { s = 0; ph ^= 1; }
      }
    }
  } else {
    // ============================== consumers ==============================
    const int ctid = tid - 32;
    const int items = p.rows_unit * p.cpr;
    int s = 0;
    uint32_t ph = 0;
    for (long long k = 0; k < mine; ++k) {
      mbar_wait(full + s, ph);
      uint8_t* dst = meta[s].dst;
      const uint32_t sh = meta[s].sh;
      const uint32_t stage = smem_u32(smem + static_cast<size_t>(s) * p.stage_bytes);
      const uint32_t r = (sh & 3u) * 8u;
      switch (sh >> 2) {                     // unit-uniform
        case 0: rows_copy_unit<CPR, 0>(stage, dst, items, p.pitch, p.cpr, r, ctid); break;
        case 1: rows_copy_unit<CPR, 1>(stage, dst, items, p.pitch, p.cpr, r, ctid); break;
        case 2: rows_copy_unit<CPR, 2>(stage, dst, items, p.pitch, p.cpr, r, ctid); break;
        default: rows_copy_unit<CPR, 3>(stage, dst, items, p.pitch, p.cpr, r, ctid); break;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty + s);     // this warp has read everything it needs from the stage
      if (++s == NS) { s = 0; ph ^= 1; }
    }
  }
}

// ------------------------------------------------------------------------------------------ scatter (disjoint patches)
// Mirror image: contiguous bulk load of a unit, TMA box store into the volume when the corner is
// 16-byte aligned, otherwise the threads write the rows with the widest aligned stores available.
template <int ES>
__global__ void __launch_bounds__(kThreadsG, 1)
scatter_tma_kernel(const __grid_constant__ CUtensorMap tmap, const TmaGatherParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) mbar_init(full + s, 1);
    fence_barrier_init();
    prefetch_tmap(&tmap);
  }
  __syncthreads();
  const long long units = p.n * p.units_per_patch;
  const long long first = blockIdx.x, step = gridDim.x;
  const long long mine = first < units ? (units - first + step - 1) / step : 0;
  const int chunks_per_row = p.px * ES / 16;
  const int chunks = p.pzc * p.py * chunks_per_row;
  long long issued = 0;
  auto issue = [&](long long k) {
    if (tid < 32) {
      const long long u = first + k * step;
      const int s = static_cast<int>(k % kStages);
      if (elect_one()) {
        mbar_expect_tx(full + s, p.dense_bytes);
        bulk_load(smem + s * kStageBytes, p.patches + u * static_cast<long long>(p.dense_bytes), p.dense_bytes, full + s);
      }
      __syncwarp();
    }
  };
  for (; issued < mine && issued < kStages - 1; ++issued) issue(issued);
  for (long long k = 0; k < mine; ++k) {
    const int s = static_cast<int>(k % kStages);
    mbar_wait(full + s, static_cast<uint32_t>((k / kStages) & 1));
    const Unit t = load_unit(p, first + k * step);
    const uint8_t* stage = smem + s * kStageBytes;
    if ((t.x0 * ES) % 16 == 0) {
      if (tid < 32) {
        if (elect_one())
          tma_store_3d(&tmap, stage, static_cast<int>(t.x0), static_cast<int>(t.y0), static_cast<int>(t.z0) + t.zc * p.pzc);
        __syncwarp();
      }
    } else {
      for (int c = tid; c < chunks; c += kThreadsG) {
        const int row = c / chunks_per_row, cx = c - row * chunks_per_row;
        const int iy = row % p.py, iz = row / p.py + t.zc * p.pzc;
        const uint4 v = *reinterpret_cast<const uint4*>(stage + static_cast<size_t>(c) * 16);
        uint8_t* d = p.vol + (((static_cast<long long>(t.z0) + iz) * p.Y + t.y0 + iy) * p.X + t.x0) * ES +
                     static_cast<long long>(cx) * 16;
        const uintptr_t a = reinterpret_cast<uintptr_t>(d);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
        if ((a & 3) == 0) {
#pragma unroll
          for (int j = 0; j < 4; ++j) reinterpret_cast<uint32_t*>(d)[j] = w[j];
        } else if (ES == 2 && (a & 1) == 0) {
#pragma unroll
          for (int e = 0; e < 8; ++e) reinterpret_cast<uint16_t*>(d)[e] = static_cast<uint16_t>(w[e / 2] >> ((e & 1) * 16));
        } else {
#pragma unroll
          for (int e = 0; e < 16; ++e) d[e] = static_cast<uint8_t>(w[e / 4] >> ((e & 3) * 8));
        }
      }
    }
    if (tid < 32) {
      if (elect_one()) {
        bulk_commit();
        if (issued < mine) bulk_wait_read<1>();
      }
      __syncwarp();
    }
    __syncthreads();
    if (issued < mine) { issue(issued); ++issued; }
  }
  if (tid < 32) {
    if (elect_one()) bulk_wait_all();
    __syncwarp();
  }
}

// Chooses the z-chunk of a unit and builds the dense and the padded tensor map; returns false when
// the TMA path does not apply.
bool plan(const void* vol, int es, const int64_t vs[3], const int32_t patch[3], CUtensorMap* tm_dense,
          CUtensorMap* tm_pad, TmaGatherParams* p) {
  if (es != 1 && es != 2 && es != 4) return false;
  const long long Z = vs[0], Y = vs[1], X = vs[2];
  if ((reinterpret_cast<uintptr_t>(vol) & 15) != 0 || (X * es) % 16 != 0) return false;
  if (X >= (1ll << 31) || Y >= (1ll << 31) || Z >= (1ll << 31)) return false;
  const int pz = patch[0], py = patch[1], px = patch[2];
  const int pade = 16 / es;
  if (px + pade > 256 || py > 256 || px < 1 || py < 1 || pz < 1) return false;
  if ((static_cast<long long>(px) * es) % 16 != 0) return false;
  const long long plane_pad = static_cast<long long>(px + pade) * py * es;
  if (plane_pad > kStageBytes) return false;
  int pzc = static_cast<int>(kStageBytes / plane_pad);
  if (pzc > pz) pzc = pz;
  if (pzc > 256) pzc = 256;
  while (pz % pzc) --pzc;
  EncodeTiledFn enc = get_encode();
  if (!enc) return false;
  cuuint64_t gdim[3] = {(cuuint64_t)X, (cuuint64_t)Y, (cuuint64_t)Z};
  cuuint64_t gstr[2] = {(cuuint64_t)X * es, (cuuint64_t)X * Y * es};
  cuuint32_t estr[3] = {1, 1, 1};
  const CUtensorMapDataType dt = es == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT8
                                 : (es == 2 ? CU_TENSOR_MAP_DATA_TYPE_UINT16 : CU_TENSOR_MAP_DATA_TYPE_UINT32);
  cuuint32_t box_d[3] = {(cuuint32_t)px, (cuuint32_t)py, (cuuint32_t)pzc};
  cuuint32_t box_p[3] = {(cuuint32_t)(px + pade), (cuuint32_t)py, (cuuint32_t)pzc};
  if (enc(tm_dense, dt, 3, const_cast<void*>(vol), gdim, gstr, box_d, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return false;
  if (enc(tm_pad, dt, 3, const_cast<void*>(vol), gdim, gstr, box_p, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return false;
  p->Y = Y; p->X = X;
  p->pz = pz; p->py = py; p->px = px;
  p->pzc = pzc;
  p->units_per_patch = pz / pzc;
  p->dense_bytes = static_cast<uint32_t>(static_cast<long long>(px) * py * es * pzc);
  p->padded_bytes = static_cast<uint32_t>(plane_pad * pzc);
  p->es = es;
  p->vol = static_cast<uint8_t*>(const_cast<void*>(vol));
  return true;
}

int grid_for_units(long long units) {
  return static_cast<int>(units < kNumSMs ? units : kNumSMs);
}

template <typename K>
cudaError_t launch_with_smem(K kernel, unsigned long long* attr_set) {
  return smem_attr_once(kernel, kSmemBytes, attr_set);
}

template <typename T>
cudaError_t launch_raw(const CUtensorMap& td, const CUtensorMap& tp, const TmaGatherParams& p, int grid, cudaStream_t st) {
  static unsigned long long attr = 0;
  cudaError_t e = launch_with_smem(gather_tma_kernel<T, T, true>, &attr);
  if (e != cudaSuccess) return e;
  gather_tma_kernel<T, T, true><<<grid, kThreadsG, kSmemBytes, st>>>(td, tp, p);
  return cudaGetLastError();
}
template <typename TIn, typename TOut>
cudaError_t launch_norm(const CUtensorMap& tp, const TmaGatherParams& p, int grid, cudaStream_t st) {
  static unsigned long long attr = 0;
  cudaError_t e = launch_with_smem(gather_tma_kernel<TIn, TOut, false>, &attr);
  if (e != cudaSuccess) return e;
  gather_tma_kernel<TIn, TOut, false><<<grid, kThreadsG, kSmemBytes, st>>>(tp, tp, p);
  return cudaGetLastError();
}
template <typename TIn>
cudaError_t launch_norm_in(const CUtensorMap& tp, const TmaGatherParams& p, int out_dtype, int grid, cudaStream_t st) {
  if (out_dtype == TE_BF16) return launch_norm<TIn, __nv_bfloat16>(tp, p, grid, st);
  if (out_dtype == TE_F16) return launch_norm<TIn, __half>(tp, p, grid, st);
  return launch_norm<TIn, float>(tp, p, grid, st);
}
template <int ES>
cudaError_t launch_scatter(const CUtensorMap& td, const TmaGatherParams& p, int grid, cudaStream_t st) {
  static unsigned long long attr = 0;
  cudaError_t e = launch_with_smem(scatter_tma_kernel<ES>, &attr);
  if (e != cudaSuccess) return e;
  scatter_tma_kernel<ES><<<grid, kThreadsG, kSmemBytes, st>>>(td, p);
  return cudaGetLastError();
}

}  // namespace

bool gather_tma(const void* vol, int es, const int64_t vs[3], const uint32_t* points, const uint32_t* widths,
                int64_t n, const int32_t patch[3], void* out, cudaStream_t st, cudaError_t* err) {
  CUtensorMap td, tp;
  TmaGatherParams p{};
  if ((reinterpret_cast<uintptr_t>(out) & 15) != 0 || !plan(vol, es, vs, patch, &td, &tp, &p)) return false;
  p.points = points; p.widths = widths; p.patches = static_cast<uint8_t*>(out); p.n = n;
  const int grid = grid_for_units(n * p.units_per_patch);
  *err = es == 1 ? launch_raw<uint8_t>(td, tp, p, grid, st)
                 : (es == 2 ? launch_raw<uint16_t>(td, tp, p, grid, st) : launch_raw<uint32_t>(td, tp, p, grid, st));
  return true;
}

template <int CPR>
static cudaError_t launch_rows(const CUtensorMap& tm, const RowsParams& p, int grid, cudaStream_t st) {
  constexpr int smem_bytes = kRowsSmemData + 1024 /*align*/ + 512 /*barriers + unit descriptors*/;
  static unsigned long long attr = 0;
  cudaError_t e = smem_attr_once(gather_rows_tma_kernel<CPR>, smem_bytes, &attr);
  if (e != cudaSuccess) return e;
  gather_rows_tma_kernel<CPR><<<grid, kRowsThreads, smem_bytes, st>>>(tm, p);
  return cudaGetLastError();
}

bool gather_rows_tma(const void* vol, int es, const int64_t vs[3], const uint32_t* points, const uint32_t* perm, int64_t n,
                     const int32_t patch[3], void* out, cudaStream_t st, cudaError_t* err) {
  if (es != 1 && es != 2 && es != 4) return false;
  const long long Z = vs[0], Y = vs[1], X = vs[2];
  const long long xb = X * es;
  if ((reinterpret_cast<uintptr_t>(vol) & 15) != 0 || (reinterpret_cast<uintptr_t>(out) & 15) != 0 || xb % 16 != 0) return false;
  if (xb >= (1ll << 31) || Y >= (1ll << 31) || Z >= (1ll << 31)) return false;
  const int pz = patch[0], py = patch[1], px = patch[2];
  const long long row_bytes = static_cast<long long>(px) * es;
  if (row_bytes % 16 != 0 || row_bytes + 16 > 256 || py > 256 || pz < 1 || py < 1) return false;
  const long long plane = (row_bytes + 16) * py;
  if (plane > kRowsStageMax) return false;
  int pzc = static_cast<int>(kRowsStageMax / plane);
  if (pzc > pz) pzc = pz;
  if (pzc > 256) pzc = 256;
  while (pz % pzc) --pzc;
  EncodeTiledFn enc = get_encode();
  if (!enc) return false;
  // byte-typed map: coordinates and extents of the innermost dimension are bytes, whatever the element size
  CUtensorMap tm;
  cuuint64_t gdim[3] = {(cuuint64_t)xb, (cuuint64_t)Y, (cuuint64_t)Z};
  cuuint64_t gstr[2] = {(cuuint64_t)xb, (cuuint64_t)xb * Y};
  cuuint32_t estr[3] = {1, 1, 1};
  cuuint32_t box[3] = {(cuuint32_t)(row_bytes + 16), (cuuint32_t)py, (cuuint32_t)pzc};
  if (enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void*>(vol), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return false;
  RowsParams p{};
  p.points = points; p.perm = perm; p.out = static_cast<uint8_t*>(out); p.n = n;
  p.es = es; p.pzc = pzc; p.upp = pz / pzc;
  p.rows_unit = py * pzc;
  p.cpr = static_cast<int>(row_bytes / 16);
  p.pitch = static_cast<uint32_t>(row_bytes + 16);
  p.box_bytes = static_cast<uint32_t>(plane * pzc);
  p.stage_bytes = (p.box_bytes + 127u) & ~127u;
  p.nstages = static_cast<int>(kRowsSmemData / p.stage_bytes);
  if (p.nstages > kRowsMaxStages) p.nstages = kRowsMaxStages;
  if (p.nstages < 2) return false;
  p.unit_out_bytes = static_cast<uint32_t>(row_bytes * p.rows_unit);
  const int grid = grid_for_units(n * p.upp);
  *err = p.cpr == 2 ? launch_rows<2>(tm, p, grid, st)
                    : (p.cpr == 4 ? launch_rows<4>(tm, p, grid, st) : (p.cpr == 8 ? launch_rows<8>(tm, p, grid, st) : launch_rows<0>(tm, p, grid, st)));
  return true;
}

bool scatter_tma(const void* sub_vols, int es, const uint32_t* points, int64_t n, const int32_t width[3], void* vol_out,
                 const int64_t vs[3], cudaStream_t st, cudaError_t* err) {
  CUtensorMap td, tp;
  TmaGatherParams p{};
  if ((reinterpret_cast<uintptr_t>(sub_vols) & 15) != 0 || !plan(vol_out, es, vs, width, &td, &tp, &p)) return false;
  p.points = points; p.widths = nullptr; p.patches = const_cast<uint8_t*>(static_cast<const uint8_t*>(sub_vols)); p.n = n;
  const int grid = grid_for_units(n * p.units_per_patch);
  *err = es == 1 ? launch_scatter<1>(td, p, grid, st) : (es == 2 ? launch_scatter<2>(td, p, grid, st) : launch_scatter<4>(td, p, grid, st));
  return true;
}

bool gather_norm_tma(const void* vol, int vol_dtype, const int64_t vs[3], const uint32_t* points,
                     const uint32_t* widths, int64_t n, const int32_t patch[3], int rescale, int do_clip, float lo,
                     float hi, void* out, int out_dtype, cudaStream_t st, cudaError_t* err) {
  CUtensorMap td, tp;
  TmaGatherParams p{};
  const int es = dtype_size(vol_dtype);
  if (patch[2] % 8 != 0 || (reinterpret_cast<uintptr_t>(out) & 15) != 0 || !plan(vol, es, vs, patch, &td, &tp, &p)) return false;
  p.points = points; p.widths = widths; p.patches = static_cast<uint8_t*>(out); p.n = n;
  p.rescale = rescale; p.do_clip = do_clip; p.lo = lo; p.hi = hi;
  const int grid = grid_for_units(n * p.units_per_patch);
  switch (vol_dtype) {
    case TE_U8: *err = launch_norm_in<uint8_t>(tp, p, out_dtype, grid, st); break;
    case TE_U16: *err = launch_norm_in<uint16_t>(tp, p, out_dtype, grid, st); break;
    case TE_F16: *err = launch_norm_in<__half>(tp, p, out_dtype, grid, st); break;
    case TE_F32: *err = launch_norm_in<float>(tp, p, out_dtype, grid, st); break;
    default: return false;
  }
  return true;
}

}  // namespace te

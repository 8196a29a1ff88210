// Volume normalisation / contrast / masking kernels run immediately before (and after) the hot path
// (SURVEY.md section 8f row 1).  Reference: /root/reference/tomo_encoders/misc/voxel_processing.py
//   te_minmax_nd, te_histogram_nd   modified_autocontrast :160-204 (np.histogram over a sub-sampled view)
//   te_rescale                      _rescale_data :81-89, normalize_volume_gpu :207-253
//   te_edge_map                     edge_map :106-122
//   te_cylindrical_mask             cylindrical_mask :124-140
//   te_cyl_values                   get_values_cyl_mask :142-157
//
// All of them are HBM-bound byte / element work: coalesced 128-bit accesses where the layout allows
// it (contiguous innermost dimension, 16-byte aligned rows), per-warp privatised histograms in
// shared memory, grids sized as multiples of the SM count.  No tensor cores.
#include "te_common.cuh"

#include <type_traits>
#include <utility>

namespace te {
namespace {

constexpr int kSeg = 4096;       // elements of the innermost dimension one block takes at a time
constexpr int kMaxBins = 1024;

// short rows: a block takes this many consecutive rows at a time (>= one per warp, >= 4 * kSeg elements)
__host__ __device__ inline int rows_per_unit(long long n3) {
  const long long r = (4LL * kSeg) / (n3 > 0 ? n3 : 1);
  return static_cast<int>(r < 8 ? 8 : r);
}

struct View4 {
  const uint8_t* base;
  long long n[4];    // logical extents (outer -> inner)
  long long st[4];   // strides in ELEMENTS
};

template <typename T> __device__ __forceinline__ float elem_f32(T v);
template <> __device__ __forceinline__ float elem_f32<uint8_t>(uint8_t v) { return static_cast<float>(v); }
template <> __device__ __forceinline__ float elem_f32<uint16_t>(uint16_t v) { return static_cast<float>(v); }
template <> __device__ __forceinline__ float elem_f32<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float elem_f32<float>(float v) { return v; }

// Calls f(value as fp32) for every element of the view.  Long rows (n[3] >= kSeg): a block takes one
// (row, segment of kSeg elements) at a time, all threads striding the segment.  Short rows: a block
// takes kSeg / n[3] consecutive rows, one warp per row, so that every lane has work whatever the row
// length (a 64^3 patch row is 64 elements).  VEC = true reads 16 bytes per lane (innermost stride 1,
// 16-byte aligned rows).
// A functor may also take a whole 16-byte word of packed integers (f.vec16(q)): the uint8 / uint16 min-max uses the
// packed-SIMD integer instructions instead of converting every element to fp32.
template <typename F, typename = void> struct has_vec16 : std::false_type {};
template <typename F>
struct has_vec16<F, std::void_t<decltype(std::declval<F&>().vec16(std::declval<const uint4&>()))>> : std::true_type {};

template <typename T, bool VEC, typename F>
__device__ __forceinline__ void for_each_in_row(const T* rp, long long sx, int len, int first, int step, F&& f) {
  constexpr int EPV = 16 / sizeof(T);
  if (VEC) {
    const int nvec = len / EPV;
    for (int c = first; c < nvec; c += step) {
      const uint4 q = __ldg(reinterpret_cast<const uint4*>(rp) + c);
      if constexpr (has_vec16<std::decay_t<F>>::value) {
        f.vec16(q);
      } else {
        const T* e = reinterpret_cast<const T*>(&q);
#pragma unroll
        for (int i = 0; i < EPV; ++i) f(elem_f32<T>(e[i]));
      }
    }
    for (int x = nvec * EPV + first; x < len; x += step) f(elem_f32<T>(rp[x]));
  } else {
    for (int x = first; x < len; x += step) f(elem_f32<T>(rp[x * sx]));
  }
}

template <typename T, bool VEC, typename F>
__device__ __forceinline__ void for_each_elem(const View4& v, F&& f) {
  const long long rows = v.n[0] * v.n[1] * v.n[2];
  const T* base = reinterpret_cast<const T*>(v.base);
  if (v.n[3] >= kSeg) {
    const long long segs = (v.n[3] + kSeg - 1) / kSeg;
    const long long units = rows * segs;
    for (long long u = blockIdx.x; u < units; u += gridDim.x) {
      const long long row = u / segs, seg = u - row * segs;
      const long long i2 = row % v.n[2], t = row / v.n[2];
      const long long i1 = t % v.n[1], i0 = t / v.n[1];
      const long long x0 = seg * kSeg;                 // x0 * sizeof(T) is a multiple of 16
      const int len = static_cast<int>(v.n[3] - x0 < kSeg ? v.n[3] - x0 : kSeg);
      const T* rp = base + i0 * v.st[0] + i1 * v.st[1] + i2 * v.st[2] + x0 * v.st[3];
      for_each_in_row<T, VEC>(rp, v.st[3], len, threadIdx.x, blockDim.x, f);
    }
  } else {
    const int rpu = rows_per_unit(v.n[3]);
    const long long units = (rows + rpu - 1) / rpu;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int len = static_cast<int>(v.n[3]);
    for (long long u = blockIdx.x; u < units; u += gridDim.x) {
      const long long r0 = u * rpu;
      const long long r1 = r0 + rpu < rows ? r0 + rpu : rows;
      long long row = r0 + warp;
      if (row >= r1) continue;
      long long i2 = row % v.n[2], t = row / v.n[2];
      long long i1 = t % v.n[1], i0 = t / v.n[1];
      for (; row < r1; row += nwarps) {
        const T* rp = base + i0 * v.st[0] + i1 * v.st[1] + i2 * v.st[2];
        for_each_in_row<T, VEC>(rp, v.st[3], len, lane, 32, f);
        i2 += nwarps;                                  // advance the row index with carries (no division)
        while (i2 >= v.n[2]) { i2 -= v.n[2]; if (++i1 >= v.n[1]) { i1 = 0; ++i0; } }
      }
    }
  }
}

template <typename T> struct MinMaxAcc {
  float lo = INFINITY, hi = -INFINITY;
  __device__ __forceinline__ void operator()(float a) { lo = fminf(lo, a); hi = fmaxf(hi, a); }
  __device__ __forceinline__ void finish() {}
};
template <> struct MinMaxAcc<uint8_t> {
  float lo = INFINITY, hi = -INFINITY;
  unsigned lo4 = 0xffffffffu, hi4 = 0u;
  __device__ __forceinline__ void operator()(float a) { lo = fminf(lo, a); hi = fmaxf(hi, a); }
  __device__ __forceinline__ void vec16(const uint4& q) {
    lo4 = __vminu4(__vminu4(lo4, q.x), __vminu4(__vminu4(q.y, q.z), q.w));
    hi4 = __vmaxu4(__vmaxu4(hi4, q.x), __vmaxu4(__vmaxu4(q.y, q.z), q.w));
  }
  __device__ __forceinline__ void finish() {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      lo = fminf(lo, static_cast<float>((lo4 >> (8 * i)) & 255u));
      hi = fmaxf(hi, static_cast<float>((hi4 >> (8 * i)) & 255u));
    }
  }
};
template <> struct MinMaxAcc<uint16_t> {
  float lo = INFINITY, hi = -INFINITY;
  unsigned lo2 = 0xffffffffu, hi2 = 0u;
  __device__ __forceinline__ void operator()(float a) { lo = fminf(lo, a); hi = fmaxf(hi, a); }
  __device__ __forceinline__ void vec16(const uint4& q) {
    lo2 = __vminu2(__vminu2(lo2, q.x), __vminu2(__vminu2(q.y, q.z), q.w));
    hi2 = __vmaxu2(__vmaxu2(hi2, q.x), __vmaxu2(__vmaxu2(q.y, q.z), q.w));
  }
  __device__ __forceinline__ void finish() {
    lo = fminf(lo, fminf(static_cast<float>(lo2 & 0xffffu), static_cast<float>(lo2 >> 16)));
    hi = fmaxf(hi, fmaxf(static_cast<float>(hi2 & 0xffffu), static_cast<float>(hi2 >> 16)));
  }
};

template <typename T, bool VEC>
__global__ void __launch_bounds__(256) minmax_nd_kernel(View4 v, float* partial) {
  MinMaxAcc<T> acc;
  for_each_elem<T, VEC>(v, acc);
  acc.finish();
  float lo = acc.lo, hi = acc.hi;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  __shared__ float slo[8], shi[8];
  if ((threadIdx.x & 31) == 0) { slo[threadIdx.x >> 5] = lo; shi[threadIdx.x >> 5] = hi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) { lo = fminf(lo, slo[w]); hi = fmaxf(hi, shi[w]); }
    partial[2 * blockIdx.x] = lo;
    partial[2 * blockIdx.x + 1] = hi;
  }
}

// C-contiguous, 16-byte aligned data: plain grid-stride loop over 16-byte words, four independent loads per thread and trip
// (the per-row traversal above pays its index arithmetic for every 1 KB row of a uint8 volume: 2.5 TB/s)
template <typename T>
__global__ void __launch_bounds__(256) minmax_flat_kernel(const T* __restrict__ data, long long n, float* partial) {
  constexpr int EPV = 16 / sizeof(T);
  MinMaxAcc<T> acc;
  const long long nvec = n / EPV;
  const uint4* p = reinterpret_cast<const uint4*>(data);
  const long long stride = static_cast<long long>(gridDim.x) * 256;
  long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  auto take = [&](const uint4& q) {
    if constexpr (has_vec16<MinMaxAcc<T>>::value) {
      acc.vec16(q);
    } else {
      const T* e = reinterpret_cast<const T*>(&q);
#pragma unroll
      for (int k = 0; k < EPV; ++k) acc(elem_f32<T>(e[k]));
    }
  };
  for (; i + 3 * stride < nvec; i += 4 * stride) {
    const uint4 q0 = __ldg(p + i), q1 = __ldg(p + i + stride), q2 = __ldg(p + i + 2 * stride), q3 = __ldg(p + i + 3 * stride);
    take(q0); take(q1); take(q2); take(q3);
  }
  for (; i < nvec; i += stride) take(__ldg(p + i));
  if (blockIdx.x == 0) for (long long x = nvec * EPV + threadIdx.x; x < n; x += 256) acc(elem_f32<T>(data[x]));
  acc.finish();
  float lo = acc.lo, hi = acc.hi;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  __shared__ float slo[8], shi[8];
  if ((threadIdx.x & 31) == 0) { slo[threadIdx.x >> 5] = lo; shi[threadIdx.x >> 5] = hi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) { lo = fminf(lo, slo[w]); hi = fmaxf(hi, shi[w]); }
    partial[2 * blockIdx.x] = lo;
    partial[2 * blockIdx.x + 1] = hi;
  }
}

__global__ void minmax_nd_final_kernel(const float* partial, int nblocks, float* out2) {
  float lo = INFINITY, hi = -INFINITY;
  for (int i = threadIdx.x; i < nblocks; i += 32) {
    lo = fminf(lo, partial[2 * i]);
    hi = fmaxf(hi, partial[2 * i + 1]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if (threadIdx.x == 0) { out2[0] = lo; out2[1] = hi; }
}

// np.histogram with explicit, monotone bin edges: bin i holds edges[i] <= a < edges[i+1], the last
// bin is closed (numpy/lib/_histograms_impl.py: the uniform-bin estimate followed by the correction
// against the edges -- the corrected index IS this definition, so the estimate's rounding is free).
// Values outside [edges[0], edges[nbins]] and NaNs are dropped, like numpy's `keep` mask.
// All comparisons run in fp32 and are still EXACT: every supported element is a float, and for a float a and a double
// edge e,  a >= e  <=>  a >= ru(e)  (ru = e rounded UP to float), hence a < e <=> a < ru(e); the closed upper end
// a <= eN <=> a <= rd(eN).  No fp64 instruction in the element loop.
template <typename T, bool VEC>
__global__ void __launch_bounds__(256) histogram_nd_kernel(View4 v, const double* __restrict__ edges, int nbins,
                                                           unsigned long long* __restrict__ counts) {
  extern __shared__ double hsm[];
  float* ed = reinterpret_cast<float*>(hsm);                            // nbins + 1 edges, rounded up
  unsigned int* hist = reinterpret_cast<unsigned int*>(ed + nbins + 2);  // 8 warps x nbins
  for (int i = threadIdx.x; i <= nbins; i += blockDim.x) ed[i] = __double2float_ru(edges[i]);
  for (int i = threadIdx.x; i < 8 * nbins; i += blockDim.x) hist[i] = 0u;
  __syncthreads();
  const double e0d = edges[0], eNd = edges[nbins];
  const float e0 = ed[0], eN = __double2float_rd(eNd);
  const float scale = eNd > e0d ? static_cast<float>(nbins / (eNd - e0d)) : 0.f;
  unsigned int* mine = hist + (threadIdx.x >> 5) * nbins;
  // neighbouring voxels of a tomogram mostly fall into the same bin: a thread keeps the bin of its previous
  // element and a run count, and touches shared memory only when the bin changes
  int cur = -1;
  unsigned int run = 0;
  float clo = 0.f, chi = 0.f;
  for_each_elem<T, VEC>(v, [&](float a) {
    if (cur >= 0 && a >= clo && a < chi) { ++run; return; }
    if (!(a >= e0 && a <= eN)) return;
    int i = static_cast<int>((a - e0) * scale);        // estimate; the comparisons below decide
    i = i < 0 ? 0 : (i > nbins - 1 ? nbins - 1 : i);
    while (i > 0 && a < ed[i]) --i;
    while (i < nbins - 1 && a >= ed[i + 1]) ++i;
    if (i == cur) { ++run; return; }                   // closed last bin: a == eN
    if (run) atomicAdd(mine + cur, run);
    cur = i; run = 1; clo = ed[i]; chi = ed[i + 1];
  });
  if (run) atomicAdd(mine + cur, run);
  __syncthreads();
  for (int i = threadIdx.x; i < nbins; i += blockDim.x) {
    unsigned long long s = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += hist[w * nbins + i];
    if (s) atomicAdd(counts + i, s);
  }
}

// uint8 volumes have 256 possible values: count them directly (per-warp privatised 256-entry tables, one shared-memory
// atomic per element, no bin search), then thread k maps value k to its bin with the same exact comparisons.
template <bool VEC>
__global__ void __launch_bounds__(256) histogram_u8_kernel(View4 v, const double* __restrict__ edges, int nbins,
                                                           unsigned long long* __restrict__ counts) {
  __shared__ unsigned int hist[8][256];
  for (int i = threadIdx.x; i < 8 * 256; i += blockDim.x) (&hist[0][0])[i] = 0u;
  __syncthreads();
  unsigned int* mine = hist[threadIdx.x >> 5];
  const long long rows = v.n[0] * v.n[1] * v.n[2];
  const uint8_t* base = v.base;
  if (VEC && v.n[3] >= kSeg) {
    // raw 16-byte words: four bytes at a time, equal neighbours merged before the atomic
    const long long segs = (v.n[3] + kSeg - 1) / kSeg;
    const long long units = rows * segs;
    for (long long u = blockIdx.x; u < units; u += gridDim.x) {
      const long long row = u / segs, seg = u - row * segs;
      const long long i2 = row % v.n[2], t = row / v.n[2];
      const long long i1 = t % v.n[1], i0 = t / v.n[1];
      const long long x0 = seg * kSeg;
      const int len = static_cast<int>(v.n[3] - x0 < kSeg ? v.n[3] - x0 : kSeg);
      const uint8_t* rp = base + i0 * v.st[0] + i1 * v.st[1] + i2 * v.st[2] + x0;
      const int nvec = len / 16;
      for (int c = threadIdx.x; c < nvec; c += blockDim.x) {
        const uint4 q = __ldg(reinterpret_cast<const uint4*>(rp) + c);
        const unsigned w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const unsigned b0 = w[k] & 255u, b1 = (w[k] >> 8) & 255u, b2 = (w[k] >> 16) & 255u, b3 = w[k] >> 24;
          if (b0 == b1 && b1 == b2 && b2 == b3) { atomicAdd(mine + b0, 4u); continue; }
          atomicAdd(mine + b0, 1u); atomicAdd(mine + b1, 1u); atomicAdd(mine + b2, 1u); atomicAdd(mine + b3, 1u);
        }
      }
      for (int x = nvec * 16 + threadIdx.x; x < len; x += blockDim.x) atomicAdd(mine + rp[x], 1u);
    }
  } else {
    for_each_elem<uint8_t, VEC>(v, [&](float a) { atomicAdd(mine + static_cast<int>(a), 1u); });
  }
  __syncthreads();
  // thread k owns value k
  const int k = threadIdx.x;
  unsigned long long s = 0;
#pragma unroll
  for (int w = 0; w < 8; ++w) s += hist[w][k];
  if (s == 0) return;
  const double a = static_cast<double>(k);
  const double e0 = edges[0], eN = edges[nbins];
  if (!(a >= e0 && a <= eN)) return;
  int i = eN > e0 ? static_cast<int>((a - e0) * (nbins / (eN - e0))) : 0;
  i = i < 0 ? 0 : (i > nbins - 1 ? nbins - 1 : i);
  while (i > 0 && a < edges[i]) --i;
  while (i < nbins - 1 && a >= edges[i + 1]) ++i;
  atomicAdd(counts + i, s);
}

// out = (in - lo) / denom in fp32 (IEEE round-to-nearest subtract and divide: bit-exact against numpy
// float32 arithmetic), converted to the output type.  in == out is allowed for equal element sizes.
template <typename TIn, typename TOut>
__global__ void __launch_bounds__(256) rescale_kernel(const TIn* in, TOut* out, long long n, float lo, float denom) {
  constexpr int V = 16 / (sizeof(TIn) > sizeof(TOut) ? sizeof(TIn) : sizeof(TOut));   // elements per thread
  const long long nv = n / V;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long c = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; c < nv; c += stride) {
    alignas(16) TIn a[V];
    alignas(16) TOut b[V];
    if (sizeof(TIn) * V == 16) *reinterpret_cast<uint4*>(a) = *(reinterpret_cast<const uint4*>(in) + c);
    else if (sizeof(TIn) * V == 8) *reinterpret_cast<uint2*>(a) = *(reinterpret_cast<const uint2*>(in) + c);
    else *reinterpret_cast<uint32_t*>(a) = *(reinterpret_cast<const uint32_t*>(in) + c);
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const float r = __fdiv_rn(__fsub_rn(elem_f32<TIn>(a[i]), lo), denom);
      if (sizeof(TOut) == 4) reinterpret_cast<float*>(b)[i] = r;
      else reinterpret_cast<__half*>(b)[i] = __float2half_rn(r);
    }
    if (sizeof(TOut) * V == 16) *(reinterpret_cast<uint4*>(out) + c) = *reinterpret_cast<uint4*>(b);
    else *(reinterpret_cast<uint2*>(out) + c) = *reinterpret_cast<uint2*>(b);
  }
  for (long long e = nv * V + static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < n; e += stride) {
    const float r = __fdiv_rn(__fsub_rn(elem_f32<TIn>(in[e]), lo), denom);
    if (sizeof(TOut) == 4) reinterpret_cast<float*>(out)[e] = r;
    else reinterpret_cast<__half*>(out)[e] = __float2half_rn(r);
  }
}

// ---- edge_map -------------------------------------------------------------------------------------
// out[z,y,x] = 1 when any in-bounds 6-neighbour holds a different value (bit pattern), else 0.
// Vector kernel: a thread owns 16 bytes of a row; x neighbours come from funnel shifts of the
// adjacent 32-bit words, comparisons are SIMD-in-register (__vcmpne4 / __vcmpne2 / !=).
// per-element "is non-zero" of a 32-bit word holding 4 / 2 / 1 elements -> 1 in the lowest bit of each element
template <int EB> __device__ __forceinline__ uint32_t nonzero_elems(uint32_t d);
template <> __device__ __forceinline__ uint32_t nonzero_elems<1>(uint32_t d) {
  return ((((d & 0x7f7f7f7fu) + 0x7f7f7f7fu) | d) >> 7) & 0x01010101u;
}
template <> __device__ __forceinline__ uint32_t nonzero_elems<2>(uint32_t d) {
  return ((((d & 0x7fff7fffu) + 0x7fff7fffu) | d) >> 15) & 0x00010001u;
}
template <> __device__ __forceinline__ uint32_t nonzero_elems<4>(uint32_t d) { return d != 0u ? 1u : 0u; }

// A block takes 8 consecutive rows (one warp per row: the y neighbours of a row are read by the adjacent
// warps of the same block and hit L1, the z neighbours hit L2); a lane takes 16 bytes of the row at a time.
template <int EB>
__global__ void __launch_bounds__(256) edge_map_vec_kernel(const uint8_t* __restrict__ Y, long long Z, long long Yn,
                                                           long long row_bytes, uint8_t* __restrict__ out) {
  const int cpr = static_cast<int>(row_bytes / 16);          // chunks per row
  const long long rows = Z * Yn;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (long long row = static_cast<long long>(blockIdx.x) * 8 + warp; row < rows; row += static_cast<long long>(gridDim.x) * 8) {
    const long long y = row % Yn, z = row / Yn;
    const uint8_t* rp = Y + row * row_bytes;
    const bool has_ym = y > 0, has_yp = y + 1 < Yn, has_zm = z > 0, has_zp = z + 1 < Z;
    const long long plane = Yn * row_bytes;
    uint8_t* orow = out + row * (row_bytes / EB);
    for (int c = lane; c < cpr; c += 32) {
      const uint8_t* p = rp + c * 16;
      const uint4 q = __ldg(reinterpret_cast<const uint4*>(p));
      const uint4 ym = has_ym ? __ldg(reinterpret_cast<const uint4*>(p - row_bytes)) : q;
      const uint4 yp = has_yp ? __ldg(reinterpret_cast<const uint4*>(p + row_bytes)) : q;
      const uint4 zm = has_zm ? __ldg(reinterpret_cast<const uint4*>(p - plane)) : q;
      const uint4 zp = has_zp ? __ldg(reinterpret_cast<const uint4*>(p + plane)) : q;
      // word before / after the chunk; at the row ends replicate the border element (no difference)
      const uint32_t prev = c > 0 ? __ldg(reinterpret_cast<const uint32_t*>(p) - 1)
                                  : (EB == 4 ? q.x : q.x << (32 - 8 * EB));
      const uint32_t next = c + 1 < cpr ? __ldg(reinterpret_cast<const uint32_t*>(p) + 4)
                                        : (EB == 4 ? q.w : q.w >> (32 - 8 * EB));
      const uint32_t w[6] = {prev, q.x, q.y, q.z, q.w, next};
      const uint32_t a[4] = {ym.x, ym.y, ym.z, ym.w}, b[4] = {yp.x, yp.y, yp.z, yp.w};
      const uint32_t d[4] = {zm.x, zm.y, zm.z, zm.w}, e[4] = {zp.x, zp.y, zp.z, zp.w};
      uint32_t m[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint32_t cur = w[i + 1];
        const uint32_t left = EB == 4 ? w[i] : __funnelshift_l(w[i], cur, 8 * EB);      // element x-1 under element x
        const uint32_t right = EB == 4 ? w[i + 2] : __funnelshift_r(cur, w[i + 2], 8 * EB);
        const uint32_t diff = (cur ^ left) | (cur ^ right) | (cur ^ a[i]) | (cur ^ b[i]) | (cur ^ d[i]) | (cur ^ e[i]);
        m[i] = nonzero_elems<EB>(diff);
      }
      uint8_t* o = orow + c * (16 / EB);
      if (EB == 1) {
        *reinterpret_cast<uint4*>(o) = make_uint4(m[0], m[1], m[2], m[3]);
      } else if (EB == 2) {
        // 8 elements -> 8 bytes (element flags sit at bits 0 and 16 of each word)
        *reinterpret_cast<uint2*>(o) = make_uint2((m[0] & 1u) | ((m[0] >> 8) & 0x100u) | ((m[1] & 1u) << 16) | ((m[1] << 8) & 0x01000000u),
                                                  (m[2] & 1u) | ((m[2] >> 8) & 0x100u) | ((m[3] & 1u) << 16) | ((m[3] << 8) & 0x01000000u));
      } else {
        *reinterpret_cast<uint32_t*>(o) = m[0] | (m[1] << 8) | (m[2] << 16) | (m[3] << 24);
      }
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256) edge_map_scalar_kernel(const T* __restrict__ Y, long long Z, long long Yn, long long X,
                                                              uint8_t* __restrict__ out) {
  const long long total = Z * Yn * X;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long g = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; g < total; g += stride) {
    const long long x = g % X, t = g / X;
    const long long y = t % Yn, z = t / Yn;
    const T c = Y[g];
    bool d = false;
    if (x > 0) d |= Y[g - 1] != c;
    if (x + 1 < X) d |= Y[g + 1] != c;
    if (y > 0) d |= Y[g - X] != c;
    if (y + 1 < Yn) d |= Y[g + X] != c;
    if (z > 0) d |= Y[g - Yn * X] != c;
    if (z + 1 < Z) d |= Y[g + Yn * X] != c;
    out[g] = d ? 1 : 0;
  }
}

// ---- cylindrical mask -----------------------------------------------------------------------------
// inside(y, x) = (y - n//2)^2 + (x - n//2)^2 < rad^2 (rad > 0), n = shape[1] = shape[2].

// A warp takes one row: the inside interval [lo, hi) is found analytically (float sqrt + integer fix-up) and
// only the outside voxels are written (the inside ones are never touched: HBM traffic = bytes written).
template <typename T>
__global__ void __launch_bounds__(256) cyl_mask_kernel(T* __restrict__ vol, long long Z, long long n, long long rad, T val) {
  constexpr int V = 16 / sizeof(T);
  const long long half = n / 2;
  const long long rows = Z * n;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool vec_ok = (n % V) == 0 && (reinterpret_cast<uintptr_t>(vol) & 15) == 0;
  alignas(16) T fill[V];
#pragma unroll
  for (int i = 0; i < V; ++i) fill[i] = val;
  for (long long row = static_cast<long long>(blockIdx.x) * 8 + warp; row < rows; row += static_cast<long long>(gridDim.x) * 8) {
    const long long y = row % n;
    const long long yy = y - half;
    long long lo = 0, hi = 0;                                   // empty interval: everything outside
    const long long rem = rad > 0 ? rad * rad - yy * yy : 0;    // inside iff xx^2 < rem
    if (rem > 0) {
      long long w = static_cast<long long>(sqrt(static_cast<double>(rem)));      // largest w with w^2 <= rem (after fix-up)
      while (w * w > rem) --w;
      while ((w + 1) * (w + 1) <= rem) ++w;
      const long long wi = w * w == rem ? w - 1 : w;            // largest |xx| strictly inside
      lo = half - wi < 0 ? 0 : half - wi;
      hi = half + wi + 1 > n ? n : half + wi + 1;
    }
    T* rp = vol + row * n;
    // two outside runs: [0, lo) and [hi, n)
#pragma unroll
    for (int side = 0; side < 2; ++side) {
      const long long a = side == 0 ? 0 : hi, b = side == 0 ? lo : n;
      if (a >= b) continue;
      if (vec_ok) {
        const long long va = (a + V - 1) / V * V, vb = b / V * V;      // aligned core
        if (va < vb) {
          for (long long x = a + lane; x < va; x += 32) rp[x] = val;
          for (long long x = va + lane * V; x < vb; x += 32 * V) *reinterpret_cast<uint4*>(rp + x) = *reinterpret_cast<const uint4*>(fill);
          for (long long x = vb + lane; x < b; x += 32) rp[x] = val;
          continue;
        }
      }
      for (long long x = a + lane; x < b; x += 32) rp[x] = val;
    }
  }
}

// out[z * per_slice + row_off[y] + (x - x_lo[y])] = vol[z, y, x] for x in [x_lo[y], x_hi[y])  (C order of vol[cyl > 0])
template <typename T>
__global__ void __launch_bounds__(256) cyl_values_kernel(const T* __restrict__ vol, long long Z, long long n,
                                                         long long sz, long long sy, long long sx,
                                                         const int* __restrict__ x_lo, const int* __restrict__ x_hi,
                                                         const long long* __restrict__ row_off, long long per_slice,
                                                         T* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const long long warp = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  for (long long row = warp; row < Z * n; row += nwarps) {
    const long long y = row % n, z = row / n;
    const int lo = x_lo[y], hi = x_hi[y];
    const T* src = vol + z * sz + y * sy;
    T* dst = out + z * per_slice + row_off[y] - lo;
    for (int x = lo + lane; x < hi; x += 32) dst[x] = src[x * sx];
  }
}

static int blocks_for(long long items, int per_block, int waves = 8) {
  long long b = (items + per_block - 1) / per_block;
  const long long cap = static_cast<long long>(kNumSMs) * waves;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return static_cast<int>(b);
}

static bool make_view(View4& v, const void* data, int dtype, const int64_t shape[4], const int64_t* strides, bool* vec) {
  v.base = static_cast<const uint8_t*>(data);
  long long total = 1;
  for (int i = 0; i < 4; ++i) {
    if (shape[i] < 0 || strides[i] < 0) return false;
    v.n[i] = shape[i];
    v.st[i] = strides[i];
    total *= shape[i];
  }
  const int es = dtype_size(dtype);
  *vec = v.st[3] == 1 && (reinterpret_cast<uintptr_t>(data) & 15) == 0 && (v.st[0] * es) % 16 == 0 &&
         (v.st[1] * es) % 16 == 0 && (v.st[2] * es) % 16 == 0;
  return total >= 0;
}

static long long view_units(const View4& v) {
  const long long rows = v.n[0] * v.n[1] * v.n[2];
  if (v.n[3] >= kSeg) return rows * ((v.n[3] + kSeg - 1) / kSeg);
  if (v.n[3] == 0) return 0;
  const int rpu = rows_per_unit(v.n[3]);
  return (rows + rpu - 1) / rpu;
}

}  // namespace
}  // namespace te

using namespace te;

#define TE_DISPATCH_DTYPE(dt, ...)                        \
  switch (dt) {                                           \
    case TE_U8: { using T = uint8_t; __VA_ARGS__; break; }       \
    case TE_U16: { using T = uint16_t; __VA_ARGS__; break; }     \
    case TE_F16: { using T = __half; __VA_ARGS__; break; }       \
    case TE_F32: { using T = float; __VA_ARGS__; break; }        \
    default: ::te::set_error("unsupported dtype %d", dt); return TE_ERR_ARG; \
  }

extern "C" int te_minmax_nd(const void* data, int dtype, const int64_t shape[4], const int64_t strides[4],
                            float* out2, float* scratch, void* stream) {
  TE_CHECK_ARG(data && out2 && scratch, "te_minmax_nd: null pointer");
  View4 v;
  bool vec;
  TE_CHECK_ARG(make_view(v, data, dtype, shape, strides, &vec), "te_minmax_nd: bad shape / strides");
  TE_CHECK_ARG(view_units(v) > 0, "te_minmax_nd: empty view");
  int blocks = blocks_for(view_units(v), 1, 4);   // <= 592: scratch holds 2 floats per block
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  bool flat = (reinterpret_cast<uintptr_t>(data) & 15) == 0;
  long long total = 1;
  for (int i = 3; i >= 0; --i) {
    if (v.n[i] > 1 && v.st[i] != total) flat = false;
    total *= v.n[i];
  }
  if (flat) {
    blocks = kNumSMs * 4;
    TE_DISPATCH_DTYPE(dtype, { minmax_flat_kernel<T><<<blocks, 256, 0, st>>>(reinterpret_cast<const T*>(data), total, scratch); });
  } else {
    TE_DISPATCH_DTYPE(dtype, {
      if (vec) minmax_nd_kernel<T, true><<<blocks, 256, 0, st>>>(v, scratch);
      else minmax_nd_kernel<T, false><<<blocks, 256, 0, st>>>(v, scratch);
    });
  }
  TE_LAUNCH_CHECK();
  minmax_nd_final_kernel<<<1, 32, 0, st>>>(scratch, blocks, out2);
  TE_LAUNCH_CHECK();
  return TE_OK;
}

extern "C" int te_histogram_nd(const void* data, int dtype, const int64_t shape[4], const int64_t strides[4],
                               const double* edges, int32_t nbins, unsigned long long* counts, void* stream) {
  TE_CHECK_ARG(data && edges && counts, "te_histogram_nd: null pointer");
  TE_CHECK_ARG(nbins >= 1 && nbins <= kMaxBins, "te_histogram_nd: nbins must be in [1, %d]", kMaxBins);
  View4 v;
  bool vec;
  TE_CHECK_ARG(make_view(v, data, dtype, shape, strides, &vec), "te_histogram_nd: bad shape / strides");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TE_CUDA(cudaMemsetAsync(counts, 0, sizeof(unsigned long long) * nbins, st));
  if (view_units(v) == 0) return TE_OK;
  const int blocks = blocks_for(view_units(v), 1, 4);
  const size_t smem = sizeof(float) * (nbins + 2) + sizeof(unsigned int) * 8 * nbins;
  if (dtype == TE_U8) {
    if (vec) histogram_u8_kernel<true><<<blocks, 256, 0, st>>>(v, edges, nbins, counts);
    else histogram_u8_kernel<false><<<blocks, 256, 0, st>>>(v, edges, nbins, counts);
    TE_LAUNCH_CHECK();
    return TE_OK;
  }
  TE_DISPATCH_DTYPE(dtype, {
    if (vec) histogram_nd_kernel<T, true><<<blocks, 256, smem, st>>>(v, edges, nbins, counts);
    else histogram_nd_kernel<T, false><<<blocks, 256, smem, st>>>(v, edges, nbins, counts);
  });
  TE_LAUNCH_CHECK();
  return TE_OK;
}

extern "C" int te_rescale(const void* in, int in_dtype, void* out, int out_dtype, int64_t n, float lo, float denom,
                          void* stream) {
  TE_CHECK_ARG(n >= 0, "te_rescale: n < 0");
  if (n == 0) return TE_OK;
  TE_CHECK_ARG(in && out, "te_rescale: null pointer");
  TE_CHECK_ARG(out_dtype == TE_F32 || out_dtype == TE_F16, "te_rescale: out_dtype must be TE_F32 or TE_F16");
  TE_CHECK_ARG((reinterpret_cast<uintptr_t>(in) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
               "te_rescale: buffers must be 16-byte aligned");
  TE_CHECK_ARG(in != out || dtype_size(in_dtype) == dtype_size(out_dtype), "te_rescale: in-place needs equal element sizes");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int blocks = blocks_for(n, 256 * 8);
  TE_DISPATCH_DTYPE(in_dtype, {
    if (out_dtype == TE_F32) rescale_kernel<T, float><<<blocks, 256, 0, st>>>(static_cast<const T*>(in), static_cast<float*>(out), n, lo, denom);
    else rescale_kernel<T, __half><<<blocks, 256, 0, st>>>(static_cast<const T*>(in), static_cast<__half*>(out), n, lo, denom);
  });
  TE_LAUNCH_CHECK();
  return TE_OK;
}

extern "C" int te_edge_map(const void* Y, int elem_size, const int64_t shape[3], uint8_t* out, void* stream) {
  TE_CHECK_ARG(elem_size == 1 || elem_size == 2 || elem_size == 4, "te_edge_map: elem_size must be 1, 2 or 4");
  TE_CHECK_ARG(shape[0] >= 0 && shape[1] >= 0 && shape[2] >= 0, "te_edge_map: negative extent");
  const long long total = static_cast<long long>(shape[0]) * shape[1] * shape[2];
  if (total == 0) return TE_OK;
  TE_CHECK_ARG(Y && out, "te_edge_map: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long row_bytes = static_cast<long long>(shape[2]) * elem_size;
  const bool vec = row_bytes % 16 == 0 && (reinterpret_cast<uintptr_t>(Y) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(out) & 15) == 0;
  if (vec) {
    const int blocks = blocks_for(static_cast<long long>(shape[0]) * shape[1], 8, 16);
    const uint8_t* y8 = static_cast<const uint8_t*>(Y);
    if (elem_size == 1) edge_map_vec_kernel<1><<<blocks, 256, 0, st>>>(y8, shape[0], shape[1], row_bytes, out);
    else if (elem_size == 2) edge_map_vec_kernel<2><<<blocks, 256, 0, st>>>(y8, shape[0], shape[1], row_bytes, out);
    else edge_map_vec_kernel<4><<<blocks, 256, 0, st>>>(y8, shape[0], shape[1], row_bytes, out);
  } else {
    const int blocks = blocks_for(total, 256);
    if (elem_size == 1) edge_map_scalar_kernel<uint8_t><<<blocks, 256, 0, st>>>(static_cast<const uint8_t*>(Y), shape[0], shape[1], shape[2], out);
    else if (elem_size == 2) edge_map_scalar_kernel<uint16_t><<<blocks, 256, 0, st>>>(static_cast<const uint16_t*>(Y), shape[0], shape[1], shape[2], out);
    else edge_map_scalar_kernel<uint32_t><<<blocks, 256, 0, st>>>(static_cast<const uint32_t*>(Y), shape[0], shape[1], shape[2], out);
  }
  TE_LAUNCH_CHECK();
  return TE_OK;
}

extern "C" int te_cylindrical_mask(void* vol, int elem_size, const int64_t shape[3], int64_t rad, uint32_t value_bits,
                                   void* stream) {
  TE_CHECK_ARG(elem_size == 1 || elem_size == 2 || elem_size == 4, "te_cylindrical_mask: elem_size must be 1, 2 or 4");
  TE_CHECK_ARG(shape[0] >= 0 && shape[1] >= 0 && shape[1] == shape[2],
               "te_cylindrical_mask: must be a tomographic volume where shape y = shape x");
  const long long total = static_cast<long long>(shape[0]) * shape[1] * shape[2];
  if (total == 0) return TE_OK;
  TE_CHECK_ARG(vol, "te_cylindrical_mask: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int blocks = blocks_for(static_cast<long long>(shape[0]) * shape[1], 8, 16);
  if (elem_size == 1) cyl_mask_kernel<uint8_t><<<blocks, 256, 0, st>>>(static_cast<uint8_t*>(vol), shape[0], shape[1], rad, static_cast<uint8_t>(value_bits));
  else if (elem_size == 2) cyl_mask_kernel<uint16_t><<<blocks, 256, 0, st>>>(static_cast<uint16_t*>(vol), shape[0], shape[1], rad, static_cast<uint16_t>(value_bits));
  else cyl_mask_kernel<uint32_t><<<blocks, 256, 0, st>>>(static_cast<uint32_t*>(vol), shape[0], shape[1], rad, value_bits);
  TE_LAUNCH_CHECK();
  return TE_OK;
}

extern "C" int te_cyl_values(const void* vol, int elem_size, const int64_t shape[3], const int64_t strides[3],
                             const int32_t* x_lo, const int32_t* x_hi, const int64_t* row_off, int64_t per_slice,
                             void* out, void* stream) {
  TE_CHECK_ARG(elem_size == 1 || elem_size == 2 || elem_size == 4, "te_cyl_values: elem_size must be 1, 2 or 4");
  TE_CHECK_ARG(shape[0] >= 0 && shape[1] >= 0 && shape[1] == shape[2],
               "te_cyl_values: must be a tomographic volume where shape y = shape x");
  if (static_cast<long long>(shape[0]) * per_slice == 0) return TE_OK;
  TE_CHECK_ARG(vol && x_lo && x_hi && row_off && out, "te_cyl_values: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int blocks = blocks_for(static_cast<long long>(shape[0]) * shape[1], 8);
  const long long* ro = reinterpret_cast<const long long*>(row_off);
  if (elem_size == 1) cyl_values_kernel<uint8_t><<<blocks, 256, 0, st>>>(static_cast<const uint8_t*>(vol), shape[0], shape[1], strides[0], strides[1], strides[2], x_lo, x_hi, ro, per_slice, static_cast<uint8_t*>(out));
  else if (elem_size == 2) cyl_values_kernel<uint16_t><<<blocks, 256, 0, st>>>(static_cast<const uint16_t*>(vol), shape[0], shape[1], strides[0], strides[1], strides[2], x_lo, x_hi, ro, per_slice, static_cast<uint16_t*>(out));
  else cyl_values_kernel<uint32_t><<<blocks, 256, 0, st>>>(static_cast<const uint32_t*>(vol), shape[0], shape[1], strides[0], strides[1], strides[2], x_lo, x_hi, ro, per_slice, static_cast<uint32_t*>(out));
  TE_LAUNCH_CHECK();
  return TE_OK;
}

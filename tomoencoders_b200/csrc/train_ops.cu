// Training-step operators of the U-net segmenter:
//   SurfaceSegmenter.train / Keras Model.fit with Adam + BinaryCrossentropy, BatchNormalization in
//   training mode (/root/reference/tomo_encoders/neural_nets/surface_segmenter.py:43-66, 322-323;
//   Unet3D.py:44-194 for the layer semantics; SURVEY.md appendix A.3 for the Keras conventions).
// One C-ABI entry per operator (include/tomoenc.h, "training operators"); the graph is walked by the
// host code (tomoencoders_b200/neural_nets/_train.py).  Activations are NDHWC in fp32 (check mode)
// or bf16 / fp16; master weights, gradients and every reduction are fp32 / fp64.
//
// Tensor-core use: the forward 3x3x3 convolutions and their data gradients (a 3x3x3 convolution of
// dy with the flipped, transposed kernel) run on conv_umma_kernel (tcgen05) through te_tr_conv3 with
// device-side weight packing; the weight gradients use warp-level mma.sync tiles (wgrad_mma_kernel);
// normalisation / activation / pooling / loss operators are HBM-bound element-wise kernels.
#include "conv_umma.cuh"
#include "layers_simt.cuh"
#include "sm100_ptx.cuh"
#include "te_common.cuh"

#include <cstdlib>
#include <type_traits>

namespace te {
bool wgrad_umma_ok(const te_tensor& x, const te_tensor& dy);                                                    // wgrad_umma.cu
int wgrad_umma(const te_tensor& x, const te_tensor& dy, float* dw, int dw_cin, int ci_off, cudaStream_t st);
namespace {
// t / b with the remainder, through a 32-bit division whenever the dividend fits (it practically always does): the element
// and tile decodes below chain 3-5 of these per item, and a 64-bit division is ~5x the instructions of a 32-bit one
__device__ __forceinline__ long long divmod32(long long a, int b, int& rem) {
  if (static_cast<unsigned long long>(a) <= 0xffffffffull) {
    const unsigned q = static_cast<unsigned>(a) / static_cast<unsigned>(b);
    rem = static_cast<int>(static_cast<unsigned>(a) - q * static_cast<unsigned>(b));
    return q;
  }
  const long long q = a / b;
  rem = static_cast<int>(a - q * b);
  return q;
}


template <typename T> __device__ __forceinline__ float ldv(const T* p);
template <> __device__ __forceinline__ float ldv<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ldv<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <> __device__ __forceinline__ float ldv<__half>(const __half* p) { return __half2float(*p); }
template <typename T> __device__ __forceinline__ void stv(T* p, float v);
template <> __device__ __forceinline__ void stv<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void stv<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ void stv<__half>(__half* p, float v) { *p = __float2half_rn(v); }

inline int blocks_for(long long items, int threads, int waves = 16) {
  long long b = (items + threads - 1) / threads;
  const long long cap = static_cast<long long>(kNumSMs) * waves;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return static_cast<int>(b);
}

TView view_of(const te_tensor& t) {
  TView v;
  v.ptr = t.ptr; v.ctot = t.ctot; v.coff = t.coff; v.C = t.C; v.N = t.N; v.D = t.D; v.H = t.H; v.W = t.W;
  return v;
}
int act_of(int dtype) { return dtype == TE_F32 ? ACT_F32 : (dtype == TE_F16 ? ACT_F16 : ACT_BF16); }

#define TE_DISPATCH_T(dtype, ...)                                         \
  do {                                                                    \
    if ((dtype) == TE_F32) { using T = float; __VA_ARGS__; }              \
    else if ((dtype) == TE_F16) { using T = __half; __VA_ARGS__; }        \
    else if ((dtype) == TE_BF16) { using T = __nv_bfloat16; __VA_ARGS__; } \
    else { set_error("unsupported tensor dtype %d", (int)(dtype)); return TE_ERR_ARG; } \
  } while (0)

// ------------------------------------------------------------------------------------------------
// Per-channel reductions: every thread owns channel (tid % C) of a strided set of voxels (needs
// blockDim % C == 0 or C % blockDim... handled by the generic path: thread t walks elements
// e = t, t + stride, ... with stride a multiple of C so that its channel is fixed).
// ------------------------------------------------------------------------------------------------
template <typename T, typename F>
__device__ __forceinline__ void per_channel_reduce2(const TView& v, double* out, F f) {
  // out[c] += sum a, out[C + c] += sum b where (a, b) = f(voxel, channel)
  const int C = v.C;
  const long long total = v.voxels() * C;
  long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  stride -= stride % C;                       // host guarantees gridDim * blockDim >= C
  const long long start = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (start >= stride) return;
  const int c = static_cast<int>(start % C);
  float sa = 0.f, sb = 0.f;
  double da = 0.0, db = 0.0;
  int cnt = 0;
  for (long long e = start; e < total; e += stride) {
    float a, b;
    f(e / C, c, a, b);
    sa += a; sb += b;
    if (++cnt == 64) { da += sa; db += sb; sa = sb = 0.f; cnt = 0; }   // bounded fp32 partial sums
  }
  da += sa; db += sb;
  atomicAdd(out + c, da);
  atomicAdd(out + C + c, db);
}

template <typename T>
__global__ void __launch_bounds__(256) bn_stats_kernel(TView y, double* stats) {
  const T* p = static_cast<const T*>(y.ptr);
  per_channel_reduce2<T>(y, stats, [&](long long vox, int c, float& a, float& b) {
    const float v = ldv<T>(p + vox * y.ctot + y.coff + c);
    a = v; b = v * v;
  });
}

// a = lrelu(scale[c] * y + shift[c])
template <typename T>
__global__ void __launch_bounds__(256) bn_act_fwd_kernel(TView y, TView a, const float* __restrict__ scale,
                                                          const float* __restrict__ shift, float alpha) {
  const int C = y.C;
  const long long total = y.voxels() * C;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const T* yp = static_cast<const T*>(y.ptr);
  T* ap = static_cast<T*>(a.ptr);
  for (long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < total; e += stride) {
    const long long vox = e / C;
    const int c = static_cast<int>(e - vox * C);
    float z = fmaf(scale[c], ldv<T>(yp + vox * y.ctot + y.coff + c), shift[c]);
    z = z > 0.f ? z : alpha * z;
    stv<T>(ap + vox * a.ctot + a.coff + c, z);
  }
}

// red[c] += sum dz, red[C + c] += sum dz * xhat, with z = scale*y + shift, dz = da * lrelu'(z),
// xhat = (y - mean) * rstd
template <typename T>
__global__ void __launch_bounds__(256) bn_act_bwd_reduce_kernel(TView y, TView da, const float* __restrict__ scale,
                                                                 const float* __restrict__ shift,
                                                                 const float* __restrict__ mean,
                                                                 const float* __restrict__ rstd, float alpha, double* red) {
  const T* yp = static_cast<const T*>(y.ptr);
  const T* dp = static_cast<const T*>(da.ptr);
  per_channel_reduce2<T>(y, red, [&](long long vox, int c, float& a, float& b) {
    const float yv = ldv<T>(yp + vox * y.ctot + y.coff + c);
    const float z = fmaf(scale[c], yv, shift[c]);
    const float dz = ldv<T>(dp + vox * da.ctot + da.coff + c) * (z > 0.f ? 1.f : alpha);
    a = dz; b = dz * (yv - mean[c]) * rstd[c];
  });
}

// dy = gamma * rstd * (dz - dbeta / n - xhat * dgamma / n)
template <typename T>
__global__ void __launch_bounds__(256) bn_act_bwd_apply_kernel(TView y, TView da, TView dy, const float* __restrict__ scale,
                                                                const float* __restrict__ shift,
                                                                const float* __restrict__ mean,
                                                                const float* __restrict__ rstd,
                                                                const float* __restrict__ dbeta_n,
                                                                const float* __restrict__ dgamma_n, float alpha) {
  const int C = y.C;
  const long long total = y.voxels() * C;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const T* yp = static_cast<const T*>(y.ptr);
  const T* dp = static_cast<const T*>(da.ptr);
  T* op = static_cast<T*>(dy.ptr);
  for (long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < total; e += stride) {
    const long long vox = e / C;
    const int c = static_cast<int>(e - vox * C);
    const float yv = ldv<T>(yp + vox * y.ctot + y.coff + c);
    const float z = fmaf(scale[c], yv, shift[c]);
    const float dz = ldv<T>(dp + vox * da.ctot + da.coff + c) * (z > 0.f ? 1.f : alpha);
    const float xhat = (yv - mean[c]) * rstd[c];
    stv<T>(op + vox * dy.ctot + dy.coff + c, scale[c] * (dz - dbeta_n[c] - xhat * dgamma_n[c]));
  }
}

// ------------------------------------------------------------------------------------------------
// 8-channel vectorised versions (16-bit tensors, C a power-of-two multiple of 8 up to 256, 16-byte aligned
// channel slices): one 128-bit access per tensor and voxel; per-channel sums are reduced across the warp
// (lanes with the same channel group), across the block in shared memory, then 2C atomics per block.
// ------------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ void load8v(const T* p, float (&f)[8]) {
  const uint4 r = *reinterpret_cast<const uint4*>(p);
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if constexpr (std::is_same<T, __half>::value) {
      const float2 v = __half22float2(*reinterpret_cast<const __half2*>(&w[j]));
      f[2 * j] = v.x; f[2 * j + 1] = v.y;
    } else {
      const float2 v = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[j]));
      f[2 * j] = v.x; f[2 * j + 1] = v.y;
    }
  }
}
// r[j] = f[j] rounded to T (what store8v writes)
template <typename T>
__device__ __forceinline__ void load8r(const float (&f)[8], float (&r)[8]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    if constexpr (std::is_same<T, __half>::value) r[j] = __half2float(__float2half_rn(f[j]));
    else r[j] = __bfloat162float(__float2bfloat16_rn(f[j]));
  }
}
template <typename T>
__device__ __forceinline__ void store8v(T* p, const float (&f)[8]) {
  uint32_t w[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if constexpr (std::is_same<T, __half>::value) { __half2 h = __floats2half2_rn(f[2 * j], f[2 * j + 1]); w[j] = *reinterpret_cast<uint32_t*>(&h); }
    else { __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]); w[j] = *reinterpret_cast<uint32_t*>(&h); }
  }
  *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
}

// out[c] += sum a, out[C + c] += sum b with (a[8], b[8]) = f(voxel, first channel of the group); `items` work items,
// item -> (voxel index passed to f) is the identity: the caller decodes it.
template <typename F>
__device__ __forceinline__ void reduce2_v8(int C, long long items, double* out, F f) {
  __shared__ float red[8][2][256];
  const int G = C >> 3;
  const long long nthreads = static_cast<long long>(gridDim.x) * blockDim.x;
  const long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int g = static_cast<int>(t % G);
  const long long istride = nthreads / G;
  float sa[8], sb[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) sa[j] = sb[j] = 0.f;
  long long it = t / G;
  for (; it + 3 * istride < items; it += 4 * istride) {          // four independent items: their loads are issued together
    float a[4][8], b[4][8];
#pragma unroll
    for (int u = 0; u < 4; ++u) f(it + u * istride, g * 8, a[u], b[u]);
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int j = 0; j < 8; ++j) { sa[j] += a[u][j]; sb[j] += b[u][j]; }
  }
  for (; it < items; it += istride) {
    float a[8], b[8];
    f(it, g * 8, a, b);
#pragma unroll
    for (int j = 0; j < 8; ++j) { sa[j] += a[j]; sb[j] += b[j]; }
  }
  for (int off = 16; off >= G; off >>= 1) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      sa[j] += __shfl_xor_sync(0xffffffffu, sa[j], off);
      sb[j] += __shfl_xor_sync(0xffffffffu, sb[j], off);
    }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane < G) {
#pragma unroll
    for (int j = 0; j < 8; ++j) { red[warp][0][lane * 8 + j] = sa[j]; red[warp][1][lane * 8 + j] = sb[j]; }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
    const int which = i / C, c = i - which * C;
    float tot = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) tot += red[w][which][c];
    atomicAdd(out + which * C + c, static_cast<double>(tot));
  }
}

bool v8_ok(const te_tensor& t) {
  const int G = t.C >> 3;
  return t.dtype != TE_F32 && t.C % 8 == 0 && t.C <= 256 && G >= 1 && (G & (G - 1)) == 0 && G <= 32 && t.ctot % 8 == 0 && t.coff % 8 == 0;
}

template <typename T>
__global__ void __launch_bounds__(256) bn_stats_v8_kernel(TView y, double* stats) {
  const T* p = static_cast<const T*>(y.ptr);
  reduce2_v8(y.C, y.voxels(), stats, [&](long long vox, int c0, float (&a)[8], float (&b)[8]) {
    load8v<T>(p + vox * y.ctot + y.coff + c0, a);
#pragma unroll
    for (int j = 0; j < 8; ++j) b[j] = a[j] * a[j];
  });
}

template <typename T>
__global__ void __launch_bounds__(256) bn_act_fwd_v8_kernel(TView y, TView a, const float* __restrict__ scale,
                                                             const float* __restrict__ shift, float alpha) {
  const int G = y.C >> 3;
  const long long total = y.voxels() * G;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const T* yp = static_cast<const T*>(y.ptr);
  T* ap = static_cast<T*>(a.ptr);
  for (long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < total; e += stride) {
    const long long vox = e / G;
    const int c0 = static_cast<int>(e - vox * G) * 8;
    float f[8];
    load8v<T>(yp + vox * y.ctot + y.coff + c0, f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float z = fmaf(scale[c0 + j], f[j], shift[c0 + j]);
      f[j] = z > 0.f ? z : alpha * z;
    }
    store8v<T>(ap + vox * a.ctot + a.coff + c0, f);
  }
}

// The channel group of a thread is fixed over its whole loop (the thread stride is a multiple of the group count), so the
// per-channel constants are hoisted into registers once.
template <typename T>
__global__ void __launch_bounds__(256) bn_act_bwd_reduce_v8_kernel(TView y, TView da, const float* __restrict__ scale,
                                                                    const float* __restrict__ shift, const float* __restrict__ mean,
                                                                    const float* __restrict__ rstd, float alpha, double* red) {
  const T* yp = static_cast<const T*>(y.ptr);
  const T* dp = static_cast<const T*>(da.ptr);
  const int G = y.C >> 3;
  const int cg = static_cast<int>((static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) % G) * 8;
  float sc[8], sh[8], mu[8], rs[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { sc[j] = scale[cg + j]; sh[j] = shift[cg + j]; mu[j] = mean[cg + j]; rs[j] = rstd[cg + j]; }
  reduce2_v8(y.C, y.voxels(), red, [&](long long vox, int c0, float (&a)[8], float (&b)[8]) {
    float yv[8], d[8];
    load8v<T>(yp + vox * y.ctot + y.coff + c0, yv);
    load8v<T>(dp + vox * da.ctot + da.coff + c0, d);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float z = fmaf(sc[j], yv[j], sh[j]);
      const float dz = d[j] * (z > 0.f ? 1.f : alpha);
      a[j] = dz;
      b[j] = dz * (yv[j] - mu[j]) * rs[j];
    }
  });
}

// dy = scale * (dz - dbeta_n - xhat * dgamma_n), possibly in place (dy == da); dsum2c (optional): [0, C) += sum_v dy, the
// gradient of the bias of the convolution feeding this BatchNormalization (analytically zero; Keras computes it, so it is
// produced here from the values already in registers instead of by another pass over dy).
template <typename T>
__global__ void __launch_bounds__(256) bn_act_bwd_apply_v8_kernel(TView y, TView da, TView dy, const float* __restrict__ scale,
                                                                   const float* __restrict__ shift, const float* __restrict__ mean,
                                                                   const float* __restrict__ rstd, const float* __restrict__ dbeta_n,
                                                                   const float* __restrict__ dgamma_n, float alpha, double* dsum2c) {
  const T* yp = static_cast<const T*>(y.ptr);
  const T* dp = static_cast<const T*>(da.ptr);
  T* op = static_cast<T*>(dy.ptr);
  const int G = y.C >> 3;
  const int cg = static_cast<int>((static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) % G) * 8;
  float sc[8], sh[8], mu[8], rs[8], db[8], dg[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    sc[j] = scale[cg + j]; sh[j] = shift[cg + j]; mu[j] = mean[cg + j]; rs[j] = rstd[cg + j];
    db[j] = dbeta_n[cg + j]; dg[j] = dgamma_n[cg + j];
  }
  // reduce2_v8 hands four independent items per iteration to the functor calls; each call loads, computes and stores its own
  // voxel row (a thread never touches another thread's rows, so the in-place update is safe)
  auto body = [&](long long vox, int c0, float (&a)[8], float (&b)[8]) {
    float yv[8], d[8];
    load8v<T>(yp + vox * y.ctot + y.coff + c0, yv);
    load8v<T>(dp + vox * da.ctot + da.coff + c0, d);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float z = fmaf(sc[j], yv[j], sh[j]);
      const float dz = d[j] * (z > 0.f ? 1.f : alpha);
      const float xhat = (yv[j] - mu[j]) * rs[j];
      d[j] = sc[j] * (dz - db[j] - xhat * dg[j]);
    }
    store8v<T>(op + vox * dy.ctot + dy.coff + c0, d);
    if (dsum2c) {
      float r[8];
      load8r<T>(d, r);                      // the values as stored (rounded to the tensor dtype)
#pragma unroll
      for (int j = 0; j < 8; ++j) { a[j] = r[j]; b[j] = 0.f; }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) { a[j] = 0.f; b[j] = 0.f; }
    }
  };
  if (dsum2c) {
    reduce2_v8(y.C, y.voxels(), dsum2c, body);
  } else {
    const long long nthreads = static_cast<long long>(gridDim.x) * blockDim.x;
    const long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const long long items = y.voxels(), istride = nthreads / G;
    float a[8], b[8];
    for (long long it = t / G; it < items; it += istride) body(it, cg, a, b);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) channel_sum_v8_kernel(TView dy, double* out2) {
  const T* p = static_cast<const T*>(dy.ptr);
  reduce2_v8(dy.C, dy.voxels(), out2, [&](long long vox, int c0, float (&a)[8], float (&b)[8]) {
    load8v<T>(p + vox * dy.ctot + dy.coff + c0, a);
#pragma unroll
    for (int j = 0; j < 8; ++j) b[j] = 0.f;
  });
}

template <typename T>
__global__ void __launch_bounds__(256) head_wgrad_v8_kernel(TView a, const float* __restrict__ dl, double* dwb2c) {
  const T* p = static_cast<const T*>(a.ptr);
  reduce2_v8(a.C, a.voxels(), dwb2c, [&](long long vox, int c0, float (&s0)[8], float (&s1)[8]) {
    const float d = dl[vox];
    load8v<T>(p + vox * a.ctot + a.coff + c0, s0);
#pragma unroll
    for (int j = 0; j < 8; ++j) { s0[j] *= d; s1[j] = d; }
  });
}

// space-to-depth transposed-conv backward, vectorised: item = (low-res voxel, sub-position)
template <typename T>
__global__ void __launch_bounds__(256) lrelu_bwd_s2d_v8_kernel(TView a, TView da, TView dz2, float alpha, double* dbias2) {
  const int C = a.C, Dl = a.D / 2, Hl = a.H / 2, Wl = a.W / 2;
  const T* ap = static_cast<const T*>(a.ptr);
  const T* dp = static_cast<const T*>(da.ptr);
  T* op = static_cast<T*>(dz2.ptr);
  reduce2_v8(C, static_cast<long long>(a.N) * Dl * Hl * Wl * 8, dbias2, [&](long long item, int c0, float (&s0)[8], float (&s1)[8]) {
    long long t = item;
    const int sub = static_cast<int>(t & 7); t >>= 3;
    const long long lv = t;
    int x_r; t = divmod32(t, Wl, x_r); const int x = x_r;
    int y_r; t = divmod32(t, Hl, y_r); const int y = y_r;
    const int z = static_cast<int>(t % Dl);
    const long long n = t / Dl;
    const long long hv = ((n * a.D + 2 * z + (sub >> 2)) * a.H + 2 * y + ((sub >> 1) & 1)) * a.W + 2 * x + (sub & 1);
    float av[8];
    load8v<T>(ap + hv * a.ctot + a.coff + c0, av);
    load8v<T>(dp + hv * da.ctot + da.coff + c0, s0);
#pragma unroll
    for (int j = 0; j < 8; ++j) { s0[j] *= (av[j] > 0.f ? 1.f : alpha); s1[j] = 0.f; }
    store8v<T>(op + lv * dz2.ctot + dz2.coff + sub * C + c0, s0);
  });
}

// MaxPool3D(p) backward: the gradient of a pooled voxel goes to the FIRST window element (z, y, x
// scan order) equal to the maximum (torch / TF argmax convention); every input voxel is written.
// MaxPool3D(2) backward, 8 channels per thread: the eight window positions are loaded as 128-bit rows (all
// in flight), first-argmax per channel (Keras / TF tie rule of the scalar kernel below), eight 128-bit stores.
template <typename T>
__global__ void __launch_bounds__(256) maxpool2_bwd_v8_kernel(TView x, TView dpool, TView dx, int accumulate) {
  const int G = x.C >> 3;
  const int Do = x.D / 2, Ho = x.H / 2, Wo = x.W / 2;
  const long long total = static_cast<long long>(x.N) * Do * Ho * Wo * G;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const T* xp = static_cast<const T*>(x.ptr);
  const T* gp = static_cast<const T*>(dpool.ptr);
  T* op = static_cast<T*>(dx.ptr);
  for (long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < total; e += stride) {
    const int c0 = static_cast<int>(e % G) * 8;
    long long t = e / G;
    int ox_r; t = divmod32(t, Wo, ox_r); const int ox = ox_r;
    int oy_r; t = divmod32(t, Ho, oy_r); const int oy = oy_r;
    const int oz = static_cast<int>(t % Do);
    const long long n = t / Do;
    float v[8][8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const long long vox = ((n * x.D + oz * 2 + (k >> 2)) * x.H + oy * 2 + ((k >> 1) & 1)) * x.W + ox * 2 + (k & 1);
      load8v<T>(xp + vox * x.ctot + x.coff + c0, v[k]);
    }
    float g[8];
    load8v<T>(gp + (((n * Do + oz) * Ho + oy) * Wo + ox) * dpool.ctot + dpool.coff + c0, g);
    int arg[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float best = -INFINITY;
      arg[j] = 0;
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (v[k][j] > best) { best = v[k][j]; arg[j] = k; }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const long long vox = ((n * x.D + oz * 2 + (k >> 2)) * x.H + oy * 2 + ((k >> 1) & 1)) * x.W + ox * 2 + (k & 1);
      T* dst = op + vox * dx.ctot + dx.coff + c0;
      float o[8];
      if (accumulate) load8v<T>(dst, o);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float add = arg[j] == k ? g[j] : 0.f;
        o[j] = accumulate ? o[j] + add : add;
      }
      store8v<T>(dst, o);
    }
  }
}

// accumulate = 1 adds to dx (the skip connection already deposited its gradient there).
template <typename T>
__global__ void __launch_bounds__(256) maxpool_bwd_kernel(TView x, TView dpool, TView dx, int p, int accumulate) {
  const int C = x.C;
  const int Do = x.D / p, Ho = x.H / p, Wo = x.W / p;
  const long long total = static_cast<long long>(x.N) * Do * Ho * Wo * C;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const T* xp = static_cast<const T*>(x.ptr);
  const T* gp = static_cast<const T*>(dpool.ptr);
  T* op = static_cast<T*>(dx.ptr);
  for (long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < total; e += stride) {
    const int c = static_cast<int>(e % C);
    long long t = e / C;
    int ox_r; t = divmod32(t, Wo, ox_r); const int ox = ox_r;
    int oy_r; t = divmod32(t, Ho, oy_r); const int oy = oy_r;
    const int oz = static_cast<int>(t % Do);
    const long long n = t / Do;
    float best = -INFINITY;
    int arg = 0;
    for (int k = 0; k < p * p * p; ++k) {
      const int dz = k / (p * p), dy = (k / p) % p, dxx = k % p;
      const long long vox = ((n * x.D + oz * p + dz) * x.H + oy * p + dy) * x.W + ox * p + dxx;
      const float v = ldv<T>(xp + vox * x.ctot + x.coff + c);
      if (v > best) { best = v; arg = k; }
    }
    const float g = ldv<T>(gp + (((n * Do + oz) * Ho + oy) * Wo + ox) * dpool.ctot + dpool.coff + c);
    for (int k = 0; k < p * p * p; ++k) {
      const int dz = k / (p * p), dy = (k / p) % p, dxx = k % p;
      const long long vox = ((n * x.D + oz * p + dz) * x.H + oy * p + dy) * x.W + ox * p + dxx;
      T* dst = op + vox * dx.ctot + dx.coff + c;
      const float add = k == arg ? g : 0.f;
      stv<T>(dst, accumulate ? ldv<T>(dst) + add : add);
    }
  }
}

// Transposed conv backward on the tensor cores: dz2[v][sub * C + c] = da[2v + sub][c] * lrelu'(a[2v + sub][c])
// ("space to depth": the 8 sub-positions of a low-resolution voxel become channel groups), so that
//   dx = pointwise_conv(dz2, w viewed as [8 Cout][Cin])   and   dw = pointwise_wgrad(dz2, x);
// dbias2[c] += sum dz.  Thread = (low-res voxel, sub, 8-channel group).
template <typename T>
__global__ void __launch_bounds__(256) lrelu_bwd_s2d_kernel(TView a, TView da, TView dz2, float alpha, double* dbias2) {
  const int C = a.C, Dl = a.D / 2, Hl = a.H / 2, Wl = a.W / 2;
  const long long total = static_cast<long long>(a.N) * Dl * Hl * Wl * 8 * C;
  long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  stride -= stride % C;
  const long long start = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (start >= stride) return;
  const int c = static_cast<int>(start % C);
  const T* ap = static_cast<const T*>(a.ptr);
  const T* dp = static_cast<const T*>(da.ptr);
  T* op = static_cast<T*>(dz2.ptr);
  float s = 0.f;
  double ds = 0.0;
  int cnt = 0;
  for (long long e = start; e < total; e += stride) {
    long long t = e / C;
    const int sub = static_cast<int>(t & 7); t >>= 3;
    int x_r; t = divmod32(t, Wl, x_r); const int x = x_r;
    int y_r; t = divmod32(t, Hl, y_r); const int y = y_r;
    const int z = static_cast<int>(t % Dl);
    const long long n = t / Dl;
    const long long hv = ((n * a.D + 2 * z + (sub >> 2)) * a.H + 2 * y + ((sub >> 1) & 1)) * a.W + 2 * x + (sub & 1);
    const float av = ldv<T>(ap + hv * a.ctot + a.coff + c);
    const float dz = ldv<T>(dp + hv * da.ctot + da.coff + c) * (av > 0.f ? 1.f : alpha);
    stv<T>(op + (e / (8LL * C)) * dz2.ctot + dz2.coff + sub * C + c, dz);
    s += dz;
    if (++cnt == 64) { ds += s; s = 0.f; cnt = 0; }
  }
  ds += s;
  atomicAdd(dbias2 + c, ds);
}

// Transposed conv (k = 2, stride s) backward, part 1: dz = da * lrelu'(a) in place (a and z have the
// same sign) and dbias[c] += sum dz.
template <typename T>
__global__ void __launch_bounds__(256) lrelu_bwd_bias_kernel(TView a, TView da, float alpha, double* dbias2) {
  const T* ap = static_cast<const T*>(a.ptr);
  T* dp = static_cast<T*>(da.ptr);
  per_channel_reduce2<T>(a, dbias2, [&](long long vox, int c, float& s0, float& s1) {
    const float av = ldv<T>(ap + vox * a.ctot + a.coff + c);
    T* d = dp + vox * da.ctot + da.coff + c;
    const float dz = ldv<T>(d) * (av > 0.f ? 1.f : alpha);
    stv<T>(d, dz);
    s0 = dz; s1 = 0.f;
  });
}

// part 2: dx[v][ci] = sum_sub sum_co dz[s*v + sub][co] * w[sub][co][ci]   (w Keras layout (2,2,2,Cout,Cin))
template <typename T>
__global__ void __launch_bounds__(256) upconv_dgrad_kernel(TView dz, TView dx, const float* __restrict__ w, int s) {
  const int Ci = dx.C, Co = dz.C;
  const long long total = dx.voxels() * Ci;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const T* gp = static_cast<const T*>(dz.ptr);
  T* op = static_cast<T*>(dx.ptr);
  for (long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < total; e += stride) {
    const int ci = static_cast<int>(e % Ci);
    long long t = e / Ci;
    int x_r; t = divmod32(t, dx.W, x_r); const int x = x_r;
    int y_r; t = divmod32(t, dx.H, y_r); const int y = y_r;
    const int z = static_cast<int>(t % dx.D);
    const long long n = t / dx.D;
    float acc = 0.f;
    for (int sub = 0; sub < 8; ++sub) {
      const int a = sub >> 2, b = (sub >> 1) & 1, c = sub & 1;
      const long long vox = ((n * dz.D + z * s + a) * dz.H + y * s + b) * dz.W + x * s + c;
      const T* g = gp + vox * dz.ctot + dz.coff;
      const float* ws = w + static_cast<long long>(sub) * Co * Ci + ci;
      for (int co = 0; co < Co; ++co) acc = fmaf(ldv<T>(g + co), ws[static_cast<long long>(co) * Ci], acc);
    }
    stv<T>(op + (e / Ci) * dx.ctot + dx.coff + ci, acc);
  }
}

// part 3: dw[sub][co][ci] += sum_v dz[s*v + sub][co] * x[v][ci].  One block per (sub, co-tile of 8),
// threads over ci x voxel slices; block-level reduction in shared memory, one atomicAdd per element.
template <typename T>
__global__ void __launch_bounds__(256) upconv_wgrad_kernel(TView x, TView dz, float* dw, int s, int vox_splits) {
  const int Ci = x.C, Co = dz.C;
  const int sub = blockIdx.y, split = blockIdx.z;
  const int a = sub >> 2, b = (sub >> 1) & 1, c = sub & 1;
  const T* xp = static_cast<const T*>(x.ptr);
  const T* gp = static_cast<const T*>(dz.ptr);
  const long long nvox = x.voxels();
  const long long per = (nvox + vox_splits - 1) / vox_splits;
  const long long v0 = split * per, v1 = v0 + per < nvox ? v0 + per : nvox;
  // element (co, ci) handled by thread: idx = blockIdx.x * 256 + tid over Co * Ci
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<long long>(Co) * Ci) return;
  const int co = static_cast<int>(idx / Ci), ci = static_cast<int>(idx % Ci);
  float acc = 0.f;
  for (long long v = v0; v < v1; ++v) {
    long long t = v;
    int xx_r; t = divmod32(t, x.W, xx_r); const int xx = xx_r;
    int yy_r; t = divmod32(t, x.H, yy_r); const int yy = yy_r;
    const int zz = static_cast<int>(t % x.D);
    const long long n = t / x.D;
    const long long hv = ((n * dz.D + zz * s + a) * dz.H + yy * s + b) * dz.W + xx * s + c;
    acc = fmaf(ldv<T>(gp + hv * dz.ctot + dz.coff + co), ldv<T>(xp + v * x.ctot + x.coff + ci), acc);
  }
  atomicAdd(dw + (static_cast<long long>(sub) * Co + co) * Ci + ci, acc);
}

// 3x3x3 conv weight gradient, CUDA-core version (fp32 check mode and shapes the tensor-core kernel
// does not take): dw[tap][ci][co] += sum_v x[v + tap][ci] * dy[v][co].  Block = (tap, voxel split);
// thread = one (ci, co) element group.
template <typename T>
__global__ void __launch_bounds__(256) conv3_wgrad_simt_kernel(TView x, TView dy, float* dw, int dw_cin, int ci_off,
                                                                int vox_splits) {
  const int Ci = x.C, Co = dy.C;
  const int tap = blockIdx.y, split = blockIdx.z;
  const int dz = tap / 9 - 1, dyy = (tap / 3) % 3 - 1, dxx = tap % 3 - 1;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<long long>(Ci) * Co) return;
  const int ci = static_cast<int>(idx / Co), co = static_cast<int>(idx % Co);
  const T* xp = static_cast<const T*>(x.ptr);
  const T* gp = static_cast<const T*>(dy.ptr);
  const long long nvox = dy.voxels();
  const long long per = (nvox + vox_splits - 1) / vox_splits;
  const long long v0 = split * per, v1 = v0 + per < nvox ? v0 + per : nvox;
  float acc = 0.f;
  for (long long v = v0; v < v1; ++v) {
    long long t = v;
    int xx_r; t = divmod32(t, dy.W, xx_r); const int xx = xx_r + dxx;
    int yy_r; t = divmod32(t, dy.H, yy_r); const int yy = yy_r + dyy;
    const int zz = static_cast<int>(t % dy.D) + dz;
    const long long n = t / dy.D;
    if (xx < 0 || xx >= x.W || yy < 0 || yy >= x.H || zz < 0 || zz >= x.D) continue;
    const long long xv = ((n * x.D + zz) * x.H + yy) * x.W + xx;
    acc = fmaf(ldv<T>(xp + xv * x.ctot + x.coff + ci), ldv<T>(gp + v * dy.ctot + dy.coff + co), acc);
  }
  atomicAdd(dw + (static_cast<long long>(tap) * dw_cin + ci_off + ci) * Co + co, acc);
}

// First layer (Cin = 1): dw[tap][co] += sum_v x[v + tap] dy[v][co].  Thread = one (tap, co) pair; a block
// stages a (4, 16, 8) voxel tile of dy and the haloed x tile in shared memory as fp32 and walks it.
template <typename T>
__global__ void __launch_bounds__(1024) conv3_wgrad_c1_kernel(TView x, TView dy, float* dw, int dw_cin, int ci_off) {
  constexpr int TZ = 4;
  extern __shared__ float sh[];
  float* xs = sh;                         // [TZ + 2][18][10]
  float* ds = sh + (TZ + 2) * 180;        // [TZ * 128][Cout]
  const int Co = dy.C;
  const int tid = threadIdx.x;
  const bool active = tid < 27 * Co;
  const int tap = active ? tid / Co : 0, co = active ? tid - tap * Co : 0;
  const int dz = tap / 9, dyy = (tap / 3) % 3, dxx = tap % 3;            // +1 biased
  const int tiles_x = (dy.W + 7) / 8, tiles_y = (dy.H + 15) / 16, tiles_z = (dy.D + TZ - 1) / TZ;
  const long long total = static_cast<long long>(dy.N) * tiles_z * tiles_y * tiles_x;
  const T* xp = static_cast<const T*>(x.ptr);
  const T* gp = static_cast<const T*>(dy.ptr);
  float acc = 0.f;
  for (long long tile = blockIdx.x; tile < total; tile += gridDim.x) {
    long long t = tile;
    int x0_r; t = divmod32(t, tiles_x, x0_r); const int x0 = x0_r * 8;
    int y0_r; t = divmod32(t, tiles_y, y0_r); const int y0 = y0_r * 16;
    int z0_r; t = divmod32(t, tiles_z, z0_r); const int z0 = z0_r * TZ;
    const long long n = t;
    __syncthreads();
    for (int i = tid; i < (TZ + 2) * 180; i += blockDim.x) {
      const int bx = i % 10, by = (i / 10) % 18, bz = i / 180;
      const int xx = x0 + bx - 1, yy = y0 + by - 1, zz = z0 + bz - 1;
      const bool ok = xx >= 0 && xx < x.W && yy >= 0 && yy < x.H && zz >= 0 && zz < x.D;
      xs[i] = ok ? ldv<T>(xp + (((n * x.D + zz) * x.H + yy) * x.W + xx) * x.ctot + x.coff) : 0.f;
    }
    for (int i = tid; i < TZ * 128 * Co; i += blockDim.x) {
      const int c = i % Co, v = i / Co;
      const int bx = v & 7, by = (v >> 3) & 15, bz = v >> 7;
      const int xx = x0 + bx, yy = y0 + by, zz = z0 + bz;
      const bool ok = xx < dy.W && yy < dy.H && zz < dy.D;
      ds[i] = ok ? ldv<T>(gp + (((n * dy.D + zz) * dy.H + yy) * dy.W + xx) * dy.ctot + dy.coff + c) : 0.f;
    }
    __syncthreads();
    if (active) {
      for (int pz = 0; pz < TZ; ++pz)
        for (int py = 0; py < 16; ++py) {
          const float* xr = xs + ((pz + dz) * 18 + py + dyy) * 10 + dxx;
          const float* dr = ds + ((pz * 16 + py) * 8) * Co + co;
#pragma unroll
          for (int px = 0; px < 8; ++px) acc = fmaf(xr[px], dr[px * Co], acc);
        }
    }
  }
  if (active) atomicAdd(dw + (static_cast<long long>(tap) * dw_cin + ci_off) * Co + co, acc);
}

// First layer, register-tiled (Cout = 8 or 16, 16-bit tensors).  Warp = (ty tap row, group of 4 output channels):
// 9 taps (tz, tx) x 4 channels = 36 accumulators per lane; the 32 lanes split the voxels of a (8, 16, 8) tile as
// (4 y rows, 8 x) and walk z with a sliding 3 x 3 window of x values, so one voxel costs 3 scalar + 1 128-bit
// shared loads for 36 FMAs (the kernel above: 16 loads per 8 FMAs).  Row pitch 40 floats = 8 banks per y row and a
// voxel-major float4 layout of dy keep every load conflict-free.  Lanes are reduced with shuffles at the very end.
template <typename T>
__global__ void __launch_bounds__(384, 2) conv3_wgrad_c1_tiled_kernel(TView x, TView dy, float* dw, int dw_cin, int ci_off) {
  constexpr int TZ = 8, XP = 40, NVOX = TZ * 128;
  extern __shared__ __align__(16) float sh[];
  float* xs = sh;                                                  // [TZ + 2][18][XP]
  float4* ds = reinterpret_cast<float4*>(sh + (TZ + 2) * 18 * XP);  // [Co / 4][NVOX]
  const int Co = dy.C, CG = Co >> 2, H8 = Co >> 3;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nwarps_active = 3 * CG;
  const int tg = warp % 3, cg = warp / 3;
  const int px = lane & 7, pyq = lane >> 3;
  const int tiles_x = (dy.W + 7) / 8, tiles_y = (dy.H + 15) / 16, tiles_z = (dy.D + TZ - 1) / TZ;
  const long long total = static_cast<long long>(dy.N) * tiles_z * tiles_y * tiles_x;
  const T* xp = static_cast<const T*>(x.ptr);
  const T* gp = static_cast<const T*>(dy.ptr);
  float acc[3][3][4];
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = 0; b < 3; ++b) acc[a][b][0] = acc[a][b][1] = acc[a][b][2] = acc[a][b][3] = 0.f;
  for (long long tile = blockIdx.x; tile < total; tile += gridDim.x) {
    long long t = tile;
    int x0_r; t = divmod32(t, tiles_x, x0_r); const int x0 = x0_r * 8;
    int y0_r; t = divmod32(t, tiles_y, y0_r); const int y0 = y0_r * 16;
    int z0_r; t = divmod32(t, tiles_z, z0_r); const int z0 = z0_r * TZ;
    const long long n = t;
    __syncthreads();
    for (int i = tid; i < (TZ + 2) * 180; i += blockDim.x) {
      const int bx = i % 10, by = (i / 10) % 18, bz = i / 180;
      const int xx = x0 + bx - 1, yy = y0 + by - 1, zz = z0 + bz - 1;
      const bool ok = xx >= 0 && xx < x.W && yy >= 0 && yy < x.H && zz >= 0 && zz < x.D;
      xs[(bz * 18 + by) * XP + bx] = ok ? ldv<T>(xp + (((n * x.D + zz) * x.H + yy) * x.W + xx) * x.ctot + x.coff) : 0.f;
    }
    for (int i = tid; i < NVOX * H8; i += blockDim.x) {
      const int h = i % H8, v = i / H8;
      const int bx = v & 7, by = (v >> 3) & 15, bz = v >> 7;
      const int xx = x0 + bx, yy = y0 + by, zz = z0 + bz;
      float f[8];
      if (xx < dy.W && yy < dy.H && zz < dy.D) load8v<T>(gp + (((n * dy.D + zz) * dy.H + yy) * dy.W + xx) * dy.ctot + dy.coff + h * 8, f);
      else {
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = 0.f;
      }
      ds[(2 * h) * NVOX + v] = make_float4(f[0], f[1], f[2], f[3]);
      ds[(2 * h + 1) * NVOX + v] = make_float4(f[4], f[5], f[6], f[7]);
    }
    __syncthreads();
    if (warp < nwarps_active) {
#pragma unroll 1
      for (int pyb = 0; pyb < 4; ++pyb) {
        const int py = pyb * 4 + pyq;
        const float* xr = xs + (py + tg) * XP + px;
        float w0[3], w1[3], w2[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) { w0[d] = xr[d]; w1[d] = xr[18 * XP + d]; }
        const float4* dr = ds + cg * NVOX + py * 8 + px;
#pragma unroll
        for (int pz = 0; pz < TZ; ++pz) {
#pragma unroll
          for (int d = 0; d < 3; ++d) w2[d] = xr[(pz + 2) * 18 * XP + d];
          const float4 g = dr[pz * 128];
#pragma unroll
          for (int d = 0; d < 3; ++d) {
            acc[0][d][0] = fmaf(w0[d], g.x, acc[0][d][0]); acc[0][d][1] = fmaf(w0[d], g.y, acc[0][d][1]);
            acc[0][d][2] = fmaf(w0[d], g.z, acc[0][d][2]); acc[0][d][3] = fmaf(w0[d], g.w, acc[0][d][3]);
            acc[1][d][0] = fmaf(w1[d], g.x, acc[1][d][0]); acc[1][d][1] = fmaf(w1[d], g.y, acc[1][d][1]);
            acc[1][d][2] = fmaf(w1[d], g.z, acc[1][d][2]); acc[1][d][3] = fmaf(w1[d], g.w, acc[1][d][3]);
            acc[2][d][0] = fmaf(w2[d], g.x, acc[2][d][0]); acc[2][d][1] = fmaf(w2[d], g.y, acc[2][d][1]);
            acc[2][d][2] = fmaf(w2[d], g.z, acc[2][d][2]); acc[2][d][3] = fmaf(w2[d], g.w, acc[2][d][3]);
          }
#pragma unroll
          for (int d = 0; d < 3; ++d) { w0[d] = w1[d]; w1[d] = w2[d]; }
        }
      }
    }
  }
  if (warp < nwarps_active) {
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float v = acc[a][b][c];
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
          if (lane == 0) atomicAdd(dw + (static_cast<long long>((a * 3 + tg) * 3 + b) * dw_cin + ci_off) * Co + cg * 4 + c, v);
        }
  }
}

// dbias[c] += sum_v dy[v][c]
template <typename T>
__global__ void __launch_bounds__(256) channel_sum_kernel(TView dy, double* out2) {
  const T* p = static_cast<const T*>(dy.ptr);
  per_channel_reduce2<T>(dy, out2, [&](long long vox, int c, float& a, float& b) {
    a = ldv<T>(p + vox * dy.ctot + dy.coff + c); b = 0.f;
  });
}

// ------------------------------------------------------------------------------------------------
// 3x3x3 conv weight gradient on the tensor cores (warp-level mma.sync m16n8k16, fp32 accumulate).
// CTA = (32 input channels) x (32 output channels) x all 27 taps, persistent over voxel tiles of
// TZ z-planes x 16 x 8; the 27 x 32 x 32 fp32 accumulators live in the registers of 8 warps
// (warp w owns the (16 ci) x (8 co) mma tile (w & 1, w >> 1) of every tap: 108 registers per lane)
// and are flushed with atomicAdd once at the end.  Both operands have the voxel index as the GEMM K
// dimension, i.e. they are "transposed" in shared memory ([voxel][channel]) and are fetched with
// ldmatrix.trans.  Voxel rows are padded to 80 bytes (32 ch x 2 B + 16) to spread the 8 row
// addresses of an ldmatrix over distinct banks.
// ------------------------------------------------------------------------------------------------
constexpr int kWgTZ = 4;
constexpr int kWgThreads = 288;                                      // 9 warps: one (dz, dy) tap row each

__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t saddr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(saddr));
}
template <bool FP16>
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  if constexpr (FP16)
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  else
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

struct WgParams {
  float* dw; int dw_cin, ci_off, Cout;        // dw[ntaps][dw_cin][Cout]
  int Ci, x_coff, dy_coff;
  int N, D, H, W;
  int tiles_x, tiles_y, tiles_z;
  long long total_tiles;
};

// Shared-memory row (voxel) of 32 channels = 64 bytes, 64-byte TMA swizzle: 16-byte chunk c of voxel
// row v sits at v * 64 + ((c ^ ((v >> 1) & 3)) << 4) relative to a 512-byte aligned base.  The 8 row
// addresses of an ldmatrix (8 consecutive voxels, fixed c) then fall into 8 distinct 16-byte bank
// groups, for any starting voxel.
__device__ __forceinline__ uint32_t sw64(uint32_t base, int vox, int chunk) {
  return base + static_cast<uint32_t>(vox) * 64u + (static_cast<uint32_t>(chunk ^ ((vox >> 1) & 3)) << 4);
}

// NTAPS = 27: 3x3x3 conv (haloed x box); warp w owns the tap row (dz, dy) = (w / 3, w % 3), i.e. the
// three dx taps, for the whole 32 (ci) x 32 (co) block: 3 x 2 x 4 mma tiles = 96 accumulator registers.
// NTAPS = 1: pointwise product (transposed-conv backward); the 9 warps split the k-steps.
// Operands arrive by TMA (one 5-D box per operand and tile, zero-filled outside the patch / beyond the
// channel count), double-buffered; fragments are fetched with ldmatrix.trans because the GEMM K
// dimension (voxels) is the slow dimension of both NDHWC operands.
template <bool FP16, int NTAPS>
__global__ void __launch_bounds__(kWgThreads, 1)
wgrad_mma_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_dy, const WgParams p) {
  constexpr int HL = NTAPS == 27 ? 1 : 0;
  constexpr int BXX = 8 + 2 * HL, BYY = 16 + 2 * HL, BZZ = kWgTZ + 2 * HL;
  constexpr uint32_t X_BYTES = BZZ * BYY * BXX * 64, D_BYTES = kWgTZ * 16 * 8 * 64;
  constexpr uint32_t X_STAGE = (X_BYTES + 1023) / 1024 * 1024, D_STAGE = (D_BYTES + 1023) / 1024 * 1024;
  constexpr int TPW = NTAPS == 27 ? 3 : 1;              // taps per warp
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + 2 * (X_STAGE + D_STAGE));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ci0 = blockIdx.y * 32, co0 = blockIdx.z * 32;
  if (tid == 0) {
    mbar_init(full, 1); mbar_init(full + 1, 1);
    fence_barrier_init();
    prefetch_tmap(&tmap_x); prefetch_tmap(&tmap_dy);
  }
  __syncthreads();
  float acc[TPW][2][4][4];
#pragma unroll
  for (int t = 0; t < TPW; ++t)
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
      for (int n = 0; n < 4; ++n) { acc[t][m][n][0] = acc[t][m][n][1] = acc[t][m][n][2] = acc[t][m][n][3] = 0.f; }

  auto issue = [&](long long tile, int stage) {          // warp 0, warp-uniform; one elected lane issues
    long long t = tile;
    int x0_r; t = divmod32(t, p.tiles_x, x0_r); const int x0 = x0_r * 8;
    int y0_r; t = divmod32(t, p.tiles_y, y0_r); const int y0 = y0_r * 16;
    int z0_r; t = divmod32(t, p.tiles_z, z0_r); const int z0 = z0_r * kWgTZ;
    const int n = static_cast<int>(t);
    if (elect_one()) {
      mbar_expect_tx(full + stage, X_BYTES + D_BYTES);
      tma_load_5d(smem + stage * (X_STAGE + D_STAGE), &tmap_x, full + stage, p.x_coff + ci0, x0 - HL, y0 - HL, z0 - HL, n);
      tma_load_5d(smem + stage * (X_STAGE + D_STAGE) + X_STAGE, &tmap_dy, full + stage, p.dy_coff + co0, x0, y0, z0, n);
    }
    __syncwarp();
  };

  const long long first = blockIdx.x, step = gridDim.x;
  int stage = 0;
  uint32_t phases = 0;                                     // bit s = parity of stage s
  if (warp == 0 && first < p.total_tiles) issue(first, 0);
  const int wdz = warp / 3, wdy = warp % 3;               // this warp's tap row (NTAPS = 27)
  for (long long tile = first; tile < p.total_tiles; tile += step) {
    const long long next = tile + step;
    if (warp == 0 && next < p.total_tiles) issue(next, stage ^ 1);
    mbar_wait(full + stage, (phases >> stage) & 1u);
    phases ^= 1u << stage;
    const uint32_t sx = smem_u32(smem + stage * (X_STAGE + D_STAGE));
    const uint32_t sd = sx + X_STAGE;
    // k-step = 16 voxels = two x rows (y, y+1) of one z plane
#pragma unroll 1
    for (int ks = (NTAPS == 27 ? 0 : warp); ks < kWgTZ * 8; ks += (NTAPS == 27 ? 1 : 9)) {
      const int pz = ks >> 3, yr = (ks & 7) * 2;
      const int mi = lane >> 3, r = lane & 7;
      // B fragments (dy): matrices (k-half kh, chunk c): lane group mi -> kh = mi & 1, chunk = 2 j + (mi >> 1)
      uint32_t bf[2][4];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int vox = (pz * 16 + yr + (mi & 1)) * 8 + r;
        ldmatrix_x4_trans(bf[j], sw64(sd, vox, 2 * j + (mi >> 1)));
      }
#pragma unroll
      for (int t = 0; t < TPW; ++t) {
        const int dz = NTAPS == 27 ? wdz : 0, dyy = NTAPS == 27 ? wdy : 0, dxx = NTAPS == 27 ? t : 0;   // +HL biased
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          // A fragment (x window): a0: m 0-7,k 0-7  a1: m 8-15,k 0-7  a2: m 0-7,k 8-15  a3: m 8-15,k 8-15
          uint32_t af[4];
          const int vox = ((pz + dz) * BYY + (yr + (mi >> 1) + dyy)) * BXX + (r + dxx);
          ldmatrix_x4_trans(af, sw64(sx, vox, mt * 2 + (mi & 1)));
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) mma16816<FP16>(acc[t][mt][nt], af, bf[nt >> 1][(nt & 1) * 2], bf[nt >> 1][(nt & 1) * 2 + 1]);
        }
      }
    }
    __syncthreads();                                       // every warp is done with this stage before it is refilled
    stage ^= 1;
  }
  // flush: c0,c1 -> (row g, cols 2 tq, 2 tq + 1); c2,c3 -> (row g + 8)
  const int g = lane >> 2, tq = lane & 3;
#pragma unroll
  for (int t = 0; t < TPW; ++t) {
    const int tap = NTAPS == 27 ? warp * 3 + t : 0;
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int ci = ci0 + mt * 16 + g + (e >> 1) * 8, co = co0 + nt * 8 + tq * 2 + (e & 1);
          if (ci < p.Ci && co < p.Cout)
            atomicAdd(p.dw + (static_cast<long long>(tap) * p.dw_cin + p.ci_off + ci) * p.Cout + co, acc[t][mt][nt][e]);
        }
  }
}

// ------------------------------------------------------------------------------------------------
// 1x1x1 head + sigmoid + binary cross-entropy, forward and backward in one pass:
//   logit = a . w + b ; p = sigmoid(logit)
//   loss += -[y log(clip(p) + eps) + (1 - y) log(1 - clip(p) + eps)] / n_total   (Keras backend, eps = 1e-7)
//   dl = dloss/dlogit ; dw[c] += a[c] dl ; db += dl ; da[c] = dl w[c]
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) head_loss_kernel(TView a, const float* __restrict__ w, const float* __restrict__ b,
                                                         const uint8_t* __restrict__ y, TView da, double* loss,
                                                         float inv_n, float* __restrict__ dl_out, float* probs) {
  const int C = a.C;
  const long long nvox = a.voxels();
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const T* ap = static_cast<const T*>(a.ptr);
  T* dp = static_cast<T*>(da.ptr);
  extern __shared__ float sw[];                 // [C] weights
  for (int i = threadIdx.x; i < C; i += blockDim.x) sw[i] = w[i];
  __syncthreads();
  const float bias = b[0];
  double lsum = 0.0;
  for (long long v = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; v < nvox; v += stride) {
    const T* row = ap + v * a.ctot + a.coff;
    float logit = bias;
    for (int c = 0; c < C; ++c) logit = fmaf(ldv<T>(row + c), sw[c], logit);
    const float pr = 1.f / (1.f + expf(-logit));
    if (probs) probs[v] = pr;
    const float yt = static_cast<float>(y[v]);
    const float eps = 1e-7f;
    const float pc = fminf(fmaxf(pr, eps), 1.f - eps);
    lsum += -(yt * logf(pc + eps) + (1.f - yt) * logf(1.f - pc + eps));
    // d/dp of the clipped expression (zero outside the clip range), times dp/dlogit = p (1 - p)
    const float dldp = (pr > eps && pr < 1.f - eps) ? -(yt / (pc + eps) - (1.f - yt) / (1.f - pc + eps)) : 0.f;
    const float dl = dldp * pr * (1.f - pr) * inv_n;
    dl_out[v] = dl;
    T* drow = dp + v * da.ctot + da.coff;
    for (int c = 0; c < C; ++c) stv<T>(drow + c, dl * sw[c]);
  }
  __shared__ double sl[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, o);
  if ((threadIdx.x & 31) == 0) sl[threadIdx.x >> 5] = lsum;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int i = 0; i < 8; ++i) tot += sl[i];
    atomicAdd(loss, tot * inv_n);
  }
}

// Fused, vectorised head: G = C / 8 lanes share a voxel (one 128-bit load of `a` and one 128-bit store of `da`
// per lane, logit reduced with shuffles), and the head's weight / bias gradient (sum_v a[v][c] dl[v], sum_v dl[v])
// is accumulated in the same pass, so `a` is read once and dl never re-read.
template <typename T>
__global__ void __launch_bounds__(256) head_loss_v8_kernel(TView a, const float* __restrict__ w, const float* __restrict__ b,
                                                            const uint8_t* __restrict__ y, TView da, double* loss, double* dwb2c,
                                                            float inv_n, float* __restrict__ dl_out, float* probs) {
  __shared__ float red[8][264];
  __shared__ double sl[8];
  const int C = a.C, G = C >> 3;
  const long long nvox = a.voxels();
  const long long nthreads = static_cast<long long>(gridDim.x) * blockDim.x;
  const long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int g = static_cast<int>(t % G);
  const long long istride = nthreads / G;
  const T* ap = static_cast<const T*>(a.ptr);
  T* dp = static_cast<T*>(da.ptr);
  float wr[8], sa[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { wr[j] = w[g * 8 + j]; sa[j] = 0.f; }
  const float bias = b[0];
  float lsum = 0.f, dsum = 0.f;
  for (long long v = t / G; v < nvox; v += istride) {
    float f[8];
    load8v<T>(ap + v * a.ctot + a.coff + g * 8, f);
    float part = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) part = fmaf(f[j], wr[j], part);
    for (int off = G >> 1; off > 0; off >>= 1) part += __shfl_xor_sync(0xffffffffu, part, off);
    const float logit = part + bias;
    const float pr = 1.f / (1.f + expf(-logit));
    const float yt = static_cast<float>(y[v]);
    const float eps = 1e-7f;
    const float pc = fminf(fmaxf(pr, eps), 1.f - eps);
    const float dldp = (pr > eps && pr < 1.f - eps) ? -(yt / (pc + eps) - (1.f - yt) / (1.f - pc + eps)) : 0.f;
    const float dl = dldp * pr * (1.f - pr) * inv_n;
    if (g == 0) {
      if (probs) probs[v] = pr;
      dl_out[v] = dl;
      lsum += -(yt * logf(pc + eps) + (1.f - yt) * logf(1.f - pc + eps));
      dsum += dl;
    }
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { o[j] = dl * wr[j]; sa[j] = fmaf(f[j], dl, sa[j]); }
    store8v<T>(dp + v * da.ctot + da.coff + g * 8, o);
  }
  for (int off = 16; off >= G; off >>= 1) {
#pragma unroll
    for (int j = 0; j < 8; ++j) sa[j] += __shfl_xor_sync(0xffffffffu, sa[j], off);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    lsum += __shfl_xor_sync(0xffffffffu, lsum, off);
    dsum += __shfl_xor_sync(0xffffffffu, dsum, off);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane < G) {
#pragma unroll
    for (int j = 0; j < 8; ++j) red[warp][lane * 8 + j] = sa[j];
  }
  if (lane == 0) { red[warp][256] = dsum; sl[warp] = static_cast<double>(lsum); }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    float tot = 0.f, dtot = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) { tot += red[k][i]; dtot += red[k][256]; }
    atomicAdd(dwb2c + i, static_cast<double>(tot));
    atomicAdd(dwb2c + C + i, static_cast<double>(dtot));
  }
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int i = 0; i < 8; ++i) tot += sl[i];
    atomicAdd(loss, tot * inv_n);
  }
}

// dwb[c] += sum_v a[v][c] dl[v]  (c < C);  dwb[C + c] += sum_v dl[v]  (identical for every c; the host reads c = 0)
template <typename T>
__global__ void __launch_bounds__(256) head_wgrad_kernel(TView a, const float* __restrict__ dl, double* dwb2c) {
  const T* p = static_cast<const T*>(a.ptr);
  per_channel_reduce2<T>(a, dwb2c, [&](long long vox, int c, float& s0, float& s1) {
    const float d = dl[vox];
    s0 = ldv<T>(p + vox * a.ctot + a.coff + c) * d; s1 = d;
  });
}

// ------------------------------------------------------------------------------------------------
// Autoencoder loss (RegularizedAutoencoder.train_step, /root/reference/tomo_encoders/neural_nets/autoencoders.py:498-526):
//   logit = a . w + b ; p = sigmoid(logit)
//   total = mean((x - p)^2) + weight * mean(KL(x || p)),  KL = clip(x, eps, 1) * log(clip(x, eps, 1) / clip(p, eps, 1))
// (keras.losses.mean_squared_error / kl_divergence over the single channel, eps = 1e-7).  loss3 += {total, mse, kl}.
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) head_ae_loss_kernel(TView a, const float* __restrict__ w, const float* __restrict__ b,
                                                            const T* __restrict__ target, TView da, double* loss3, float inv_n,
                                                            float weight, float* __restrict__ dl_out, float* probs) {
  const int C = a.C;
  const long long nvox = a.voxels();
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const T* ap = static_cast<const T*>(a.ptr);
  T* dp = static_cast<T*>(da.ptr);
  extern __shared__ float sw[];
  for (int i = threadIdx.x; i < C; i += blockDim.x) sw[i] = w[i];
  __syncthreads();
  const float bias = b[0];
  double s_mse = 0.0, s_kl = 0.0;
  for (long long v = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; v < nvox; v += stride) {
    const T* row = ap + v * a.ctot + a.coff;
    float logit = bias;
    for (int c = 0; c < C; ++c) logit = fmaf(ldv<T>(row + c), sw[c], logit);
    const float pr = 1.f / (1.f + expf(-logit));
    if (probs) probs[v] = pr;
    const float xt = ldv<T>(target + v);
    const float eps = 1e-7f;
    const float yt = fminf(fmaxf(xt, eps), 1.f), yp = fminf(fmaxf(pr, eps), 1.f);
    const float diff = xt - pr;
    s_mse += diff * diff;
    s_kl += yt * logf(yt / yp);
    float dldp = -2.f * diff;
    if (pr > eps && pr < 1.f) dldp += weight * (-yt / yp);
    const float dl = dldp * pr * (1.f - pr) * inv_n;
    dl_out[v] = dl;
    T* drow = dp + v * da.ctot + da.coff;
    for (int c = 0; c < C; ++c) stv<T>(drow + c, dl * sw[c]);
  }
  __shared__ double sl[2][8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { s_mse += __shfl_xor_sync(0xffffffffu, s_mse, o); s_kl += __shfl_xor_sync(0xffffffffu, s_kl, o); }
  if ((threadIdx.x & 31) == 0) { sl[0][threadIdx.x >> 5] = s_mse; sl[1][threadIdx.x >> 5] = s_kl; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double m = 0.0, k = 0.0;
    for (int i = 0; i < 8; ++i) { m += sl[0][i]; k += sl[1][i]; }
    atomicAdd(loss3, (m + weight * k) * inv_n);
    atomicAdd(loss3 + 1, m * inv_n);
    atomicAdd(loss3 + 2, k * inv_n);
  }
}

// Dense layer backward (fp32, small): dw[i][j] += sum_n x[n][i] dy[n][j]; db2[j] += sum_n dy[n][j]; dx[n][i] = sum_j dy[n][j] w[i][j]
__global__ void __launch_bounds__(256) dense_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy, long long n, int nin,
                                                           int nout, float* __restrict__ dw, double* __restrict__ db2) {
  const long long total = static_cast<long long>(nin) * nout;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < total; e += stride) {
    const int j = static_cast<int>(e % nout);
    const long long i = e / nout;
    float acc = 0.f;
    for (long long r = 0; r < n; ++r) acc = fmaf(x[r * nin + i], dy[r * nout + j], acc);
    dw[e] += acc;
    if (i == 0) {
      float sb = 0.f;
      for (long long r = 0; r < n; ++r) sb += dy[r * nout + j];
      atomicAdd(db2 + j, static_cast<double>(sb));
    }
  }
}
__global__ void __launch_bounds__(256) dense_dgrad_kernel(const float* __restrict__ dy, const float* __restrict__ w, long long n, int nin,
                                                           int nout, float* __restrict__ dx) {
  // one warp per (row, input index): lanes split the output range (coalesced reads of w[i][:]), shuffle reduction
  const long long warps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  const int lane = threadIdx.x & 31;
  for (long long e = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; e < n * nin; e += warps) {
    const long long r = e / nin, i = e - r * nin;
    float acc = 0.f;
    for (int j = lane; j < nout; j += 32) acc = fmaf(dy[r * nout + j], w[i * nout + j], acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) dx[e] = acc;
  }
}

// Keras Adam: lr_t = lr sqrt(1 - b2^t) / (1 - b1^t); theta -= lr_t m / (sqrt(v) + eps)
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ w, const float* __restrict__ g, float* __restrict__ m,
                                                    float* __restrict__ v, long long n, float lr_t, float b1, float b2,
                                                    float eps, float grad_scale) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float gi = g[i] * grad_scale;
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi; v[i] = vi;
    w[i] -= lr_t * mi / (sqrtf(vi) + eps);
  }
}

// w[tap][ci][co] -> wf[26 - tap][co][ci]  (the data gradient of a 'same' 3x3x3 conv is the same conv
// of dy with the spatially flipped, channel-transposed kernel)
__global__ void __launch_bounds__(256) flip_weights_kernel(const float* __restrict__ w, float* __restrict__ wf, int Ci, int Co,
                                                            int ci_lo, int ci_n) {
  // output covers input channels [ci_lo, ci_lo + ci_n) of the forward conv: wf[27][Co][ci_n]
  const long long total = 27LL * Co * ci_n;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < total; e += stride) {
    const int ci = static_cast<int>(e % ci_n);
    const int co = static_cast<int>((e / ci_n) % Co);
    const int tap = static_cast<int>(e / (static_cast<long long>(ci_n) * Co));
    wf[e] = w[(static_cast<long long>(26 - tap) * Ci + ci_lo + ci) * Co + co];
  }
}

// Device-side twin of conv_umma_pack().  ntaps = 27: fp32 w[tap][cin][cout] (Keras Conv3D layout).
// ntaps = 1, nsub = 1: pointwise conv, w[cin][cout].  ntaps = 1, nsub = 8: transposed conv,
// w[sub][cout][cin] (Keras Conv3DTranspose layout).  Output: swizzled 16-bit slabs.
__global__ void __launch_bounds__(256) pack_umma_kernel(const float* __restrict__ w, uint8_t* __restrict__ dst, int Cin, int Cout,
                                                         int CB, int NT, int fuse, int fp16, int ntaps, int nsub) {
  const long long total = static_cast<long long>(ntaps) * nsub * Cin * Cout;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const int rowb = CB * 2;
  const uint32_t mask = CB == 16 ? 1u : (CB == 32 ? 3u : 7u);
  const int kchunks = Cin / CB;
  const int taps_g = ntaps == 27 ? 3 : 1, ngroups = ntaps == 27 ? 9 : 1;
  const long long slab = static_cast<long long>(taps_g) * NT * rowb;
  for (long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < total; e += stride) {
    int cout, cin, tap = 0, sub = 0;
    if (nsub == 8) {                       // [sub][cout][cin]
      cin = static_cast<int>(e % Cin);
      cout = static_cast<int>((e / Cin) % Cout);
      sub = static_cast<int>(e / (static_cast<long long>(Cin) * Cout));
    } else {                               // [tap][cin][cout]
      cout = static_cast<int>(e % Cout);
      cin = static_cast<int>((e / Cout) % Cin);
      tap = static_cast<int>(e / (static_cast<long long>(Cin) * Cout));
    }
    const int nt = cout / NT, r = cout % NT, kc = cin / CB, kk = cin % CB;
    int g = 0, t = 0;
    if (ntaps == 27) { if (fuse) { g = tap % 9; t = 2 - tap / 9; } else { g = tap / 3; t = tap % 3; } }
    uint8_t* base = dst + (((static_cast<long long>(nt) * nsub + sub) * kchunks + kc) * ngroups + g) * slab;
    const uint32_t Lo = static_cast<uint32_t>((t * NT + r) * rowb + kk * 2);
    const uint32_t P = Lo ^ (((Lo >> 7) & mask) << 4);
    const float val = w[e];
    uint16_t bits;
    if (fp16) { __half h = __float2half_rn(val); bits = *reinterpret_cast<uint16_t*>(&h); }
    else { __nv_bfloat16 h = __float2bfloat16_rn(val); bits = *reinterpret_cast<uint16_t*>(&h); }
    *reinterpret_cast<uint16_t*>(base + P) = bits;
  }
}

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

int make_tmap5(CUtensorMap* tm, const te_tensor& v, const ConvUmmaConfig& cfg, int halo = 1) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled is not available"); return TE_ERR_CUDA; }
  cuuint64_t gdim[5] = {(cuuint64_t)v.ctot, (cuuint64_t)v.W, (cuuint64_t)v.H, (cuuint64_t)v.D, (cuuint64_t)v.N};
  cuuint64_t gstr[4] = {(cuuint64_t)v.ctot * 2, (cuuint64_t)v.W * v.ctot * 2, (cuuint64_t)v.H * v.W * v.ctot * 2,
                        (cuuint64_t)v.D * v.H * v.W * v.ctot * 2};
  cuuint32_t box[5] = {(cuuint32_t)cfg.CB, (cuuint32_t)(8 + 2 * halo), (cuuint32_t)(16 + 2 * halo), (cuuint32_t)(cfg.TZ + 2 * halo), 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  const CUtensorMapSwizzle sw = cfg.CB == 16 ? CU_TENSOR_MAP_SWIZZLE_32B
                                : (cfg.CB == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B);
  CUresult r = enc(tm, v.dtype == TE_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, v.ptr, gdim,
                   gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return TE_ERR_CUDA; }
  return TE_OK;
}

// Packs the weights on the device and launches conv_umma_kernel.  ntaps 27: 3x3x3 'same' conv
// (w[27][Cin][Cout]); ntaps 1, up_stride 0: pointwise conv (w[Cin][Cout]); ntaps 1, up_stride 2:
// transposed conv k = s = 2 (w[8][Cout][Cin]).
int umma_conv(const te_tensor& x, const te_tensor* x2, const float* w, const float* bias, const te_tensor& y, void* wpack,
              const ConvUmmaConfig& cfg, int ntaps, int up_stride, int lrelu, float alpha, cudaStream_t st) {
  const int Cin = x.C + (x2 ? x2->C : 0), Cout = y.C;
  const int nsub = up_stride ? 8 : 1;
  const int fuse = ntaps == 27 && 3 * cfg.NT <= 256;
  pack_umma_kernel<<<blocks_for(static_cast<long long>(ntaps) * nsub * Cin * Cout, 256), 256, 0, st>>>(
      w, static_cast<uint8_t*>(wpack), Cin, Cout, cfg.CB, cfg.NT, fuse, x.dtype == TE_F16, ntaps, nsub);
  TE_LAUNCH_CHECK();
  ConvUmmaLaunch C{};
  const int halo = ntaps == 27 ? 1 : 0;
  int rc = make_tmap5(&C.tmap_in, x, cfg, halo);
  if (rc != TE_OK) return rc;
  if (x2) { rc = make_tmap5(&C.tmap_in2, *x2, cfg, halo); if (rc != TE_OK) return rc; } else C.tmap_in2 = C.tmap_in;
  C.Cin2 = x2 ? x2->C : 0;
  C.w_packed = wpack; C.bias = bias;
  C.out = y.ptr; C.out_ctot = y.ctot; C.out_coff = y.coff;
  C.N = x.N; C.D = x.D; C.H = x.H; C.W = x.W;
  C.Cin = Cin; C.Cout = Cout; C.ntaps = ntaps; C.up_stride = up_stride; C.act_lrelu = lrelu; C.alpha = alpha;
  C.fp16 = x.dtype == TE_F16;
  C.CB = cfg.CB; C.NT = cfg.NT; C.TZ = cfg.TZ;
  conv_umma_make_out_maps(&C);
  cudaError_t e = conv_umma_launch(C, st);
  count_launch();
  if (e != cudaSuccess) { set_error("tensor-core conv launch failed: %s", cudaGetErrorString(e)); return TE_ERR_CUDA; }
  return TE_OK;
}

bool dense(const te_tensor& t) { return t.coff == 0 && t.C == t.ctot; }

bool same_grid(const te_tensor& a, const te_tensor& b) {
  return a.N == b.N && a.D == b.D && a.H == b.H && a.W == b.W;
}

}  // namespace
}  // namespace te

using namespace te;

#define TE_TR_CHECK_T(t, name) TE_CHECK_ARG((t) && (t)->ptr && (t)->C > 0 && (t)->ctot >= (t)->coff + (t)->C, name ": bad tensor")

// y = conv3x3x3_same(x [, x2]) + bias.  w: device fp32 [27][Cin][Cout] with Cin = x.C (+ x2.C).
// 16-bit tensors with Cin, Cout multiples of 16 (and dense inputs) run on the tcgen05 kernel: the
// weights are packed on the device into wpack (te_tr_conv3_wpack_bytes) first.
extern "C" int64_t te_tr_conv3_wpack_bytes(int32_t Cin, int32_t Cout) {
  ConvUmmaConfig cfg;
  if (!conv_umma_config(Cin, Cout, &cfg)) return 0;
  return static_cast<int64_t>(conv_umma_packed_bytes(Cin, Cout, 27, 1, cfg));
}

extern "C" int te_tr_conv3(const te_tensor* x, const te_tensor* x2, const float* w, const float* bias, const te_tensor* y,
                           void* wpack, void* stream) {
  TE_TR_CHECK_T(x, "te_tr_conv3"); TE_TR_CHECK_T(y, "te_tr_conv3");
  TE_CHECK_ARG(w && bias, "te_tr_conv3: null pointer");
  TE_CHECK_ARG(same_grid(*x, *y) && x->dtype == y->dtype, "te_tr_conv3: x and y must share grid and dtype");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int Cin = x->C + (x2 ? x2->C : 0), Cout = y->C;
  ConvUmmaConfig cfg;
  const bool dense_in = x->coff == 0 && x->C == x->ctot && (!x2 || (x2->coff == 0 && x2->C == x2->ctot));
  const bool umma = x->dtype != TE_F32 && wpack && dense_in && conv_umma_config(Cin, Cout, &cfg) &&
                    (!x2 || (x->C % cfg.CB == 0 && x2->C % cfg.CB == 0));
  if (umma) return umma_conv(*x, x2, w, bias, *y, wpack, cfg, 27, 0, 0, 0.f, st);
  TE_CHECK_ARG(!x2, "te_tr_conv3: two-input form needs the tensor-core path (16-bit, dense, channel counts % 16)");
  cudaError_t e = simt_conv3(view_of(*x), view_of(*y), w, bias, 0, 0.f, act_of(x->dtype), st);
  count_launch();
  if (e != cudaSuccess) { set_error("te_tr_conv3: launch failed: %s", cudaGetErrorString(e)); return TE_ERR_CUDA; }
  return TE_OK;
}

extern "C" int te_conv1_kernel_kind(const te_tensor* x, const te_tensor* y) {
  if (!x || !y || x->C != 1 || x->ctot != 1 || x->dtype == TE_F32 || x->dtype != y->dtype) return 0;
  if (conv1_tz_applies(view_of(*x), view_of(*y), act_of(x->dtype))) return 2;
  static const bool off = getenv("TOMOENC_NO_C1_UMMA") != nullptr;
  return (!off && (y->C == 16 || y->C == 32) && y->ctot % 8 == 0 && y->coff % 8 == 0) ? 1 : 0;
}

// wf[27][Cout][ci_n] = flipped / transposed slice of w[27][Cin][Cout] (dgrad weights)
extern "C" int te_tr_flip_weights(const float* w, int32_t Cin, int32_t Cout, int32_t ci_lo, int32_t ci_n, float* wf, void* stream) {
  TE_CHECK_ARG(w && wf && ci_lo >= 0 && ci_n > 0 && ci_lo + ci_n <= Cin, "te_tr_flip_weights: bad argument");
  flip_weights_kernel<<<blocks_for(27LL * Cout * ci_n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(w, wf, Cin, Cout, ci_lo, ci_n);
  TE_LAUNCH_CHECK();
  return TE_OK;
}

extern "C" int te_tr_bn_stats(const te_tensor* y, double* stats2c, void* stream) {
  TE_TR_CHECK_T(y, "te_tr_bn_stats");
  TE_CHECK_ARG(stats2c && y->C <= (1 << 20), "te_tr_bn_stats: bad argument");
  const TView v = view_of(*y);
  int grid = blocks_for(v.voxels() * v.C, 256, 8);
  if (static_cast<long long>(grid) * 256 < v.C) grid = (v.C + 255) / 256;
  if (v8_ok(*y)) {
    const int g8 = blocks_for(v.voxels() * (v.C >> 3), 256, 8);
    if (y->dtype == TE_F16) bn_stats_v8_kernel<__half><<<g8, 256, 0, static_cast<cudaStream_t>(stream)>>>(v, stats2c);
    else bn_stats_v8_kernel<__nv_bfloat16><<<g8, 256, 0, static_cast<cudaStream_t>(stream)>>>(v, stats2c);
    TE_LAUNCH_CHECK();
    return TE_OK;
  }
  TE_DISPATCH_T(y->dtype, (bn_stats_kernel<T><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(v, stats2c)));
  TE_LAUNCH_CHECK();
  return TE_OK;
}

extern "C" int te_tr_bn_act_fwd(const te_tensor* y, const te_tensor* a, const float* scale, const float* shift, float alpha,
                                void* stream) {
  TE_TR_CHECK_T(y, "te_tr_bn_act_fwd"); TE_TR_CHECK_T(a, "te_tr_bn_act_fwd");
  TE_CHECK_ARG(scale && shift && same_grid(*y, *a) && y->C == a->C && y->dtype == a->dtype, "te_tr_bn_act_fwd: bad argument");
  const TView vy = view_of(*y), va = view_of(*a);
  if (v8_ok(*y) && v8_ok(*a)) {
    const int g8 = blocks_for(vy.voxels() * (vy.C >> 3), 256);
    if (y->dtype == TE_F16) bn_act_fwd_v8_kernel<__half><<<g8, 256, 0, static_cast<cudaStream_t>(stream)>>>(vy, va, scale, shift, alpha);
    else bn_act_fwd_v8_kernel<__nv_bfloat16><<<g8, 256, 0, static_cast<cudaStream_t>(stream)>>>(vy, va, scale, shift, alpha);
    TE_LAUNCH_CHECK();
    return TE_OK;
  }
  TE_DISPATCH_T(y->dtype, (bn_act_fwd_kernel<T><<<blocks_for(vy.voxels() * vy.C, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
                              vy, va, scale, shift, alpha)));
  TE_LAUNCH_CHECK();
  return TE_OK;
}

extern "C" int te_tr_bn_act_bwd_reduce(const te_tensor* y, const te_tensor* da, const float* scale, const float* shift,
                                       const float* mean, const float* rstd, float alpha, double* red2c, void* stream) {
  TE_TR_CHECK_T(y, "te_tr_bn_act_bwd_reduce"); TE_TR_CHECK_T(da, "te_tr_bn_act_bwd_reduce");
  TE_CHECK_ARG(scale && shift && mean && rstd && red2c && same_grid(*y, *da) && y->C == da->C && y->dtype == da->dtype,
               "te_tr_bn_act_bwd_reduce: bad argument");
  const TView vy = view_of(*y), vd = view_of(*da);
  if (v8_ok(*y) && v8_ok(*da)) {
    const int g8 = blocks_for(vy.voxels() * (vy.C >> 3), 256, 8);
    cudaStream_t st8 = static_cast<cudaStream_t>(stream);
    if (y->dtype == TE_F16) bn_act_bwd_reduce_v8_kernel<__half><<<g8, 256, 0, st8>>>(vy, vd, scale, shift, mean, rstd, alpha, red2c);
    else bn_act_bwd_reduce_v8_kernel<__nv_bfloat16><<<g8, 256, 0, st8>>>(vy, vd, scale, shift, mean, rstd, alpha, red2c);
    TE_LAUNCH_CHECK();
    return TE_OK;
  }
  int grid = blocks_for(vy.voxels() * vy.C, 256, 8);
  if (static_cast<long long>(grid) * 256 < vy.C) grid = (vy.C + 255) / 256;
  TE_DISPATCH_T(y->dtype, (bn_act_bwd_reduce_kernel<T><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
                              vy, vd, scale, shift, mean, rstd, alpha, red2c)));
  TE_LAUNCH_CHECK();
  return TE_OK;
}

extern "C" int te_tr_bn_act_bwd_apply(const te_tensor* y, const te_tensor* da, const te_tensor* dy, const float* scale,
                                      const float* shift, const float* mean, const float* rstd, const float* dbeta_n,
                                      const float* dgamma_n, float alpha, double* dsum2c, void* stream) {
  TE_TR_CHECK_T(y, "te_tr_bn_act_bwd_apply"); TE_TR_CHECK_T(da, "te_tr_bn_act_bwd_apply"); TE_TR_CHECK_T(dy, "te_tr_bn_act_bwd_apply");
  TE_CHECK_ARG(scale && shift && mean && rstd && dbeta_n && dgamma_n && same_grid(*y, *da) && same_grid(*y, *dy) &&
                   y->C == da->C && y->C == dy->C && y->dtype == da->dtype && y->dtype == dy->dtype,
               "te_tr_bn_act_bwd_apply: bad argument");
  const TView vy = view_of(*y), vd = view_of(*da), vo = view_of(*dy);
  cudaStream_t st8 = static_cast<cudaStream_t>(stream);
  if (v8_ok(*y) && v8_ok(*da) && v8_ok(*dy)) {
    const int g8 = blocks_for(vy.voxels() * (vy.C >> 3), 256, 8);
    if (y->dtype == TE_F16) bn_act_bwd_apply_v8_kernel<__half><<<g8, 256, 0, st8>>>(vy, vd, vo, scale, shift, mean, rstd, dbeta_n, dgamma_n, alpha, dsum2c);
    else bn_act_bwd_apply_v8_kernel<__nv_bfloat16><<<g8, 256, 0, st8>>>(vy, vd, vo, scale, shift, mean, rstd, dbeta_n, dgamma_n, alpha, dsum2c);
    TE_LAUNCH_CHECK();
    return TE_OK;
  }
  TE_DISPATCH_T(y->dtype, (bn_act_bwd_apply_kernel<T><<<blocks_for(vy.voxels() * vy.C, 256), 256, 0, st8>>>(
                              vy, vd, vo, scale, shift, mean, rstd, dbeta_n, dgamma_n, alpha)));
  TE_LAUNCH_CHECK();
  if (dsum2c) {
    int grid = blocks_for(vo.voxels() * vo.C, 256, 8);
    if (static_cast<long long>(grid) * 256 < vo.C) grid = (vo.C + 255) / 256;
    TE_DISPATCH_T(dy->dtype, (channel_sum_kernel<T><<<grid, 256, 0, st8>>>(vo, dsum2c)));
    TE_LAUNCH_CHECK();
  }
  return TE_OK;
}

extern "C" int te_tr_maxpool_fwd(const te_tensor* x, const te_tensor* out, int32_t pool, void* stream) {
  TE_TR_CHECK_T(x, "te_tr_maxpool_fwd"); TE_TR_CHECK_T(out, "te_tr_maxpool_fwd");
  TE_CHECK_ARG(pool >= 1 && x->D % pool == 0 && x->H % pool == 0 && x->W % pool == 0 && x->C == out->C && x->dtype == out->dtype,
               "te_tr_maxpool_fwd: bad argument");
  cudaError_t e = simt_maxpool(view_of(*x), view_of(*out), pool, act_of(x->dtype), static_cast<cudaStream_t>(stream));
  count_launch();
  if (e != cudaSuccess) { set_error("te_tr_maxpool_fwd: %s", cudaGetErrorString(e)); return TE_ERR_CUDA; }
  return TE_OK;
}

extern "C" int te_tr_maxpool_bwd(const te_tensor* x, const te_tensor* dpool, const te_tensor* dx, int32_t pool, int32_t accumulate,
                                 void* stream) {
  TE_TR_CHECK_T(x, "te_tr_maxpool_bwd"); TE_TR_CHECK_T(dpool, "te_tr_maxpool_bwd"); TE_TR_CHECK_T(dx, "te_tr_maxpool_bwd");
  TE_CHECK_ARG(pool >= 1 && same_grid(*x, *dx) && x->C == dx->C && x->C == dpool->C && x->dtype == dx->dtype &&
                   x->dtype == dpool->dtype, "te_tr_maxpool_bwd: bad argument");
  const TView vx = view_of(*x), vg = view_of(*dpool), vd = view_of(*dx);
  if (pool == 2 && v8_ok(*x) && v8_ok(*dpool) && v8_ok(*dx) && x->D % 2 == 0 && x->H % 2 == 0 && x->W % 2 == 0) {
    const int grid = blocks_for(vg.voxels() * (vg.C >> 3), 256, 8);
    if (x->dtype == TE_F16) maxpool2_bwd_v8_kernel<__half><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(vx, vg, vd, accumulate);
    else maxpool2_bwd_v8_kernel<__nv_bfloat16><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(vx, vg, vd, accumulate);
    TE_LAUNCH_CHECK();
    return TE_OK;
  }
  TE_DISPATCH_T(x->dtype, (maxpool_bwd_kernel<T><<<blocks_for(vg.voxels() * vg.C, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
                              vx, vg, vd, pool, accumulate)));
  TE_LAUNCH_CHECK();
  return TE_OK;
}

// a = lrelu(convT_{k=2,stride}(x) + bias); w Keras layout (2,2,2,Cout,Cin) fp32 on the device.  16-bit dense
// tensors with channel counts % 16 and stride 2 run on the tcgen05 kernel (wpack as for te_tr_conv3).
extern "C" int te_tr_upconv_fwd(const te_tensor* x, const float* w, const float* bias, const te_tensor* a, int32_t stride,
                                float alpha, void* wpack, void* stream) {
  TE_TR_CHECK_T(x, "te_tr_upconv_fwd"); TE_TR_CHECK_T(a, "te_tr_upconv_fwd");
  TE_CHECK_ARG(w && bias && stride >= 2 && a->D == x->D * stride && a->H == x->H * stride && a->W == x->W * stride &&
                   a->N == x->N && x->dtype == a->dtype, "te_tr_upconv_fwd: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ConvUmmaConfig cfg;
  if (x->dtype != TE_F32 && wpack && stride == 2 && dense(*x) && conv_umma_config(x->C, a->C, &cfg))
    return umma_conv(*x, nullptr, w, bias, *a, wpack, cfg, 1, 2, 1, alpha, st);
  cudaError_t e = simt_upconv(view_of(*x), view_of(*a), w, bias, stride, 1, alpha, act_of(x->dtype), st);
  count_launch();
  if (e != cudaSuccess) { set_error("te_tr_upconv_fwd: %s", cudaGetErrorString(e)); return TE_ERR_CUDA; }
  return TE_OK;
}

namespace te {
namespace {
constexpr int kWgSmem = 2 * (((kWgTZ + 2) * 18 * 10 * 64 + 1023) / 1024 * 1024 + kWgTZ * 16 * 8 * 64) + 1024 + 64;

template <bool FP16, int NTAPS>
cudaError_t launch_wgrad_mma_t(const CUtensorMap& tx, const CUtensorMap& td, const WgParams& p, dim3 g, cudaStream_t st) {
  static unsigned long long attr = 0;
  cudaError_t e = smem_attr_once(wgrad_mma_kernel<FP16, NTAPS>, kWgSmem, &attr);
  if (e != cudaSuccess) return e;
  wgrad_mma_kernel<FP16, NTAPS><<<g, kWgThreads, kWgSmem, st>>>(tx, td, p);
  return cudaGetLastError();
}

bool wgrad_mma_ok(const te_tensor& x, const te_tensor& dy) {
  static const bool no_mma = getenv("TOMOENC_NO_MMA_WGRAD") != nullptr;
  return !no_mma && x.dtype != TE_F32 && x.C % 8 == 0 && dy.C % 8 == 0 && x.ctot % 8 == 0 && dy.ctot % 8 == 0 &&
         x.coff % 8 == 0 && dy.coff % 8 == 0 && x.C >= 8;
}

// 5-D map (C, W, H, D, N) with a box of 32 channels x (8 + 2h) x (16 + 2h) x (TZ + 2h) voxels, 64-byte swizzle
int make_tmap_wg(CUtensorMap* tm, const te_tensor& v, int halo) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled is not available"); return TE_ERR_CUDA; }
  cuuint64_t gdim[5] = {(cuuint64_t)v.ctot, (cuuint64_t)v.W, (cuuint64_t)v.H, (cuuint64_t)v.D, (cuuint64_t)v.N};
  cuuint64_t gstr[4] = {(cuuint64_t)v.ctot * 2, (cuuint64_t)v.W * v.ctot * 2, (cuuint64_t)v.H * v.W * v.ctot * 2,
                        (cuuint64_t)v.D * v.H * v.W * v.ctot * 2};
  cuuint32_t box[5] = {32, (cuuint32_t)(8 + 2 * halo), (cuuint32_t)(16 + 2 * halo), (cuuint32_t)(kWgTZ + 2 * halo), 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(tm, v.dtype == TE_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, v.ptr, gdim,
                   gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return TE_ERR_CUDA; }
  return TE_OK;
}

// dw[ntaps][dw_cin][dy.C] += wgrad(x, dy) on mma.sync tiles
int wgrad_mma(const te_tensor& x, const te_tensor& dy, float* dw, int dw_cin, int ci_off, int ntaps, cudaStream_t st) {
  WgParams p;
  p.dw = dw; p.dw_cin = dw_cin; p.ci_off = ci_off; p.Cout = dy.C;
  p.Ci = x.C; p.x_coff = x.coff; p.dy_coff = dy.coff;
  p.N = x.N; p.D = x.D; p.H = x.H; p.W = x.W;
  p.tiles_x = (x.W + 7) / 8; p.tiles_y = (x.H + 15) / 16; p.tiles_z = (x.D + kWgTZ - 1) / kWgTZ;
  p.total_tiles = static_cast<long long>(p.N) * p.tiles_z * p.tiles_y * p.tiles_x;
  CUtensorMap tx, td;
  int rc = make_tmap_wg(&tx, x, ntaps == 27 ? 1 : 0);
  if (rc != TE_OK) return rc;
  rc = make_tmap_wg(&td, dy, 0);
  if (rc != TE_OK) return rc;
  const int cib = (x.C + 31) / 32, cob = (dy.C + 31) / 32;
  long long gx = (static_cast<long long>(kNumSMs) + cib * cob - 1) / (cib * cob);
  if (gx < 1) gx = 1;
  if (gx > p.total_tiles) gx = p.total_tiles;
  dim3 g(static_cast<unsigned>(gx), cib, cob);
  const bool fp16 = x.dtype == TE_F16;
  cudaError_t e = ntaps == 27 ? (fp16 ? launch_wgrad_mma_t<true, 27>(tx, td, p, g, st) : launch_wgrad_mma_t<false, 27>(tx, td, p, g, st))
                              : (fp16 ? launch_wgrad_mma_t<true, 1>(tx, td, p, g, st) : launch_wgrad_mma_t<false, 1>(tx, td, p, g, st));
  count_launch();
  if (e != cudaSuccess) { set_error("wgrad launch failed: %s", cudaGetErrorString(e)); return TE_ERR_CUDA; }
  return TE_OK;
}
}  // namespace
}  // namespace te

// Transposed conv backward.  dbias2c[0..C) += sum dz; dw (2,2,2,Cout,Cin) += weight gradient; dx = data gradient
// (optional).  Tensor-core path (16-bit, dense, channels % 16, dz2 = scratch tensor (N, D/2, H/2, W/2, 8 Cout) of the
// same dtype, wpack): da is left untouched and dz goes, rearranged, to dz2.  Otherwise (fp32 check mode) da is
// overwritten with dz in place and CUDA-core kernels are used.
extern "C" int te_tr_upconv_bwd(const te_tensor* x, const te_tensor* a, const te_tensor* da, const float* w, int32_t stride,
                                float alpha, float* dw, double* dbias2c, const te_tensor* dx, const te_tensor* dz2, void* wpack,
                                void* stream) {
  TE_TR_CHECK_T(x, "te_tr_upconv_bwd"); TE_TR_CHECK_T(a, "te_tr_upconv_bwd"); TE_TR_CHECK_T(da, "te_tr_upconv_bwd");
  TE_CHECK_ARG(w && dw && dbias2c && stride == 2 && same_grid(*a, *da) && a->C == da->C && x->dtype == a->dtype && a->dtype == da->dtype,
               "te_tr_upconv_bwd: bad argument (stride 2 only)");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const TView vx = view_of(*x), va = view_of(*a), vd = view_of(*da);
  ConvUmmaConfig cfg;
  const bool tc = a->dtype != TE_F32 && dz2 && dz2->ptr && wpack && dense(*dz2) && dz2->C == 8 * a->C && same_grid(*dz2, *x) &&
                  dz2->dtype == a->dtype && wgrad_mma_ok(*dz2, *x) && (!dx || (dense(*dx) && conv_umma_config(dz2->C, x->C, &cfg)));
  if (tc) {
    const TView vz = view_of(*dz2);
    int grid = blocks_for(va.voxels() * va.C, 256, 8);
    if (static_cast<long long>(grid) * 256 < va.C) grid = (va.C + 255) / 256;
    if (v8_ok(*a) && v8_ok(*da) && dz2->ctot % 8 == 0) {
      const int g8 = blocks_for(va.voxels() * (va.C >> 3), 256, 8);
      if (a->dtype == TE_F16) lrelu_bwd_s2d_v8_kernel<__half><<<g8, 256, 0, st>>>(va, vd, vz, alpha, dbias2c);
      else lrelu_bwd_s2d_v8_kernel<__nv_bfloat16><<<g8, 256, 0, st>>>(va, vd, vz, alpha, dbias2c);
    } else {
      TE_DISPATCH_T(a->dtype, (lrelu_bwd_s2d_kernel<T><<<grid, 256, 0, st>>>(va, vd, vz, alpha, dbias2c)));
    }
    TE_LAUNCH_CHECK();
    // dw[(sub, co)][ci] = sum_v dz2[v][(sub, co)] x[v][ci]
    int rc = wgrad_mma(*dz2, *x, dw, dz2->C, 0, 1, st);
    if (rc != TE_OK) return rc;
    if (dx) {
      TE_CHECK_ARG(same_grid(*x, *dx) && x->C == dx->C && x->dtype == dx->dtype, "te_tr_upconv_bwd: dx must match x");
      // dx[v][ci] = sum_(sub, co) dz2[v][(sub, co)] w[(sub, co)][ci]: pointwise conv, w already in [Cin' = 8 Cout][Cout' = Cin] order
      static float* zero_bias = nullptr;
      if (!zero_bias) { TE_CUDA(cudaMalloc(&zero_bias, 4096 * sizeof(float))); TE_CUDA(cudaMemset(zero_bias, 0, 4096 * sizeof(float))); }
      TE_CHECK_ARG(dx->C <= 4096, "te_tr_upconv_bwd: too many channels");
      rc = umma_conv(*dz2, nullptr, w, zero_bias, *dx, wpack, cfg, 1, 0, 0, 0.f, st);
      if (rc != TE_OK) return rc;
    }
    return TE_OK;
  }
  int grid = blocks_for(va.voxels() * va.C, 256, 8);
  if (static_cast<long long>(grid) * 256 < va.C) grid = (va.C + 255) / 256;
  TE_DISPATCH_T(a->dtype, (lrelu_bwd_bias_kernel<T><<<grid, 256, 0, st>>>(va, vd, alpha, dbias2c)));
  TE_LAUNCH_CHECK();
  const long long elems = static_cast<long long>(va.C) * vx.C;
  const long long nvox = vx.voxels();
  int splits = static_cast<int>(nvox / 2048);
  if (splits < 1) splits = 1;
  if (splits > 512) splits = 512;
  dim3 g(static_cast<unsigned>((elems + 255) / 256), 8, splits);
  TE_DISPATCH_T(a->dtype, (upconv_wgrad_kernel<T><<<g, 256, 0, st>>>(vx, vd, dw, stride, splits)));
  TE_LAUNCH_CHECK();
  if (dx) {
    TE_TR_CHECK_T(dx, "te_tr_upconv_bwd");
    TE_CHECK_ARG(same_grid(*x, *dx) && x->C == dx->C && x->dtype == dx->dtype, "te_tr_upconv_bwd: dx must match x");
    const TView vo = view_of(*dx);
    TE_DISPATCH_T(a->dtype, (upconv_dgrad_kernel<T><<<blocks_for(vo.voxels() * vo.C, 256, 32), 256, 0, st>>>(vd, vo, w, stride)));
    TE_LAUNCH_CHECK();
  }
  return TE_OK;
}

// dw[27][dw_cin][Cout] (+)= wgrad(x, dy) for the input channels [ci_off, ci_off + x.C); dbias2c may be null
extern "C" int te_tr_conv3_wgrad(const te_tensor* x, const te_tensor* dy, float* dw, int32_t dw_cin, int32_t ci_off,
                                 double* dbias2c, void* stream) {
  TE_TR_CHECK_T(x, "te_tr_conv3_wgrad"); TE_TR_CHECK_T(dy, "te_tr_conv3_wgrad");
  TE_CHECK_ARG(dw && same_grid(*x, *dy) && x->dtype == dy->dtype && ci_off >= 0 && ci_off + x->C <= dw_cin,
               "te_tr_conv3_wgrad: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const TView vx = view_of(*x), vd = view_of(*dy);
  if (te::wgrad_umma_ok(*x, *dy)) {
    const int rc = te::wgrad_umma(*x, *dy, dw, dw_cin, ci_off, st);
    if (rc != TE_OK) return rc;
  } else if (wgrad_mma_ok(*x, *dy)) {
    const int rc = wgrad_mma(*x, *dy, dw, dw_cin, ci_off, 27, st);
    if (rc != TE_OK) return rc;
  } else if (x->C == 1 && x->dtype != TE_F32 && (dy->C == 8 || dy->C == 16) && dy->ctot % 8 == 0 && dy->coff % 8 == 0) {
    const size_t shb = (10 * 18 * 40 + 1024 * static_cast<size_t>(dy->C)) * sizeof(float);
    const long long tiles = static_cast<long long>(vd.N) * ((vd.D + 7) / 8) * ((vd.H + 15) / 16) * ((vd.W + 7) / 8);
    const int grid = static_cast<int>(tiles < 2LL * kNumSMs ? tiles : 2LL * kNumSMs);
    if (x->dtype == TE_F16) {
      TE_CUDA(cudaFuncSetAttribute(conv3_wgrad_c1_tiled_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(shb)));
      conv3_wgrad_c1_tiled_kernel<__half><<<grid, 384, shb, st>>>(vx, vd, dw, dw_cin, ci_off);
    } else {
      TE_CUDA(cudaFuncSetAttribute(conv3_wgrad_c1_tiled_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(shb)));
      conv3_wgrad_c1_tiled_kernel<__nv_bfloat16><<<grid, 384, shb, st>>>(vx, vd, dw, dw_cin, ci_off);
    }
    TE_LAUNCH_CHECK();
  } else if (x->C == 1 && dy->C <= 37) {
    const int threads = (27 * dy->C + 31) / 32 * 32;
    const size_t shb = (6 * 180 + 4 * 128 * static_cast<size_t>(dy->C)) * sizeof(float);
    const long long tiles = static_cast<long long>(vd.N) * ((vd.D + 3) / 4) * ((vd.H + 15) / 16) * ((vd.W + 7) / 8);
    const int grid = static_cast<int>(tiles < 2LL * kNumSMs ? tiles : 2LL * kNumSMs);
    if (x->dtype == TE_F32) {
      TE_CUDA(cudaFuncSetAttribute(conv3_wgrad_c1_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(shb)));
      conv3_wgrad_c1_kernel<float><<<grid, threads, shb, st>>>(vx, vd, dw, dw_cin, ci_off);
    } else if (x->dtype == TE_F16) {
      TE_CUDA(cudaFuncSetAttribute(conv3_wgrad_c1_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(shb)));
      conv3_wgrad_c1_kernel<__half><<<grid, threads, shb, st>>>(vx, vd, dw, dw_cin, ci_off);
    } else {
      TE_CUDA(cudaFuncSetAttribute(conv3_wgrad_c1_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(shb)));
      conv3_wgrad_c1_kernel<__nv_bfloat16><<<grid, threads, shb, st>>>(vx, vd, dw, dw_cin, ci_off);
    }
    TE_LAUNCH_CHECK();
  } else {
    const long long elems = static_cast<long long>(vx.C) * vd.C;
    const long long nvox = vd.voxels();
    int splits = static_cast<int>(nvox / (elems >= 256 ? 4096 : 512));
    if (splits < 1) splits = 1;
    if (splits > 2048) splits = 2048;
    dim3 g(static_cast<unsigned>((elems + 255) / 256), 27, splits);
    TE_DISPATCH_T(x->dtype, (conv3_wgrad_simt_kernel<T><<<g, 256, 0, st>>>(vx, vd, dw, dw_cin, ci_off, splits)));
    TE_LAUNCH_CHECK();
  }
  if (dbias2c) {
    int grid = blocks_for(vd.voxels() * vd.C, 256, 8);
    if (static_cast<long long>(grid) * 256 < vd.C) grid = (vd.C + 255) / 256;
    if (v8_ok(*dy)) {
      const int g8 = blocks_for(vd.voxels() * (vd.C >> 3), 256, 8);
      if (dy->dtype == TE_F16) channel_sum_v8_kernel<__half><<<g8, 256, 0, st>>>(vd, dbias2c);
      else channel_sum_v8_kernel<__nv_bfloat16><<<g8, 256, 0, st>>>(vd, dbias2c);
    } else {
      TE_DISPATCH_T(dy->dtype, (channel_sum_kernel<T><<<grid, 256, 0, st>>>(vd, dbias2c)));
    }
    TE_LAUNCH_CHECK();
  }
  return TE_OK;
}

// loss (double, accumulated); dwb2c (double [2C], accumulated): [0, C) = d head weights, [C] = d head bias;
// da written; dl = device fp32 scratch of one value per voxel; probs optional
extern "C" int te_tr_head_loss(const te_tensor* a, const float* w, const float* b, const uint8_t* y, const te_tensor* da,
                               double* loss, double* dwb2c, double n_total, float* dl, float* probs, void* stream) {
  TE_TR_CHECK_T(a, "te_tr_head_loss"); TE_TR_CHECK_T(da, "te_tr_head_loss");
  TE_CHECK_ARG(w && b && y && loss && dwb2c && dl && n_total > 0 && same_grid(*a, *da) && a->C == da->C && a->dtype == da->dtype &&
                   a->C <= 1024, "te_tr_head_loss: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const TView va = view_of(*a), vd = view_of(*da);
  if (v8_ok(*a) && v8_ok(*da)) {
    const int g8 = blocks_for(va.voxels() * (va.C >> 3), 256, 4);
    const float inv = static_cast<float>(1.0 / n_total);
    if (a->dtype == TE_F16) head_loss_v8_kernel<__half><<<g8, 256, 0, st>>>(va, w, b, y, vd, loss, dwb2c, inv, dl, probs);
    else head_loss_v8_kernel<__nv_bfloat16><<<g8, 256, 0, st>>>(va, w, b, y, vd, loss, dwb2c, inv, dl, probs);
    TE_LAUNCH_CHECK();
    return TE_OK;
  }
  const size_t sh = static_cast<size_t>(a->C) * sizeof(float);
  TE_DISPATCH_T(a->dtype, (head_loss_kernel<T><<<blocks_for(va.voxels(), 256, 8), 256, sh, st>>>(
                              va, w, b, y, vd, loss, static_cast<float>(1.0 / n_total), dl, probs)));
  TE_LAUNCH_CHECK();
  int grid = blocks_for(va.voxels() * va.C, 256, 8);
  if (static_cast<long long>(grid) * 256 < va.C) grid = (va.C + 255) / 256;
  if (v8_ok(*a)) {
    const int g8 = blocks_for(va.voxels() * (va.C >> 3), 256, 8);
    if (a->dtype == TE_F16) head_wgrad_v8_kernel<__half><<<g8, 256, 0, st>>>(va, dl, dwb2c);
    else head_wgrad_v8_kernel<__nv_bfloat16><<<g8, 256, 0, st>>>(va, dl, dwb2c);
  } else {
    TE_DISPATCH_T(a->dtype, (head_wgrad_kernel<T><<<grid, 256, 0, st>>>(va, dl, dwb2c)));
  }
  TE_LAUNCH_CHECK();
  return TE_OK;
}

// ---- small per-layer "finalize" steps of the BatchNormalization forward / backward: one launch instead of a dozen
// element-wise host-framework kernels over C values.  Each consumes the fp64 sums left in `stats` by the reduction
// kernel (after the optional SyncBN all-reduce) and ZEROES them, so the buffer is ready for its next user.
namespace te {
namespace {
__global__ void bn_finalize_kernel(double* stats, int C, double inv_n, const float* __restrict__ gamma, const float* __restrict__ beta,
                                   float eps, float* mean, float* var, float* rstd, float* scale, float* shift) {
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < C; c += gridDim.x * blockDim.x) {
    const double m = stats[c] * inv_n;
    double v = stats[C + c] * inv_n - m * m;
    if (v < 0.0) v = 0.0;
    const float vf = static_cast<float>(v);
    const double r = 1.0 / sqrt(v + static_cast<double>(eps));
    const double gr = static_cast<double>(gamma[c]) * r;
    mean[c] = static_cast<float>(m); var[c] = vf; rstd[c] = static_cast<float>(r);
    scale[c] = static_cast<float>(gr); shift[c] = static_cast<float>(static_cast<double>(beta[c]) - m * gr);
    stats[c] = 0.0; stats[C + c] = 0.0;
  }
}
__global__ void bn_bwd_finalize_kernel(double* stats, const double* stats_local, int C, double inv_n, float* g_beta, float* g_gamma,
                                       float* dbeta_n, float* dgamma_n) {
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < C; c += gridDim.x * blockDim.x) {
    const double sb = stats[c], sg = stats[C + c];
    const double lb = stats_local ? stats_local[c] : sb, lg = stats_local ? stats_local[C + c] : sg;
    g_beta[c] = static_cast<float>(lb); g_gamma[c] = static_cast<float>(lg);
    dbeta_n[c] = static_cast<float>(sb * inv_n); dgamma_n[c] = static_cast<float>(sg * inv_n);
    stats[c] = 0.0; stats[C + c] = 0.0;
  }
}
__global__ void take_stats_kernel(double* stats, int n, int n_zero, float* dst) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_zero; i += gridDim.x * blockDim.x) {
    if (i < n && dst) dst[i] = static_cast<float>(stats[i]);
    stats[i] = 0.0;
  }
}
}  // namespace
}  // namespace te

extern "C" int te_tr_bn_finalize(double* stats2c, int32_t C, double n_total, const float* gamma, const float* beta, float eps,
                                 float* mean, float* var, float* rstd, float* scale, float* shift, void* stream) {
  TE_CHECK_ARG(stats2c && C > 0 && n_total > 0 && gamma && beta && mean && var && rstd && scale && shift, "te_tr_bn_finalize: bad argument");
  te::bn_finalize_kernel<<<(C + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(stats2c, C, 1.0 / n_total, gamma, beta, eps, mean,
                                                                                       var, rstd, scale, shift);
  TE_LAUNCH_CHECK();
  return TE_OK;
}

extern "C" int te_tr_bn_bwd_finalize(double* stats2c, const double* stats2c_local, int32_t C, double n_total, float* g_beta,
                                     float* g_gamma, float* dbeta_n, float* dgamma_n, void* stream) {
  TE_CHECK_ARG(stats2c && C > 0 && n_total > 0 && g_beta && g_gamma && dbeta_n && dgamma_n, "te_tr_bn_bwd_finalize: bad argument");
  te::bn_bwd_finalize_kernel<<<(C + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(stats2c, stats2c_local, C, 1.0 / n_total,
                                                                                           g_beta, g_gamma, dbeta_n, dgamma_n);
  TE_LAUNCH_CHECK();
  return TE_OK;
}

extern "C" int te_tr_take_stats(double* stats, int32_t n, int32_t n_zero, float* dst, void* stream) {
  TE_CHECK_ARG(stats && n >= 0 && n_zero >= n, "te_tr_take_stats: bad argument");
  if (n_zero == 0) return TE_OK;
  te::take_stats_kernel<<<(n_zero + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(stats, n, n_zero, dst);
  TE_LAUNCH_CHECK();
  return TE_OK;
}

extern "C" int te_tr_adam(float* w, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                          int64_t step, float grad_scale, void* stream) {
  TE_CHECK_ARG(w && g && m && v && n >= 0 && step >= 1, "te_tr_adam: bad argument");
  if (n == 0) return TE_OK;
  const double lr_t = static_cast<double>(lr) * sqrt(1.0 - pow(static_cast<double>(beta2), static_cast<double>(step))) /
                      (1.0 - pow(static_cast<double>(beta1), static_cast<double>(step)));
  adam_kernel<<<blocks_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(w, g, m, v, n, static_cast<float>(lr_t), beta1, beta2,
                                                                                  eps, grad_scale);
  TE_LAUNCH_CHECK();
  return TE_OK;
}

// ---- dense layers of the autoencoder (fp32 activations: these tensors are a few KB per sample)
extern "C" int te_tr_dense_fwd(const float* x, int64_t n, int32_t nin, int32_t nout, const float* w, const float* bias, float* y,
                               void* stream) {
  TE_CHECK_ARG(x && w && bias && y && n >= 0 && nin > 0 && nout > 0, "te_tr_dense_fwd: bad argument");
  cudaError_t e = simt_dense(x, ACT_F32, n, nin, nout, w, bias, 0, 0.f, y, ACT_F32, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) { set_error("te_tr_dense_fwd: %s", cudaGetErrorString(e)); return TE_ERR_CUDA; }
  return TE_OK;
}

extern "C" int te_tr_dense_bwd(const float* x, const float* dy, int64_t n, int32_t nin, int32_t nout, const float* w, float* dw,
                               double* dbias2c, float* dx, void* stream) {
  TE_CHECK_ARG(x && dy && w && dw && dbias2c && n >= 0 && nin > 0 && nout > 0, "te_tr_dense_bwd: bad argument");
  if (n == 0) return TE_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  dense_wgrad_kernel<<<blocks_for(static_cast<long long>(nin) * nout, 256), 256, 0, st>>>(x, dy, n, nin, nout, dw, dbias2c);
  TE_LAUNCH_CHECK();
  if (dx) {
    dense_dgrad_kernel<<<blocks_for(n * nin * 32, 256), 256, 0, st>>>(dy, w, n, nin, nout, dx);
    TE_LAUNCH_CHECK();
  }
  return TE_OK;
}

// loss3 (double [3], accumulated) = {total, mse, kl}; dwb2c as for te_tr_head_loss; target: (n_voxels) of a's dtype
extern "C" int te_tr_head_ae_loss(const te_tensor* a, const float* w, const float* b, const void* target, const te_tensor* da,
                                  double* loss3, double* dwb2c, double n_total, float weight, float* dl, float* probs, void* stream) {
  TE_TR_CHECK_T(a, "te_tr_head_ae_loss"); TE_TR_CHECK_T(da, "te_tr_head_ae_loss");
  TE_CHECK_ARG(w && b && target && loss3 && dwb2c && dl && n_total > 0 && same_grid(*a, *da) && a->C == da->C && a->dtype == da->dtype &&
                   a->C <= 1024, "te_tr_head_ae_loss: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const TView va = view_of(*a), vd = view_of(*da);
  const size_t sh = static_cast<size_t>(a->C) * sizeof(float);
  TE_DISPATCH_T(a->dtype, (head_ae_loss_kernel<T><<<blocks_for(va.voxels(), 256, 8), 256, sh, st>>>(
                              va, w, b, static_cast<const T*>(target), vd, loss3, static_cast<float>(1.0 / n_total), weight, dl, probs)));
  TE_LAUNCH_CHECK();
  if (v8_ok(*a)) {
    const int g8 = blocks_for(va.voxels() * (va.C >> 3), 256, 8);
    if (a->dtype == TE_F16) head_wgrad_v8_kernel<__half><<<g8, 256, 0, st>>>(va, dl, dwb2c);
    else head_wgrad_v8_kernel<__nv_bfloat16><<<g8, 256, 0, st>>>(va, dl, dwb2c);
  } else {
    int grid = blocks_for(va.voxels() * va.C, 256, 8);
    if (static_cast<long long>(grid) * 256 < va.C) grid = (va.C + 255) / 256;
    TE_DISPATCH_T(a->dtype, (head_wgrad_kernel<T><<<grid, 256, 0, st>>>(va, dl, dwb2c)));
  }
  TE_LAUNCH_CHECK();
  return TE_OK;
}

// CUDA-core layer kernels (see layers_simt.cuh).  Semantics follow the Keras layers the reference
// builds its graphs from (/root/reference/tomo_encoders/neural_nets/Unet3D.py:44-194,
// neural_nets/autoencoders.py:60-243): Conv3D(k=3,'same'), Conv3DTranspose(k=2, strides=s, 'same'),
// MaxPool3D(p, 'same' on divisible extents), Dense, LeakyReLU(0.2), sigmoid.
#include "layers_simt.cuh"
#include "te_common.cuh"

namespace te {
namespace {

template <typename T> __device__ __forceinline__ float ld(const T* p);
template <> __device__ __forceinline__ float ld<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ld<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <> __device__ __forceinline__ float ld<__half>(const __half* p) { return __half2float(*p); }
template <typename T> __device__ __forceinline__ void st(T* p, float v);
template <> __device__ __forceinline__ void st<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void st<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ void st<__half>(__half* p, float v) { *p = __float2half_rn(v); }

inline int blocks_for(long long items, int threads) {
  long long b = (items + threads - 1) / threads;
  const long long cap = static_cast<long long>(kNumSMs) * 32;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return static_cast<int>(b);
}

template <typename T>
__global__ void __launch_bounds__(256) conv3_kernel(TView in, TView out, const float* __restrict__ w,
                                                     const float* __restrict__ bias, int lrelu, float alpha) {
  const int Cin = in.C, Cout = out.C;
  const long long total = out.voxels() * Cout;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const T* ip = static_cast<const T*>(in.ptr);
  T* op = static_cast<T*>(out.ptr);
  for (long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < total; e += stride) {
    const int co = static_cast<int>(e % Cout);
    long long v = e / Cout;
    const int x = static_cast<int>(v % in.W); v /= in.W;
    const int y = static_cast<int>(v % in.H); v /= in.H;
    const int z = static_cast<int>(v % in.D);
    const long long n = v / in.D;
    float acc = bias[co];
    for (int dz = -1; dz <= 1; ++dz) {
      const int zz = z + dz;
      if (zz < 0 || zz >= in.D) continue;
      for (int dy = -1; dy <= 1; ++dy) {
        const int yy = y + dy;
        if (yy < 0 || yy >= in.H) continue;
        for (int dx = -1; dx <= 1; ++dx) {
          const int xx = x + dx;
          if (xx < 0 || xx >= in.W) continue;
          const int tap = ((dz + 1) * 3 + (dy + 1)) * 3 + (dx + 1);
          const T* src = ip + (((n * in.D + zz) * in.H + yy) * in.W + xx) * in.ctot + in.coff;
          const float* wt = w + (static_cast<long long>(tap) * Cin) * Cout + co;
          for (int ci = 0; ci < Cin; ++ci) acc = fmaf(ld<T>(src + ci), wt[static_cast<long long>(ci) * Cout], acc);
        }
      }
    }
    if (lrelu) acc = acc > 0.f ? acc : alpha * acc;
    st<T>(op + (e / Cout) * out.ctot + out.coff + co, acc);
  }
}

// First layer of every graph: Cin = 1.  One thread computes XV consecutive x voxels: the 3 x 3 x
// (XV + 2) input window lives in registers and every weight vector fetched from shared memory is
// reused for the XV voxels; all Cout channels are produced 16 at a time and written as 32-byte
// vectors of the 16-bit NDHWC output.
template <typename T, int XV>
__global__ void __launch_bounds__(256) conv3_c1_kernel(TView in, TView out, const float* __restrict__ w,
                                                        const float* __restrict__ bias, int lrelu, float alpha) {
  extern __shared__ float sw[];                 // [27][Cout] weights followed by [Cout] bias
  const int Cout = out.C;
  for (int i = threadIdx.x; i < 27 * Cout; i += blockDim.x) sw[i] = w[i];
  for (int i = threadIdx.x; i < Cout; i += blockDim.x) sw[27 * Cout + i] = bias[i];
  __syncthreads();
  const float* sb = sw + 27 * Cout;
  const int W = in.W, H = in.H, D = in.D;
  const int WX = W / XV;
  const long long total = static_cast<long long>(in.N) * D * H * WX;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const T* ip = static_cast<const T*>(in.ptr);
  T* op = static_cast<T*>(out.ptr);
  for (long long v = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; v < total; v += stride) {
    const int x0 = static_cast<int>(v % WX) * XV;
    long long t = v / WX;
    const int y = static_cast<int>(t % H); t /= H;
    const int z = static_cast<int>(t % D);
    const long long n = t / D;
    float val[9][XV + 2];
#pragma unroll
    for (int dz = 0; dz < 3; ++dz)
#pragma unroll
      for (int dy = 0; dy < 3; ++dy) {
        const int zz = z + dz - 1, yy = y + dy - 1;
        const bool okr = zz >= 0 && zz < D && yy >= 0 && yy < H;
        const T* row = ip + (((n * D + zz) * H + yy) * W) * in.ctot + in.coff;
#pragma unroll
        for (int dx = 0; dx < XV + 2; ++dx) {
          const int xx = x0 + dx - 1;
          val[dz * 3 + dy][dx] = (okr && xx >= 0 && xx < W) ? ld<T>(row + static_cast<long long>(xx) * in.ctot) : 0.f;
        }
      }
    T* dst = op + ((((n * D + z) * H + y) * W) + x0) * out.ctot + out.coff;
    for (int c0 = 0; c0 < Cout; c0 += 16) {
      float acc[XV][16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float bj = sb[c0 + j];
#pragma unroll
        for (int i = 0; i < XV; ++i) acc[i][j] = bj;
      }
#pragma unroll
      for (int r = 0; r < 9; ++r)
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
          const float4* wr = reinterpret_cast<const float4*>(sw + (r * 3 + dx) * Cout + c0);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 ww = wr[q];
#pragma unroll
            for (int i = 0; i < XV; ++i) {
              const float a = val[r][i + dx];
              acc[i][4 * q + 0] = fmaf(a, ww.x, acc[i][4 * q + 0]);
              acc[i][4 * q + 1] = fmaf(a, ww.y, acc[i][4 * q + 1]);
              acc[i][4 * q + 2] = fmaf(a, ww.z, acc[i][4 * q + 2]);
              acc[i][4 * q + 3] = fmaf(a, ww.w, acc[i][4 * q + 3]);
            }
          }
        }
#pragma unroll
      for (int i = 0; i < XV; ++i) {
        if (lrelu) {
#pragma unroll
          for (int j = 0; j < 16; ++j) acc[i][j] = acc[i][j] > 0.f ? acc[i][j] : alpha * acc[i][j];
        }
        uint32_t pk[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if constexpr (std::is_same<T, __half>::value) {
            __half2 h = __floats2half2_rn(acc[i][2 * j], acc[i][2 * j + 1]);
            pk[j] = *reinterpret_cast<uint32_t*>(&h);
          } else {
            __nv_bfloat162 h = __floats2bfloat162_rn(acc[i][2 * j], acc[i][2 * j + 1]);
            pk[j] = *reinterpret_cast<uint32_t*>(&h);
          }
        }
        uint4* d4 = reinterpret_cast<uint4*>(dst + static_cast<long long>(i) * out.ctot + c0);
        d4[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        d4[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
      }
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256) upconv_kernel(TView in, TView out, const float* __restrict__ w,
                                                      const float* __restrict__ bias, int s, int lrelu, float alpha) {
  const int Cin = in.C, Cout = out.C;
  const long long total = out.voxels() * Cout;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const T* ip = static_cast<const T*>(in.ptr);
  T* op = static_cast<T*>(out.ptr);
  for (long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < total; e += stride) {
    const int co = static_cast<int>(e % Cout);
    long long v = e / Cout;
    const int x = static_cast<int>(v % out.W); v /= out.W;
    const int y = static_cast<int>(v % out.H); v /= out.H;
    const int z = static_cast<int>(v % out.D);
    const long long n = v / out.D;
    const int a = z % s, b = y % s, c = x % s;
    float acc = bias[co];
    if (a < 2 && b < 2 && c < 2) {
      const int sub = (a * 2 + b) * 2 + c;
      const T* src = ip + (((n * in.D + z / s) * in.H + y / s) * in.W + x / s) * in.ctot + in.coff;
      const float* wt = w + (static_cast<long long>(sub) * Cout + co) * Cin;
      for (int ci = 0; ci < Cin; ++ci) acc = fmaf(ld<T>(src + ci), wt[ci], acc);
    }
    if (lrelu) acc = acc > 0.f ? acc : alpha * acc;
    st<T>(op + (e / Cout) * out.ctot + out.coff + co, acc);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) maxpool_kernel(TView in, TView out, int p) {
  const int C = out.C;
  const long long total = out.voxels() * C;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const T* ip = static_cast<const T*>(in.ptr);
  T* op = static_cast<T*>(out.ptr);
  for (long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < total; e += stride) {
    const int c = static_cast<int>(e % C);
    long long v = e / C;
    const int x = static_cast<int>(v % out.W); v /= out.W;
    const int y = static_cast<int>(v % out.H); v /= out.H;
    const int z = static_cast<int>(v % out.D);
    const long long n = v / out.D;
    float m = -INFINITY;
    for (int dz = 0; dz < p; ++dz)
      for (int dy = 0; dy < p; ++dy)
        for (int dx = 0; dx < p; ++dx)
          m = fmaxf(m, ld<T>(ip + (((n * in.D + z * p + dz) * in.H + y * p + dy) * in.W + x * p + dx) * in.ctot +
                             in.coff + c));
    st<T>(op + (e / C) * out.ctot + out.coff + c, m);
  }
}

// 16-bit fast path: one thread per (output voxel, 8-channel group), 128-bit loads/stores.
template <typename T2>
__global__ void __launch_bounds__(256) maxpool_vec8_kernel(TView in, TView out, int p) {
  const int G = out.C / 8;
  const long long total = out.voxels() * G;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const uint8_t* ip = static_cast<const uint8_t*>(in.ptr);
  uint8_t* op = static_cast<uint8_t*>(out.ptr);
  for (long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < total; e += stride) {
    const int g = static_cast<int>(e % G);
    long long v = e / G;
    const int x = static_cast<int>(v % out.W); v /= out.W;
    const int y = static_cast<int>(v % out.H); v /= out.H;
    const int z = static_cast<int>(v % out.D);
    const long long n = v / out.D;
    T2 m[4];
    bool first = true;
    for (int dz = 0; dz < p; ++dz)
      for (int dy = 0; dy < p; ++dy)
        for (int dx = 0; dx < p; ++dx) {
          const long long vox = ((n * in.D + z * p + dz) * in.H + y * p + dy) * in.W + x * p + dx;
          const uint4 r = *reinterpret_cast<const uint4*>(ip + (vox * in.ctot + in.coff + g * 8) * 2);
          const T2* h = reinterpret_cast<const T2*>(&r);
          if (first) { m[0] = h[0]; m[1] = h[1]; m[2] = h[2]; m[3] = h[3]; first = false; }
          else { m[0] = __hmax2(m[0], h[0]); m[1] = __hmax2(m[1], h[1]); m[2] = __hmax2(m[2], h[2]); m[3] = __hmax2(m[3], h[3]); }
        }
    *reinterpret_cast<uint4*>(op + ((e / G) * out.ctot + out.coff + g * 8) * 2) = *reinterpret_cast<uint4*>(m);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) head_kernel(TView in, const float* __restrict__ w, float b, uint8_t* labels,
                                                    float* probs, float* logits) {
  const long long total = in.voxels();
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const T* ip = static_cast<const T*>(in.ptr);
  for (long long v = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; v < total; v += stride) {
    const T* src = ip + v * in.ctot + in.coff;
    float acc = b;
    for (int c = 0; c < in.C; ++c) acc = fmaf(ld<T>(src + c), w[c], acc);
    if (labels) labels[v] = acc > 0.f ? 1 : 0;     // sigmoid(x) > 0.5 <=> x > 0; np.round(0.5) == 0
    if (probs) probs[v] = 1.f / (1.f + expf(-acc));
    if (logits) logits[v] = acc;
  }
}

// One warp per (patch n, 32 outputs): lanes own consecutive outputs j (coalesced weight rows), the
// block's 8 warps split the input range and reduce through shared memory.
template <typename TIn, typename TOut>
__global__ void __launch_bounds__(256) dense_kernel(const TIn* __restrict__ x, long long n, int nin, int nout,
                                                     const float* __restrict__ w, const float* __restrict__ b, int lrelu,
                                                     float alpha, TOut* __restrict__ out) {
  // block = 8 rows x 32 outputs; its 8 warps split the input range; every weight fetched (coalesced over
  // the lanes) feeds 8 independent accumulators, and the loop is unrolled 4 x for memory-level parallelism
  constexpr int R = 8;
  __shared__ float part[8][R][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row0 = static_cast<long long>(blockIdx.y) * R;
  const int j = blockIdx.x * 32 + lane;
  const int chunk = (nin + 7) / 8;
  const int i0 = warp * chunk, i1 = min(nin, i0 + chunk);
  float acc[R];
#pragma unroll
  for (int r = 0; r < R; ++r) acc[r] = 0.f;
  const int jj = j < nout ? j : nout - 1;
  int i = i0;
  for (; i + 4 <= i1; i += 4) {
    float wv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) wv[u] = __ldg(w + static_cast<long long>(i + u) * nout + jj);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      if (row0 + r < n) {
        const TIn* xr = x + (row0 + r) * nin + i;
#pragma unroll
        for (int u = 0; u < 4; ++u) acc[r] = fmaf(ld<TIn>(xr + u), wv[u], acc[r]);
      }
    }
  }
  for (; i < i1; ++i) {
    const float wv = __ldg(w + static_cast<long long>(i) * nout + jj);
#pragma unroll
    for (int r = 0; r < R; ++r)
      if (row0 + r < n) acc[r] = fmaf(ld<TIn>(x + (row0 + r) * nin + i), wv, acc[r]);
  }
#pragma unroll
  for (int r = 0; r < R; ++r) part[warp][r][lane] = acc[r];
  __syncthreads();
  if (warp < R && j < nout && row0 + warp < n) {
    float sum = b[j];
#pragma unroll
    for (int k = 0; k < 8; ++k) sum += part[k][warp][lane];
    if (lrelu) sum = sum > 0.f ? sum : alpha * sum;
    st<TOut>(out + (row0 + warp) * nout + j, sum);
  }
}

// Split-K variant for long input vectors (the 8192 -> 128 layer behind Flatten): grid.z blocks share the
// input range and add their partial sums into a zeroed fp32 scratch; dense_finish applies bias / activation.
template <typename TIn>
__global__ void __launch_bounds__(256) dense_splitk_kernel(const TIn* __restrict__ x, long long n, int nin, int nout,
                                                            const float* __restrict__ w, float* __restrict__ acc_out, int kper) {
  constexpr int R = 8;
  __shared__ float part[8][R][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row0 = static_cast<long long>(blockIdx.y) * R;
  const int j = blockIdx.x * 32 + lane;
  const int k0 = blockIdx.z * kper, k1 = min(nin, k0 + kper);
  const int chunk = (k1 - k0 + 7) / 8;
  const int i0 = k0 + warp * chunk, i1 = min(k1, i0 + chunk);
  float acc[R];
#pragma unroll
  for (int r = 0; r < R; ++r) acc[r] = 0.f;
  const int jj = j < nout ? j : nout - 1;
  int i = i0;
  for (; i + 8 <= i1; i += 8) {
    float wv[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) wv[u] = __ldg(w + static_cast<long long>(i + u) * nout + jj);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      if (row0 + r < n) {
        const TIn* xr = x + (row0 + r) * nin + i;
#pragma unroll
        for (int u = 0; u < 8; ++u) acc[r] = fmaf(ld<TIn>(xr + u), wv[u], acc[r]);
      }
    }
  }
  for (; i < i1; ++i) {
    const float wv = __ldg(w + static_cast<long long>(i) * nout + jj);
#pragma unroll
    for (int r = 0; r < R; ++r)
      if (row0 + r < n) acc[r] = fmaf(ld<TIn>(x + (row0 + r) * nin + i), wv, acc[r]);
  }
#pragma unroll
  for (int r = 0; r < R; ++r) part[warp][r][lane] = acc[r];
  __syncthreads();
  if (warp < R && j < nout && row0 + warp < n) {
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) sum += part[k][warp][lane];
    atomicAdd(acc_out + (row0 + warp) * nout + j, sum);
  }
}

template <typename TOut>
__global__ void __launch_bounds__(256) dense_finish_kernel(const float* __restrict__ acc, const float* __restrict__ b, long long n,
                                                            int nout, int lrelu, float alpha, TOut* __restrict__ out) {
  const long long total = n * nout;
  for (long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < total;
       e += static_cast<long long>(gridDim.x) * blockDim.x) {
    float sum = acc[e] + b[e % nout];
    if (lrelu) sum = sum > 0.f ? sum : alpha * sum;
    st<TOut>(out + e, sum);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) to_f32_kernel(TView in, float* out) {
  const long long total = in.voxels() * in.C;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const T* ip = static_cast<const T*>(in.ptr);
  for (long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < total; e += stride)
    out[e] = ld<T>(ip + (e / in.C) * in.ctot + in.coff + (e % in.C));
}
template <typename T>
__global__ void __launch_bounds__(256) from_f32_kernel(const float* in, T* out, long long count) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < count; e += stride)
    st<T>(out + e, in[e]);
}

}  // namespace

#define TE_DISPATCH_ACT(act_type, CALL)                      \
  switch (act_type) {                                        \
    case ACT_F32: { using T = float; CALL; } break;          \
    case ACT_BF16: { using T = __nv_bfloat16; CALL; } break; \
    case ACT_F16: { using T = __half; CALL; } break;         \
    default: return cudaErrorInvalidValue;                   \
  }

cudaError_t simt_conv3(const TView& in, const TView& out, const float* w, const float* bias, int lrelu, float alpha,
                       int act_type, cudaStream_t st, int* overflow) {
  {
    cudaError_t e = cudaSuccess;
    if (conv1_tz_try(in, out, w, bias, lrelu, alpha, act_type, st, &e, overflow)) { count_launch(); return e; }
    if (conv1_umma_try(in, out, w, bias, lrelu, alpha, act_type, st, &e, overflow)) { count_launch(); return e; }
  }
  if (in.C == 1 && act_type != ACT_F32 && out.C % 16 == 0 && out.ctot % 8 == 0 && out.coff % 8 == 0 && out.C <= 256) {
    const size_t smem = static_cast<size_t>(28) * out.C * sizeof(float);
    if (in.W % 4 == 0) {
      const int grid = blocks_for(out.voxels() / 4, 128);
      if (act_type == ACT_BF16) conv3_c1_kernel<__nv_bfloat16, 4><<<grid, 128, smem, st>>>(in, out, w, bias, lrelu, alpha);
      else conv3_c1_kernel<__half, 4><<<grid, 128, smem, st>>>(in, out, w, bias, lrelu, alpha);
    } else {
      const int grid = blocks_for(out.voxels(), 256);
      if (act_type == ACT_BF16) conv3_c1_kernel<__nv_bfloat16, 1><<<grid, 256, smem, st>>>(in, out, w, bias, lrelu, alpha);
      else conv3_c1_kernel<__half, 1><<<grid, 256, smem, st>>>(in, out, w, bias, lrelu, alpha);
    }
    count_launch();
    return cudaGetLastError();
  }
  const int grid = blocks_for(out.voxels() * out.C, 256);
  TE_DISPATCH_ACT(act_type, (conv3_kernel<T><<<grid, 256, 0, st>>>(in, out, w, bias, lrelu, alpha)));
  count_launch();
  return cudaGetLastError();
}

cudaError_t simt_upconv(const TView& in, const TView& out, const float* w, const float* bias, int stride, int lrelu,
                        float alpha, int act_type, cudaStream_t st) {
  const int grid = blocks_for(out.voxels() * out.C, 256);
  TE_DISPATCH_ACT(act_type, (upconv_kernel<T><<<grid, 256, 0, st>>>(in, out, w, bias, stride, lrelu, alpha)));
  count_launch();
  return cudaGetLastError();
}

cudaError_t simt_maxpool(const TView& in, const TView& out, int pool, int act_type, cudaStream_t st) {
  const bool vec = act_type != ACT_F32 && out.C % 8 == 0 && in.ctot % 8 == 0 && in.coff % 8 == 0 &&
                   out.ctot % 8 == 0 && out.coff % 8 == 0;
  if (vec) {
    const int grid = blocks_for(out.voxels() * (out.C / 8), 256);
    if (act_type == ACT_BF16) maxpool_vec8_kernel<__nv_bfloat162><<<grid, 256, 0, st>>>(in, out, pool);
    else maxpool_vec8_kernel<__half2><<<grid, 256, 0, st>>>(in, out, pool);
  } else {
    const int grid = blocks_for(out.voxels() * out.C, 256);
    TE_DISPATCH_ACT(act_type, (maxpool_kernel<T><<<grid, 256, 0, st>>>(in, out, pool)));
  }
  count_launch();
  return cudaGetLastError();
}

cudaError_t simt_head(const TView& in, const float* w, float b, uint8_t* labels, float* probs, float* logits,
                      int act_type, cudaStream_t st) {
  const int grid = blocks_for(in.voxels(), 256);
  TE_DISPATCH_ACT(act_type, (head_kernel<T><<<grid, 256, 0, st>>>(in, w, b, labels, probs, logits)));
  count_launch();
  return cudaGetLastError();
}

cudaError_t simt_dense(const void* x, int in_type, long long n, int nin, int nout, const float* w, const float* b,
                       int lrelu, float alpha, void* out, int out_type, cudaStream_t st, float* acc_scratch) {
  if (n <= 0) return cudaSuccess;
  if (acc_scratch != nullptr && nin >= 2048) {
    const int kper = 512, ks = (nin + kper - 1) / kper;
    cudaError_t e = cudaMemsetAsync(acc_scratch, 0, static_cast<size_t>(n) * nout * sizeof(float), st);
    if (e != cudaSuccess) return e;
    dim3 g3((nout + 31) / 32, static_cast<unsigned>((n + 7) / 8), ks);
    if (in_type == ACT_F32) dense_splitk_kernel<float><<<g3, 256, 0, st>>>(static_cast<const float*>(x), n, nin, nout, w, acc_scratch, kper);
    else if (in_type == ACT_BF16) dense_splitk_kernel<__nv_bfloat16><<<g3, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(x), n, nin, nout, w, acc_scratch, kper);
    else dense_splitk_kernel<__half><<<g3, 256, 0, st>>>(static_cast<const __half*>(x), n, nin, nout, w, acc_scratch, kper);
    count_launch();
    const int gf = blocks_for(n * nout, 256);
    if (out_type == ACT_F32) dense_finish_kernel<float><<<gf, 256, 0, st>>>(acc_scratch, b, n, nout, lrelu, alpha, static_cast<float*>(out));
    else if (out_type == ACT_BF16) dense_finish_kernel<__nv_bfloat16><<<gf, 256, 0, st>>>(acc_scratch, b, n, nout, lrelu, alpha, static_cast<__nv_bfloat16*>(out));
    else dense_finish_kernel<__half><<<gf, 256, 0, st>>>(acc_scratch, b, n, nout, lrelu, alpha, static_cast<__half*>(out));
    count_launch();
    return cudaGetLastError();
  }
  dim3 grid((nout + 31) / 32, static_cast<unsigned>((n + 7) / 8));
#define TE_DENSE(TI, TO) \
  dense_kernel<TI, TO><<<grid, 256, 0, st>>>(static_cast<const TI*>(x), n, nin, nout, w, b, lrelu, alpha, static_cast<TO*>(out))
  if (in_type == ACT_F32 && out_type == ACT_F32) TE_DENSE(float, float);
  else if (in_type == ACT_F32 && out_type == ACT_BF16) TE_DENSE(float, __nv_bfloat16);
  else if (in_type == ACT_F32 && out_type == ACT_F16) TE_DENSE(float, __half);
  else if (in_type == ACT_BF16 && out_type == ACT_F32) TE_DENSE(__nv_bfloat16, float);
  else if (in_type == ACT_F16 && out_type == ACT_F32) TE_DENSE(__half, float);
  else return cudaErrorInvalidValue;
#undef TE_DENSE
  count_launch();
  return cudaGetLastError();
}

cudaError_t simt_to_f32(const TView& in, float* out, int act_type, cudaStream_t st) {
  const int grid = blocks_for(in.voxels() * in.C, 256);
  TE_DISPATCH_ACT(act_type, (to_f32_kernel<T><<<grid, 256, 0, st>>>(in, out)));
  count_launch();
  return cudaGetLastError();
}

cudaError_t simt_from_f32(const float* in, void* out, long long count, int act_type, cudaStream_t st) {
  const int grid = blocks_for(count, 256);
  TE_DISPATCH_ACT(act_type, (from_f32_kernel<T><<<grid, 256, 0, st>>>(in, static_cast<T*>(out), count)));
  count_launch();
  return cudaGetLastError();
}

}  // namespace te

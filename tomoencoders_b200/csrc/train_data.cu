// Training data generator of the segmenter on the GPU.
//
// Reference (paths relative to /root/reference/tomo_encoders/):
//   _find_blanks                  neural_nets/surface_segmenter.py:75-82   per-patch np.std of the binned label patches
//   extract_training_patch_pairs  neural_nets/keras_processor.py:213-232   binned gather of x and y, per-patch Gaussian
//                                                                          noise N(0, sigma_i), np.rot90(k_i, axes_i)
// te_patch_std: one block per patch, population standard deviation (ddof = 0) from float64 sums.
// te_training_pairs: ONE kernel writes the rotated, noised x (float32) and the rotated y (uint8) of every patch: the
// thread of an OUTPUT voxel inverts the rotation to find its source voxel (so the writes stay coalesced), decimates
// by the patch's binning, and adds sigma_i * N(0, 1) drawn from Philox4x32-10 keyed by (seed, patch, voxel).  The noise
// field is i.i.d., so drawing it at the output position instead of before the rotation is the same distribution; the
// deterministic part (gather + rot90) is bit-exact against numpy.
#include "te_common.cuh"

namespace te {
namespace {

template <typename T> __device__ __forceinline__ float ld_f32(const void* base, long long i);
template <> __device__ __forceinline__ float ld_f32<uint8_t>(const void* b, long long i) { return static_cast<float>(static_cast<const uint8_t*>(b)[i]); }
template <> __device__ __forceinline__ float ld_f32<uint16_t>(const void* b, long long i) { return static_cast<float>(static_cast<const uint16_t*>(b)[i]); }
template <> __device__ __forceinline__ float ld_f32<__half>(const void* b, long long i) { return __half2float(static_cast<const __half*>(b)[i]); }
template <> __device__ __forceinline__ float ld_f32<float>(const void* b, long long i) { return static_cast<const float*>(b)[i]; }

struct PatchTable {
  const uint32_t* points;   // (n, 3) z, y, x
  const uint32_t* widths;   // (n, 3)
  long long n;
  long long Y, X;           // volume extents
  int P0, P1, P2;           // patch extents
};

template <typename T>
__global__ void __launch_bounds__(256) patch_std_kernel(const void* vol, PatchTable t, float* __restrict__ out) {
  const long long i = blockIdx.x;
  const uint32_t z0 = t.points[3 * i], y0 = t.points[3 * i + 1], x0 = t.points[3 * i + 2];
  const uint32_t b = t.widths[3 * i] / static_cast<uint32_t>(t.P0);
  const int rows = t.P0 * t.P1;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double s = 0.0, s2 = 0.0;
  for (int r = warp; r < rows; r += 8) {
    const int iz = r / t.P1, iy = r - iz * t.P1;
    const long long row = ((static_cast<long long>(z0) + static_cast<long long>(iz) * b) * t.Y + y0 + static_cast<long long>(iy) * b) * t.X + x0;
    for (int ix = lane; ix < t.P2; ix += 32) {
      const double v = ld_f32<T>(vol, row + static_cast<long long>(ix) * b);
      s += v;
      s2 += v * v;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  __shared__ double sh[2][8];
  if (lane == 0) { sh[0][warp] = s; sh[1][warp] = s2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) { s += sh[0][w]; s2 += sh[1][w]; }
    const double cnt = static_cast<double>(rows) * t.P2;
    const double mean = s / cnt;
    double var = s2 / cnt - mean * mean;
    if (var < 0.0) var = 0.0;
    out[i] = static_cast<float>(sqrt(var));
  }
}

// Philox4x32-10 (Salmon et al., SC'11): counter (c0..c3), key (k0, k1)
__device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
    c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}

struct PairParams {
  const void* X;
  const void* Y;
  PatchTable t;
  const float* sigma;      // (n) noise standard deviation per patch (nullptr = no noise)
  const int32_t* nrots;    // (n) k of np.rot90 (nullptr = no rotation)
  const int32_t* axes;     // (n, 2) ordered axis pair of np.rot90
  unsigned long long seed;
  float* x_out;
  uint8_t* y_out;
};

template <typename TX, typename TY>
__global__ void __launch_bounds__(256) training_pairs_kernel(PairParams p) {
  const long long per = static_cast<long long>(p.t.P0) * p.t.P1 * p.t.P2;
  const long long total = p.t.n * per;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < total; e += stride) {
    const long long i = e / per;
    const unsigned r = static_cast<unsigned>(e - i * per);
    int o[3];
    o[2] = static_cast<int>(r % static_cast<unsigned>(p.t.P2));
    const unsigned r1 = r / static_cast<unsigned>(p.t.P2);
    o[1] = static_cast<int>(r1 % static_cast<unsigned>(p.t.P1));
    o[0] = static_cast<int>(r1 / static_cast<unsigned>(p.t.P1));
    int s[3] = {o[0], o[1], o[2]};
    const int k = p.nrots ? (p.nrots[i] & 3) : 0;
    if (k) {
      // np.rot90(m, k, axes=(a, b)) on a cubic patch: k = 1: out[p, q] = m[q, P-1-p]; k = 2: m[P-1-p, P-1-q];
      // k = 3: m[P-1-q, p]  (p, q = the output indices along a and b)
      const int a = p.axes[2 * i], bb = p.axes[2 * i + 1];
      const int P = p.t.P0;
      const int pa = o[a], qb = o[bb];
      if (k == 1) { s[a] = qb; s[bb] = P - 1 - pa; }
      else if (k == 2) { s[a] = P - 1 - pa; s[bb] = P - 1 - qb; }
      else { s[a] = P - 1 - qb; s[bb] = pa; }
    }
    const uint32_t b = p.t.widths[3 * i] / static_cast<uint32_t>(p.t.P0);
    const long long src = ((static_cast<long long>(p.t.points[3 * i]) + static_cast<long long>(s[0]) * b) * p.t.Y +
                           p.t.points[3 * i + 1] + static_cast<long long>(s[1]) * b) * p.t.X +
                          p.t.points[3 * i + 2] + static_cast<long long>(s[2]) * b;
    float x = ld_f32<TX>(p.X, src);
    if (p.sigma != nullptr) {
      const float sg = p.sigma[i];
      if (sg != 0.f) {
        uint32_t c[4] = {r, static_cast<uint32_t>(i), static_cast<uint32_t>(i >> 32), 0x746f6d6fu};
        philox4x32_10(c, static_cast<uint32_t>(p.seed), static_cast<uint32_t>(p.seed >> 32));
        // Box-Muller on two uniforms in (0, 1]
        const float u1 = (static_cast<float>(c[0] >> 8) + 1.0f) * (1.0f / 16777216.0f);
        const float u2 = static_cast<float>(c[1] >> 8) * (1.0f / 16777216.0f);
        x += sg * sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
      }
    }
    p.x_out[e] = x;
    p.y_out[e] = static_cast<uint8_t>(ld_f32<TY>(p.Y, src));
  }
}

int grid_for(long long items) {
  long long blocks = (items + 255) / 256;
  const long long cap = 16LL * kNumSMs;
  if (blocks > cap) blocks = cap;
  return static_cast<int>(blocks < 1 ? 1 : blocks);
}

template <typename TX>
int launch_pairs_y(const PairParams& p, int y_dtype, cudaStream_t st) {
  const int grid = grid_for(p.t.n * p.t.P0 * p.t.P1 * p.t.P2);
  switch (y_dtype) {
    case TE_U8: training_pairs_kernel<TX, uint8_t><<<grid, 256, 0, st>>>(p); break;
    case TE_U16: training_pairs_kernel<TX, uint16_t><<<grid, 256, 0, st>>>(p); break;
    case TE_F16: training_pairs_kernel<TX, __half><<<grid, 256, 0, st>>>(p); break;
    case TE_F32: training_pairs_kernel<TX, float><<<grid, 256, 0, st>>>(p); break;
    default: set_error("te_training_pairs: unsupported y_dtype %d", y_dtype); return TE_ERR_ARG;
  }
  TE_LAUNCH_CHECK();
  return TE_OK;
}

}  // namespace
}  // namespace te

using namespace te;

extern "C" int te_patch_std(const void* vol, int vol_dtype, const int64_t vol_shape[3], const uint32_t* points,
                            const uint32_t* widths, int64_t n, const int32_t patch[3], float* out_std, void* stream) {
  TE_CHECK_ARG(n >= 0 && n < (1LL << 31), "te_patch_std: bad n");
  if (n == 0) return TE_OK;
  TE_CHECK_ARG(vol && points && widths && out_std, "te_patch_std: null pointer");
  TE_CHECK_ARG(patch[0] > 0 && patch[1] > 0 && patch[2] > 0, "te_patch_std: patch extents must be positive");
  PatchTable t{points, widths, n, vol_shape[1], vol_shape[2], patch[0], patch[1], patch[2]};
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const unsigned g = static_cast<unsigned>(n);
  switch (vol_dtype) {
    case TE_U8: patch_std_kernel<uint8_t><<<g, 256, 0, st>>>(vol, t, out_std); break;
    case TE_U16: patch_std_kernel<uint16_t><<<g, 256, 0, st>>>(vol, t, out_std); break;
    case TE_F16: patch_std_kernel<__half><<<g, 256, 0, st>>>(vol, t, out_std); break;
    case TE_F32: patch_std_kernel<float><<<g, 256, 0, st>>>(vol, t, out_std); break;
    default: set_error("te_patch_std: unsupported vol_dtype %d", vol_dtype); return TE_ERR_ARG;
  }
  TE_LAUNCH_CHECK();
  return TE_OK;
}

extern "C" int te_training_pairs(const void* X, int x_dtype, const void* Y, int y_dtype, const int64_t vol_shape[3],
                                 const uint32_t* points, const uint32_t* widths, int64_t n, const int32_t patch[3],
                                 const float* sigma, const int32_t* nrots, const int32_t* axes, uint64_t seed,
                                 float* x_out, uint8_t* y_out, void* stream) {
  TE_CHECK_ARG(n >= 0, "te_training_pairs: n < 0");
  if (n == 0) return TE_OK;
  TE_CHECK_ARG(X && Y && points && widths && x_out && y_out, "te_training_pairs: null pointer");
  TE_CHECK_ARG(patch[0] > 0 && patch[1] > 0 && patch[2] > 0, "te_training_pairs: patch extents must be positive");
  TE_CHECK_ARG(nrots == nullptr || axes != nullptr, "te_training_pairs: nrots needs axes");
  TE_CHECK_ARG(nrots == nullptr || (patch[0] == patch[1] && patch[1] == patch[2]),
               "te_training_pairs: rot90 augmentation needs cubic patches (np.rot90 would change the shape otherwise)");
  TE_CHECK_ARG(static_cast<long long>(patch[0]) * patch[1] * patch[2] < (1LL << 32), "te_training_pairs: patch too large");
  PairParams p{X, Y, PatchTable{points, widths, n, vol_shape[1], vol_shape[2], patch[0], patch[1], patch[2]},
               sigma, nrots, axes, seed, x_out, y_out};
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (x_dtype) {
    case TE_U8: return launch_pairs_y<uint8_t>(p, y_dtype, st);
    case TE_U16: return launch_pairs_y<uint16_t>(p, y_dtype, st);
    case TE_F16: return launch_pairs_y<__half>(p, y_dtype, st);
    case TE_F32: return launch_pairs_y<float>(p, y_dtype, st);
    default: set_error("te_training_pairs: unsupported x_dtype %d", x_dtype); return TE_ERR_ARG;
  }
}

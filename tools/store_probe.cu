// Store-path probe for the conv epilogues: how fast can 148 persistent CTAs write an NDHWC activation tensor when every
// epilogue warp owns 32 voxels (8 x * 4 y of one z plane, like conv_umma's TMEM lane quarters) and writes C channels of
// 16 bits per voxel?  No compute, no TMEM: staging buffers are filled once, then only the store path runs.
//   mode 0: TMA tensor box stores (C, 8, 4, 1, 1), one per warp and tile plane (the conv_umma epilogue of today)
//   mode 1: the same bytes with per-thread 128-bit global stores, 8 lanes per voxel row (C = 64) / 4 lanes (C = 32): a
//           warp instruction covers 4 (8) complete rows
//   mode 2: TMA box stores of (C, 8, 16, 1, 1): one per CTA and plane (128 rows per instruction) from one elected thread
// `xs` = 1: dense conv output (rows of C channels contiguous along x); 2: the sub-position pattern of a transposed conv
// with stride 2 (every second voxel in x, y, z: rows of 2 C bytes separated by gaps).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o store_probe store_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct P {
  uint8_t* out;
  int C, W, H, D, N, xs;      // logical tensor written by this launch: (N, D, H, W) voxels of C channels; xs = voxel stride in the buffer
  long long tiles;            // (N, D / TZ, H / 16, W / 8) tiles of TZ planes
  int TZ;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

template <int MODE>
__global__ void __launch_bounds__(256) store_kernel(const __grid_constant__ CUtensorMap tm, const P p) {
  extern __shared__ __align__(1024) uint8_t sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rowb = p.C * 2;
  for (int i = threadIdx.x; i < 128 * rowb / 16; i += 256) reinterpret_cast<uint4*>(sm)[i] = make_uint4(i, 2, 3, 4);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  const int tx = p.W / 8, ty = p.H / 16, tz = p.D / p.TZ;
  const int q = warp & 3, eg = warp >> 2;                       // two groups of four warps: each takes half of the planes
  for (long long t = blockIdx.x; t < p.tiles; t += gridDim.x) {
    unsigned u = (unsigned)t;
    const int x0 = (u % tx) * 8; u /= tx;
    const int y0 = (u % ty) * 16; u /= ty;
    const int z0 = (u % tz) * p.TZ; u /= tz;
    const int n = u;
    for (int pz = eg * (p.TZ / 2); pz < (eg + 1) * (p.TZ / 2); ++pz) {
      if (MODE == 0) {
        if (lane == 0) {
          asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];" ::"l"(
                           reinterpret_cast<uint64_t>(&tm)),
                       "r"(smem_u32(sm + q * 32 * rowb)), "r"(0), "r"(x0), "r"(y0 + 4 * q), "r"(z0 + pz), "r"(n)
                       : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        }
        __syncwarp();
      } else if (MODE == 2) {
        if (q == 0 && lane == 0) {
          asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];" ::"l"(
                           reinterpret_cast<uint64_t>(&tm)),
                       "r"(smem_u32(sm)), "r"(0), "r"(x0), "r"(y0), "r"(z0 + pz), "r"(n)
                       : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        }
        __syncwarp();
      } else {
        const int lpr = rowb / 16;                               // lanes per voxel row
        const int rpi = 32 / lpr;                                // rows per warp instruction
        const long long vstride = (long long)p.xs * rowb;        // bytes between consecutive voxels along x in the buffer
        for (int r0 = 0; r0 < 32; r0 += rpi) {
          const int r = r0 + lane / lpr, c = lane % lpr;
          const int vy = y0 + 4 * q + (r >> 3), vx = x0 + (r & 7);
          const long long vox = (((long long)n * p.D * p.xs + (long long)(z0 + pz) * p.xs) * (p.H * p.xs) + (long long)vy * p.xs) * (p.W * p.xs) + (long long)vx * p.xs;
          const uint4 v = *reinterpret_cast<const uint4*>(sm + (q * 32 + r) * rowb + c * 16);
          *reinterpret_cast<uint4*>(p.out + vox * rowb + c * 16) = v;
          (void)vstride;
        }
      }
    }
  }
  if (MODE != 1 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main() {
  EncodeTiledFn enc = nullptr;
  {
    cudaDriverEntryPointQueryResult q;
    void* ptr = nullptr;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q));
    enc = (EncodeTiledFn)ptr;
  }
  const size_t nbytes = 2ull << 30;
  uint8_t* buf; CK(cudaMalloc(&buf, nbytes));
  uint8_t* fl; CK(cudaMalloc(&fl, 512 << 20));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  struct Case { const char* name; int C, W, H, D, N, xs, TZ; };
  const Case cases[] = {
      {"dense 64^3 x 32 patches, C = 32 (enc0.conv2 output)", 32, 64, 64, 64, 32, 1, 8},
      {"dense 64^3 x 32 patches, C = 64", 64, 64, 64, 64, 32, 1, 4},
      {"dense 64^3 x 32 patches, C = 16 (conv1 output)", 16, 64, 64, 64, 32, 1, 8},
      {"stride-2 sub-position, 32^3 -> 64^3, C = 64 (1/8 of dec0.upconv)", 64, 32, 32, 32, 32, 2, 4},
  };
  for (const Case& c : cases) {
    P p;
    p.out = buf; p.C = c.C; p.W = c.W; p.H = c.H; p.D = c.D; p.N = c.N; p.xs = c.xs; p.TZ = c.TZ;
    p.tiles = (long long)c.N * (c.D / c.TZ) * (c.H / 16) * (c.W / 8);
    const double bytes = (double)c.N * c.D * c.H * c.W * c.C * 2;
    for (int mode = 0; mode < 3; ++mode) {
      CUtensorMap tm;
      const long long s = c.xs, ct = c.C;
      cuuint64_t gdim[5] = {(cuuint64_t)c.C, (cuuint64_t)c.W, (cuuint64_t)c.H, (cuuint64_t)c.D, (cuuint64_t)c.N};
      cuuint64_t gstr[4] = {(cuuint64_t)(s * ct * 2), (cuuint64_t)(s * c.W * s * ct * 2), (cuuint64_t)(s * c.H * s * c.W * s * ct * 2),
                            (cuuint64_t)(c.D * s * c.H * s * c.W * s * ct * 2)};
      cuuint32_t box[5] = {(cuuint32_t)c.C, 8, (cuuint32_t)(mode == 2 ? 16 : 4), 1, 1};
      cuuint32_t estr[5] = {1, 1, 1, 1, 1};
      const CUtensorMapSwizzle sw = c.C * 2 == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (c.C * 2 == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
      if (enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, buf, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode failed\n"); return 1; }
      float best = 1e9;
      for (int r = 0; r < 4; ++r) {
        cudaMemset(fl, r, 512 << 20);
        cudaEventRecord(e0);
        const int smem = 128 * c.C * 2 + 1024;
        if (mode == 0) store_kernel<0><<<148, 256, smem>>>(tm, p);
        else if (mode == 1) store_kernel<1><<<148, 256, smem>>>(tm, p);
        else store_kernel<2><<<148, 256, smem>>>(tm, p);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
      }
      cudaError_t e = cudaGetLastError();
      printf("%-68s mode %d: %8.1f us %7.0f GB/s %s\n", c.name, mode, best * 1e3, bytes / best / 1e6, e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
  }
  return 0;
}

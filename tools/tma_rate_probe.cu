// Hardware probe (not product code): throughput of 5-D TMA box loads shaped like the conv kernel's
// haloed input boxes (CB channels x 10 x 18 x 6 voxels) from an L2-resident tensor, one CTA per SM,
// as a function of the innermost row size (CB * 2 bytes) and of the number of loads kept in flight.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cuda_fp16.h>
#include "../tomoencoders_b200/csrc/sm100_ptx.cuh"
using namespace te;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(1);} } while (0)

struct P { int iters, inflight, stage_bytes, box_bytes, nz, ny, nx; };

__global__ void __launch_bounds__(32, 1) rate(const __grid_constant__ CUtensorMap tmap, P p, long long* cyc) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 4 * p.stage_bytes);
  if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) mbar_init(bar + i, 1); fence_barrier_init(); }
  __syncwarp();
  const int b = blockIdx.x;
  long long t0 = clock64();
  int issued = 0;
  for (; issued < p.inflight && issued < p.iters; ++issued) {
    if (elect_one()) {
      mbar_expect_tx(bar + issued % 4, p.box_bytes);
      tma_load_5d(smem + (issued % 4) * p.stage_bytes, &tmap, bar + issued % 4, 0, ((b + issued) * 8) % p.nx - 1,
                  ((b * 3 + issued) * 16) % p.ny - 1, ((b + issued * 5) * 4) % p.nz - 1, (b + issued) % 4);
    }
    __syncwarp();
  }
  for (int k = 0; k < p.iters; ++k) {
    mbar_wait(bar + k % 4, (k / 4) & 1);
    if (issued < p.iters) {
      if (elect_one()) {
        mbar_expect_tx(bar + issued % 4, p.box_bytes);
        tma_load_5d(smem + (issued % 4) * p.stage_bytes, &tmap, bar + issued % 4, 0, ((b + issued) * 8) % p.nx - 1,
                    ((b * 3 + issued) * 16) % p.ny - 1, ((b + issued * 5) * 4) % p.nz - 1, (b + issued) % 4);
      }
      __syncwarp();
      ++issued;
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  EncodeTiledFn enc = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &q));
  const int N = 4, Z = 64, Y = 64, X = 64;
  long long* d_cyc; CK(cudaMalloc(&d_cyc, 148 * 8));
  CK(cudaFuncSetAttribute(rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
  for (int C : {16, 32, 64, 128}) {
    // tensor (N, Z, Y, X, C) fp16; box = (CBOX, 10, 18, 6, 1) with CBOX = min(C, 64)
    for (int cbox : {16, 32, 64}) {
      if (cbox > C) continue;
      __half* dx; size_t bytes = (size_t)N * Z * Y * X * C * 2;
      CK(cudaMalloc(&dx, bytes)); CK(cudaMemset(dx, 0, bytes));
      CUtensorMap tm;
      cuuint64_t gdim[5] = {(cuuint64_t)C, X, Y, Z, N};
      cuuint64_t gstr[4] = {(cuuint64_t)C * 2, (cuuint64_t)X * C * 2, (cuuint64_t)Y * X * C * 2, (cuuint64_t)Z * Y * X * C * 2};
      cuuint32_t box[5] = {(cuuint32_t)cbox, 10, 18, 6, 1};
      cuuint32_t estr[5] = {1, 1, 1, 1, 1};
      CUtensorMapSwizzle sw = cbox == 16 ? CU_TENSOR_MAP_SWIZZLE_32B : cbox == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B;
      CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, dx, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
      const int box_bytes = cbox * 2 * 10 * 18 * 6;
      const int stage = (box_bytes + 1023) / 1024 * 1024;
      for (int inflight : {1, 2, 3}) {
        if ((size_t)stage * 4 > 210 * 1024) continue;
        P p{200, inflight, stage, box_bytes, Z, Y, X};
        rate<<<148, 32, 4 * stage + 2048>>>(tm, p, d_cyc);   // warm (L2)
        rate<<<148, 32, 4 * stage + 2048>>>(tm, p, d_cyc);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("FAILED %s\n", cudaGetErrorString(e)); return 2; }
        std::vector<long long> h(148); CK(cudaMemcpy(h.data(), d_cyc, 148 * 8, cudaMemcpyDeviceToHost));
        double avg = 0; for (auto v : h) avg += v; avg /= 148;
        printf("C=%3d box inner=%3d B (rows of %3d B, %d rows, %6d B/box) inflight=%d : %8.0f clk/box  %5.1f B/clk/SM  %5.1f clk/row\n",
               C, cbox * 2, cbox * 2, 1080, box_bytes, inflight, avg / 200, box_bytes / (avg / 200), avg / 200 / 1080);
      }
      CK(cudaFree(dx));
    }
  }
  printf("probe done\n");
  return 0;
}

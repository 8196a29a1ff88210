// Hardware probe (not product code): pins down the MN-major ("transposed") tcgen05 operand
// descriptors the tensor-core weight-gradient kernel relies on, on a real B200.
//
// Weight gradient of a 3x3x3 conv on NDHWC tensors: dW[tap][ci][co] = sum_v X[v+tap][ci] dY[v][co].
// The GEMM K dimension is the voxel index, i.e. the SLOW dimension of both operands in shared memory
// (rows = voxels, 32/64-byte rows of channels) -> both operands are MN-major.
//   A = dY plane box, rows (y, x') with x' extent XE, 32 channels (64-byte rows, SWIZZLE_64B);
//       M = 128 = 4 atoms of 32 channels, atom j = the box shifted by j rows (LBO = 64 B) -> the
//       x-taps ride in M.
//   B = X plane box, rows (y, x) with x extent 8, CB channels (SW64 for CB = 32, SW32 for CB = 16);
//       N = 3 atoms of CB channels, atom i = the box shifted by i y-rows (LBO = 8 rows) -> y-taps in N.
//   K = 16 voxels = two groups of 8 consecutive x of y rows 2s and 2s+1 (SBO = one y row pitch).
// Expected: D[j*32+co][i*CB+ci] = sum_{k<16} dY[(2s+k/8)*XE + k%8 + j][co] * X[(2s+k/8+i)*8 + k%8][ci].
// Variants: "canon" = LBO/SBO as documented for MN-major swizzled layouts, "swap" = exchanged.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o wgrad_probe wgrad_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_fp16.h>
#include "../tomoencoders_b200/csrc/sm100_ptx.cuh"

using namespace te;

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e_ = (x);                                                              \
    if (e_ != cudaSuccess) {                                                           \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);  \
      exit(1);                                                                         \
    }                                                                                  \
  } while (0)

__host__ __device__ inline uint32_t swz_abs(uint32_t addr, uint32_t layout) {
  uint32_t bits = layout == kLayoutSW128 ? 7u : layout == kLayoutSW64 ? 3u : layout == kLayoutSW32 ? 1u : 0u;
  return addr ^ (((addr >> 7) & bits) << 4);
}

constexpr int XE = 11;              // x extent of the dY box (8 + 3 shifts)
constexpr int A_ROWS = 16 * XE + 8; // + slack rows read by the junk atom
constexpr int B_ROWS = 18 * 8;

struct P {
  int cb;        // X channels per row: 16 or 32
  int s;         // k-step 0..7
  int swap;      // exchange LBO / SBO
  int a_off;     // extra byte offset of the A box inside its 1024-aligned slot (0 or 512 ...)
};

__global__ void __launch_bounds__(128, 1) probe(P p, const __half* __restrict__ ga, const __half* __restrict__ gb, float* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* A = smem + p.a_off;             // <= 16 KB
  uint8_t* B = smem + 16384;               // <= 9216 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 16384 + 10240);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t a_base = smem_u32(A), b_base = smem_u32(B);
  const uint32_t b_layout = p.cb == 32 ? kLayoutSW64 : kLayoutSW32;
  const uint32_t b_rowB = p.cb * 2;
  // the way TMA would write the boxes: logical byte L of the box lands at swizzle(abs address)
  for (int i = tid; i < A_ROWS * 32; i += 128) {
    uint32_t L = i * 2;
    uint32_t phys = swz_abs(a_base + L, kLayoutSW64) - a_base;
    *reinterpret_cast<__half*>(A + phys) = ga[i];
  }
  for (int i = tid; i < B_ROWS * p.cb; i += 128) {
    uint32_t L = i * 2;
    uint32_t phys = swz_abs(b_base + L, b_layout) - b_base;
    *reinterpret_cast<__half*>(B + phys) = gb[i];
  }
  if (tid == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(tmem_slot, 128); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (tid == 0) {
    uint32_t a_lbo = 64, a_sbo = XE * 64;
    uint32_t b_lbo = 8 * b_rowB, b_sbo = 8 * b_rowB;
    if (p.swap) { uint32_t t = a_lbo; a_lbo = a_sbo; a_sbo = t; }
    uint64_t da = make_smem_desc(a_base + (2 * p.s * XE) * 64, a_lbo, a_sbo, kLayoutSW64, 0);
    uint64_t db = make_smem_desc(b_base + (2 * p.s * 8) * b_rowB, b_lbo, b_sbo, b_layout, 0);
    uint32_t idesc = make_idesc_f16(128, 3 * p.cb, 0) | (1u << 15) | (1u << 16);
    umma_ss(tmem, da, db, idesc, 0);
    umma_commit(bar);
  }
  mbar_wait(bar, 0);
  tc_fence_after();
  for (int c0 = 0; c0 < 3 * p.cb; c0 += 16) {
    uint32_t r[16];
    tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, r);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j) out[(size_t)tid * 96 + c0 + j] = __uint_as_float(r[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 128);
}


// ------------------------------------------------------------------------------------------ issue rate
// cycles per UMMA (M 128, N 3 cb, K 16) with the product kernel's operand layouts: 8 k-steps x 3 accumulators per
// "plane", back to back from one thread, commit at the end.
struct PerfP { int cb; int a_mn, b_mn; int iters; };
__global__ void __launch_bounds__(128, 1) perf(PerfP p, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* A = smem;                       // 16 KB
  uint8_t* B = smem + 16384;               // 3 x 9216
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 16384 + 3 * 9216 + 1024);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (uint32_t i = tid; i < (16384 + 3 * 9216) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (tid == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (warp == 1) {
    const uint32_t rowB = p.cb * 2;
    const uint32_t blay = p.cb == 32 ? kLayoutSW64 : kLayoutSW32;
    uint32_t idesc = make_idesc_f16(128, 3 * p.cb, 1);
    uint64_t da0, db0;
    uint32_t a_step, b_step;
    if (p.a_mn) { idesc |= 1u << 15; da0 = make_smem_desc(smem_u32(A), 64, XE * 64, kLayoutSW64, 0); a_step = (2 * XE * 64) >> 4; }
    else { da0 = make_smem_desc(smem_u32(A), 16, 8 * 64, kLayoutSW64, 0); a_step = 2; }       // K-major 128 rows x 64 B, k-step = +32 B
    if (p.b_mn) { idesc |= 1u << 16; db0 = make_smem_desc(smem_u32(B), 8 * rowB, 8 * rowB, blay, 0); b_step = (16 * rowB) >> 4; }
    else { db0 = make_smem_desc(smem_u32(B), 16, 8 * 64, kLayoutSW64, 0); b_step = 2; }
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < p.iters; ++it) {
      if (elect_one()) {
#pragma unroll
        for (int dz = 0; dz < 3; ++dz)
#pragma unroll
          for (int s = 0; s < 8; ++s)
            umma_ss(tmem + dz * 3 * p.cb, da0 + (uint64_t)((s & (p.a_mn ? 7 : 1)) * a_step),
                    db0 + (uint64_t)(((dz * 9216) >> 4) + (s & (p.b_mn ? 7 : 1)) * b_step), idesc, 1);
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(bar);
    mbar_wait(bar, 0);
    long long t1 = clock64();
    if ((tid & 31) == 0) cycles[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

int main() {
  std::vector<__half> ha(A_ROWS * 32), hb(B_ROWS * 32);
  std::vector<float> fa(ha.size()), fb(hb.size());
  srand(1);
  for (size_t i = 0; i < ha.size(); ++i) { fa[i] = (float)(rand() % 7 - 3); ha[i] = __float2half(fa[i]); }
  for (size_t i = 0; i < hb.size(); ++i) { fb[i] = (float)(rand() % 5 - 2); hb[i] = __float2half(fb[i]); }
  __half *da, *db;
  float* dout;
  CK(cudaMalloc(&da, ha.size() * 2));
  CK(cudaMalloc(&db, hb.size() * 2));
  CK(cudaMalloc(&dout, 128 * 96 * 4));
  CK(cudaMemcpy(da, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 40960));
  std::vector<float> out(128 * 96);
  int bad_total = 0;
  for (int cb : {32, 16})
    for (int swap : {0, 1})
      for (int a_off : {0, 512, 64 * 3})
        for (int s = 0; s < 8; ++s) {
          P p{cb, s, swap, a_off};
          CK(cudaMemset(dout, 0xff, 128 * 96 * 4));
          probe<<<1, 128, 40960>>>(p, da, db, dout);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("cb %d swap %d a_off %d s %d: launch failed: %s\n", cb, swap, a_off, s, cudaGetErrorString(e)); return 1; }
          CK(cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost));
          int bad = 0, bad_junk = 0;
          double maxerr = 0;
          for (int m = 0; m < 128; ++m)
            for (int n = 0; n < 3 * cb; ++n) {
              int j = m / 32, co = m % 32, i = n / cb, ci = n % cb;
              double ref = 0;
              for (int k = 0; k < 16; ++k) {
                int ra = (2 * s + k / 8) * XE + k % 8 + j;
                int rb = (2 * s + k / 8 + i) * 8 + k % 8;
                // B rows hold cb channels each in the device box
                ref += (double)fa[ra * 32 + co] * (double)fb[rb * cb + ci];
              }
              double err = fabs(ref - out[m * 96 + n]);
              if (err > 1e-3) { (j < 3 ? bad : bad_junk)++; }
              if (j < 3 && err > maxerr) maxerr = err;
            }
          if (s == 0 || bad) printf("cb %2d %s a_off %3d s %d: mismatches %5d (junk atom %d) max err %.3g\n", cb, swap ? "swap " : "canon", a_off, s, bad, bad_junk, maxerr);
          if (!swap) bad_total += bad;
        }
  printf("canonical descriptors: %s\n", bad_total == 0 ? "ALL MATCH" : "MISMATCH");
  long long* dcyc;
  CK(cudaMalloc(&dcyc, 16));
  CK(cudaFuncSetAttribute(perf, cudaFuncAttributeMaxDynamicSharedMemorySize, 50 * 1024));
  for (int cb : {32, 16})
    for (int mode = 0; mode < 4; ++mode) {
      PerfP pp{cb, mode & 1, (mode >> 1) & 1, 2000};
      perf<<<1, 128, 50 * 1024>>>(pp, dcyc);
      CK(cudaDeviceSynchronize());
      long long c;
      CK(cudaMemcpy(&c, dcyc, 8, cudaMemcpyDeviceToHost));
      printf("issue rate N=%3d A %s B %s: %.1f cycles / UMMA\n", 3 * cb, pp.a_mn ? "MN-major" : "K-major ", pp.b_mn ? "MN-major" : "K-major ",
             (double)c / (2000.0 * 24));
    }
  return 0;
}

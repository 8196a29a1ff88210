// Hardware probe (not product code): tcgen05.mma.cta_group::2 (CTA pairs) on a real B200.
//   Q1  semantics: cluster of 2 CTAs; each CTA holds its own A tile (128 x 16, SW32 rows) and HALF of the B rows (N / 2) at
//       the same shared-memory offsets; the leader issues ONE M = 256 instruction; each CTA reads its 128 accumulator rows
//       from its own TMEM.  Checked against the host for N = 96 / 64 / 32 (the dz-fused shapes of the Cout = 32 conv layers).
//   Q2  issue rate (cycles / instruction) of the pair for N in {32 .. 256}: the single-CTA model is max(N/2, 32 + N/4)
//       (tools/umma_probe.cu); with the B rows split over the pair the operand read per CTA is 32 + N/8 cycles.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma2_probe umma2_probe.cu
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

__device__ __forceinline__ uint32_t sw32(uint32_t L) { return L ^ (((L >> 7) & 1u) << 4); }

// A_rank[m][k] = ((m * 3 + k * 5 + rank * 7) % 13) - 6 ; B[n][k] = ((n * 2 + k * 3) % 11) - 5   (exact in fp16 / fp32)
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) probe_pair(int N, float* out, int mode) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* A = smem;                    // 128 rows x 32 B = 4 KB
  uint8_t* B = smem + 4096;             // N / 2 rows x 32 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 8192);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 8192 + 64);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int rank = static_cast<int>(cluster_ctarank());
  for (int i = tid; i < 128 * 16; i += 128) {
    const int m = i / 16, k = i % 16;
    *reinterpret_cast<__half*>(A + sw32(m * 32 + k * 2)) = __float2half(static_cast<float>(((m * 3 + k * 5 + rank * 7) % 13) - 6));
  }
  const int nh = N / 2;
  for (int i = tid; i < nh * 16; i += 128) {
    const int nl = i / 16, k = i % 16, n = rank * nh + nl;
    *reinterpret_cast<__half*>(B + sw32(nl * 32 + k * 2)) = __float2half(static_cast<float>(((n * 2 + k * 3) % 11) - 5));
  }
  if (tid == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc2(tmem_slot, 128); tmem_relinquish2(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                   // the leader's MMA reads the peer's shared memory too
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (mode == 1) {
    // harness check: every CTA multiplies its A tile with ITS half of B through a plain cta_group::1 instruction
    if (tid == 0) {
      const uint64_t da = make_smem_desc(smem_u32(A), 16, 256, kLayoutSW32, 0);
      const uint64_t db = make_smem_desc(smem_u32(B), 16, 256, kLayoutSW32, 0);
      umma_ss(tmem, da, db, make_idesc_f16(128, N / 2, 0), 0);
      umma_commit(bar);
    }
  } else if (rank == 0 && tid == 0) {
    const uint64_t da = make_smem_desc(smem_u32(A), 16, 256, kLayoutSW32, 0);
    const uint64_t db = make_smem_desc(smem_u32(B), 16, 256, kLayoutSW32, 0);
    umma_ss2(tmem, da, db, make_idesc_f16(256, N, 0), 0);
    umma_commit2(bar, 3);
  }
  mbar_wait(bar, 0);
  tc_fence_after();
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t r[16];
    tmem_ld16(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c0, r);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j) out[(static_cast<size_t>(rank) * 128 + tid) * 256 + c0 + j] = __uint_as_float(r[j]);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) tmem_dealloc2(tmem, 128);
}

struct PerfParams { uint32_t n, iters; };
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) probe_pair_perf(PerfParams p, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* A = smem;               // 64 KB window
  uint8_t* B = smem + 65536;       // 32 KB
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 65536 + 32768);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t rank = cluster_ctarank();
  for (uint32_t i = tid; i < (65536 + 32768) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (tid == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc2(tmem_slot, 512); tmem_relinquish2(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (warp == 1 && rank == 0) {
    const uint32_t idesc = make_idesc_f16(256, p.n, 1);
    const uint64_t da0 = make_smem_desc(smem_u32(A), 16, 256, kLayoutSW32, 0);
    const uint64_t db0 = make_smem_desc(smem_u32(B), 16, 256, kLayoutSW32, 0);
    long long t0 = clock64();
#pragma unroll 1
    for (uint32_t it = 0; it < p.iters; it += 4) {
#pragma unroll
      for (int h = 0; h < 4; ++h)
        if (elect_one()) umma_ss2(tmem + (h & 1) * 256, da0 + static_cast<uint64_t>(h * 256), db0, idesc, 1);
    }
    long long t1 = clock64();
    if (elect_one()) umma_commit2(bar, 3);
    mbar_wait(bar, 0);
    long long t2 = clock64();
    if ((tid & 31) == 0) { cycles[0] = t2 - t0; cycles[1] = t1 - t0; }
  } else {
    mbar_wait(bar, 0);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) tmem_dealloc2(tmem, 512);
}

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  printf("device: %s sm_%d%d SMs=%d\n", prop.name, prop.major, prop.minor, prop.multiProcessorCount);
  CK(cudaFuncSetAttribute(probe_pair, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384));
  CK(cudaFuncSetAttribute(probe_pair_perf, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 32768 + 2048));
  float* d_out;
  CK(cudaMalloc(&d_out, 2 * 128 * 256 * 4));
  for (int N : {64, 96}) {
    CK(cudaMemset(d_out, 0xff, 2 * 128 * 256 * 4));
    probe_pair<<<2, 128, 16384>>>(N, d_out, 1);
    CK(cudaDeviceSynchronize());
    std::vector<float> h(2 * 128 * 256);
    CK(cudaMemcpy(h.data(), d_out, h.size() * 4, cudaMemcpyDeviceToHost));
    long ok = 0, tot = 0;
    for (int rank = 0; rank < 2; ++rank)
      for (int m = 0; m < 128; ++m)
        for (int nl = 0; nl < N / 2; ++nl) {
          const int n = rank * (N / 2) + nl;
          float acc = 0.f;
          for (int k = 0; k < 16; ++k) acc += static_cast<float>(((m * 3 + k * 5 + rank * 7) % 13) - 6) * static_cast<float>(((n * 2 + k * 3) % 11) - 5);
          ok += (h[(static_cast<size_t>(rank) * 128 + m) * 256 + nl] == acc);
          ++tot;
        }
    printf("Q0 harness (cta_group::1 per CTA, N/2 = %d): %ld/%ld match; first raw words %08x %08x\n", N / 2, ok, tot,
           *reinterpret_cast<uint32_t*>(&h[0]), *reinterpret_cast<uint32_t*>(&h[1]));
  }
  for (int N : {96, 64, 32, 128, 256}) {
    CK(cudaMemset(d_out, 0xff, 2 * 128 * 256 * 4));
    probe_pair<<<2, 128, 16384>>>(N, d_out, 0);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { printf("Q1 N=%d launch error: %s\n", N, cudaGetErrorString(e)); return 2; }
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("Q1 N=%d LAUNCH FAILED: %s\n", N, cudaGetErrorString(e)); return 2; }
    std::vector<float> h(2 * 128 * 256);
    CK(cudaMemcpy(h.data(), d_out, h.size() * 4, cudaMemcpyDeviceToHost));
    long ok = 0, tot = 0;
    for (int rank = 0; rank < 2; ++rank)
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < N; ++n) {
          float acc = 0.f;
          for (int k = 0; k < 16; ++k) acc += static_cast<float>(((m * 3 + k * 5 + rank * 7) % 13) - 6) * static_cast<float>(((n * 2 + k * 3) % 11) - 5);
          ok += (h[(static_cast<size_t>(rank) * 128 + m) * 256 + n] == acc);
          ++tot;
        }
    printf("Q1 cta_group::2 M=256 N=%3d K=16 (A per CTA, B rows split %d + %d): %ld/%ld accumulator values match the host\n", N, N / 2, N / 2, ok, tot);
    if (ok != tot) {
      for (int rank = 0; rank < 2; ++rank)
        for (int m : {0, 1, 9}) {
          printf("   rank %d m %d raw %08x got:", rank, m, *reinterpret_cast<uint32_t*>(&h[(static_cast<size_t>(rank) * 128 + m) * 256]));
          for (int n = 0; n < 8; ++n) printf(" %7.0f", h[(static_cast<size_t>(rank) * 128 + m) * 256 + n]);
          printf(" | n=%d..:", N / 2);
          for (int n = N / 2; n < N / 2 + 4; ++n) printf(" %7.0f", h[(static_cast<size_t>(rank) * 128 + m) * 256 + n]);
          printf("\n   expected       :");
          for (int n = 0; n < 8; ++n) {
            float acc = 0.f;
            for (int k = 0; k < 16; ++k) acc += static_cast<float>(((m * 3 + k * 5 + rank * 7) % 13) - 6) * static_cast<float>(((n * 2 + k * 3) % 11) - 5);
            printf(" %7.0f", acc);
          }
          printf("\n");
        }
    }
  }
  long long* d_cyc;
  CK(cudaMalloc(&d_cyc, 16));
  for (uint32_t n : {32u, 64u, 96u, 128u, 192u, 256u}) {
    PerfParams pp{n, 8192};
    probe_pair_perf<<<2, 128, 65536 + 32768 + 2048>>>(pp, d_cyc);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { printf("Q2 launch error: %s\n", cudaGetErrorString(e)); return 2; }
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("Q2 LAUNCH FAILED: %s\n", cudaGetErrorString(e)); return 2; }
    long long cyc[2];
    CK(cudaMemcpy(cyc, d_cyc, 16, cudaMemcpyDeviceToHost));
    const double per = static_cast<double>(cyc[0]) / 8192.0;
    printf("Q2 cta_group::2 M=256 N=%3u K=16 : %.1f cycles/UMMA (issue loop alone %.1f; math N/2 = %.0f; single-CTA model max(N/2, 32+N/4) = %.0f) -> %.0f%% of the pair's tensor peak\n",
           n, per, static_cast<double>(cyc[1]) / 8192.0, n / 2.0, n / 2.0 > 32 + n / 4.0 ? n / 2.0 : 32 + n / 4.0, 100.0 * (n / 2.0) / per);
  }
  printf("probe done\n");
  return 0;
}

// Hardware probe (not product code): pins down the tcgen05 shared-memory-descriptor semantics the
// conv kernels rely on, on a real B200:
//   P1  K-major swizzled A tiles whose start address is NOT aligned to the swizzle repeat and whose
//       8-row-group stride (SBO) is NOT a multiple of it  -> "shifted window into a halo tile".
//   P2  a 5-D TMA box load with negative start coordinates (zero fill) + SWIZZLE_64B, followed by
//       UMMA reads through shifted descriptors (the actual conv access pattern).
//   P3  UMMA issue throughput (cycles / instruction) for M=128, N in {16..256}, SS operands.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_probe umma_probe.cu
// Output is plain text on stdout; the summary is kept under profiles/.
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

__device__ __forceinline__ uint32_t swz_abs(uint32_t addr, uint32_t layout) {
  uint32_t bits = layout == kLayoutSW128 ? 7u : layout == kLayoutSW64 ? 3u : layout == kLayoutSW32 ? 1u : 0u;
  return addr ^ (((addr >> 7) & bits) << 4);
}

struct DescParams {
  uint32_t layout, rowB, start_off, sbo, base_off, lbo;
};

// ------------------------------------------------------------------------------------------ P1
__global__ void __launch_bounds__(128, 1) probe_desc(DescParams p, float* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* A = smem;                    // 32 KB
  uint8_t* B = smem + 32768;            // 512 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 32768 + 1024);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 32768 + 1024 + 64);
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t a_base = smem_u32(A);
  const uint32_t chunks_per_row = p.rowB / 16;
  // fill A: 2048 chunks of 16 B, value = logical chunk index (= pixel*chunks_per_row + chunk)
  for (uint32_t ch = tid; ch < 2048; ch += 128) {
    uint32_t L = ch * 16;
    uint32_t phys = swz_abs(a_base + L, p.layout) - a_base;
    __half v = __float2half((float)ch);
    __half* dst = reinterpret_cast<__half*>(A + phys);
    for (int e = 0; e < 8; ++e) dst[e] = v;
  }
  // fill B: identity 16x16 fp16, no swizzle, core matrices 8 rows x 16 B, LBO(K)=128, SBO(N)=256
  for (int i = tid; i < 256; i += 128) {
    int n = i / 16, k = i % 16;
    uint32_t off = (n % 8) * 16 + (n / 8) * 256 + (k / 8) * 128 + (k % 8) * 2;
    *reinterpret_cast<__half*>(B + off) = __float2half(n == k ? 1.f : 0.f);
  }
  if (tid == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(tmem_slot, 32); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (tid == 0) {
    uint64_t da = make_smem_desc(a_base + p.start_off, p.lbo, p.sbo, p.layout, p.base_off);
    uint64_t db = make_smem_desc(smem_u32(B), 128, 256, kLayoutNone, 0);
    uint32_t idesc = make_idesc_f16(128, 16, 0);
    umma_ss(tmem, da, db, idesc, 0);
    umma_commit(bar);
  }
  mbar_wait(bar, 0);
  tc_fence_after();
  uint32_t r[16];
  tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16), r);
  tmem_ld_wait();
  for (int j = 0; j < 16; ++j) out[(size_t)tid * 16 + j] = __uint_as_float(r[j]);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 32);
}

// ------------------------------------------------------------------------------------------ P2
// tensor X: (N=1, Z=8, Y=32, X=32, C=32) fp16, NDHWC.  box = (32 ch, 10 x, 18 y, 3 z) at
// (0, x0-1, y0-1, z0-1).  Then for each of the 27 taps one UMMA (M=128 = 16 y-rows of 8 x, N=16,
// K=16, second 16-channel half via +32 B start) reads the shifted window; identity B extracts it.
struct TmaParams { int x0, y0, z0; };
__global__ void __launch_bounds__(128, 1)
probe_tma(const __grid_constant__ CUtensorMap tmap, TmaParams tp, float* out, uint8_t* raw_dump) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr uint32_t kPlaneB = 18 * 10 * 64;       // 11520
  constexpr uint32_t kABytes = 3 * kPlaneB;        // 34560
  uint8_t* A = smem;
  uint8_t* B = smem + 35840;                        // 1024-aligned
  uint64_t* bar_tma = reinterpret_cast<uint64_t*>(smem + 35840 + 1024);
  uint64_t* bar_mma = bar_tma + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_tma + 2);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 256; i += 128) {
    int n = i / 16, k = i % 16;
    uint32_t off = (n % 8) * 16 + (n / 8) * 256 + (k / 8) * 128 + (k % 8) * 2;
    *reinterpret_cast<__half*>(B + off) = __float2half(n == k ? 1.f : 0.f);
  }
  if (tid == 0) { mbar_init(bar_tma, 1); mbar_init(bar_mma, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(tmem_slot, 32); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (tid == 0) {
    mbar_expect_tx(bar_tma, kABytes);
    tma_load_5d(A, &tmap, bar_tma, 0, tp.x0 - 1, tp.y0 - 1, tp.z0 - 1, 0);
  }
  mbar_wait(bar_tma, 0);
  for (uint32_t i = tid; i < kABytes; i += 128) raw_dump[i] = A[i];
  __syncthreads();
  uint32_t phase = 0;
  for (int tap = 0; tap < 27; ++tap) {
    for (int kh = 0; kh < 2; ++kh) {
      const int dz = tap / 9, dy = (tap / 3) % 3, dx = tap % 3;   // already +1 biased
      if (tid == 0) {
        uint32_t start = smem_u32(A) + dz * kPlaneB + (dy * 10 + dx) * 64 + kh * 32;
        uint64_t da = make_smem_desc(start, 16, 640, kLayoutSW64, 0);
        uint64_t db = make_smem_desc(smem_u32(B), 128, 256, kLayoutNone, 0);
        umma_ss(tmem, da, db, make_idesc_f16(128, 16, 0), 0);
        umma_commit(bar_mma);
      }
      mbar_wait(bar_mma, phase);
      phase ^= 1;
      tc_fence_after();
      uint32_t r[16];
      tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16), r);
      tmem_ld_wait();
      float* o = out + ((size_t)(tap * 2 + kh) * 128 + tid) * 16;
      for (int j = 0; j < 16; ++j) o[j] = __uint_as_float(r[j]);
      tc_fence_before();
      __syncthreads();
      tc_fence_after();
    }
  }
  if (warp == 0) tmem_dealloc(tmem, 32);
}

// ------------------------------------------------------------------------------------------ P3
// Issue loop shaped like the product kernel: the whole MMA warp runs the (uniform) loop, one elected
// lane issues; descriptors are precomputed and only their low word advances.
struct PerfParams { uint32_t n, layout, rowB, iters, k_steps, same_acc; };
template <int KS>
__device__ __forceinline__ void perf_loop(const PerfParams& p, uint32_t tmem, uint32_t a_addr, uint32_t b_addr) {
  const uint32_t idesc = make_idesc_f16(128, p.n, 1);
  const uint32_t sbo = 8 * p.rowB;
  const uint64_t da0 = make_smem_desc(a_addr, 16, sbo, p.layout, 0);
  const uint64_t db0 = make_smem_desc(b_addr, 16, sbo, p.layout, 0);
  const uint32_t acc_stride = p.same_acc ? 0u : 256u;
#pragma unroll 1
  for (uint32_t it = 0; it < p.iters; it += 2) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
#pragma unroll
      for (int k = 0; k < KS; ++k) {
        if (elect_one()) umma_ss(tmem + h * acc_stride, da0 + (uint64_t)(k * 2 + h * 64), db0 + (uint64_t)(k * 2), idesc, 1);
      }
    }
  }
}
__global__ void __launch_bounds__(128, 1) probe_perf(PerfParams p, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* A = smem;               // 64 KB window
  uint8_t* B = smem + 65536;       // 256 rows x 128 B = 32 KB
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 65536 + 32768);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (uint32_t i = tid; i < (65536 + 32768) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (tid == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (warp == 1) {
    long long t0 = clock64();
    if (p.k_steps == 4) perf_loop<4>(p, tmem, smem_u32(A), smem_u32(B));
    else if (p.k_steps == 2) perf_loop<2>(p, tmem, smem_u32(A), smem_u32(B));
    else perf_loop<1>(p, tmem, smem_u32(A), smem_u32(B));
    long long t1 = clock64();
    if (elect_one()) umma_commit(bar);
    mbar_wait(bar, 0);
    long long t2 = clock64();
    if ((tid & 31) == 0) { cycles[0] = t2 - t0; cycles[1] = t1 - t0; }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------ host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static void run_desc(const char* name, DescParams p, float* d_out, bool verbose) {
  CK(cudaMemset(d_out, 0xff, 128 * 16 * 4));
  probe_desc<<<1, 128, 40960>>>(p, d_out);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("P1 %-34s LAUNCH FAILED: %s\n", name, cudaGetErrorString(e)); exit(2); }
  std::vector<float> h(128 * 16);
  CK(cudaMemcpy(h.data(), d_out, h.size() * 4, cudaMemcpyDeviceToHost));
  int ok = 0, tot = 0;
  for (int m = 0; m < 128; ++m)
    for (int half = 0; half < 2; ++half) {
      uint32_t L = p.start_off + (m % 8) * p.rowB + (m / 8) * p.sbo + half * 16;
      float expect = (float)(L / 16);
      bool good = true;
      for (int e2 = 0; e2 < 8; ++e2) good &= (h[m * 16 + half * 8 + e2] == expect);
      ok += good; ++tot;
    }
  printf("P1 %-34s layout=%u rowB=%3u start=%5u sbo=%5u base_off=%u : %3d/%3d chunks match absolute-address model\n",
         name, p.layout, p.rowB, p.start_off, p.sbo, p.base_off, ok, tot);
  if (ok != tot && verbose) {
    printf("    row: got(chunk idx k0-7, k8-15) expected\n");
    for (int m = 0; m < 24; ++m) {
      uint32_t L = p.start_off + (m % 8) * p.rowB + (m / 8) * p.sbo;
      printf("    m=%3d got %6.0f %6.0f  exp %4u %4u\n", m, h[m * 16], h[m * 16 + 8], L / 16, L / 16 + 1);
    }
  }
}

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  printf("device: %s sm_%d%d SMs=%d\n", prop.name, prop.major, prop.minor, prop.multiProcessorCount);
  CK(cudaFuncSetAttribute(probe_desc, cudaFuncAttributeMaxDynamicSharedMemorySize, 40960));
  CK(cudaFuncSetAttribute(probe_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, 40960));
  CK(cudaFuncSetAttribute(probe_perf, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 32768 + 2048));
  float* d_out;
  CK(cudaMalloc(&d_out, 27 * 2 * 128 * 16 * 4));

  // ---- P1
  struct { const char* name; DescParams p; } cases[] = {
      {"sw128 aligned canonical", {kLayoutSW128, 128, 0, 1024, 0, 16}},
      {"sw128 k-advance +32B", {kLayoutSW128, 128, 32, 1024, 0, 16}},
      {"sw128 start +1 row, bo=0", {kLayoutSW128, 128, 128, 1024, 0, 16}},
      {"sw128 start +1 row, bo=1", {kLayoutSW128, 128, 128, 1024, 1, 16}},
      {"sw128 start +3 rows, bo=0", {kLayoutSW128, 128, 384, 1024, 0, 16}},
      {"sw128 start +3 rows, bo=3", {kLayoutSW128, 128, 384, 1024, 3, 16}},
      {"sw128 sbo=1280 aligned start", {kLayoutSW128, 128, 0, 1280, 0, 16}},
      {"sw128 sbo=1280 start +11 rows bo=0", {kLayoutSW128, 128, 11 * 128, 1280, 0, 16}},
      {"sw128 sbo=1280 start +11 rows bo=3", {kLayoutSW128, 128, 11 * 128, 1280, 3, 16}},
      {"sw128 sbo=2048 start +11 rows bo=0", {kLayoutSW128, 128, 11 * 128, 2048, 0, 16}},
      {"sw128 sbo=2048 start +11 rows bo=3", {kLayoutSW128, 128, 11 * 128, 2048, 3, 16}},
      {"sw64 aligned canonical", {kLayoutSW64, 64, 0, 512, 0, 16}},
      {"sw64 k-advance +32B", {kLayoutSW64, 64, 32, 512, 0, 16}},
      {"sw64 start +1 row bo=0", {kLayoutSW64, 64, 64, 512, 0, 16}},
      {"sw64 start +2 rows bo=0", {kLayoutSW64, 64, 128, 512, 0, 16}},
      {"sw64 start +2 rows bo=1", {kLayoutSW64, 64, 128, 512, 1, 16}},
      {"sw64 sbo=640 aligned start", {kLayoutSW64, 64, 0, 640, 0, 16}},
      {"sw64 sbo=640 start +11 rows bo=0", {kLayoutSW64, 64, 11 * 64, 640, 0, 16}},
      {"sw64 sbo=640 start +11 rows +32B", {kLayoutSW64, 64, 11 * 64 + 32, 640, 0, 16}},
      {"sw64 sbo=640 start +11 rows bo=5", {kLayoutSW64, 64, 11 * 64, 640, 5, 16}},
      {"sw32 aligned canonical", {kLayoutSW32, 32, 0, 256, 0, 16}},
      {"sw32 start +1 row bo=0", {kLayoutSW32, 32, 32, 256, 0, 16}},
      {"sw32 sbo=320 start +11 rows bo=0", {kLayoutSW32, 32, 11 * 32, 320, 0, 16}},
  };
  for (auto& c : cases) run_desc(c.name, c.p, d_out, true);

  // ---- P2
  {
    EncodeTiledFn encode = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres));
    if (!encode) { printf("P2 no cuTensorMapEncodeTiled\n"); return 3; }
    const int Z = 8, Y = 32, X = 32, C = 32;
    std::vector<__half> hx((size_t)Z * Y * X * C);
    for (int z = 0; z < Z; ++z) for (int y = 0; y < Y; ++y) for (int x = 0; x < X; ++x) for (int c = 0; c < C; ++c)
      hx[(((size_t)z * Y + y) * X + x) * C + c] = __float2half((float)(((z * 7 + y * 3 + x) % 61) * 32 + c + 1));
    __half* dx;
    CK(cudaMalloc(&dx, hx.size() * 2));
    CK(cudaMemcpy(dx, hx.data(), hx.size() * 2, cudaMemcpyHostToDevice));
    CUtensorMap tmap;
    cuuint64_t gdim[5] = {(cuuint64_t)C, (cuuint64_t)X, (cuuint64_t)Y, (cuuint64_t)Z, 1};
    cuuint64_t gstr[4] = {(cuuint64_t)C * 2, (cuuint64_t)X * C * 2, (cuuint64_t)Y * X * C * 2, (cuuint64_t)Z * Y * X * C * 2};
    cuuint32_t box[5] = {32, 10, 18, 3, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, dx, gdim, gstr, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("P2 cuTensorMapEncodeTiled failed %d\n", (int)r); return 3; }
    uint8_t* d_raw;
    CK(cudaMalloc(&d_raw, 3 * 18 * 10 * 64));
    struct { int x0, y0, z0; } tiles[] = {{0, 0, 0}, {24, 16, 7}, {8, 16, 3}};
    for (auto& t : tiles) {
      TmaParams tp{t.x0, t.y0, t.z0};
      probe_tma<<<1, 128, 40960>>>(tmap, tp, d_out, d_raw);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("P2 LAUNCH FAILED: %s\n", cudaGetErrorString(e)); return 2; }
      std::vector<float> h((size_t)27 * 2 * 128 * 16);
      CK(cudaMemcpy(h.data(), d_out, h.size() * 4, cudaMemcpyDeviceToHost));
      std::vector<uint8_t> raw(3 * 18 * 10 * 64);
      CK(cudaMemcpy(raw.data(), d_raw, raw.size(), cudaMemcpyDeviceToHost));
      // (a) raw smem layout check under the absolute-address SW64 model (tile base is 1024-aligned)
      long ok_raw = 0, tot_raw = 0;
      for (int pz = 0; pz < 3; ++pz) for (int py = 0; py < 18; ++py) for (int px = 0; px < 10; ++px) for (int c = 0; c < 32; ++c) {
        uint32_t L = ((pz * 18 + py) * 10 + px) * 64 + c * 2;
        uint32_t P = L ^ (((L >> 7) & 3) << 4);
        __half got; memcpy(&got, &raw[P], 2);
        int z = t.z0 - 1 + pz, y = t.y0 - 1 + py, x = t.x0 - 1 + px;
        float expect = 0.f;
        if (z >= 0 && z < Z && y >= 0 && y < Y && x >= 0 && x < X) expect = __half2float(hx[(((size_t)z * Y + y) * X + x) * C + c]);
        ok_raw += (__half2float(got) == expect); ++tot_raw;
      }
      // (b) UMMA shifted-window reads
      long ok = 0, tot = 0;
      for (int tap = 0; tap < 27; ++tap) for (int kh = 0; kh < 2; ++kh) for (int m = 0; m < 128; ++m) for (int n = 0; n < 16; ++n) {
        int dz = tap / 9 - 1, dy = (tap / 3) % 3 - 1, dx2 = tap % 3 - 1;
        int z = t.z0 + dz, y = t.y0 + (m / 8) + dy, x = t.x0 + (m % 8) + dx2, c = kh * 16 + n;
        float expect = 0.f;
        if (z >= 0 && z < Z && y >= 0 && y < Y && x >= 0 && x < X) expect = __half2float(hx[(((size_t)z * Y + y) * X + x) * C + c]);
        ok += (h[((size_t)(tap * 2 + kh) * 128 + m) * 16 + n] == expect); ++tot;
      }
      printf("P2 tile(x0=%d,y0=%d,z0=%d): raw TMA smem image %ld/%ld ok ; shifted-window UMMA reads %ld/%ld ok\n",
             t.x0, t.y0, t.z0, ok_raw, tot_raw, ok, tot);
    }
  }

  // ---- P3
  {
    long long* d_cyc;
    CK(cudaMalloc(&d_cyc, 16));
    struct { uint32_t layout, rowB, ksteps; const char* nm; } lay[] = {
        {kLayoutSW128, 128, 4, "SW128 k64"}, {kLayoutSW64, 64, 2, "SW64 k32"}, {kLayoutSW32, 32, 1, "SW32 k16"}};
    uint32_t ns[] = {16, 32, 64, 96, 128, 256};
    for (uint32_t same_acc = 0; same_acc < 2; ++same_acc)
    for (auto& l : lay)
      for (uint32_t n : ns) {
        PerfParams pp{n, l.layout, l.rowB, 4096, l.ksteps, same_acc};
        probe_perf<<<1, 128, 65536 + 32768 + 2048>>>(pp, d_cyc);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("P3 LAUNCH FAILED: %s\n", cudaGetErrorString(e)); return 2; }
        long long cyc[2];
        CK(cudaMemcpy(cyc, d_cyc, 16, cudaMemcpyDeviceToHost));
        double per = (double)cyc[0] / (4096.0 * l.ksteps);
        double per_issue = (double)cyc[1] / (4096.0 * l.ksteps);
        printf("P3 %-10s acc=%s M=128 N=%3u K=16 : %.1f cycles/UMMA (issue loop alone %.1f; ideal N/2 = %.0f) -> %.0f%% of tensor peak\n",
               l.nm, same_acc ? "same" : "alt ", n, per, per_issue, n / 2.0, 100.0 * (n / 2.0) / per);
      }
  }
  printf("probe done\n");
  return 0;
}

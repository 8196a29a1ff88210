// First layer of every graph (Conv3D(f, 3, 'same') on ONE input channel,
// /root/reference/tomo_encoders/neural_nets/Unet3D.py:117-122, autoencoders.py:168-173) on the tensor
// cores.  K = 27 taps is padded to 32: four "builder" warps assemble the im2col operand of one z plane
// (128 voxels x 32 taps, 64-byte rows, 64-byte swizzle) in shared memory from the haloed single-channel
// input tile, one warp issues two K = 16 UMMAs per plane into TMEM, four warps run the epilogue
// (bias + LeakyReLU -> 16-bit NDHWC).  The CUDA-core version of this layer needed 432 FMAs per voxel
// and was 8 % of a U-net forward / 22 % of an encoder forward for 0.2 % of the FLOPs.
#include "layers_simt.cuh"
#include "sm100_ptx.cuh"
#include "te_common.cuh"

#include <cstdlib>

namespace te {
namespace {

constexpr int kTZ = 8;                       // z planes per tile (TMEM: 2 x 8 x Cout columns -> Cout <= 32)
constexpr int kStagesA = kTZ;                // im2col ring: one plane (128 rows x 64 B = 8 KB) per stage; plane pz of
                                             // the k-th tile of a CTA uses stage pz with phase parity k & 1
constexpr int kThreadsC1 = 448;              // warp 0 UMMA, warp 1 idle, warps 2-5 epilogue, warps 6-13 builders (two groups of 128)
constexpr int kInTile = (kTZ + 2) * 18 * 10; // haloed input tile (elements)

struct C1Params {
  const uint16_t* in;       // (N, D, H, W) 16-bit
  uint8_t* out;             // NDHWC 16-bit
  int out_ctot, out_coff;
  const float* w;           // [27][Cout] fp32 (Keras (3,3,3,1,Cout))
  const float* bias;
  int N, D, H, W;
  int tiles_x, tiles_y, tiles_z;
  long long total_tiles;
  int lrelu, fp16;
  float alpha;
  int* overflow;            // fp16: OR-ed with 1 when an output leaves the finite fp16 range (nullptr = not tracked)
};

__device__ __forceinline__ uint16_t to_bits(float v, int fp16) {
  if (fp16) { __half h = __float2half_rn(v); return *reinterpret_cast<uint16_t*>(&h); }
  __nv_bfloat16 h = __float2bfloat16_rn(v);
  return *reinterpret_cast<uint16_t*>(&h);
}
__device__ __forceinline__ uint32_t pack2f(float a, float b, int fp16) {
  if (fp16) { __half2 h = __floats2half2_rn(a, b); return *reinterpret_cast<uint32_t*>(&h); }
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

template <int NT>
// NT = 16: two CTAs per SM (72 registers, 2 x 74 KB shared memory, 2 x 256 TMEM columns): builders, issuer and epilogue of
// one CTA are each latency-bound, a second CTA fills the gaps.  NT = 32 needs all 512 TMEM columns: one CTA per SM.
__global__ void __launch_bounds__(kThreadsC1, NT <= 16 ? 2 : 1) conv1_umma_kernel(const C1Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_stage = smem;                                   // kStagesA x 8 KB
  uint8_t* b_tile = smem + kStagesA * 8192;                  // NT rows x 64 B (K-major, SW64), 1024-aligned
  uint16_t* in_tile = reinterpret_cast<uint16_t*>(b_tile + 4096);   // [kInTile]
  float* s_bias = reinterpret_cast<float*>(b_tile + 4096 + 4096);
  uint64_t* bars = reinterpret_cast<uint64_t*>(b_tile + 4096 + 4096 + 512);
  uint64_t* a_full = bars;                    // [kStagesA], 128 arrivals
  uint64_t* a_empty = a_full + kStagesA;      // [kStagesA]
  uint64_t* acc_full = a_empty + kStagesA;    // [2]
  uint64_t* acc_empty = acc_full + 2;         // [2], 4 arrivals
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int TMEM_COLS = 2 * kTZ * NT <= 128 ? 128 : (2 * kTZ * NT <= 256 ? 256 : 512);

  if (threadIdx.x == 0) {
    for (int i = 0; i < kStagesA; ++i) { mbar_init(a_full + i, 128); mbar_init(a_empty + i, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(acc_full + i, 1); mbar_init(acc_empty + i, 4); }
    fence_barrier_init();
  }
  if (warp == 0) { tmem_alloc(tmem_slot, TMEM_COLS); tmem_relinquish(); }
  // weights: K slot k = 10 * dz + (3 * dy + dx) holds tap (dz, dy, dx); slots 9, 19, 29, 30, 31 are zero padding (each input
  // plane owns 10 slots = 5 packed 32-bit words, so the builders keep a plane's taps packed while the window rolls);
  // 64-byte rows, SW64 swizzle
  for (int i = threadIdx.x; i < NT * 32; i += kThreadsC1) {
    const int n = i >> 5, k = i & 31;
    const int kz = k / 10, kr = k - kz * 10;
    const float v = (k < 30 && kr < 9) ? __ldg(p.w + (kz * 9 + kr) * NT + n) : 0.f;
    const uint32_t lo = static_cast<uint32_t>(n * 64 + k * 2);
    *reinterpret_cast<uint16_t*>(b_tile + (lo ^ (((lo >> 7) & 3u) << 4))) = to_bits(v, p.fp16);
  }
  for (int i = threadIdx.x; i < NT; i += kThreadsC1) s_bias[i] = __ldg(p.bias + i);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ============================== UMMA issuer ==============================
    const uint32_t idesc = make_idesc_f16(128, NT, p.fp16 ? 0u : 1u);
    const uint64_t bdesc = make_smem_desc(smem_u32(b_tile), 16, 512, kLayoutSW64);
    uint32_t as = 0, aph = 0, ab = 0, abph = 0;
    for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      mbar_wait(acc_empty + ab, abph ^ 1);
      tc_fence_after();
#pragma unroll 1
      for (int pz = 0; pz < kTZ; ++pz) {
        mbar_wait(a_full + as, aph);
        tc_fence_after();
        const uint64_t adesc = make_smem_desc(smem_u32(a_stage + as * 8192), 16, 512, kLayoutSW64);
        if (elect_one()) {
          const uint32_t acc = tmem_base + ab * (kTZ * NT) + pz * NT;
          umma_ss(acc, adesc, bdesc, idesc, 0u);
          umma_ss(acc, adesc + 2, bdesc + 2, idesc, 1u);       // second K = 16 step: +32 bytes
          umma_commit(a_empty + as);
          if (pz == kTZ - 1) umma_commit(acc_full + ab);
        }
        __syncwarp();
        if (++as == kStagesA) { as = 0; aph ^= 1; }
      }
      if (++ab == 2) { ab = 0; abph ^= 1; }
    }
  } else if (warp >= 6) {
    // ============================== im2col builders (2 groups x 128 threads) ==============================
    // group g builds planes g and g + 2 of every tile; plane pz of the k-th tile of this CTA goes to ring
    // stage (4 k + pz) % 8.  The haloed input tile of the NEXT tile is prefetched into registers while the
    // current one is being expanded, so the global-memory latency is off the critical path.
    const int bt = threadIdx.x - 6 * 32;              // 0..255
    const int grp = bt >> 7, row = bt & 127;          // row = voxel of the plane
    const int my = row >> 3, mx = row & 7;
    constexpr int kPre = (kInTile + 255) / 256;       // input elements per builder thread
    uint16_t pre[kPre];
    auto fetch = [&](long long tile) {
      unsigned t = static_cast<unsigned>(tile);           // 32-bit divisions (the launch checks total_tiles < 2^31)
      const unsigned ntx = p.tiles_x, nty = p.tiles_y, ntz = p.tiles_z;
      const int x0 = static_cast<int>(t % ntx) * 8; t /= ntx;
      const int y0 = static_cast<int>(t % nty) * 16; t /= nty;
      const int z0 = static_cast<int>(t % ntz) * kTZ; t /= ntz;
      const long long n = t;
#pragma unroll
      for (int j = 0; j < kPre; ++j) {
        const int i = bt + 256 * j;
        const int bx = i % 10, by = (i / 10) % 18, bz = i / 180;
        const int xx = x0 + bx - 1, yy = y0 + by - 1, zz = z0 + bz - 1;
        const bool ok = i < kInTile && xx >= 0 && xx < p.W && yy >= 0 && yy < p.H && zz >= 0 && zz < p.D;
        pre[j] = ok ? __ldg(p.in + ((n * p.D + zz) * p.H + yy) * p.W + xx) : static_cast<uint16_t>(0);
      }
    };
    if (static_cast<long long>(blockIdx.x) < p.total_tiles) fetch(blockIdx.x);
    uint32_t k = 0;
    for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++k) {
      named_bar_sync(1, 256);                           // previous tile fully expanded by both groups
#pragma unroll
      for (int j = 0; j < kPre; ++j) {
        const int i = bt + 256 * j;
        if (i < kInTile) in_tile[i] = pre[j];
      }
      named_bar_sync(1, 256);
      if (tile + gridDim.x < p.total_tiles) fetch(tile + gridDim.x);
      // group g expands planes [g kTZ/2, (g + 1) kTZ/2) one after the other and keeps the 3 x 3 in-plane
      // neighbourhoods of the last two input planes in registers: 9 shared-memory loads per output plane
      // instead of 27
      constexpr int PPG = kTZ / 2;
      const int pz0 = grp * PPG;
      // pk[q][0..4] = the 3 x 3 in-plane neighbourhood of input plane q as five packed words (taps 0..8 + one zero)
      uint32_t pk[3][5];
      auto load_plane = [&](int pzi, uint32_t (&w)[5]) {
        const uint16_t* base = in_tile + (pzi * 18 + my) * 10 + mx;
        uint16_t t9[9];
#pragma unroll
        for (int t = 0; t < 9; ++t) t9[t] = base[(t / 3) * 10 + t % 3];
#pragma unroll
        for (int j = 0; j < 4; ++j) w[j] = static_cast<uint32_t>(t9[2 * j]) | (static_cast<uint32_t>(t9[2 * j + 1]) << 16);
        w[4] = static_cast<uint32_t>(t9[8]);
      };
      load_plane(pz0, pk[0]);
      load_plane(pz0 + 1, pk[1]);
#pragma unroll
      for (int pp = 0; pp < PPG; ++pp) {
        const int pz = pz0 + pp;
        load_plane(pz + 2, pk[2]);
        const uint32_t as = pz, aph = k & 1u;
        mbar_wait(a_empty + as, aph ^ 1);
        uint8_t* dst = a_stage + as * 8192 + row * 64;
        const int sw = (row >> 1) & 3;
        // 16 words: plane 0 (5), plane 1 (5), plane 2 (5), one zero word
        *reinterpret_cast<uint4*>(dst + ((0 ^ sw) << 4)) = make_uint4(pk[0][0], pk[0][1], pk[0][2], pk[0][3]);
        *reinterpret_cast<uint4*>(dst + ((1 ^ sw) << 4)) = make_uint4(pk[0][4], pk[1][0], pk[1][1], pk[1][2]);
        *reinterpret_cast<uint4*>(dst + ((2 ^ sw) << 4)) = make_uint4(pk[1][3], pk[1][4], pk[2][0], pk[2][1]);
        *reinterpret_cast<uint4*>(dst + ((3 ^ sw) << 4)) = make_uint4(pk[2][2], pk[2][3], pk[2][4], 0u);
        fence_proxy_async_smem();                       // generic-proxy writes -> visible to the UMMA (async proxy)
        mbar_arrive(a_full + as);
#pragma unroll
        for (int j = 0; j < 5; ++j) { pk[0][j] = pk[1][j]; pk[1][j] = pk[2][j]; }
      }
    }
  } else if (warp >= 2) {
    // ============================== epilogue warps ==============================
    const int q = warp & 3;
    const int m = q * 32 + lane, my = m >> 3, mx = m & 7;
    uint32_t ab = 0, abph = 0;
    float amax = 0.f;                                     // see conv_umma.cu: one FMNMX per element detects fp16 overflow
    for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      unsigned t = static_cast<unsigned>(tile);           // 32-bit divisions (the launch checks total_tiles < 2^31)
      const unsigned ntx = p.tiles_x, nty = p.tiles_y, ntz = p.tiles_z;
      const int x0 = static_cast<int>(t % ntx) * 8; t /= ntx;
      const int y0 = static_cast<int>(t % nty) * 16; t /= nty;
      const int z0 = static_cast<int>(t % ntz) * kTZ; t /= ntz;
      const long long n = t;
      mbar_wait(acc_full + ab, abph);
      tc_fence_after();
      const uint32_t tbase = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + ab * (kTZ * NT);
      const int iy = y0 + my, ix = x0 + mx;
#pragma unroll 1
      for (int pz = 0; pz < kTZ; pz += 2) {
#pragma unroll
        for (int c0 = 0; c0 < NT; c0 += 16) {
          uint32_t r0[16], r1[16];
          tmem_ld16(tbase + pz * NT + c0, r0);
          tmem_ld16(tbase + (pz + 1) * NT + c0, r1);
          tmem_ld_wait();
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int iz = z0 + pz + h;
            float v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              float x = __uint_as_float(h ? r1[j] : r0[j]) + s_bias[c0 + j];
              if (p.lrelu) x = x > 0.f ? x : p.alpha * x;
              amax = fmaxf(amax, fabsf(x));
              v[j] = x;
            }
            if (iz < p.D && iy < p.H && ix < p.W) {
              const long long ov = ((n * p.D + iz) * p.H + iy) * p.W + ix;
              uint4* dst = reinterpret_cast<uint4*>(p.out + (ov * p.out_ctot + p.out_coff + c0) * 2);
              dst[0] = make_uint4(pack2f(v[0], v[1], p.fp16), pack2f(v[2], v[3], p.fp16), pack2f(v[4], v[5], p.fp16),
                                  pack2f(v[6], v[7], p.fp16));
              dst[1] = make_uint4(pack2f(v[8], v[9], p.fp16), pack2f(v[10], v[11], p.fp16), pack2f(v[12], v[13], p.fp16),
                                  pack2f(v[14], v[15], p.fp16));
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty + ab);
      if (++ab == 2) { ab = 0; abph ^= 1; }
    }
    if (p.fp16 && p.overflow != nullptr && amax > 65504.f) atomicOr(p.overflow, 1);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, TMEM_COLS);
}

constexpr int kSmemC1 = kStagesA * 8192 + 4096 + 4096 + 512 + 512 + 1024;

template <int NT>
cudaError_t launch_nt(const C1Params& p, int grid, cudaStream_t st) {
  static unsigned long long attr = 0;
  cudaError_t e = smem_attr_once(conv1_umma_kernel<NT>, kSmemC1, &attr);
  if (e != cudaSuccess) return e;
  conv1_umma_kernel<NT><<<grid, kThreadsC1, kSmemC1, st>>>(p);
  return cudaGetLastError();
}

}  // namespace

// Returns true when the layer was taken (Cin = 1, 16-bit activations, Cout in {16, 32}); *err = launch status.
bool conv1_umma_try(const TView& in, const TView& out, const float* w, const float* bias, int lrelu, float alpha, int act_type,
                    cudaStream_t st, cudaError_t* err, int* overflow) {
  static const bool off = getenv("TOMOENC_NO_C1_UMMA") != nullptr;
  if (off || in.C != 1 || in.ctot != 1 || act_type == ACT_F32) return false;
  if (out.C != 16 && out.C != 32) return false;             // TMEM: 2 x kTZ x Cout <= 512 columns
  if (out.ctot % 8 != 0 || out.coff % 8 != 0 || kInTile * 2 > 4096) return false;
  C1Params p;
  p.in = static_cast<const uint16_t*>(in.ptr);
  p.out = static_cast<uint8_t*>(out.ptr);
  p.out_ctot = out.ctot; p.out_coff = out.coff;
  p.w = w; p.bias = bias;
  p.N = in.N; p.D = in.D; p.H = in.H; p.W = in.W;
  p.tiles_x = (in.W + 7) / 8; p.tiles_y = (in.H + 15) / 16; p.tiles_z = (in.D + kTZ - 1) / kTZ;
  p.total_tiles = static_cast<long long>(in.N) * p.tiles_z * p.tiles_y * p.tiles_x;
  p.lrelu = lrelu; p.alpha = alpha; p.fp16 = act_type == ACT_F16;
  p.overflow = overflow;
  if (p.total_tiles <= 0) { *err = cudaSuccess; return true; }
  if (p.total_tiles >= (1LL << 31)) return false;       // 32-bit tile indices in the kernel: let the caller fall back
  const int ctas = (out.C == 16 ? 2 : 1) * kNumSMs;
  const int grid = static_cast<int>(p.total_tiles < ctas ? p.total_tiles : ctas);
  *err = out.C == 16 ? launch_nt<16>(p, grid, st) : launch_nt<32>(p, grid, st);
  return true;
}

}  // namespace te

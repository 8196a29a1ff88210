// First layer of every graph (Conv3D(f, 3, 'same') on ONE input channel,
// /root/reference/tomo_encoders/neural_nets/Unet3D.py:117-122, autoencoders.py:168-173) as a Toeplitz product on the
// tensor cores -- no im2col builder at all (conv1_umma.cu spends its time in the integer pipe expanding 27 taps per voxel).
//
// With one input channel the contraction index can be the input x position itself:
//     out[z, y, x0 + xo, co] = sum_{dz, dy} sum_{x'} in[z + dz - 1, y + dy - 1, x'] * Wb[dz, dy][x' - x0][(xo, co)],
//     Wb[dz, dy][x' - x0][(xo, co)] = w[dz, dy, x' - x0 - xo + 1, co]   (a banded matrix; zero outside the 3 x taps).
// GEMM view: M = 128 rows = (16 z) x (8 y) of the raw 16-bit input, K = 16 consecutive input x positions, N = 4 output x
// positions x Cout channels; nine UMMAs (one per (dz, dy)) per tile of 16 x 8 x 4 voxels.
//
// Operand layout.  The single-channel input has 2-byte elements, so a haloed x window never starts on a 16-byte boundary
// (TMA faults on such boxes, tools/tma_copy_probe.cu).  Instead the input is brought in as ALIGNED 8-element x groups:
// "plane" p of a work unit is the TMA box (8 x, 8 + 2 y, 16 + 2 z) at x = 8 (p - 1), which lands in shared memory as
// [z][y][8 x] = rows of 16 bytes.  That is exactly the un-swizzled K-major UMMA layout: a core matrix = 8 consecutive y rows
// (128 contiguous bytes), the stride between 8-row groups (SBO) = one z step = 10 rows, and the stride between the two core
// matrices of a K = 16 instruction (LBO) = the distance between two planes.  A (dz, dy) tap is a different start address.
// An output tile of 4 x positions at x0 needs input x0 - 1 .. x0 + 4: for x0 = 8 g that is inside planes (g - 1, g), for
// x0 = 8 g + 4 inside planes (g, g + 1) -- two "phases" with their own banded weight tiles, one K = 16 step each.
// Out-of-range planes / rows are zero-filled by the TMA unit (the patch index is its own tensor dimension: Keras' per-patch
// 'same' padding).
//
// A work unit = (patch, 16 z, 8 y) walks the whole x extent (W / 4 tiles) on ONE input load of W / 8 + 2 planes; persistent
// CTAs: warp 0 TMA producer (double-buffered units), warps 1-2 UMMA issuers (alternate tiles), warps 4..19 two epilogue groups of eight (TMEM ->
// bias + LeakyReLU -> swizzled staging of two adjacent tiles -> coalesced 256-byte row stores), four accumulators in TMEM.
// Algorithmic bytes: 2 B read + 2 Cout B written per voxel; the kernel is bound by the output write.
#include "layers_simt.cuh"
#include "sm100_ptx.cuh"
#include "te_common.cuh"

#include <cstdlib>
#include <type_traits>

namespace te {
namespace {

constexpr int kXT = 4;                         // output x positions per tile
constexpr int kMaxPlanes = 10;                 // W <= 64
constexpr int kPlaneRows = 18 * 10;            // haloed (z, y) rows of one plane
constexpr int kPlaneStride = 2944;             // bytes between planes: 180 rows x 16 B rounded up to 128 (TMA destination alignment)
constexpr int kUnitBytes = kMaxPlanes * kPlaneStride;
constexpr int kAcc = 4;                        // accumulator stages in TMEM, each the two tiles of a pair (2 N columns)
constexpr int kThreadsTz = 640;                // warp 0 producer, warps 1-2 issuers, warp 3 idle, warps 4-19 epilogue

struct TzParams {
  const float* w;           // [27][Cout] fp32 (Keras (3,3,3,1,Cout), BN folded)
  const float* bias;
  uint8_t* out;             // (N, D, H, W, Cout) 16-bit, dense
  int N, D, H, W;
  int units_y, units_z;
  int total_units;
  int lrelu, fp16;
  float alpha;
  int* overflow;
};

__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      :
      : "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ uint16_t tz_bits(float v, int fp16) {
  if (fp16) { __half h = __float2half_rn(v); return *reinterpret_cast<uint16_t*>(&h); }
  __nv_bfloat16 h = __float2bfloat16_rn(v);
  return *reinterpret_cast<uint16_t*>(&h);
}
template <bool FP16>
__device__ __forceinline__ uint32_t tz_pack2(float a, float b) {
  if (FP16) { __half2 h = __floats2half2_rn(a, b); return *reinterpret_cast<uint32_t*>(&h); }
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

template <int NT>
struct TzCfg {
  static constexpr int N = kXT * NT;                       // UMMA N: (xo, co)
  static constexpr int B_TILE = N * 32;                    // one banded weight tile: [2 K core matrices][N rows][16 B]
  static constexpr int B_BYTES = 2 * 9 * B_TILE;           // two phases x nine (dz, dy)
  static constexpr int ROW_HALF = kXT * NT;                // bytes one epilogue thread writes per tile: 2 of the 4 x positions of its (z, y) row
  static_assert(ROW_HALF == 64, "the staging rows are one 64-byte swizzle atom wide");
  static constexpr int WBUF = 32 * 4 * ROW_HALF;           // staging buffer of a warp pair (one TMEM lane quarter, two adjacent tiles): 32 rows x 256 bytes
  static constexpr int STG_BYTES = 8 * WBUF;               // eight warp pairs
  static constexpr int IN_BYTES = 2 * kUnitBytes;
  static constexpr int TMEM_COLS = kAcc * 2 * N;
  static_assert(TMEM_COLS == 512, "four accumulator pairs fill TMEM");
  static constexpr int SMEM_BYTES = STG_BYTES + IN_BYTES + 128 + B_BYTES + 1024 /*bias + barriers*/ + 1024 /*align*/;
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget exceeded");
};

template <int NT>
__global__ void __launch_bounds__(kThreadsTz, 1)
conv1_tz_kernel(const __grid_constant__ CUtensorMap tmap_in, const TzParams p) {
  using C = TzCfg<NT>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stg = smem;                                              // [8 warp pairs][32 rows x 256 B]
  uint8_t* in_buf = smem + C::STG_BYTES;                            // [2][kMaxPlanes][kPlaneStride], 128-aligned
  uint8_t* b_tiles = in_buf + ((C::IN_BYTES + 127) / 128) * 128;    // [2 phases][9][2][N][16 B]
  float* s_bias = reinterpret_cast<float*>(b_tiles + C::B_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(b_tiles + C::B_BYTES + 256);
  uint64_t* in_full = bars;                   // [2]
  uint64_t* in_empty = in_full + 2;           // [2]
  uint64_t* acc_full = in_empty + 2;          // [kAcc]
  uint64_t* acc_empty = acc_full + kAcc;      // [kAcc], 8 arrivals (the warps of the owning epilogue group)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + kAcc);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int planes = p.W / 8 + 2;
  const int steps = p.W / kXT;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) { mbar_init(in_full + i, 1); mbar_init(in_empty + i, 2); }
    for (int i = 0; i < kAcc; ++i) { mbar_init(acc_full + i, 1); mbar_init(acc_empty + i, 8); }
    fence_barrier_init();
    prefetch_tmap(&tmap_in);
  }
  if (warp == 1) { tmem_alloc(tmem_slot, C::TMEM_COLS); tmem_relinquish(); }
  // banded weight tiles.  Phase 0 (x0 = 8 g): K slot k holds input x0 - 8 + k; phase 1 (x0 = 8 g + 4): x0 - 4 + k.
  // B[n = (xo, co)][k] = w[dz, dy, dx, co] with dx = x' - xo + 1 for the input offset x' = k - 8 (or k - 4) relative to x0.
  // One thread per B row (t = phase * 9 + dz * 3 + dy, n = (xo, co)): its 16 K slots are two 16-byte stores; at most three
  // slots are non-zero.  (An element-wise loop with its divisions and 2-byte stores cost every CTA ~16 us of prologue.)
  for (int row = threadIdx.x; row < 2 * 9 * C::N; row += kThreadsTz) {
    const int n = row % C::N, t = row / C::N;
    const int ph = t / 9, tap = t - ph * 9;
    const int xo = n / NT, co = n - xo * NT;
    const int k0 = (ph ? 4 : 8) + xo - 1;                         // K slot of dx = 0
    float wv[3];
#pragma unroll
    for (int dx = 0; dx < 3; ++dx) wv[dx] = __ldg(p.w + (tap * 3 + dx) * NT + co);
    uint32_t pkw[8];
#pragma unroll
    for (int k = 0; k < 16; k += 2) {
      const int d0 = k - k0, d1 = k + 1 - k0;
      const float v0 = d0 == 0 ? wv[0] : (d0 == 1 ? wv[1] : (d0 == 2 ? wv[2] : 0.f));
      const float v1 = d1 == 0 ? wv[0] : (d1 == 1 ? wv[1] : (d1 == 2 ? wv[2] : 0.f));
      pkw[k >> 1] = static_cast<uint32_t>(tz_bits(v0, p.fp16)) | (static_cast<uint32_t>(tz_bits(v1, p.fp16)) << 16);
    }
    uint8_t* dst = b_tiles + t * C::B_TILE + n * 16;
    *reinterpret_cast<uint4*>(dst) = make_uint4(pkw[0], pkw[1], pkw[2], pkw[3]);
    *reinterpret_cast<uint4*>(dst + C::N * 16) = make_uint4(pkw[4], pkw[5], pkw[6], pkw[7]);
  }
  for (int i = threadIdx.x; i < NT; i += kThreadsTz) s_bias[i] = __ldg(p.bias + i);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ============================== TMA producer: one unit = `planes` aligned boxes of 8 x ==============================
    uint32_t k = 0;
    for (int unit = blockIdx.x; unit < p.total_units; unit += gridDim.x, ++k) {
      const uint32_t b = k & 1u, ph = (k >> 1) & 1u;
      unsigned u = static_cast<unsigned>(unit);
      const int y0 = static_cast<int>(u % p.units_y) * 8; u /= p.units_y;
      const int z0 = static_cast<int>(u % p.units_z) * 16; u /= p.units_z;
      const int n = static_cast<int>(u);
      mbar_wait(in_empty + b, ph ^ 1u);
      if (elect_one()) {
        mbar_expect_tx(in_full + b, static_cast<uint32_t>(planes) * (kPlaneRows * 16));
        for (int pl = 0; pl < planes; ++pl)
          tma_load_4d(in_buf + b * kUnitBytes + pl * kPlaneStride, &tmap_in, in_full + b, 8 * (pl - 1), y0 - 1, z0 - 1, n);
      }
      __syncwarp();
    }
  } else if (warp == 1 || warp == 2) {
    // ============================== UMMA issuers: warp 1 takes the even tile pairs (accumulators 0, 2), warp 2 the odd ones ==============================
    // One issuer is bound by its own instruction stream (ncu: ~130 instructions per tile at ~10 cycles each, the barriers
    // never waited on), so two warps on different schedulers share the tiles, and the per-tap descriptors are the tile's
    // base descriptor plus a compile-time constant in the start-address field (low word; the high word never changes).
    const uint32_t par = static_cast<uint32_t>(warp - 1);
    const uint32_t idesc = make_idesc_f16(128, C::N, p.fp16 ? 0u : 1u);
    const uint64_t a_proto = make_smem_desc(0, kPlaneStride, 160, kLayoutNone);
    const uint64_t b_proto = make_smem_desc(smem_u32(b_tiles), C::N * 16, 128, kLayoutNone);
    const uint32_t a_hi = static_cast<uint32_t>(a_proto >> 32), b_hi = static_cast<uint32_t>(b_proto >> 32);
    const uint32_t a_lo_proto = static_cast<uint32_t>(a_proto), b_lo0 = static_cast<uint32_t>(b_proto);
    uint32_t k = 0;                              // unit counter of this CTA
    const int pairs = steps >> 1;                // tile pairs per unit (W / 8)
    for (int unit = blockIdx.x; unit < p.total_units; unit += gridDim.x, ++k) {
      const uint32_t b = k & 1u;
      mbar_wait(in_full + b, (k >> 1) & 1u);
      tc_fence_after();
      const uint32_t a_lo_unit = a_lo_proto | ((smem_u32(in_buf + b * kUnitBytes) >> 4) & 0x3FFFu);
#pragma unroll 1
      for (int g = static_cast<int>((par ^ (k * static_cast<uint32_t>(pairs))) & 1u); g < pairs; g += 2) {
        // ONE handshake per tile pair (x group g: the phase-0 and phase-1 tiles side by side in one 128-column accumulator);
        // per-tile handshakes cost ~350 cycles each even with nothing else to do
        const uint32_t pi = k * static_cast<uint32_t>(pairs) + static_cast<uint32_t>(g);   // pair counter of this CTA: pi & 1 == par
        const uint32_t s = pi % kAcc, sph = (pi / kAcc) & 1u;
        mbar_wait(acc_empty + s, sph ^ 1u);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_lo = a_lo_unit + static_cast<uint32_t>(g) * (kPlaneStride >> 4);      // phase 0: planes (g, g + 1); phase 1: (g + 1, g + 2)
          const uint32_t acc = tmem_base + s * (2 * C::N);
#pragma unroll
          for (int j = 0; j < 18; ++j) {
            const int phs = j / 9, tap = j % 9;
            asm volatile(
                "{\n\t.reg .b64 da, db;\n\t.reg .pred p;\n\t"
                "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\tsetp.ne.b32 p, %6, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}\n"
                :
                : "r"(acc + phs * C::N), "r"(a_lo + static_cast<uint32_t>(phs * (kPlaneStride >> 4) + (tap / 3) * 10 + tap % 3)), "r"(a_hi),
                  "r"(b_lo0 + static_cast<uint32_t>(j * (C::B_TILE >> 4))), "r"(b_hi), "r"(idesc), "r"(tap ? 1u : 0u)
                : "memory");
          }
          umma_commit(acc_full + s);
        }
        __syncwarp();
      }
      // this issuer's MMAs on the unit's input are all committed behind this one: the buffer is free when both have arrived
      if (elect_one()) umma_commit(in_empty + b);
      __syncwarp();
    }
  } else if (warp >= 4) {
    // ============================== epilogue: two groups of eight warps; group eg takes the tile pairs with (pair counter & 1) == eg ==============================
    // Two warps share a TMEM lane quarter and take 32 accumulator columns each (2 of a tile's 4 x positions): the epilogue
    // is a serial chain of ~4 instructions per element, so it wants many warps with few elements each (8 warps x 64
    // elements ran at 36 % issue utilisation, stalled on fixed-latency dependencies).
    // Output path: the two tiles of a pair are adjacent in x, so a (z, y) row of the pair is 8 voxels x 16 channels = 256
    // contiguous bytes.  The warp pair stages its 32 rows in shared memory (XOR-swizzled 16-byte chunks: conflict-free both
    // ways) and writes them back with fully coalesced 128-bit stores, a half-warp per row.  (TMA box stores of 128-byte
    // rows were paced by the store engine at ~14 B/clk/SM -- the previous round's store had not even finished READING its
    // buffer one tile later, profiles/r02_notes.md -- while plain contiguous stores fill HBM at 6.2-7.2 TB/s,
    // profiles/write_probe_r02.txt.)
    const int e = warp - 4, q = e & 3, half = (e >> 2) & 1, eg = e >> 3;
    const int pair = eg * 4 + q;
    const uint32_t buf = smem_u32(stg + static_cast<size_t>(pair) * C::WBUF);
    const uint32_t rowoff = static_cast<uint32_t>(lane) * 256u;
    const uint32_t sw = static_cast<uint32_t>(lane) & 7u;                   // chunk ^= row & 7
    const float al = p.lrelu ? p.alpha : 1.f;                              // 0 <= alpha <= 1 (checked at launch): lrelu(x) = max(x, alpha x)
    float bs[NT];
#pragma unroll
    for (int j = 0; j < NT; ++j) bs[j] = s_bias[j];
    float amax = 0.f;
    uint32_t k = 0;                                                        // unit counter of this CTA
    const size_t row_bytes = static_cast<size_t>(p.W) * NT * 2;
    const int hl = lane >> 4, c = lane & 15;
    auto run = [&](auto fp_tag) {
      constexpr bool FP16 = decltype(fp_tag)::value;
      for (int unit = blockIdx.x; unit < p.total_units; unit += gridDim.x, ++k) {
        unsigned u = static_cast<unsigned>(unit);
        const int y0 = static_cast<int>(u % p.units_y) * 8; u /= p.units_y;
        const int z0 = static_cast<int>(u % p.units_z) * 16; u /= p.units_z;
        const int n = static_cast<int>(u);
        uint8_t* gunit = p.out + ((static_cast<size_t>(n) * p.D + z0 + 4 * q) * p.H + y0) * row_bytes + c * 16;
        // this group's tile pairs of the unit: the pairs of the CTA alternate between the groups (pair counter pi & 1 == eg)
#pragma unroll 1
        for (int tp = static_cast<int>((static_cast<uint32_t>(eg) ^ (k * static_cast<uint32_t>(steps >> 1))) & 1u); 2 * tp < steps; tp += 2) {
          asm volatile("bar.sync %0, 64;" ::"r"(1 + pair) : "memory");      // the warp pair has read the previous tile pair out
          const uint32_t pi = k * static_cast<uint32_t>(steps >> 1) + static_cast<uint32_t>(tp);
          const uint32_t s = pi % kAcc, sph = (pi / kAcc) & 1u;
          mbar_wait(acc_full + s, sph);
          tc_fence_after();
#pragma unroll
          for (int tt = 0; tt < 2; ++tt) {
            uint32_t r[32];
            tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + s * (2 * C::N) + tt * C::N + half * 32, r);
            tmem_ld_wait();
            if (tt == 1) {
              // both tiles' accumulator columns of this warp are in registers: hand the stage back before the math and the store
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(acc_empty + s);
            }
            uint32_t pk[16];
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              // packed fp32 pairs (FADD2 / FMUL2): bias, then lrelu(v) = max(v, alpha v); the products are exact fp32
              // operations, identical to the scalar ones
              uint64_t v2, w2, b2, a2;
              asm("mov.b64 %0, {%1, %2};" : "=l"(v2) : "r"(r[j]), "r"(r[j + 1]));
              asm("mov.b64 %0, {%1, %2};" : "=l"(b2) : "f"(bs[j % NT]), "f"(bs[(j + 1) % NT]));
              asm("mov.b64 %0, {%1, %1};" : "=l"(a2) : "f"(al));
              asm("add.rn.f32x2 %0, %1, %2;" : "=l"(v2) : "l"(v2), "l"(b2));
              asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(w2) : "l"(v2), "l"(a2));
              float v0, v1, w0, w1;
              asm("mov.b64 {%0, %1}, %2;" : "=f"(v0), "=f"(v1) : "l"(v2));
              asm("mov.b64 {%0, %1}, %2;" : "=f"(w0), "=f"(w1) : "l"(w2));
              const float x0 = fmaxf(v0, w0), x1 = fmaxf(v1, w1);
              if (FP16) amax = fmaxf(amax, fmaxf(fabsf(x0), fabsf(x1)));
              pk[j >> 1] = tz_pack2<FP16>(x0, x1);
            }
            const uint32_t c0 = static_cast<uint32_t>(tt * 8 + half * 4);
#pragma unroll
            for (int ch = 0; ch < 4; ++ch)      // row `lane` (= 4 z x 8 y of this lane quarter), 16-byte chunks c0 .. c0 + 3 of 16
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(buf + rowoff + (((c0 + ch) ^ sw) << 4)),
                           "r"(pk[4 * ch]), "r"(pk[4 * ch + 1]), "r"(pk[4 * ch + 2]), "r"(pk[4 * ch + 3])
                           : "memory");
          }
          asm volatile("bar.sync %0, 64;" ::"r"(1 + pair) : "memory");      // both tiles of the pair are staged
          uint8_t* gbase = gunit + static_cast<size_t>(2 * tp) * (kXT * NT * 2);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int row = 2 * (half * 8 + j) + hl;                        // 0..31 = (z - z0 - 4 q) * 8 + (y - y0)
            uint4 v;
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                         : "r"(buf + static_cast<uint32_t>(row) * 256u + ((static_cast<uint32_t>(c) ^ (static_cast<uint32_t>(row) & 7u)) << 4)));
            *reinterpret_cast<uint4*>(gbase + (static_cast<size_t>(row >> 3) * p.H + (row & 7)) * row_bytes) = v;
          }
        }
      }
    };
    if (p.fp16) run(std::true_type{}); else run(std::false_type{});
    if (p.fp16 && p.overflow != nullptr && amax > 65504.f) atomicOr(p.overflow, 1);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, C::TMEM_COLS);
}

typedef CUresult (*EncodeTiledFnTz)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFnTz get_encode_tz() {
  static EncodeTiledFnTz fn = nullptr;
  if (!fn) {
    cudaDriverEntryPointQueryResult q;
    void* ptr = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess) fn = (EncodeTiledFnTz)ptr;
  }
  return fn;
}

}  // namespace

// The shapes the Toeplitz kernel takes: Cin = 1, 16-bit activations, Cout = 16 written densely (no channel slice),
// W a multiple of 8 up to 64, H a multiple of 8, D a multiple of 16.
bool conv1_tz_applies(const TView& in, const TView& out, int act_type) {
  static const bool off = getenv("TOMOENC_NO_C1_TZ") != nullptr;
  if (off || in.C != 1 || in.ctot != 1 || act_type == ACT_F32) return false;
  if (out.C != 16 || out.ctot != out.C || out.coff != 0) return false;
  if (in.W % 8 != 0 || in.W > 8 * (kMaxPlanes - 2) || in.H % 8 != 0 || in.D % 16 != 0) return false;
  if (in.W * 2 * 16 /*W * Cout * 2 B*/ % 16 != 0) return false;
  const long long units = static_cast<long long>(in.N) * (in.D / 16) * (in.H / 8);
  return units > 0 && units < (1LL << 31);
}

bool conv1_tz_try(const TView& in, const TView& out, const float* w, const float* bias, int lrelu, float alpha, int act_type,
                  cudaStream_t st, cudaError_t* err, int* overflow) {
  if (!conv1_tz_applies(in, out, act_type)) return false;
  if (lrelu && !(alpha >= 0.f && alpha <= 1.f)) return false;             // epilogue: lrelu(x) = max(x, alpha x)
  if ((reinterpret_cast<uintptr_t>(in.ptr) | reinterpret_cast<uintptr_t>(out.ptr)) & 15) return false;
  EncodeTiledFnTz enc = get_encode_tz();
  if (!enc) return false;
  constexpr int NT = 16;
  const bool fp16 = act_type == ACT_F16;
  const CUtensorMapDataType dt = fp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  CUtensorMap tin;
  {
    cuuint64_t gdim[4] = {(cuuint64_t)in.W, (cuuint64_t)in.H, (cuuint64_t)in.D, (cuuint64_t)in.N};
    cuuint64_t gstr[3] = {(cuuint64_t)in.W * 2, (cuuint64_t)in.H * in.W * 2, (cuuint64_t)in.D * in.H * in.W * 2};
    cuuint32_t box[4] = {8, 10, 18, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    if (enc(&tin, dt, 4, in.ptr, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return false;
  }
  TzParams p;
  p.w = w; p.bias = bias;
  p.out = static_cast<uint8_t*>(out.ptr);
  p.N = in.N; p.D = in.D; p.H = in.H; p.W = in.W;
  p.units_y = in.H / 8; p.units_z = in.D / 16;
  p.total_units = in.N * p.units_z * p.units_y;
  p.lrelu = lrelu; p.alpha = alpha; p.fp16 = fp16 ? 1 : 0;
  p.overflow = overflow;
  static unsigned long long attr = 0;
  cudaError_t e = smem_attr_once(conv1_tz_kernel<NT>, TzCfg<NT>::SMEM_BYTES, &attr);
  if (e != cudaSuccess) { *err = e; return true; }
  const int grid = p.total_units < kNumSMs ? p.total_units : kNumSMs;
  conv1_tz_kernel<NT><<<grid, kThreadsTz, TzCfg<NT>::SMEM_BYTES, st>>>(tin, p);
  *err = cudaGetLastError();
  return true;
}

}  // namespace te

// Implicit-GEMM 3-D convolution on the 5th-generation tensor cores (tcgen05 / TMEM / TMA), sm_100a.
//
// Replaces the cuDNN calls behind L.Conv3D(f, 3, padding="same") + BN + LeakyReLU
// (/root/reference/tomo_encoders/neural_nets/Unet3D.py:44-81) and L.Conv3DTranspose(C, 2, strides=p)
// + LeakyReLU (Unet3D.py:175-177), plus the final 1x1x1 sigmoid conv (Unet3D.py:287) as a fused
// epilogue.
//
// GEMM view: M = 128 output voxels (a 16(y) x 8(x) patch of one z plane), N = NT output channels,
// K = taps x Cin, accumulated in fp32 in TMEM.  One CTA tile is TZ consecutive z planes; its
// haloed input box (TZ+2, 18, 10, CB channels) is brought into shared memory by ONE 5-D TMA load per
// CB-channel chunk -- out-of-bounds coordinates are zero-filled by the TMA unit, and because the
// patch index is its own tensor dimension this is exactly Keras' per-patch "same" padding.  Each of
// the 27 taps is then just a different START ADDRESS of the A-operand shared-memory descriptor into
// that one box (the swizzle is a function of the absolute shared-memory address, so shifted windows
// stay consistent -- verified on hardware by tools/umma_probe.cu, profiles/umma_probe_r01_run*.txt).
//
// Warp roles (224 threads): warp 0 = TMA producer of the input boxes, warp 6 = bulk-copy producer of
// the pre-swizzled weight slabs, warp 1 = UMMA issuer (warp-uniform loop, one elected lane issues),
// warps 2..5 = epilogue (TMEM -> registers -> bias + LeakyReLU -> 16-bit NDHWC store, optionally
// pixel-shuffled for the transposed conv or reduced by the fused 1x1x1 head).  Persistent CTAs loop over tiles; the accumulator is double-buffered in TMEM so the
// epilogue of tile i overlaps the MMAs of tile i+1.
#include "conv_umma.cuh"
#include "sm100_ptx.cuh"
#include "te_common.cuh"

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cstdlib>
#include <type_traits>
#include <vector>

namespace te {

namespace {

constexpr int kThreads = 224;

struct KParams {
  const uint8_t* w_packed;
  const float* bias;
  uint8_t* out;
  int out_ctot, out_coff;
  int N, D, H, W;
  int kchunks;          // Cin / CB
  int k1_chunks;        // chunks read through the first tensor map (== kchunks for a single input)
  int ngroups, taps_g;  // weight slabs per chunk, taps per slab (9 x 3 for a 3x3x3 conv, 1 x 1 for a transposed-conv sub-position)
  int halo;             // 1 for the 3x3x3 conv, 0 otherwise
  int nsub;             // 1 or 8
  int up_stride;
  int n_ntiles;         // Cout / NT
  int tiles_x, tiles_y, tiles_z;
  long long total_tiles;
  int act_lrelu;
  float alpha;
  int fp16;
  const float* head_w;
  float head_b;
  uint8_t* head_labels;
  float* head_probs;
  float* head_logits;
  uint8_t* pool_out;    // fused MaxPool3D(2) output (nullptr = none)
  int pool_ctot, pool_coff;
  int* overflow;        // fp16 only: set to 1 when an activation leaves the finite fp16 range (nullptr = not tracked)
  int tma_out;          // un-haloed kernels: epilogue writes through the output tensor maps
  int reuse_ok;         // host: enough spatial tiles to fill the GPU when the (Cout tile, sub) pairs are looped inside a CTA
};

struct TileCoord { int ntile, sub, n, z0, y0, x0; };
struct OutMaps { CUtensorMap m[8]; };

__device__ __forceinline__ void tma_store_5d(const CUtensorMap* m, const void* src, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];"
               :
               : "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
               : "memory");
}

// `reuse` (un-haloed kernels whose channel chunks all fit the input ring): the outer index is the
// spatial tile, the inner index j enumerates the (Cout tile, sub-position) pairs that share its
// input boxes.  Otherwise j = 0 and t enumerates everything.
__device__ __forceinline__ TileCoord decode_tile(const KParams& p, long long tile, int TZ, bool reuse = false, int j = 0) {
  TileCoord c;
  // 32-bit index arithmetic (the launch checks total_tiles < 2^31): the 64-bit divisions of the first version cost
  // ~100 instructions each and sat on the critical path of every role at every tile
  unsigned t = static_cast<unsigned>(tile);
  const unsigned tx = p.tiles_x, ty = p.tiles_y, tz = p.tiles_z;
  if (reuse) {
    c.x0 = static_cast<int>(t % tx) * 8;  t /= tx;
    c.y0 = static_cast<int>(t % ty) * 16; t /= ty;
    c.z0 = static_cast<int>(t % tz) * TZ; t /= tz;
    c.n = static_cast<int>(t);
    c.sub = j % p.nsub;
    c.ntile = j / p.nsub;
    return c;
  }
  c.x0 = static_cast<int>(t % tx) * 8;  t /= tx;
  c.y0 = static_cast<int>(t % ty) * 16; t /= ty;
  c.z0 = static_cast<int>(t % tz) * TZ; t /= tz;
  c.n = static_cast<int>(t % static_cast<unsigned>(p.N)); t /= static_cast<unsigned>(p.N);
  // sub-positions of a transposed conv are the slowest index: measured faster on B200 than
  // interleaving the 8 sub-position CTAs of one output block (623 vs 706 us for dec0.upconv)
  c.sub = static_cast<int>(t % static_cast<unsigned>(p.nsub)); t /= static_cast<unsigned>(p.nsub);
  c.ntile = static_cast<int>(t);
  return c;
}

__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_le1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all_groups() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void named_barrier(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// 16 consecutive floats from shared memory as four 128-bit loads (the pointer arithmetic on the aligned dynamic buffer hides
// the address space from the compiler, which would otherwise emit 16 generic loads)
__device__ __forceinline__ void lds_f16x(const float* p, float (&b)[16]) {
  const uint32_t a = smem_u32(p);
#pragma unroll
  for (int i = 0; i < 4; ++i)
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(b[4 * i]), "=f"(b[4 * i + 1]), "=f"(b[4 * i + 2]), "=f"(b[4 * i + 3])
                 : "r"(a + 16 * i));
}

// element-wise max of two packed 16-bit pairs (fp16 or bf16)
__device__ __forceinline__ uint32_t max2_16(uint32_t a, uint32_t b, int fp16) {
  if (fp16) {
    const __half2 r = __hmax2(*reinterpret_cast<const __half2*>(&a), *reinterpret_cast<const __half2*>(&b));
    return *reinterpret_cast<const uint32_t*>(&r);
  }
  const __nv_bfloat162 r = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&a), *reinterpret_cast<const __nv_bfloat162*>(&b));
  return *reinterpret_cast<const uint32_t*>(&r);
}
__device__ __forceinline__ uint32_t pack2(float a, float b, int fp16) {
  if (fp16) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// Per-CTA weight image of one (dy, dx) slab in CTA-pair mode (cta_group::2: each CTA supplies HALF of the B rows of every
// instruction, at the same shared-memory offset in both CTAs).  The dz-fused schedule issues five different row ranges
// of the 3 * NT-row slab [t = 0 (dz = +1) | t = 1 | t = 2], plus the three single-tap ranges of a tile's first slab; the
// image holds this CTA's half of each range (rows of CB * 2 bytes, offsets in rows):
//   V96  rows [0, 3 NT)        -> 3 NT / 2 rows at 0            (interior input planes)
//   V64H rows [NT, 3 NT)       -> NT rows at 3 NT / 2           (input plane 0 of a tile)
//   V64L rows [0, 2 NT)        -> NT rows at 5 NT / 2           (input plane TZ - 1)
//   V32L / V32M / V32H rows [0, NT) / [NT, 2 NT) / [2 NT, 3 NT) -> NT / 2 rows each at 7 NT / 2, 4 NT, 9 NT / 2
struct PairImage {
  template <int NT> __host__ __device__ static constexpr int v96() { return 0; }
  template <int NT> __host__ __device__ static constexpr int v64h() { return 3 * NT / 2; }
  template <int NT> __host__ __device__ static constexpr int v64l() { return 5 * NT / 2; }
  template <int NT> __host__ __device__ static constexpr int v32(int t) { return 7 * NT / 2 + t * (NT / 2); }
  template <int NT> __host__ __device__ static constexpr int rows() { return 5 * NT; }
};

template <int CB, int NT, int TZ, int HALO, bool PAIR = false>
struct Cfg {
  static constexpr int ROWB = CB * 2;                                   // bytes of one voxel row of a chunk
  static constexpr uint32_t LAYOUT = CB == 16 ? kLayoutSW32 : (CB == 32 ? kLayoutSW64 : kLayoutSW128);
  static constexpr int KSTEPS = CB / 16;                                // UMMA K = 16 for 16-bit operands
  static constexpr int A_STAGE = (((TZ + 2 * HALO) * (16 + 2 * HALO) * (8 + 2 * HALO) * ROWB) + 1023) / 1024 * 1024;
  static_assert(!PAIR || (HALO && 3 * NT <= 256 && NT % 32 == 0), "CTA pairs: dz-fused 3x3x3 convs with NT a multiple of 32");
  // bytes of one weight slab (3 taps of a conv; in pair mode this CTA's image of the slab, see PairImage)
  static constexpr int SLAB = PAIR ? PairImage::rows<NT>() * ROWB : (HALO ? 3 : 1) * NT * ROWB;
  // Weight slabs per pipeline stage.  Every stage boundary costs a barrier wait plus a drain of the
  // UMMA issue queue (ncu: the uniform registers of queued UTCHMMAs are recycled there), so the
  // thin-channel layers, whose slabs are small, take 3 or 9 slabs per stage -> bursts of 36 / 108 UMMAs.
  static constexpr int G = !HALO ? 1 : (NT >= 128 ? 1 : (PAIR ? 9 : (SLAB * 9 <= 28 * 1024 ? 9 : 3)));
  static constexpr int NGROUPS = HALO ? 9 / G : 1;                      // stages per channel chunk
  static constexpr int W_STAGE = ((G * SLAB) + 1023) / 1024 * 1024;
  // input-box ring: the haloed boxes of the 3x3x3 conv are large (2 stages); the un-haloed boxes of
  // the transposed conv are small and each feeds few MMAs, so more of them are kept in flight
  static constexpr int NUM_A = HALO ? 2 : (A_STAGE <= 16 * 1024 ? 8 : 4);
  // plane-fused MMA schedule (see the UMMA issuer): needs 3 * NT <= 256 result columns per instruction
  static constexpr bool FUSE = HALO && (3 * NT <= 256);
  static constexpr int ACC_COLS = TZ * NT;                              // fp32 columns of one accumulator set
  static constexpr int TMEM_COLS_RAW = 2 * ACC_COLS;
  static constexpr int TMEM_COLS = TMEM_COLS_RAW <= 32 ? 32 : TMEM_COLS_RAW <= 64 ? 64 : TMEM_COLS_RAW <= 128 ? 128
                                   : TMEM_COLS_RAW <= 256 ? 256 : 512;
  static_assert(TMEM_COLS_RAW <= 512, "accumulators do not fit TMEM");
  // the un-haloed kernels (transposed / pointwise convs: 8 x more output per input voxel, few MMAs) are
  // epilogue-bound (ncu: tensor pipe 6 %, DRAM 29 % on dec0.upconv), so they run two groups of four
  // epilogue warps, each taking half of the tile's z planes
  // ... and so are the thin haloed layers (NT <= 32 with one or two channel chunks: ~1.5 k cycles of epilogue per
  // plane against ~0.6 k of UMMAs per plane and chunk)
  static constexpr int EPI_GROUPS = (HALO && NT > 32) ? 1 : 2;
  static constexpr int THREADS = kThreads + (EPI_GROUPS - 1) * 128;
  // TMA-store epilogue: per epilogue group STG_NBUF staging buffers of 128 voxels x OUT_BOX channels.  The weight
  // ring gives up a stage where that is what it takes to fit at least one buffer per group.
  static constexpr int OUT_BOX = NT < 64 ? NT : 64;
  static constexpr int STG_BUF = 128 * OUT_BOX * 2;
  static constexpr int ROOM0 = 227 * 1024 - 2048 - NUM_A * A_STAGE;
  static constexpr int W_FIT2 = (ROOM0 - EPI_GROUPS * 2 * STG_BUF) / W_STAGE;   // with double-buffered staging
  static constexpr int W_FIT1 = (ROOM0 - EPI_GROUPS * STG_BUF) / W_STAGE;       // with single-buffered staging
  static constexpr int STG_NBUF = (W_FIT2 >= 2 && !PAIR) ? 2 : 1;
  static constexpr int W_FIT = STG_NBUF == 2 ? W_FIT2 : W_FIT1;
  static constexpr int NUM_W = W_FIT >= 4 ? 4 : W_FIT;
  static_assert(NUM_W >= 2, "weight ring needs two stages");
  static constexpr int STG_BYTES = EPI_GROUPS * STG_NBUF * STG_BUF;
  static constexpr int BIAS_BYTES = 1024;   // fp32 bias of up to 256 output channels behind the barriers, then 128 head weights
  static constexpr int SMEM_BYTES = STG_BYTES + NUM_A * A_STAGE + NUM_W * W_STAGE + 2048 /*barriers + bias*/ + 1024 /*align*/;
  static_assert((2 * NUM_A + 3 * NUM_W + 4) * 8 + 4 <= 512, "barriers overlap the bias staging area");
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget exceeded");
};

template <int CB, int NT, int TZ, int HALO, bool PAIR = false>
__global__ void __launch_bounds__((Cfg<CB, NT, TZ, HALO, PAIR>::THREADS), 1)
conv_umma_kernel(const __grid_constant__ CUtensorMap tmap_in, const __grid_constant__ CUtensorMap tmap_in2,
                 const __grid_constant__ OutMaps omaps, const KParams p) {
  using C = Cfg<CB, NT, TZ, HALO, PAIR>;
  // CTA-pair mode (launched as clusters of 2): the two CTAs of a pair work on two different spatial tiles with the SAME
  // weights; the leader (cluster rank 0) issues tcgen05.mma.cta_group::2 instructions of M = 256 whose B rows are read half
  // from each CTA's shared memory (per-CTA smem operand traffic 32 + N/8 instead of 32 + N/4 cycles per instruction: the
  // N = 96 dz-fused instruction becomes math-bound, tools/umma2_probe.cu).  Every CTA keeps its own producers and epilogue.
  const uint32_t pair_rank = PAIR ? cluster_ctarank() : 0u;
  const bool leader = pair_rank == 0;
  const long long tile0 = PAIR ? (2LL * (blockIdx.x >> 1) + pair_rank) : static_cast<long long>(blockIdx.x);
  const long long tile_step = PAIR ? static_cast<long long>(gridDim.x & ~1u) : static_cast<long long>(gridDim.x);
  constexpr int kNumAStages = C::NUM_A;
  constexpr int BX = 8 + 2 * HALO, BY = 16 + 2 * HALO, BZ = TZ + 2 * HALO;   // input box (voxels)
  constexpr int kNumWStages = C::NUM_W;
  constexpr int G = C::G;                  // weight slabs per pipeline stage
  constexpr int NGROUPS = C::NGROUPS;      // weight stages per channel chunk
  constexpr int TAPS_G = HALO ? 3 : 1;     // taps per slab: the three dx of a (dz, dy) row (plane-fused: the three dz of a (dy, dx))
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_stage = smem;
  uint8_t* w_stage = smem + kNumAStages * C::A_STAGE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(w_stage + kNumWStages * C::W_STAGE);
  uint64_t* a_full = bars;                       // [kNumAStages]
  uint64_t* a_empty = a_full + kNumAStages;      // [kNumAStages]
  uint64_t* w_full = a_empty + kNumAStages;      // [kNumWStages]
  uint64_t* w_empty = w_full + kNumWStages;      // [kNumWStages]
  uint64_t* acc_full = w_empty + kNumWStages;    // [2]
  uint64_t* acc_empty = acc_full + 2;            // [2]
  uint64_t* w_land = acc_empty + 2;              // [kNumWStages] pair mode, rank 1: its own weight copies land here first
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_land + kNumWStages);
  uint8_t* stg = reinterpret_cast<uint8_t*>(bars) + 2048;    // [EPI_GROUPS][STG_NBUF][STG_BUF], 1024-aligned
  // bias of all output channels: the epilogues read it 16 channels at a time as 128-bit shared-memory loads (it was one
  // global load per accumulator element: 30 % of the epilogue's stall samples in the thin layers)
  float* s_bias = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 512);
  const bool bias_all = p.n_ntiles * NT * 4 <= C::BIAS_BYTES;    // else the epilogue stages one NT slice per tile
  if (bias_all)
    for (int i = threadIdx.x; i < p.n_ntiles * NT; i += blockDim.x) s_bias[i] = __ldg(p.bias + i);
  float* s_head = s_bias + 256;                                  // weights of the fused 1x1x1 head (one Cout tile)
  if (p.head_w != nullptr)
    for (int i = threadIdx.x; i < NT && i < 128; i += blockDim.x) s_head[i] = __ldg(p.head_w + i);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    // pair mode: the "full" barriers the issuer waits on live in the leader and collect one arrival per CTA; the
    // accumulator-empty barriers collect the epilogue warps of both CTAs
    constexpr uint32_t kFullCount = PAIR ? 2 : 1;
    for (int i = 0; i < kNumAStages; ++i) { mbar_init(a_full + i, kFullCount); mbar_init(a_empty + i, 1); }
    for (int i = 0; i < kNumWStages; ++i) { mbar_init(w_full + i, kFullCount); mbar_init(w_empty + i, 1); mbar_init(w_land + i, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(acc_full + i, 1); mbar_init(acc_empty + i, kFullCount * 4 * C::EPI_GROUPS); }
    fence_barrier_init();
    prefetch_tmap(&tmap_in);
  }
  if (warp == 1) {
    if constexpr (PAIR) { tmem_alloc2(tmem_slot, C::TMEM_COLS); tmem_relinquish2(); }
    else { tmem_alloc(tmem_slot, C::TMEM_COLS); tmem_relinquish(); }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (PAIR) cluster_sync_all();        // the peer's barriers exist before anybody signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  constexpr uint32_t a_bytes = static_cast<uint32_t>(BZ * BY * BX * C::ROWB);
  constexpr uint32_t slab_bytes = static_cast<uint32_t>(C::SLAB);
  constexpr uint32_t w_bytes = static_cast<uint32_t>(G) * slab_bytes;     // one stage
  // all the weights of the layer fit the ring: load them once, never release them
  const bool w_resident = p.n_ntiles == 1 && p.nsub == 1 && p.kchunks * NGROUPS <= kNumWStages;
  // input-box reuse across the (Cout tile, sub-position) pairs of a spatial tile (transposed / pointwise convs):
  // a transposed conv would otherwise re-read its input 8 times (ncu: 1072 MB for a 134 MB tensor)
  const bool reuse = !HALO && p.reuse_ok && p.kchunks <= kNumAStages && p.n_ntiles * p.nsub > 1;
  const int J = reuse ? p.n_ntiles * p.nsub : 1;
  const long long outer_tiles = reuse ? p.total_tiles / J : p.total_tiles;

  if (warp == 0) {
    // ============================== TMA producer ==============================
    // The whole warp runs the (warp-uniform) control flow; one elected lane issues the copies and
    // runs ahead of the MMAs by the depth of the ring.
    uint32_t as = 0, aph = 0;
    for (long long tile = tile0; tile < outer_tiles; tile += tile_step) {
      const TileCoord tc = decode_tile(p, tile, TZ, reuse, 0);
#pragma unroll 1
      for (int kc = 0; kc < p.kchunks; ++kc) {
        mbar_wait(a_empty + as, aph ^ 1);
        if (elect_one()) {
          // channel chunks [0, k1) come from the first input tensor, the rest from the second
          // (planar skip concatenation: the two halves are separate dense tensors)
          const bool second = kc >= p.k1_chunks;
          if constexpr (PAIR) {
            // this CTA's box into its own shared memory, its bytes credited to the LEADER's barrier
            mbar_expect_tx_leader(a_full + as, a_bytes);
            tma_load_5d_pair(a_stage + as * C::A_STAGE, second ? &tmap_in2 : &tmap_in, a_full + as,
                             (second ? kc - p.k1_chunks : kc) * CB, tc.x0 - HALO, tc.y0 - HALO, tc.z0 - HALO, tc.n);
          } else {
            mbar_expect_tx(a_full + as, a_bytes);
            tma_load_5d(a_stage + as * C::A_STAGE, second ? &tmap_in2 : &tmap_in, a_full + as,
                        (second ? kc - p.k1_chunks : kc) * CB, tc.x0 - HALO, tc.y0 - HALO, tc.z0 - HALO, tc.n);
          }
        }
        __syncwarp();
        if (++as == kNumAStages) { as = 0; aph ^= 1; }
      }
    }
  } else if (warp == 6) {
    // ============================== weight-slab producer ==============================
    uint32_t ws = 0, wph = 0;
    for (long long tile = tile0; tile < outer_tiles; tile += tile_step) {
#pragma unroll 1
      for (int j = 0; j < J; ++j) {
        const TileCoord tc = decode_tile(p, tile, TZ, reuse, j);
        // pair mode: one image per CTA rank (this CTA's half of every B row range)
        const uint8_t* wsrc = p.w_packed + ((static_cast<size_t>(tc.ntile) * p.nsub + tc.sub) * (PAIR ? 2 : 1) + pair_rank) *
                                               p.kchunks * NGROUPS * static_cast<size_t>(w_bytes);
        const int nstages = p.kchunks * NGROUPS;
#pragma unroll 1
        for (int g = 0; g < nstages; ++g) {
          mbar_wait(w_empty + ws, wph ^ 1);
          if (elect_one()) {
            // (a plain bulk copy signals a barrier of its own CTA: the peer's copies land on w_land and its otherwise idle
            // issuer warp forwards the arrival to the leader's w_full)
            uint64_t* land = (PAIR && !leader) ? w_land + ws : w_full + ws;
            mbar_expect_tx(land, w_bytes);
            bulk_load(w_stage + ws * C::W_STAGE, wsrc + static_cast<size_t>(g) * w_bytes, w_bytes, land);
          }
          __syncwarp();
          if (++ws == kNumWStages) { ws = 0; wph ^= 1; }
        }
      }
      if (w_resident) break;
    }
  } else if (warp == 1 && PAIR && !leader) {
    // ============================== peer of a CTA pair: weight-arrival forwarder ==============================
    // (the leader issues the MMAs for both CTAs; this warp tells it when this CTA's weight stages have landed)
    uint32_t ws = 0, wph = 0;
    for (long long tile = tile0; tile < outer_tiles; tile += tile_step) {
      const int nstages = p.kchunks * NGROUPS;
#pragma unroll 1
      for (int g = 0; g < nstages; ++g) {
        mbar_wait(w_land + ws, wph);
        if (elect_one()) mbar_arrive_leader(w_full + ws);
        __syncwarp();
        if (++ws == kNumWStages) { ws = 0; wph ^= 1; }
      }
      if (w_resident) break;
    }
  } else if (warp == 1) {
    // ============================== UMMA issuer ==============================
    // Warp-uniform loop; the elected lane issues a fully unrolled burst of G x TAPS_G x TZ x KSTEPS
    // UMMAs per weight stage.  Descriptors differ from a per-stage base only by compile-time
    // constants (start-address field, 16-byte units).
    constexpr uint32_t MM = PAIR ? 256u : 128u;      // M of the instruction: both CTAs' tiles in pair mode
    const uint32_t idesc = make_idesc_f16(MM, NT, p.fp16 ? 0u : 1u);
    // one MMA (cta_group::1, or cta_group::2 over the pair) / one commit (to this CTA, or multicast to both)
    auto mma = [](uint32_t d, uint64_t a, uint64_t b, uint32_t id, uint32_t acc) {
      if constexpr (PAIR) umma_ss2(d, a, b, id, acc); else umma_ss(d, a, b, id, acc);
    };
    auto commit = [](uint64_t* bar) {
      if constexpr (PAIR) umma_commit2(bar, 3); else umma_commit(bar);
    };
    constexpr uint32_t sbo_a = BX * C::ROWB;
    constexpr uint32_t sbo_w = 8u * C::ROWB;
    uint32_t as = 0, aph = 0, ws = 0, wph = 0, ab = 0, abph = 0;
    bool w_ready = false;                       // resident weights: every stage has been waited for once
    for (long long tile = tile0; tile < outer_tiles; tile += tile_step) {
     const uint32_t as0 = as, aph0 = aph;       // ring position of this spatial tile's first input box
#pragma unroll 1
     for (int j = 0; j < J; ++j) {
      mbar_wait(acc_empty + ab, abph ^ 1);
      tc_fence_after();
      const uint32_t acc_col = tmem_base + ab * C::ACC_COLS;
      if (w_resident) ws = 0;
      as = as0; aph = aph0;
#pragma unroll 1
      for (int kc = 0; kc < p.kchunks; ++kc) {
        if (j == 0) mbar_wait(a_full + as, aph);
        const uint64_t adesc0 = make_smem_desc(smem_u32(a_stage + as * C::A_STAGE), 16, sbo_a, C::LAYOUT);
#pragma unroll 1
        for (int gs = 0; gs < NGROUPS; ++gs) {
          if (!w_ready) mbar_wait(w_full + ws, wph);
          tc_fence_after();
          const uint64_t wdesc0 = make_smem_desc(smem_u32(w_stage + ws * C::W_STAGE), 16, sbo_w, C::LAYOUT);
          // slab g = gs * G + gi of the plain schedule = (dz, dy) with the three dx taps; slab g of
          // the plane-fused schedule = (dy, dx) with the three dz taps (ordered dz = +1, 0, -1).
          // Run-time part of its input-window offset (the rest is a compile-time constant per gi):
          const int g3r = G == 1 ? gs / 3 : (G == 3 ? gs : 0), grr = G == 1 ? gs - 3 * g3r : 0;
          const uint32_t a_off_r = !HALO ? 0u
                                   : (C::FUSE ? static_cast<uint32_t>((g3r * BX + grr) * C::ROWB)
                                              : static_cast<uint32_t>((g3r * BY + grr) * BX * C::ROWB));
          const uint64_t adesc_s = adesc0 + static_cast<uint64_t>(a_off_r >> 4);
          const bool tile_start = (kc | gs) == 0;
          if (elect_one()) {
#pragma unroll
            for (int gi = 0; gi < G; ++gi) {
              const int g3c = G == 9 ? gi / 3 : 0, grc = G == 9 ? gi % 3 : (G == 3 ? gi : 0);
              const uint32_t a_off_c = !HALO ? 0u
                                       : (C::FUSE ? static_cast<uint32_t>((g3c * BX + grc) * C::ROWB)
                                                  : static_cast<uint32_t>((g3c * BY + grc) * BX * C::ROWB));
              const uint64_t adesc_g = adesc_s + static_cast<uint64_t>(a_off_c >> 4);
              const uint64_t wdesc_g = wdesc0 + static_cast<uint64_t>((gi * C::SLAB) >> 4);
              const bool first_slab = gi == 0 && tile_start;
              if constexpr (!C::FUSE) {
#pragma unroll
                for (int t = 0; t < TAPS_G; ++t) {
#pragma unroll
                  for (int pz = 0; pz < TZ; ++pz) {
#pragma unroll
                    for (int ks = 0; ks < C::KSTEPS; ++ks) {
                      const uint64_t ad = adesc_g + static_cast<uint64_t>(((pz * BY * BX + t) * C::ROWB + ks * 32) >> 4);
                      const uint64_t wd = wdesc_g + static_cast<uint64_t>((t * NT * C::ROWB + ks * 32) >> 4);
                      mma(acc_col + pz * NT, ad, wd, idesc, (t | ks) == 0 ? (first_slab ? 0u : 1u) : 1u);
                    }
                  }
                }
              } else if (first_slab) {
                // very first slab of a tile: every output plane must start from accumulate = 0, so the
                // three dz taps are issued as separate N = NT instructions per output plane
#pragma unroll
                for (int pz = 0; pz < TZ; ++pz) {
#pragma unroll
                  for (int t = 0; t < 3; ++t) {                 // dz = 1 - t reads input plane pz + dz (box plane pz + 2 - t)
#pragma unroll
                    for (int ks = 0; ks < C::KSTEPS; ++ks) {
                      const uint64_t ad = adesc_g + static_cast<uint64_t>((((pz + 2 - t) * BY * BX) * C::ROWB + ks * 32) >> 4);
                      // B rows [t NT, (t + 1) NT) of the slab; pair mode: this CTA's half of them (PairImage V32 t)
                      const int wrow = PAIR ? PairImage::v32<NT>(t) : t * NT;
                      const uint64_t wd = wdesc_g + static_cast<uint64_t>((wrow * C::ROWB + ks * 32) >> 4);
                      mma(acc_col + pz * NT, ad, wd, idesc, (t | ks) == 0 ? 0u : 1u);
                    }
                  }
                }
              } else {
                // plane-fused schedule: input plane i (box plane i + 1) is read from shared memory ONCE
                // per (dy, dx) and multiplied against the weights of all the dz taps that use it; the
                // N = NT * nblk result columns land directly in the accumulators of the adjacent output
                // planes lo..hi (their TMEM columns are contiguous), so the sum over dz happens in TMEM.
#pragma unroll
                for (int i = -1; i <= TZ; ++i) {
                  const int lo = i - 1 < 0 ? 0 : i - 1, hi = i + 1 > TZ - 1 ? TZ - 1 : i + 1;
                  const int nblk = hi - lo + 1, toff = lo - (i - 1);
                  const uint32_t idesc_n = make_idesc_f16(MM, NT * nblk, p.fp16 ? 0u : 1u);
                  // B rows [toff NT, (toff + nblk) NT) of the slab; pair mode: this CTA's half of that range
                  const int wrow = !PAIR ? toff * NT
                                   : (nblk == 3 ? PairImage::v96<NT>()
                                      : (nblk == 2 ? (toff == 1 ? PairImage::v64h<NT>() : PairImage::v64l<NT>())
                                                   : PairImage::v32<NT>(toff)));
#pragma unroll
                  for (int ks = 0; ks < C::KSTEPS; ++ks) {
                    const uint64_t ad = adesc_g + static_cast<uint64_t>((((i + 1) * BY * BX) * C::ROWB + ks * 32) >> 4);
                    const uint64_t wd = wdesc_g + static_cast<uint64_t>((wrow * C::ROWB + ks * 32) >> 4);
                    mma(acc_col + lo * NT, ad, wd, idesc_n, 1u);
                  }
                }
              }
            }
            if (!w_resident) commit(w_empty + ws);   // stage free once these MMAs retire
            if (gs == NGROUPS - 1) {
              if (!reuse) commit(a_empty + as);
              if (kc == p.kchunks - 1) {
                commit(acc_full + ab);
                if (reuse && j == J - 1) {
                  // every pair of this spatial tile has been issued: its input boxes are free once
                  // these MMAs retire (one commit per ring stage it occupied)
                  uint32_t s2 = as0;
                  for (int k2 = 0; k2 < p.kchunks; ++k2) {
                    commit(a_empty + s2);
                    if (++s2 == kNumAStages) s2 = 0;
                  }
                }
              }
            }
          }
          __syncwarp();
          if (++ws == kNumWStages) { ws = 0; wph ^= 1; }
        }
        if (++as == kNumAStages) { as = 0; aph ^= 1; }
      }
      if (w_resident) w_ready = true;
      if (++ab == 2) { ab = 0; abph ^= 1; }
     }
    }
  } else {
    // ============================== epilogue warps ==============================
    const int q = warp & 3;                       // TMEM lane quarter this warp may access
    const int eg = warp >= 7 ? 1 : 0;             // epilogue group (warps 2-5 / 7-10): its share of the z planes
    constexpr int PZ_PER_GROUP = TZ / C::EPI_GROUPS;
    const int m = q * 32 + lane;                  // accumulator row = voxel (y = m / 8, x = m % 8)
    const int my = m >> 3, mx = m & 7;
    uint32_t ab = 0, abph = 0;
    // largest |activation| this thread produced (fp32, before the rounding to 16 bits): the FIRST value of a forward pass
    // that leaves the finite fp16 range is always a finite-or-infinite fp32 number here (NaNs only descend from it), so one
    // FMNMX per element is a complete overflow detector
    float amax = 0.f;
    uint32_t stg_it = 0;                          // staging-buffer round of this warp (TMA-store epilogue)
    for (long long tile = tile0; tile < outer_tiles; tile += tile_step) {
     for (int j = 0; j < J; ++j) {
      const TileCoord tc = decode_tile(p, tile, TZ, reuse, j);
      mbar_wait(acc_full + ab, abph);
      tc_fence_after();
      const uint32_t tbase = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + ab * C::ACC_COLS;
      const int iy = tc.y0 + my, ix = tc.x0 + mx;
      const float* bias = s_bias + tc.ntile * NT;
      if (!bias_all) {                            // rare (> 384 output channels): this tile's slice, all epilogue warps together
        constexpr int NEPI = 128 * C::EPI_GROUPS;
        named_barrier(3, NEPI);                   // everybody is done with the previous tile's slice
        for (int i = eg * 128 + m; i < NT; i += NEPI) s_bias[i] = __ldg(p.bias + tc.ntile * NT + i);
        named_barrier(3, NEPI);
        bias = s_bias;
      }
      static_assert(TZ % 2 == 0, "the epilogue walks the z planes of a tile in pairs");
      bool done_by_tma = false;
      {
        if (p.tma_out) {
          // TMEM -> registers -> bias / LeakyReLU -> swizzled staging buffer -> TMA box store.  (Per-thread 16-byte stores
          // touch 32 different 128-byte lines per instruction; the L1 store path, shared by all warps of the SM, was the
          // bound of these kernels.)  Every WARP owns the 32 voxels of its TMEM lane quarter (4 y rows x 8 x), its own
          // staging buffers and its own bulk-store groups: no barrier couples the epilogue warps (the two named barriers per
          // plane of the 128-voxel version were 25 % of the stall samples of enc0.conv2), and a warp only waits for its
          // previous store's shared-memory read right before it overwrites the buffer, after the TMEM loads and the math.
          auto run = [&](auto fp_tag) {
            constexpr bool FP16 = decltype(fp_tag)::value;
            constexpr int OB = C::OUT_BOX, ROWB_O = OB * 2, WBUF = 32 * ROWB_O;   // channels / bytes per staged voxel row, bytes per warp buffer
            constexpr uint32_t SWM = ROWB_O == 128 ? 7u : (ROWB_O == 64 ? 3u : 1u);
            constexpr int HC = OB < 32 ? OB : 32;                                 // channels per TMEM round trip
            uint8_t* wbuf0 = stg + static_cast<size_t>((eg * 4 + q) * C::STG_NBUF) * WBUF;
            const uint32_t rowoff = static_cast<uint32_t>(lane) * ROWB_O;
            const uint32_t sw = (rowoff >> 7) & SWM;
            const float al = p.act_lrelu ? p.alpha : 1.f;                         // 0 <= alpha <= 1 (checked at launch): lrelu(x) = max(x, alpha x)
            const bool rows_inside = tc.y0 + 4 * q < p.H;
#pragma unroll 1
            for (int pz = eg * PZ_PER_GROUP; pz < (eg + 1) * PZ_PER_GROUP; ++pz) {
#pragma unroll 1
              for (int cg = 0; cg < NT; cg += OB, ++stg_it) {
                uint8_t* buf = wbuf0 + (C::STG_NBUF == 2 ? (stg_it & 1u) : 0u) * WBUF;
                uint32_t pk[OB / 2];
#pragma unroll
                for (int h = 0; h < OB; h += HC) {
                  uint32_t r[HC];
#pragma unroll
                  for (int c0 = 0; c0 < HC; c0 += 16) tmem_ld16(tbase + pz * NT + cg + h + c0, *reinterpret_cast<uint32_t(*)[16]>(&r[c0]));
                  tmem_ld_wait();
#pragma unroll
                  for (int c0 = 0; c0 < HC; c0 += 16) {
                    float bs[16];
                    lds_f16x(bias + cg + h + c0, bs);
#pragma unroll
                    for (int j2 = 0; j2 < 16; j2 += 2) {
                      float x0 = __uint_as_float(r[c0 + j2]) + bs[j2], x1 = __uint_as_float(r[c0 + j2 + 1]) + bs[j2 + 1];
                      x0 = fmaxf(x0, al * x0);
                      x1 = fmaxf(x1, al * x1);
                      if (FP16) amax = fmaxf(amax, fmaxf(fabsf(x0), fabsf(x1)));
                      pk[(h + c0 + j2) >> 1] = pack2(x0, x1, FP16 ? 1 : 0);
                    }
                  }
                }
                // the store issued from this buffer (two rounds ago, or the previous one with a single buffer) must have
                // finished reading it (bulk groups belong to the issuing thread: always lane 0)
                if (lane == 0) { if (C::STG_NBUF == 2) bulk_wait_read_le1(); else bulk_wait_read_0(); }
                __syncwarp();
#pragma unroll
                for (int ch = 0; ch < OB / 8; ++ch)     // row `lane` of the warp's box, 16-byte chunk ch, swizzled like the TMA expects
                  *reinterpret_cast<uint4*>(buf + rowoff + ((static_cast<uint32_t>(ch) ^ sw) << 4)) =
                      make_uint4(pk[4 * ch], pk[4 * ch + 1], pk[4 * ch + 2], pk[4 * ch + 3]);
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0 && rows_inside) {
                  tma_store_5d(&omaps.m[tc.sub], buf, tc.ntile * NT + cg, tc.x0, tc.y0 + 4 * q, tc.z0 + pz, tc.n);
                  bulk_commit_group();
                }
              }
            }
          };
          if (p.fp16) run(std::true_type{}); else run(std::false_type{});
          done_by_tma = true;
        }
      }
      // pooled-only layers (encoder: the un-pooled tensor has no consumer) and no fused head: nothing to do at full resolution
      if (!done_by_tma && (p.out != nullptr || p.head_w != nullptr)) {
        // one z plane at a time: a voxel's channels are written back to back (interleaving the
        // stores of two planes measured 10-50 % slower on the store-bound layers)
#pragma unroll 1
        for (int pz = eg * PZ_PER_GROUP; pz < (eg + 1) * PZ_PER_GROUP; ++pz) {
          const int iz = tc.z0 + pz;
          const bool valid = iz < p.D && iy < p.H && ix < p.W;
          long long ovox;
          if (p.up_stride) {
            const int s = p.up_stride;
            const int a = (tc.sub >> 2) & 1, b = (tc.sub >> 1) & 1, c = tc.sub & 1;
            ovox = ((static_cast<long long>(tc.n) * (p.D * s) + (iz * s + a)) * (p.H * s) + (iy * s + b)) * (p.W * s) +
                   (ix * s + c);
          } else {
            ovox = ((static_cast<long long>(tc.n) * p.D + iz) * p.H + iy) * p.W + ix;
          }
          float head_acc = 0.f;
#pragma unroll
          for (int c0 = 0; c0 < NT; c0 += 16) {
            uint32_t r[16];
            tmem_ld16(tbase + pz * NT + c0, r);     // warp-collective: executed by all lanes, valid or not
            tmem_ld_wait();
            float v[16], bs[16];
            lds_f16x(bias + c0, bs);
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              float x = __uint_as_float(r[j]) + bs[j];
              if (p.act_lrelu) x = x > 0.f ? x : p.alpha * x;
              amax = fmaxf(amax, fabsf(x));
              v[j] = x;
            }
            if (p.head_w != nullptr) {
              float hw[16];
              lds_f16x(s_head + c0, hw);
#pragma unroll
              for (int j = 0; j < 16; ++j) head_acc = fmaf(v[j], hw[j], head_acc);
            }
            if (p.out != nullptr && valid) {
              uint4* dst = reinterpret_cast<uint4*>(p.out + (ovox * p.out_ctot + p.out_coff + tc.ntile * NT + c0) * 2);
              dst[0] = make_uint4(pack2(v[0], v[1], p.fp16), pack2(v[2], v[3], p.fp16), pack2(v[4], v[5], p.fp16),
                                  pack2(v[6], v[7], p.fp16));
              dst[1] = make_uint4(pack2(v[8], v[9], p.fp16), pack2(v[10], v[11], p.fp16), pack2(v[12], v[13], p.fp16),
                                  pack2(v[14], v[15], p.fp16));
            }
          }
          if (p.head_w != nullptr && valid) {
            const float logit = head_acc + p.head_b;
            if (p.head_labels) p.head_labels[ovox] = logit > 0.f ? 1 : 0;
            if (p.head_probs) p.head_probs[ovox] = 1.f / (1.f + __expf(-logit));
            if (p.head_logits) p.head_logits[ovox] = logit;
          }
        }
      }
      if (p.pool_out != nullptr) {
        // fused MaxPool3D(2): a separate pass over the accumulators (TMEM reads are cheap) so that
        // the full-resolution stores above stay one voxel row at a time.  max over the z pair
        // (two accumulator planes of the same thread), the x pair (lane ^ 1) and the y pair
        // (lane ^ 8); extents are even, so the partners of a valid even voxel are valid too.
        static_assert(PZ_PER_GROUP % 2 == 0 || HALO == 0, "the fused pool walks the planes of a group in pairs");
#pragma unroll 1
        for (int pz = eg * PZ_PER_GROUP; pz < (eg + 1) * PZ_PER_GROUP; pz += 2) {
          const int iz = tc.z0 + pz;
          const bool valid = iz < p.D && iy < p.H && ix < p.W && ((mx | my) & 1) == 0;
          const long long pvox = ((static_cast<long long>(tc.n) * (p.D >> 1) + (iz >> 1)) * (p.H >> 1) + (iy >> 1)) *
                                     (p.W >> 1) + (ix >> 1);
#pragma unroll
          for (int c0 = 0; c0 < NT; c0 += 16) {
            uint32_t r0[16], r1[16];
            tmem_ld16(tbase + pz * NT + c0, r0);
            tmem_ld16(tbase + (pz + 1) * NT + c0, r1);
            tmem_ld_wait();
            // bias, LeakyReLU (alpha > 0) and the rounding to 16 bits are all monotonic, so they commute with max: take
            // the z pair in fp32, activate and round, then pool x (lane ^ 1) and y (lane ^ 8) on PACKED pairs -- half the
            // shuffles of the fp32 version, identical results
            uint32_t w[8];
            float bs[16];
            lds_f16x(bias + c0, bs);
#pragma unroll
            for (int j = 0; j < 16; j += 2) {
              float a = fmaxf(__uint_as_float(r0[j]), __uint_as_float(r1[j])) + bs[j];
              float b = fmaxf(__uint_as_float(r0[j + 1]), __uint_as_float(r1[j + 1])) + bs[j + 1];
              if (p.act_lrelu) { a = a > 0.f ? a : p.alpha * a; b = b > 0.f ? b : p.alpha * b; }
              amax = fmaxf(amax, fmaxf(fabsf(a), fabsf(b)));
              w[j >> 1] = pack2(a, b, p.fp16);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              w[j] = max2_16(w[j], __shfl_xor_sync(0xffffffffu, w[j], 1), p.fp16);
              w[j] = max2_16(w[j], __shfl_xor_sync(0xffffffffu, w[j], 8), p.fp16);
            }
            if (valid) {
              uint4* dst = reinterpret_cast<uint4*>(p.pool_out +
                                                    (pvox * p.pool_ctot + p.pool_coff + tc.ntile * NT + c0) * 2);
              dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
              dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (PAIR) mbar_arrive_leader(acc_empty + ab);   // the leader's issuer waits for both CTAs' epilogues
        else mbar_arrive(acc_empty + ab);
      }
      if (++ab == 2) { ab = 0; abph ^= 1; }
     }
    }
    if (p.fp16 && p.overflow != nullptr && amax > 65504.f) atomicOr(p.overflow, 1);
  }

  if (p.tma_out && warp >= 2 && warp != 6) {    // every epilogue warp drains its own bulk-store groups
    if (lane == 0) bulk_wait_all_groups();
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (PAIR) {
    cluster_sync_all();                          // both CTAs are done with the pair's TMEM and with each other's barriers
    if (warp == 1) tmem_dealloc2(tmem_base, C::TMEM_COLS);
  } else {
    if (warp == 1) tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

template <int CB, int NT, int TZ, int HALO>
cudaError_t launch_cfg(const ConvUmmaLaunch& L, const KParams& p, int grid, cudaStream_t stream) {
  using C = Cfg<CB, NT, TZ, HALO>;
  static unsigned long long attr_set = 0;
  cudaError_t e = smem_attr_once(conv_umma_kernel<CB, NT, TZ, HALO>, C::SMEM_BYTES, &attr_set);
  if (e != cudaSuccess) return e;
  OutMaps om;
  for (int i = 0; i < 8; ++i) om.m[i] = L.tmap_out[i];
  conv_umma_kernel<CB, NT, TZ, HALO><<<grid, C::THREADS, C::SMEM_BYTES, stream>>>(L.tmap_in, L.tmap_in2, om, p);
  return cudaGetLastError();
}

// CTA-pair variant: clusters of 2 (the two SMs of a TPC), one pair per two tiles.
template <int CB, int NT, int TZ>
cudaError_t launch_pair(const ConvUmmaLaunch& L, KParams p, cudaStream_t stream) {
  using C = Cfg<CB, NT, TZ, 1, true>;
  static unsigned long long attr_set = 0;
  cudaError_t e = smem_attr_once(conv_umma_kernel<CB, NT, TZ, 1, true>, C::SMEM_BYTES, &attr_set);
  if (e != cudaSuccess) return e;
  OutMaps om;
  for (int i = 0; i < 8; ++i) om.m[i] = L.tmap_out[i];
  p.w_packed = static_cast<const uint8_t*>(L.w_packed_pair);
  long long g = p.total_tiles < kNumSMs ? p.total_tiles : kNumSMs;
  g &= ~1LL;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(static_cast<unsigned>(g));
  cfg.blockDim = dim3(C::THREADS);
  cfg.dynamicSmemBytes = C::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, conv_umma_kernel<CB, NT, TZ, 1, true>, L.tmap_in, L.tmap_in2, om, p);
}

KParams make_params(const ConvUmmaLaunch& L) {
  KParams p;
  p.w_packed = static_cast<const uint8_t*>(L.w_packed);
  p.bias = L.bias;
  p.out = static_cast<uint8_t*>(L.out);
  p.out_ctot = L.out_ctot;
  p.out_coff = L.out_coff;
  p.N = L.N; p.D = L.D; p.H = L.H; p.W = L.W;
  p.kchunks = L.Cin / L.CB;
  p.k1_chunks = L.Cin2 > 0 ? (L.Cin - L.Cin2) / L.CB : p.kchunks;
  p.halo = L.ntaps == 27 ? 1 : 0;
  p.ngroups = L.ntaps == 27 ? 9 : 1;
  p.taps_g = L.ntaps == 27 ? 3 : 1;
  p.nsub = L.up_stride ? 8 : 1;
  p.up_stride = L.up_stride;
  p.n_ntiles = L.Cout / L.NT;
  p.tiles_x = (L.W + 7) / 8;
  p.tiles_y = (L.H + 15) / 16;
  p.tiles_z = (L.D + L.TZ - 1) / L.TZ;
  p.total_tiles = static_cast<long long>(p.n_ntiles) * p.nsub * L.N * p.tiles_z * p.tiles_y * p.tiles_x;
  p.act_lrelu = L.act_lrelu;
  p.alpha = L.alpha;
  p.fp16 = L.fp16;
  p.head_w = L.head_w;
  p.head_b = L.head_b;
  p.head_labels = L.head_labels;
  p.head_probs = L.head_probs;
  p.head_logits = L.head_logits;
  p.pool_out = static_cast<uint8_t*>(L.pool_out);
  p.pool_ctot = L.pool_ctot;
  p.pool_coff = L.pool_coff;
  p.overflow = L.overflow;
  p.tma_out = (L.use_tma_out && L.out != nullptr) ? 1 : 0;
  p.reuse_ok = !p.halo && p.total_tiles / (static_cast<long long>(p.n_ntiles) * p.nsub) >= kNumSMs ? 1 : 0;
  return p;
}

}  // namespace

bool conv_umma_config(int Cin, int Cout, ConvUmmaConfig* cfg) {
  if (Cin % 16 != 0 || Cout % 16 != 0 || Cin <= 0 || Cout <= 0) return false;
  int NT;
  if (Cout % 128 == 0) NT = 128;
  else if (Cout == 64 || Cout == 32 || Cout == 16) NT = Cout;
  else if (Cout % 64 == 0) NT = 64;
  else if (Cout % 32 == 0) NT = 32;
  else NT = 16;
  cfg->NT = NT;
  cfg->CB = (Cin % 32 == 0) ? 32 : 16;
  cfg->TZ = NT == 128 ? 2 : 4;
  // N = 32 layers are bound by the shared-memory read of the A operand (32 cycles per UMMA whatever N is): with
  // TZ = 8 planes per tile the dz-fused instructions run at N = 96 for 6 of 10 input planes instead of 2 of 6
  // (ceiling 75 % of the tensor peak instead of 67 %); the haloed box then only fits with 16-channel chunks.
  static const bool tz8 = getenv("TOMOENC_NO_TZ8") == nullptr;
  if (tz8 && NT == 32) { cfg->TZ = 8; cfg->CB = 16; }
  // ... and with the B rows of every instruction split over a CTA pair (cta_group::2) the N = 96 instruction is
  // math-bound: 49 instead of 56 cycles (tools/umma2_probe.cu), ceiling 75 -> 83 % for a TZ = 8 tile
  static const bool pair = getenv("TOMOENC_NO_PAIR") == nullptr;
  cfg->pair = (pair && NT == 32 && cfg->TZ == 8 && cfg->CB == 16 && Cout == 32) ? 1 : 0;
  return true;
}

size_t conv_umma_pair_packed_bytes(int Cin, int Cout, const ConvUmmaConfig& cfg) {
  if (!cfg.pair) return 0;
  return static_cast<size_t>(2) * (Cin / cfg.CB) * 9 * PairImage::rows<32>() * (Cout / 32) * cfg.CB * 2;
}

// Pair-mode image of a 3x3x3 conv (Cout = NT = 32): [rank][channel chunk][slab g = (dy, dx)][PairImage rows][CB], rows
// swizzled like the single-CTA slabs (the slab base is a multiple of 256 bytes).  w = [tap][cout][cin], BN folded.
void conv_umma_pack_pair(const float* w, int Cin, int Cout, const ConvUmmaConfig& cfg, int fp16, void* dst_host) {
  const int CB = cfg.CB, NT = cfg.NT;
  const int rowb = CB * 2;
  const uint32_t mask = CB == 16 ? 1u : (CB == 32 ? 3u : 7u);
  const int kchunks = Cin / CB;
  const int rows = 5 * NT;
  const size_t slab = static_cast<size_t>(rows) * rowb;
  // image row -> row of the full [t][cout] slab (t <-> dz = 1 - t), per rank
  auto full_row = [&](int rank, int ir) -> int {
    const int h = NT / 2;
    if (ir < 3 * h) return 3 * h * rank + ir;                       // V96: rows [0, 3 NT) halved
    ir -= 3 * h;
    if (ir < NT) return NT + NT * rank + ir;                        // V64H: rows [NT, 3 NT) halved
    ir -= NT;
    if (ir < NT) return NT * rank + ir;                             // V64L: rows [0, 2 NT) halved
    ir -= NT;
    const int t = ir / h;                                           // V32 t: rows [t NT, (t + 1) NT) halved
    return t * NT + h * rank + (ir - t * h);
  };
  for (int rank = 0; rank < 2; ++rank)
    for (int kc = 0; kc < kchunks; ++kc)
      for (int g = 0; g < 9; ++g) {
        uint8_t* base = static_cast<uint8_t*>(dst_host) + ((static_cast<size_t>(rank) * kchunks + kc) * 9 + g) * slab;
        for (int ir = 0; ir < rows; ++ir) {
          const int fr = full_row(rank, ir);
          const int t = fr / NT, r = fr % NT;
          const int tap = (2 - t) * 9 + g;                          // slab g = (dy, dx), t <-> dz = 1 - t
          for (int kk = 0; kk < CB; ++kk) {
            const float val = w[(static_cast<size_t>(tap) * Cout + r) * Cin + kc * CB + kk];
            uint16_t bits;
            if (fp16) { __half h = __float2half_rn(val); bits = *reinterpret_cast<uint16_t*>(&h); }
            else { __nv_bfloat16 h = __float2bfloat16_rn(val); bits = *reinterpret_cast<uint16_t*>(&h); }
            const uint32_t Lo = static_cast<uint32_t>(ir * rowb + kk * 2);
            const uint32_t P = Lo ^ (((Lo >> 7) & mask) << 4);
            *reinterpret_cast<uint16_t*>(base + P) = bits;
          }
        }
      }
}

size_t conv_umma_packed_bytes(int Cin, int Cout, int ntaps, int nsub, const ConvUmmaConfig& cfg) {
  (void)cfg;
  return static_cast<size_t>(nsub) * ntaps * Cin * Cout * 2;
}

void conv_umma_pack(const float* w, int Cin, int Cout, int ntaps, int nsub, const ConvUmmaConfig& cfg, int fp16,
                    void* dst_host) {
  const int CB = cfg.CB, NT = cfg.NT;
  const int rowb = CB * 2;
  const uint32_t mask = CB == 16 ? 1u : (CB == 32 ? 3u : 7u);
  const int kchunks = Cin / CB, n_ntiles = Cout / NT;
  const int ngroups = ntaps == 27 ? 9 : 1, taps_g = ntaps == 27 ? 3 : 1;
  const size_t slab = static_cast<size_t>(taps_g) * NT * rowb;
  uint16_t* dst = static_cast<uint16_t*>(dst_host);
  for (int nt = 0; nt < n_ntiles; ++nt)
    for (int sub = 0; sub < nsub; ++sub)
      for (int kc = 0; kc < kchunks; ++kc)
        for (int g = 0; g < ngroups; ++g) {
          uint8_t* base = reinterpret_cast<uint8_t*>(dst) +
                          (((static_cast<size_t>(nt) * nsub + sub) * kchunks + kc) * ngroups + g) * slab;
          for (int t = 0; t < taps_g; ++t)
            for (int r = 0; r < NT; ++r)
              for (int kk = 0; kk < CB; ++kk) {
                // plain schedule: slab g = (dz, dy), t = dx.  Plane-fused schedule (3 * NT <= 256):
                // slab g = (dy, dx), t <-> dz = 1 - t.
                const bool fuse = ntaps == 27 && 3 * NT <= 256;
                const int tap = fuse ? ((2 - t) * 9 + g) : (g * taps_g + t);
                const int cout = nt * NT + r, cin = kc * CB + kk;
                const float val = w[((static_cast<size_t>(sub) * ntaps + tap) * Cout + cout) * Cin + cin];
                uint16_t bits;
                if (fp16) { __half h = __float2half_rn(val); bits = *reinterpret_cast<uint16_t*>(&h); }
                else { __nv_bfloat16 h = __float2bfloat16_rn(val); bits = *reinterpret_cast<uint16_t*>(&h); }
                const uint32_t Lo = static_cast<uint32_t>((t * NT + r) * rowb + kk * 2);
                const uint32_t P = Lo ^ (((Lo >> 7) & mask) << 4);
                *reinterpret_cast<uint16_t*>(base + P) = bits;
              }
        }
}

namespace {
typedef CUresult (*EncodeTiledFn2)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn2 get_encode2() {
  static EncodeTiledFn2 fn = nullptr;
  if (!fn) {
    cudaDriverEntryPointQueryResult q;
    void* p = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess) fn = (EncodeTiledFn2)p;
  }
  return fn;
}
}  // namespace

void conv_umma_make_out_maps(ConvUmmaLaunch* L) {
  L->use_tma_out = 0;
  static const bool off = getenv("TOMOENC_NO_TMA_EPILOGUE") != nullptr;
  static const bool off_halo = getenv("TOMOENC_NO_TMA_EPILOGUE_CONV") != nullptr;
  if (off || (off_halo && L->ntaps == 27) || L->out == nullptr) return;
  if (L->out_ctot % 8 != 0 || L->out_coff % 8 != 0 || L->NT % 16 != 0) return;
  EncodeTiledFn2 enc = get_encode2();
  if (!enc) return;
  const int s = L->up_stride ? L->up_stride : 1;
  const int nsub = L->up_stride ? 8 : 1;
  const long long D2 = static_cast<long long>(L->D) * s, H2 = static_cast<long long>(L->H) * s, W2 = static_cast<long long>(L->W) * s;
  const long long ct = L->out_ctot;
  const int ob = L->NT < 64 ? L->NT : 64;
  const CUtensorMapSwizzle sw = ob * 2 == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                : (ob * 2 == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  for (int sub = 0; sub < nsub; ++sub) {
    const int a = (sub >> 2) & 1, b = (sub >> 1) & 1, c = sub & 1;
    uint8_t* base = static_cast<uint8_t*>(L->out) + ((((L->up_stride ? a : 0) * H2 + (L->up_stride ? b : 0)) * W2 + (L->up_stride ? c : 0)) * ct + L->out_coff) * 2;
    cuuint64_t gdim[5] = {(cuuint64_t)L->Cout, (cuuint64_t)L->W, (cuuint64_t)L->H, (cuuint64_t)L->D, (cuuint64_t)L->N};
    cuuint64_t gstr[4] = {(cuuint64_t)(s * ct * 2), (cuuint64_t)(s * W2 * ct * 2), (cuuint64_t)(s * H2 * W2 * ct * 2),
                          (cuuint64_t)(D2 * H2 * W2 * ct * 2)};
    cuuint32_t box[5] = {(cuuint32_t)ob, 8, 4, 1, 1};      // one epilogue warp: the 4 y rows x 8 x of its TMEM lane quarter
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    if (enc(&L->tmap_out[sub], L->fp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, base, gdim, gstr, box,
            estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return;
  }
  for (int sub = nsub; sub < 8; ++sub) L->tmap_out[sub] = L->tmap_out[0];
  L->use_tma_out = 1;
}

int conv_umma_grid(const ConvUmmaLaunch& L) {
  const KParams p = make_params(L);
  return static_cast<int>(p.total_tiles < kNumSMs ? p.total_tiles : kNumSMs);
}

cudaError_t conv_umma_launch(const ConvUmmaLaunch& L, cudaStream_t stream) {
  const KParams p = make_params(L);
  if (p.total_tiles <= 0) return cudaSuccess;
  if (p.total_tiles >= (1LL << 31)) return cudaErrorInvalidValue;   // decode_tile works on 32-bit indices
  if (p.act_lrelu && !(p.alpha >= 0.f && p.alpha <= 1.f)) return cudaErrorInvalidValue;   // epilogue: lrelu(x) = max(x, alpha x)
  // un-haloed kernels may iterate the (Cout tile, sub-position) pairs inside a spatial tile (input reuse)
  const int grid = static_cast<int>(p.total_tiles < kNumSMs ? p.total_tiles : kNumSMs);
  // CTA pairs: both CTAs of a pair must see the same number of tiles
  if (L.w_packed_pair != nullptr && p.halo && L.CB == 16 && L.NT == 32 && L.TZ == 8 && p.n_ntiles == 1 &&
      p.total_tiles % 2 == 0 && p.total_tiles >= 2)
    return launch_pair<16, 32, 8>(L, p, stream);
#define TE_CONV_CASE(cb, nt, tz)                                                          \
  if (L.CB == cb && L.NT == nt && L.TZ == tz)                                             \
    return p.halo ? launch_cfg<cb, nt, tz, 1>(L, p, grid, stream) : launch_cfg<cb, nt, tz, 0>(L, p, grid, stream);
  TE_CONV_CASE(16, 16, 4)
  TE_CONV_CASE(16, 32, 4)
  TE_CONV_CASE(16, 32, 8)
  TE_CONV_CASE(16, 64, 4)
  TE_CONV_CASE(16, 128, 2)
  TE_CONV_CASE(32, 16, 4)
  TE_CONV_CASE(32, 32, 4)
  TE_CONV_CASE(32, 64, 4)
  TE_CONV_CASE(32, 128, 2)
#undef TE_CONV_CASE
  return cudaErrorInvalidValue;
}

}  // namespace te

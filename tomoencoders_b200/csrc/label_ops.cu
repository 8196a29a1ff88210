// Post-stitch labelling: connected components of the segmented mask, per-component
// bounding boxes / sizes, component masks, and the sparse tile selection that feeds the next (finer) pass of the path.
// Replaces cupyx.scipy.ndimage.label / scipy.ndimage.find_objects as used by the reference
// (/root/reference/tomo_encoders/tasks/digital_zoom.py:41-43, structures/voids.py:188-303, 356-373).
//
// Labelling = lock-free union-find over the voxel lattice (HBM-bound integer work, no tensor cores):
//   init      parent[v] = v for foreground voxels, -1 for background
//   merge     every foreground voxel is united with its foreground neighbours of smaller linear index (13 of the 26
//             neighbours, 3 of the 6); the union always hangs the larger root under the smaller one (atomicMin), so the root
//             of a component is its FIRST voxel in C order
//   compress  parent[v] = root(v)
//   number    roots are numbered 1..n in C order by a block-wise prefix sum -> scipy.ndimage.label's numbering
// int32 indices: volumes up to 2^31 - 1 voxels (the reference labels binned maps; 1024^3 fits).
#include "te_common.cuh"

#include <cstdlib>

namespace te {
namespace {

constexpr int kScanItems = 4096;            // voxels per block in the numbering passes (256 threads x 16)

struct Dim3 { int Z, Y, X; };

// item = 32 consecutive x of one row (one warp): a foreground voxel starts under the first voxel of its run inside the
// 32-group (ballot + bit scan, no atomics), so the merge pass only has to stitch runs together
template <typename T>
__global__ void __launch_bounds__(256) label_init_kernel(const T* __restrict__ vol, int* __restrict__ parent, Dim3 d) {
  const int segs = (d.X + 31) / 32;
  const long long items = static_cast<long long>(d.Z) * d.Y * segs;
  const int lane = threadIdx.x & 31;
  const long long wstride = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  for (long long it = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; it < items; it += wstride) {
    const int sg = static_cast<int>(it % segs);
    const long long row = it / segs;
    const int x = sg * 32 + lane;
    const long long i = row * d.X + x;
    const bool fg = x < d.X && vol[i] != T(0);
    const unsigned m = __ballot_sync(0xffffffffu, fg);
    if (x < d.X) {
      const unsigned below = ~m & ((1u << lane) - 1u);               // background voxels to my left inside the group
      const int start = below ? 32 - __clz(below) : 0;
      parent[i] = fg ? static_cast<int>(row * d.X + sg * 32 + start) : -1;
    }
  }
}

// Plain (L1-cached) loads on purpose: a stale parent is still an older ancestor, a stale "root" is caught by the atomicMin
// of uf_union (performed on the live value in L2, its return value is fresh), and every step moves to a smaller index, so
// the walk terminates; the hot roots of big components are then served by L1 instead of hammering one L2 line.
__device__ __forceinline__ int uf_find(int* parent, int x) {
  int* vp = parent;
  int p = vp[x];
  while (p != x) {
    const int g = vp[p];
    if (g != p) vp[x] = g;                  // path halving: any ancestor is a valid parent (values only ever decrease)
    x = p; p = g;
  }
  return x;
}
__device__ __forceinline__ void uf_union(int* parent, int a, int b) {
  while (true) {
    a = uf_find(parent, a);
    b = uf_find(parent, b);
    if (a == b) return;
    if (a < b) { const int t = a; a = b; b = t; }          // a > b: hang a under b
    const int old = atomicMin(parent + a, b);
    if (old == a) return;                                    // a was still a root: linked
    a = old;                                                 // somebody re-parented a meanwhile: retry from there
  }
}

// ---- tile-local pre-merge.  The global merge below spends its time chasing parent chains through L2 and on atomics
// on the hot roots of big components.  Most unions join voxels of the same small neighbourhood, so they are resolved
// first inside a 32 x 16 x 16 tile in SHARED memory (same run initialisation, same pruned union rules, restricted to pairs
// whose two voxels lie in the tile); the tile then writes parent[v] = global index of its fragment's root (depth 1), and
// only pairs that straddle a tile face are left for the global pass.  The local index order (z, y, x) agrees with the
// global C order, so "smaller index wins" still makes every root the first voxel of its fragment / component.
// Neighbours outside the tile count as background in the pruning predicates: that can only add (true-adjacency) unions.
constexpr int kTX = 32, kTY = 16, kTZ = 16, kTileVox = kTX * kTY * kTZ;

// All 32 lanes call this; lanes with the same (a, b) elect one of them to perform the union.
template <typename F>
__device__ __forceinline__ void warp_union(bool valid, int a, int b, int lane, F&& do_union) {
  if (!__any_sync(0xffffffffu, valid)) return;
  const unsigned long long key = valid ? (static_cast<unsigned long long>(static_cast<unsigned>(a)) << 32) | static_cast<unsigned>(b)
                                       : (0xffffffff00000000ull | static_cast<unsigned>(lane));
  const unsigned m = __match_any_sync(0xffffffffu, key);
  if (valid && (__ffs(m) - 1) == lane) do_union(a, b);
}

__device__ __forceinline__ int sfind(volatile int* lab, int x) {
  int p = lab[x];
  while (p != x) {
    const int g = lab[p];
    if (g != p) lab[x] = g;          // path halving (values only ever decrease: any ancestor is a valid parent)
    x = p; p = g;
  }
  return x;
}
__device__ __forceinline__ void sunion(int* lab, int a, int b) {
  while (true) {
    a = sfind(lab, a);
    b = sfind(lab, b);
    if (a == b) return;
    if (a < b) { const int t = a; a = b; b = t; }
    const int old = atomicMin(lab + a, b);
    if (old == a) return;
    a = old;
  }
}

template <typename T, int CONN>
__global__ void __launch_bounds__(256) label_tile_kernel(const T* __restrict__ vol, int* __restrict__ parent, Dim3 d, int tiles_x,
                                                         int tiles_y) {
  __shared__ int lab[kTileVox];
  __shared__ unsigned rowmask[kTY * kTZ];     // foreground bits of every 32-voxel row: the merge predicates are bit tests
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int t = blockIdx.x;
  const int x0 = (t % tiles_x) * kTX; t /= tiles_x;
  const int y0 = (t % tiles_y) * kTY;
  const int z0 = (t / tiles_y) * kTZ;
  const int x = x0 + lane;
  // init: one warp per row of 32 x
  for (int r = warp; r < kTY * kTZ; r += 8) {
    const int ly = r % kTY, lz = r / kTY;
    const int y = y0 + ly, z = z0 + lz;
    const bool in = x < d.X && y < d.Y && z < d.Z;
    const bool fg = in && vol[(static_cast<long long>(z) * d.Y + y) * d.X + x] != T(0);
    const unsigned m = __ballot_sync(0xffffffffu, fg);
    const unsigned below = ~m & ((1u << lane) - 1u);
    const int start = below ? 32 - __clz(below) : 0;
    lab[r * kTX + lane] = fg ? r * kTX + start : -1;
    if (lane == 0) rowmask[r] = m;
  }
  __syncthreads();
  // merge inside the tile (voxel l = (lz * kTY + ly) * kTX + lx; neighbour rows r - kTY - 1, r - kTY, r - kTY + 1, r - 1).
  // A pair is handed to the union as (set member of l, set member of the neighbour) = their current labels, i.e. their run
  // starts.  (Electing one lane per distinct pair with __match_any_sync, as the global pass does, costs more here than
  // the duplicate shared-memory unions it saves: 12.8 vs 7.3 ms at 512^3.)
  for (int r = warp; r < kTY * kTZ; r += 8) {
    const int l = r * kTX + lane;
    const unsigned mrow = rowmask[r];
    if (mrow == 0u) continue;                                              // warp-uniform: empty row
    const bool fg = (mrow >> lane) & 1u;
    const int me = fg ? lab[l] : -1;
    const int ly = r % kTY, lz = r / kTY;
    const bool left = lane > 0 && ((mrow >> (lane - 1)) & 1u);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int dz = q < 3 ? -1 : 0, dy = q < 3 ? q - 1 : -1;
      if (CONN == 6 && !((dz == -1 && dy == 0) || (dz == 0 && dy == -1))) continue;
      if (lz + dz < 0 || ly + dy < 0 || ly + dy >= kTY) continue;          // warp-uniform
      const unsigned mj = rowmask[r + dz * kTY + dy];                       // neighbour row's foreground bits
      if (mj == 0u) continue;
      const int jb = l + (dz * kTY + dy) * kTX;
      const bool f0 = (mj >> lane) & 1u;
      const bool fm = lane > 0 && ((mj >> (lane - 1)) & 1u);
      if (CONN == 6) {
        if (fg && f0 && !(left && fm)) sunion(lab, me, lab[jb]);
      } else {
        const bool f2 = lane > 1 && ((mj >> (lane - 2)) & 1u);
        const bool fp = lane + 1 < kTX && ((mj >> (lane + 1)) & 1u);
        if (fg && fm && !(left && f2)) sunion(lab, me, lab[jb - 1]);
        if (fg && f0 && !fm) sunion(lab, me, lab[jb]);
        if (fg && fp && !f0) sunion(lab, me, lab[jb + 1]);
      }
    }
  }
  __syncthreads();
  // flatten and publish: parent = global index of the fragment root
  for (int r = warp; r < kTY * kTZ; r += 8) {
    const int ly = r % kTY, lz = r / kTY;
    const int y = y0 + ly, z = z0 + lz;
    if (x >= d.X || y >= d.Y || z >= d.Z) continue;
    const int l = r * kTX + lane;
    int v = -1;
    if (lab[l] >= 0) {
      const int root = sfind(lab, l);
      const int rx = root % kTX, rr = root / kTX;
      v = static_cast<int>((static_cast<long long>(z0 + rr / kTY) * d.Y + (y0 + rr % kTY)) * d.X + x0 + rx);
    }
    parent[(static_cast<long long>(z) * d.Y + y) * d.X + x] = v;
  }
}

// thread = one voxel.  Neighbour rows above / behind: (dz, dy) in {(-1,-1), (-1,0), (-1,1), (0,-1)} with dx in {-1,0,1}
// (26-connectivity) or (-1,0), (0,-1) with dx = 0 (6-connectivity); the x neighbour only matters across 32-groups.
// A union (i, j) is skipped when it is implied by others: j's left neighbour also touches i (same run of row'), or both
// left neighbours are foreground (the pair (i-1, j-1) makes the same link; induction on x).
// After label_tile_kernel only the pairs whose voxels lie in different tiles remain (XT = true).  Warp-uniform control
// flow: a pair is handed over as (parent[i], parent[neighbour]) -- with XT the two fragment roots -- and one lane per
// distinct pair of the warp performs the union (a flat face between two fragments costs one union per row, not 3 x 32).
template <int CONN, bool XT>
__global__ void __launch_bounds__(256) label_merge_kernel(int* parent, Dim3 d) {
  const long long n = static_cast<long long>(d.Z) * d.Y * d.X;
  const long long plane = static_cast<long long>(d.Y) * d.X;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const int lane = threadIdx.x & 31;
  auto un = [&](int a, int b) { uf_union(parent, a, b); };
  for (long long base = static_cast<long long>(blockIdx.x) * blockDim.x + (threadIdx.x & ~31); base < n; base += stride) {
    const long long i = base + lane;
    const int pi = i < n ? parent[i] : -1;
    bool act = pi >= 0;
    int x = 0, y = 0, z = 0;
    if (act) {
      x = static_cast<int>(i % d.X);
      const long long t = i / d.X;
      y = static_cast<int>(t % d.Y);
      z = static_cast<int>(t / d.Y);
      if (XT) {   // interior voxels of a tile have no cross-tile pair
        const int lx = x & (kTX - 1), ly = y & (kTY - 1), lz = z & (kTZ - 1);
        act = lx == 0 || lx == kTX - 1 || ly == 0 || ly == kTY - 1 || lz == 0;
      }
    }
    if (!__any_sync(0xffffffffu, act)) continue;
    const int pl = act && x > 0 ? parent[i - 1] : -1;
    const bool left = pl >= 0;
    warp_union(act && left && (x & 31) == 0, pi, pl, lane, un);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int dz = r < 3 ? -1 : 0, dy = r < 3 ? r - 1 : -1;
      if (CONN == 6 && !((dz == -1 && dy == 0) || (dz == 0 && dy == -1))) continue;
      const int zz = z + dz, yy = y + dy;
      const bool rv = act && zz >= 0 && yy >= 0 && yy < d.Y;
      // the neighbour ROW is in another tile (then every pair of it is), or only the pairs across an x face are
      const bool row_x = !XT || (dz < 0 && (z & (kTZ - 1)) == 0) || (dy < 0 && (y & (kTY - 1)) == 0) ||
                         (dy > 0 && (y & (kTY - 1)) == kTY - 1);
      const bool xm = row_x || (x & (kTX - 1)) == 0;             // pair (i, jb - 1) crosses a face
      const bool xp = row_x || (x & (kTX - 1)) == kTX - 1;       // pair (i, jb + 1)
      const long long jb = i + dz * plane + static_cast<long long>(dy) * d.X;     // same x in the neighbour row
      if (CONN == 6) {
        const bool need = rv && row_x;
        const int v0 = need ? parent[jb] : -1;
        const int vm = need && x > 0 && left ? parent[jb - 1] : -1;
        warp_union(need && v0 >= 0 && !(left && vm >= 0), pi, v0, lane, un);
      } else {
        const bool need = rv && (row_x || xm || xp);
        const int vm2 = need && x > 1 ? parent[jb - 2] : -1;
        const int vm = need && x > 0 ? parent[jb - 1] : -1;
        const int v0 = need ? parent[jb] : -1;
        const int vp = need && x + 1 < d.X ? parent[jb + 1] : -1;
        warp_union(need && xm && vm >= 0 && !(left && vm2 >= 0), pi, vm, lane, un);
        warp_union(need && row_x && v0 >= 0 && vm < 0, pi, v0, lane, un);
        warp_union(need && xp && vp >= 0 && v0 < 0, pi, vp, lane, un);
      }
    }
  }
}

// Read-only walk to the root (the merge pass is over: the forest is final).  No path halving here: a late halving store of
// another thread could overwrite parent[i] = root with a mere ancestor, and the numbering pass needs roots.
__global__ void __launch_bounds__(256) label_compress_kernel(int* parent, long long n) {
  const int* vp = parent;                   // the forest is final: parent[] only changes to roots here, which a walk also accepts
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    int x = vp[i];
    if (x < 0) continue;
    int q = vp[x];
    while (q != x) { x = q; q = vp[x]; }
    parent[i] = x;
  }
}

// block b owns voxels [b * kScanItems, (b + 1) * kScanItems); thread t owns 16 consecutive ones
__device__ __forceinline__ int count_roots16(const int* parent, long long base, long long n) {
  int c = 0;
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const long long i = base + k;
    c += (i < n && parent[i] == static_cast<int>(i)) ? 1 : 0;
  }
  return c;
}
__global__ void __launch_bounds__(256) label_count_kernel(const int* __restrict__ parent, long long n, int* __restrict__ counts) {
  __shared__ int ws[8];
  int c = count_roots16(parent, static_cast<long long>(blockIdx.x) * kScanItems + threadIdx.x * 16, n);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int tot = 0;
    for (int w = 0; w < 8; ++w) tot += ws[w];
    counts[blockIdx.x] = tot;
  }
}
// exclusive scan of the block counts in place (one CTA walks the array with a running carry); n_labels = total
__global__ void __launch_bounds__(1024) label_scan_kernel(int* counts, int nblocks, long long* n_labels) {
  __shared__ int ws[32];
  __shared__ int carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < nblocks; base += 1024) {
    const int i = base + threadIdx.x;
    const int v = i < nblocks ? counts[i] : 0;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if ((threadIdx.x & 31) >= o) incl += t;
    }
    if ((threadIdx.x & 31) == 31) ws[threadIdx.x >> 5] = incl;
    __syncthreads();
    if (threadIdx.x < 32) {
      int w = ws[threadIdx.x];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, w, o);
        if (threadIdx.x >= o) w += t;
      }
      ws[threadIdx.x] = w;
    }
    __syncthreads();
    const int warp_off = (threadIdx.x >> 5) ? ws[(threadIdx.x >> 5) - 1] : 0;
    const int carry = carry_s;
    if (i < nblocks) counts[i] = carry + warp_off + incl - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = carry + warp_off + incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) *n_labels = carry_s;
}
// roots get their rank (1-based, C order)
__global__ void __launch_bounds__(256) label_rank_kernel(const int* __restrict__ parent, long long n, const int* __restrict__ offsets,
                                                          int* __restrict__ labels) {
  __shared__ int ws[8];
  const long long base = static_cast<long long>(blockIdx.x) * kScanItems + threadIdx.x * 16;
  const int c = count_roots16(parent, base, n);
  int incl = c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, o);
    if ((threadIdx.x & 31) >= o) incl += t;
  }
  if ((threadIdx.x & 31) == 31) ws[threadIdx.x >> 5] = incl;
  __syncthreads();
  int off = offsets[blockIdx.x] + incl - c;
  for (int w = 0; w < (threadIdx.x >> 5); ++w) off += ws[w];
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const long long i = base + k;
    if (i < n && parent[i] == static_cast<int>(i)) labels[i] = ++off;
  }
}
// everything else reads its root's rank
__global__ void __launch_bounds__(256) label_assign_kernel(const int* __restrict__ parent, long long n, int* __restrict__ labels) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int p = parent[i];
    if (p < 0) labels[i] = 0;
    else if (p != static_cast<int>(i)) labels[i] = labels[p];       // roots were written by the rank kernel of an earlier launch
  }
}

// ---- per-label bounding boxes (find_objects) and voxel counts.  A thread walks 32 consecutive x of one row and flushes a
// run of equal labels with one set of atomics (components are mostly runs along x).
__global__ void __launch_bounds__(256) bbox_init_kernel(int* bbox, long long* sizes, long long n_labels) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n_labels; i += stride) {
    bbox[6 * i + 0] = bbox[6 * i + 1] = bbox[6 * i + 2] = 0x7fffffff;
    bbox[6 * i + 3] = bbox[6 * i + 4] = bbox[6 * i + 5] = 0;
    sizes[i] = 0;
  }
}
__device__ __forceinline__ void bbox_flush(int* bbox, long long* sizes, int lab, int z0, int y0, int x0, int z1, int y1, int x1,
                                           unsigned long long cnt) {
  int* b = bbox + 6LL * (lab - 1);
  atomicMin(b + 0, z0); atomicMin(b + 1, y0); atomicMin(b + 2, x0);
  atomicMax(b + 3, z1); atomicMax(b + 4, y1); atomicMax(b + 5, x1);
  atomicAdd(reinterpret_cast<unsigned long long*>(sizes + (lab - 1)), cnt);
}
// A thread walks 32 consecutive x of one row and emits one record per run of equal labels.  Records first go to a
// 256-entry hash table in shared memory (a block covers a few neighbouring rows, i.e. a handful of labels; shared-memory
// atomics absorb the hot labels), the table is flushed with one set of global atomics per label and block; a full table
// falls back to global atomics.
constexpr int kTab = 256;
__global__ void __launch_bounds__(256) bbox_kernel(const int* __restrict__ labels, Dim3 d, long long n_labels, int* bbox, long long* sizes) {
  __shared__ int t_lab[kTab];
  __shared__ int t_box[kTab][6];
  __shared__ unsigned long long t_cnt[kTab];
  const int segs = (d.X + 31) / 32;
  const long long items = static_cast<long long>(d.Z) * d.Y * segs;
  const long long per_block = 256LL * 8;
  for (long long base = static_cast<long long>(blockIdx.x) * per_block; base < items; base += static_cast<long long>(gridDim.x) * per_block) {
    for (int k = threadIdx.x; k < kTab; k += 256) {
      t_lab[k] = 0; t_cnt[k] = 0;
      t_box[k][0] = t_box[k][1] = t_box[k][2] = 0x7fffffff;
      t_box[k][3] = t_box[k][4] = t_box[k][5] = 0;
    }
    __syncthreads();
    for (int q = 0; q < 8; ++q) {
      const long long it = base + q * 256 + threadIdx.x;
      if (it >= items) break;
      const int sg = static_cast<int>(it % segs);
      const long long row = it / segs;
      const int y = static_cast<int>(row % d.Y), z = static_cast<int>(row / d.Y);
      const int xa = sg * 32, xb = min(d.X, xa + 32);
      const int* r = labels + row * d.X;
      int cur = 0, x0 = 0;
      for (int x = xa; x <= xb; ++x) {
        const int l = x < xb ? r[x] : 0;
        if (l == cur) continue;
        if (cur > 0 && cur <= n_labels) {
          // insert the run [x0, x) of label cur
          unsigned h = (static_cast<unsigned>(cur) * 2654435761u) >> 24;      // 8-bit hash
          bool done = false;
          for (int probe = 0; probe < 8 && !done; ++probe, h = (h + 1) & (kTab - 1)) {
            const int old = atomicCAS(&t_lab[h], 0, cur);
            if (old == 0 || old == cur) {
              atomicMin(&t_box[h][0], z); atomicMin(&t_box[h][1], y); atomicMin(&t_box[h][2], x0);
              atomicMax(&t_box[h][3], z + 1); atomicMax(&t_box[h][4], y + 1); atomicMax(&t_box[h][5], x);
              atomicAdd(&t_cnt[h], static_cast<unsigned long long>(x - x0));
              done = true;
            }
          }
          if (!done) bbox_flush(bbox, sizes, cur, z, y, x0, z + 1, y + 1, x, static_cast<unsigned long long>(x - x0));
        }
        cur = l; x0 = x;
      }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < kTab; k += 256)
      if (t_lab[k] > 0) bbox_flush(bbox, sizes, t_lab[k], t_box[k][0], t_box[k][1], t_box[k][2], t_box[k][3], t_box[k][4], t_box[k][5], t_cnt[k]);
    __syncthreads();
  }
}

// ---- component masks: out[offsets[i] + local] = labels[box_i voxel] == ids[i]   (x_voids of Voids.count_voids)
__global__ void __launch_bounds__(256) crop_masks_kernel(const int* __restrict__ labels, Dim3 d, const int* __restrict__ boxes,
                                                          const int* __restrict__ ids, const long long* __restrict__ offsets,
                                                          uint8_t* __restrict__ out) {
  const int i = blockIdx.y;
  const int* b = boxes + 6 * i;
  const int wz = b[3] - b[0], wy = b[4] - b[1], wx = b[5] - b[2];
  const long long nv = static_cast<long long>(wz) * wy * wx;
  const int id = ids[i];
  uint8_t* o = out + offsets[i];
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < nv; e += stride) {
    const int x = static_cast<int>(e % wx);
    const long long t = e / wx;
    const int y = static_cast<int>(t % wy), z = static_cast<int>(t / wy);
    o[e] = labels[(static_cast<long long>(b[0] + z) * d.Y + b[1] + y) * d.X + b[2] + x] == id ? 1 : 0;
  }
}

// ---- V[box_i] = mask_i for i = 0..n-1 IN ORDER (numpy slice assignment: a later box overwrites an earlier one, zeros
// included).  Pass 0: owner[v] = max over boxes covering v of (i + 1); pass 1: the owner writes.
template <int PASS>
__global__ void __launch_bounds__(256) paint_boxes_kernel(const uint8_t* __restrict__ masks, const long long* __restrict__ offsets,
                                                           const int* __restrict__ boxes, Dim3 d, uint8_t* vol, unsigned* owner) {
  const int i = blockIdx.y;
  const int* b = boxes + 6 * i;
  const int wz = b[3] - b[0], wy = b[4] - b[1], wx = b[5] - b[2];
  const long long nv = static_cast<long long>(wz) * wy * wx;
  const uint8_t* m = masks + offsets[i];
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < nv; e += stride) {
    const int x = static_cast<int>(e % wx);
    const long long t = e / wx;
    const int y = static_cast<int>(t % wy), z = static_cast<int>(t / wy);
    const long long v = (static_cast<long long>(b[0] + z) * d.Y + b[1] + y) * d.X + b[2] + x;
    if (PASS == 0) atomicMax(owner + v, static_cast<unsigned>(i + 1));
    else if (owner[v] == static_cast<unsigned>(i + 1)) vol[v] = m[e];
  }
}

// ---- flags[tile] = 1 if any voxel of the wd^3 tile of the regular grid is non-zero (tiles z-major, like Grid.points).
// item = the wd * ES contiguous bytes of one tile in one row, OR-reduced with the widest aligned loads; adjacent threads
// take adjacent tiles of the same row (coalesced).
__global__ void __launch_bounds__(256) tile_any_kernel(const uint8_t* __restrict__ vol, Dim3 d, int wd, int es, uint8_t* flags) {
  const int ty = d.Y / wd, tx = d.X / wd, tz = d.Z / wd;
  const long long items = static_cast<long long>(tz) * wd * ty * wd * tx;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const int nb = wd * es;
  for (long long it = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; it < items; it += stride) {
    const int c = static_cast<int>(it % tx);
    const long long row = it / tx;
    const int y = static_cast<int>(row % (static_cast<long long>(ty) * wd)), z = static_cast<int>(row / (static_cast<long long>(ty) * wd));
    const uint8_t* p = vol + ((static_cast<long long>(z) * d.Y + y) * d.X + static_cast<long long>(c) * wd) * es;
    unsigned acc = 0;
    int k = 0;
    if ((reinterpret_cast<uintptr_t>(p) & 15) == 0)
      for (; k + 16 <= nb; k += 16) { const uint4 w = *reinterpret_cast<const uint4*>(p + k); acc |= w.x | w.y | w.z | w.w; }
    else if ((reinterpret_cast<uintptr_t>(p) & 3) == 0)
      for (; k + 4 <= nb; k += 4) acc |= *reinterpret_cast<const unsigned*>(p + k);
    for (; k < nb; ++k) acc |= p[k];
    if (acc) flags[(static_cast<long long>(z / wd) * ty + y / wd) * tx + c] = 1;
  }
}

inline int grid_for(long long items, int threads) {
  long long b = (items + threads - 1) / threads;
  const long long cap = static_cast<long long>(kNumSMs) * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return static_cast<int>(b);
}
}  // namespace
}  // namespace te

using namespace te;

extern "C" int64_t te_label_workspace_bytes(int64_t nvox) {
  if (nvox < 0) return 0;
  return nvox * 4 + ((nvox + kScanItems - 1) / kScanItems + 1) * 4 + 256;
}

extern "C" int te_label(const void* vol, int vol_dtype, const int64_t vol_shape[3], int connectivity, int32_t* labels,
                        int64_t* n_labels, void* workspace, void* stream) {
  TE_CHECK_ARG(vol_shape[0] >= 0 && vol_shape[1] >= 0 && vol_shape[2] >= 0, "te_label: negative shape");
  const long long n = static_cast<long long>(vol_shape[0]) * vol_shape[1] * vol_shape[2];
  TE_CHECK_ARG(n < 2147483647LL, "te_label: volumes of up to 2^31 - 1 voxels are supported");
  TE_CHECK_ARG(connectivity == 6 || connectivity == 26, "te_label: connectivity must be 6 (faces) or 26 (faces, edges, corners)");
  TE_CHECK_ARG(n_labels, "te_label: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (n == 0) { TE_CUDA(cudaMemsetAsync(n_labels, 0, 8, st)); return TE_OK; }
  TE_CHECK_ARG(vol && labels && workspace, "te_label: null pointer");
  const int es = dtype_size(vol_dtype);
  TE_CHECK_ARG(es == 1 || es == 2 || es == 4, "te_label: unsupported vol_dtype %d", vol_dtype);
  int* parent = static_cast<int*>(workspace);
  int* counts = parent + n;
  const Dim3 d{static_cast<int>(vol_shape[0]), static_cast<int>(vol_shape[1]), static_cast<int>(vol_shape[2])};
  const int g = grid_for(n, 256);
  // non-zero test on the raw bits (-0.0 would count as foreground; masks and label images never hold it)
  // tile-local pre-merge in shared memory, then the pairs across tile faces (TOMOENC_LABEL_NO_TILES=1: the one-level
  // algorithm, kept for A/B timing)
  static const bool no_tiles = getenv("TOMOENC_LABEL_NO_TILES") != nullptr;
  if (no_tiles) {
    const int gi = grid_for(static_cast<long long>(d.Z) * d.Y * ((d.X + 31) / 32) * 32, 256);
    if (es == 1) label_init_kernel<uint8_t><<<gi, 256, 0, st>>>(static_cast<const uint8_t*>(vol), parent, d);
    else if (es == 2) label_init_kernel<uint16_t><<<gi, 256, 0, st>>>(static_cast<const uint16_t*>(vol), parent, d);
    else label_init_kernel<uint32_t><<<gi, 256, 0, st>>>(static_cast<const uint32_t*>(vol), parent, d);
    TE_LAUNCH_CHECK();
    if (connectivity == 6) label_merge_kernel<6, false><<<g, 256, 0, st>>>(parent, d);
    else label_merge_kernel<26, false><<<g, 256, 0, st>>>(parent, d);
    TE_LAUNCH_CHECK();
  } else {
    const int tx = (d.X + kTX - 1) / kTX, ty = (d.Y + kTY - 1) / kTY, tz = (d.Z + kTZ - 1) / kTZ;
    const unsigned tiles = static_cast<unsigned>(tx) * ty * tz;
#define TE_LABEL_TILE(T)                                                                                              \
    if (connectivity == 6) label_tile_kernel<T, 6><<<tiles, 256, 0, st>>>(static_cast<const T*>(vol), parent, d, tx, ty); \
    else label_tile_kernel<T, 26><<<tiles, 256, 0, st>>>(static_cast<const T*>(vol), parent, d, tx, ty);
    if (es == 1) { TE_LABEL_TILE(uint8_t) } else if (es == 2) { TE_LABEL_TILE(uint16_t) } else { TE_LABEL_TILE(uint32_t) }
#undef TE_LABEL_TILE
    TE_LAUNCH_CHECK();
    if (connectivity == 6) label_merge_kernel<6, true><<<g, 256, 0, st>>>(parent, d);
    else label_merge_kernel<26, true><<<g, 256, 0, st>>>(parent, d);
    TE_LAUNCH_CHECK();
  }
  label_compress_kernel<<<g, 256, 0, st>>>(parent, n);
  TE_LAUNCH_CHECK();
  const int nblocks = static_cast<int>((n + kScanItems - 1) / kScanItems);
  label_count_kernel<<<nblocks, 256, 0, st>>>(parent, n, counts);
  TE_LAUNCH_CHECK();
  label_scan_kernel<<<1, 1024, 0, st>>>(counts, nblocks, reinterpret_cast<long long*>(n_labels));
  TE_LAUNCH_CHECK();
  label_rank_kernel<<<nblocks, 256, 0, st>>>(parent, n, counts, labels);
  TE_LAUNCH_CHECK();
  label_assign_kernel<<<g, 256, 0, st>>>(parent, n, labels);
  TE_LAUNCH_CHECK();
  return TE_OK;
}

extern "C" int te_label_stats(const int32_t* labels, const int64_t vol_shape[3], int64_t n_labels, int32_t* bbox6, int64_t* sizes,
                              void* stream) {
  TE_CHECK_ARG(n_labels >= 0, "te_label_stats: n_labels < 0");
  if (n_labels == 0) return TE_OK;
  TE_CHECK_ARG(labels && bbox6 && sizes, "te_label_stats: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const Dim3 d{static_cast<int>(vol_shape[0]), static_cast<int>(vol_shape[1]), static_cast<int>(vol_shape[2])};
  bbox_init_kernel<<<grid_for(n_labels, 256), 256, 0, st>>>(bbox6, reinterpret_cast<long long*>(sizes), n_labels);
  TE_LAUNCH_CHECK();
  const long long items = static_cast<long long>(d.Z) * d.Y * ((d.X + 31) / 32);
  if (items > 0) {
    bbox_kernel<<<grid_for((items + 7) / 8, 256), 256, 0, st>>>(labels, d, n_labels, bbox6, reinterpret_cast<long long*>(sizes));
    TE_LAUNCH_CHECK();
  }
  return TE_OK;
}

extern "C" int te_label_crop_masks(const int32_t* labels, const int64_t vol_shape[3], const int32_t* boxes6, const int32_t* ids,
                                   int64_t n, const int64_t* offsets, int64_t max_box_voxels, uint8_t* out, void* stream) {
  TE_CHECK_ARG(n >= 0 && n <= 65535 * 1LL, "te_label_crop_masks: n must be in [0, 65535] per call");
  if (n == 0 || max_box_voxels <= 0) return TE_OK;
  TE_CHECK_ARG(labels && boxes6 && ids && offsets && out, "te_label_crop_masks: null pointer");
  const Dim3 d{static_cast<int>(vol_shape[0]), static_cast<int>(vol_shape[1]), static_cast<int>(vol_shape[2])};
  long long gx = (max_box_voxels + 255) / 256;
  if (gx > 256) gx = 256;
  crop_masks_kernel<<<dim3(static_cast<unsigned>(gx), static_cast<unsigned>(n)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      labels, d, boxes6, ids, reinterpret_cast<const long long*>(offsets), out);
  TE_LAUNCH_CHECK();
  return TE_OK;
}

extern "C" int te_paint_boxes(const uint8_t* masks, const int64_t* offsets, const int32_t* boxes6, int64_t n, int64_t max_box_voxels,
                              uint8_t* vol, const int64_t vol_shape[3], uint32_t* owner, void* stream) {
  TE_CHECK_ARG(n >= 0 && n <= 65535 * 1LL, "te_paint_boxes: n must be in [0, 65535] per call");
  if (n == 0 || max_box_voxels <= 0) return TE_OK;
  TE_CHECK_ARG(masks && offsets && boxes6 && vol && owner, "te_paint_boxes: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const Dim3 d{static_cast<int>(vol_shape[0]), static_cast<int>(vol_shape[1]), static_cast<int>(vol_shape[2])};
  const long long nvox = static_cast<long long>(d.Z) * d.Y * d.X;
  TE_CUDA(cudaMemsetAsync(owner, 0, static_cast<size_t>(nvox) * 4, st));
  long long gx = (max_box_voxels + 255) / 256;
  if (gx > 256) gx = 256;
  const dim3 g(static_cast<unsigned>(gx), static_cast<unsigned>(n));
  paint_boxes_kernel<0><<<g, 256, 0, st>>>(masks, reinterpret_cast<const long long*>(offsets), boxes6, d, vol, owner);
  TE_LAUNCH_CHECK();
  paint_boxes_kernel<1><<<g, 256, 0, st>>>(masks, reinterpret_cast<const long long*>(offsets), boxes6, d, vol, owner);
  TE_LAUNCH_CHECK();
  return TE_OK;
}

extern "C" int te_tile_any(const void* vol, int elem_size, const int64_t vol_shape[3], int32_t width, uint8_t* flags, void* stream) {
  TE_CHECK_ARG(elem_size == 1 || elem_size == 2 || elem_size == 4, "te_tile_any: elem_size must be 1, 2 or 4");
  TE_CHECK_ARG(width > 0, "te_tile_any: width must be positive");
  const Dim3 d{static_cast<int>(vol_shape[0]), static_cast<int>(vol_shape[1]), static_cast<int>(vol_shape[2])};
  const long long ntiles = static_cast<long long>(d.Z / width) * (d.Y / width) * (d.X / width);
  const long long n = static_cast<long long>(d.Z) * d.Y * d.X;
  if (ntiles == 0 || n == 0) return TE_OK;
  TE_CHECK_ARG(vol && flags, "te_tile_any: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TE_CUDA(cudaMemsetAsync(flags, 0, static_cast<size_t>(ntiles), st));
  const long long items = static_cast<long long>(d.Z / width) * width * (d.Y / width) * width * (d.X / width);
  tile_any_kernel<<<grid_for(items, 256), 256, 0, st>>>(static_cast<const uint8_t*>(vol), d, width, elem_size, flags);
  TE_LAUNCH_CHECK();
  return TE_OK;
}

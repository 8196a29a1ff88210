// One-shot all-reduce of SMALL fp64 vectors over NVLink peer memory (one process per GPU, one box).
//
// The only per-layer exchange of the data-parallel training step is the SyncBN reduction: 2 C
// per-channel sums per BatchNormalization layer, forward and backward -- 34 latency-bound all-reduces of 0.25-4 KB per
// step of M_a01, each ~20 us through NCCL even inside a CUDA graph (~1 ms of an N = 8 step of 7 ms).  Here every rank owns
// a buffer of slots that all peers map through CUDA IPC; a reduction is ONE single-CTA kernel per rank: publish the
// local vector + an epoch flag in the own slot (st.release.sys), wait for every peer's flag of the same epoch
// (ld.acquire.sys through the NVLink mapping), add the peers' vectors in RANK ORDER (every rank computes bit-identical
// sums) and write the result over the input.  No host involvement, no NCCL launch: graph-capturable (the epoch lives in
// device memory and advances with every replay).
//
// Slot reuse: slot = epoch % nslots.  A rank finishes reduction e only after every peer published e, and publishes e + 1
// only after that, so no rank is ever more than one reduction ahead of the slowest reader of its slots: a ring of 64
// slots is never overwritten under a reader.
#include "te_common.cuh"

#include <cstring>

struct te_peer {
  int rank, world;
  int slot_doubles, nslots;
  size_t slot_bytes;
  uint8_t* local;                 // [nslots][slot_bytes]: 16-byte header (epoch flag) + slot_doubles doubles
  uint8_t* peers[16];             // peers[rank] == local
  unsigned long long* epoch;      // device counter of completed reductions
  int device;
};

namespace te {
namespace {

constexpr int kMaxPeers = 16;

struct PeerArgs {
  uint8_t* bufs[kMaxPeers];
  unsigned long long* epoch;
  double* data;
  int count, rank, world, nslots;
  size_t slot_bytes;
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ double ld_volatile_f64(const double* p) {
  double v;
  asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(256) peer_allreduce_kernel(const PeerArgs a) {
  __shared__ unsigned long long s_epoch;
  if (threadIdx.x == 0) s_epoch = *a.epoch + 1;
  __syncthreads();
  const unsigned long long epoch = s_epoch;
  const size_t slot_off = static_cast<size_t>(epoch % static_cast<unsigned long long>(a.nslots)) * a.slot_bytes;
  double* mine = reinterpret_cast<double*>(a.bufs[a.rank] + slot_off + 16);
  for (int i = threadIdx.x; i < a.count; i += blockDim.x) mine[i] = a.data[i];
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) st_release_sys(reinterpret_cast<unsigned long long*>(a.bufs[a.rank] + slot_off), epoch);
  // one thread per peer waits for that peer's flag of this epoch (bounded: a lost peer ends in a trap, not a hang)
  if (threadIdx.x < a.world && threadIdx.x != a.rank) {
    const unsigned long long* flag = reinterpret_cast<const unsigned long long*>(a.bufs[threadIdx.x] + slot_off);
    const long long t0 = clock64();
    while (ld_acquire_sys(flag) < epoch) {
      if (clock64() - t0 > 20000000000ll) __trap();
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < a.count; i += blockDim.x) {
    double acc = 0.0;
    for (int r = 0; r < a.world; ++r)
      acc += r == a.rank ? a.data[i] : ld_volatile_f64(reinterpret_cast<const double*>(a.bufs[r] + slot_off + 16) + i);
    a.data[i] = acc;
  }
  __syncthreads();
  if (threadIdx.x == 0) *a.epoch = epoch;
}

}  // namespace
}  // namespace te

using namespace te;

extern "C" int te_peer_create(int32_t rank, int32_t world, int32_t slot_doubles, int32_t nslots, te_peer** out) {
  TE_CHECK_ARG(out && world >= 2 && world <= kMaxPeers && rank >= 0 && rank < world && slot_doubles > 0 && nslots >= 4,
               "te_peer_create: bad argument (2..16 ranks of one box)");
  te_peer* p = new te_peer();
  p->rank = rank; p->world = world; p->slot_doubles = slot_doubles; p->nslots = nslots;
  p->slot_bytes = 16 + static_cast<size_t>(slot_doubles) * 8;
  p->slot_bytes = (p->slot_bytes + 127) / 128 * 128;
  for (int i = 0; i < kMaxPeers; ++i) p->peers[i] = nullptr;
  cudaError_t e = cudaGetDevice(&p->device);
  // a plain cudaMalloc allocation of its own: the IPC handle then names exactly this buffer
  if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&p->local), p->slot_bytes * nslots + 256);
  if (e == cudaSuccess) e = cudaMemset(p->local, 0, p->slot_bytes * nslots + 256);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    set_error("te_peer_create: %s", cudaGetErrorString(e));
    delete p;
    return TE_ERR_CUDA;
  }
  p->epoch = reinterpret_cast<unsigned long long*>(p->local + p->slot_bytes * nslots);
  p->peers[rank] = p->local;
  *out = p;
  return TE_OK;
}

// handle64: 64 bytes (cudaIpcMemHandle_t) to be exchanged between the ranks (all-gather of bytes, done by the caller)
extern "C" int te_peer_export(te_peer* p, void* handle64) {
  TE_CHECK_ARG(p && handle64, "te_peer_export: null pointer");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  cudaIpcMemHandle_t h;
  TE_CUDA(cudaIpcGetMemHandle(&h, p->local));
  memcpy(handle64, &h, 64);
  return TE_OK;
}

// handles: world x 64 bytes in rank order (the own entry is ignored)
extern "C" int te_peer_import(te_peer* p, const void* handles) {
  TE_CHECK_ARG(p && handles, "te_peer_import: null pointer");
  for (int r = 0; r < p->world; ++r) {
    if (r == p->rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, static_cast<const uint8_t*>(handles) + 64 * r, 64);
    void* ptr = nullptr;
    TE_CUDA(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
    p->peers[r] = static_cast<uint8_t*>(ptr);
  }
  return TE_OK;
}

// data[0 .. count) (device, fp64) <- sum over the ranks, in place; every rank must make the same sequence of calls
extern "C" int te_peer_allreduce_f64(te_peer* p, double* data, int32_t count, void* stream) {
  TE_CHECK_ARG(p && data && count > 0 && count <= p->slot_doubles, "te_peer_allreduce_f64: count outside (0, slot size]");
  PeerArgs a;
  for (int r = 0; r < kMaxPeers; ++r) a.bufs[r] = r < p->world ? p->peers[r] : nullptr;
  for (int r = 0; r < p->world; ++r) TE_CHECK_ARG(a.bufs[r] != nullptr, "te_peer_allreduce_f64: peer %d not imported", r);
  a.epoch = p->epoch; a.data = data; a.count = count; a.rank = p->rank; a.world = p->world; a.nslots = p->nslots;
  a.slot_bytes = p->slot_bytes;
  peer_allreduce_kernel<<<1, 256, 0, static_cast<cudaStream_t>(stream)>>>(a);
  TE_LAUNCH_CHECK();
  return TE_OK;
}

extern "C" int te_peer_destroy(te_peer* p) {
  if (!p) return TE_OK;
  for (int r = 0; r < p->world; ++r)
    if (r != p->rank && p->peers[r]) cudaIpcCloseMemHandle(p->peers[r]);
  cudaFree(p->local);
  delete p;
  return TE_OK;
}

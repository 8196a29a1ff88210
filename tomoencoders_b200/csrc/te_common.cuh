// Shared host-side helpers of libtomoenc.so: error reporting, launch accounting, dtype helpers.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <atomic>

#include "../../include/tomoenc.h"

namespace te {

constexpr int kNumSMs = 148;  // B200

void set_error(const char* fmt, ...);
extern std::atomic<long long> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

#define TE_CHECK_ARG(cond, ...)         \
  do {                                  \
    if (!(cond)) {                      \
      ::te::set_error(__VA_ARGS__);     \
      return TE_ERR_ARG;                \
    }                                   \
  } while (0)

#define TE_CUDA(call)                                                                         \
  do {                                                                                        \
    cudaError_t e__ = (call);                                                                 \
    if (e__ != cudaSuccess) {                                                                 \
      ::te::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__,      \
                      __LINE__);                                                              \
      return TE_ERR_CUDA;                                                                     \
    }                                                                                         \
  } while (0)

#define TE_LAUNCH_CHECK()                                                                     \
  do {                                                                                        \
    ::te::count_launch();                                                                     \
    cudaError_t e__ = cudaGetLastError();                                                     \
    if (e__ != cudaSuccess) {                                                                 \
      ::te::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__), __FILE__,  \
                      __LINE__);                                                              \
      return TE_ERR_CUDA;                                                                     \
    }                                                                                         \
  } while (0)

inline int dtype_size(int dt) {
  switch (dt) {
    case TE_U8: return 1;
    case TE_U16: case TE_F16: case TE_BF16: return 2;
    case TE_F32: return 4;
    default: return 0;
  }
}


// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-DEVICE attribute of a kernel: the "already set" cache is a bit
// mask over device ordinals (a process that touches a second GPU must opt in there too), not one process-wide flag.
template <typename K>
inline cudaError_t smem_attr_once(K kernel, int bytes, unsigned long long* device_mask) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  const unsigned long long bit = 1ull << (dev & 63);
  if (dev < 64 && (*device_mask & bit)) return cudaSuccess;
  e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess && dev < 64) *device_mask |= bit;
  return e;
}

}  // namespace te

// Error reporting / bookkeeping entry points of the C ABI (include/tomoenc.h).
#include "te_common.cuh"

namespace te {
static thread_local char g_err[1024] = "";
std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace te

extern "C" int te_version(void) { return 100; }
extern "C" const char* te_last_error(void) { return te::g_err; }
extern "C" int64_t te_launch_count(void) { return te::g_launches.load(); }
extern "C" int64_t te_launch_count_add(int64_t n) {
  te::count_launch(static_cast<int>(n));                  // n < 0: a capture pass went through the counting entry points
  return te::g_launches.load();
}

// Host volume -> device copy of vol[::s, ::s, :] (every s-th row of every s-th plane, rows whole) as ONE strided 3-D DMA:
// 1 / s^2 of the bytes cross PCIe.  The min / max pre-pass of a host volume (_find_min_max, misc/voxel_processing.py:91-104)
// needs vol[::s, ::s, ::s] only; the x stride is applied on the device.
extern "C" int te_upload_sampled(const void* host_vol, int32_t elem_bytes, const int64_t vol_shape[3], int32_t s, void* dev_out,
                                 void* stream) {
  TE_CHECK_ARG(host_vol && dev_out && vol_shape && s >= 1 && elem_bytes >= 1, "te_upload_sampled: bad argument");
  const size_t Z = static_cast<size_t>(vol_shape[0]), Y = static_cast<size_t>(vol_shape[1]), X = static_cast<size_t>(vol_shape[2]);
  TE_CHECK_ARG(Z > 0 && Y > 0 && X > 0, "te_upload_sampled: empty volume");
  const size_t rowb = X * static_cast<size_t>(elem_bytes);
  const size_t nz = (Z + s - 1) / s, ny = (Y + s - 1) / s;
  cudaMemcpy3DParms p = {};
  // source: rows s * rowb apart, "slices" = pitch * ysize = s planes apart
  p.srcPtr = make_cudaPitchedPtr(const_cast<void*>(host_vol), static_cast<size_t>(s) * rowb, rowb, Y);
  p.dstPtr = make_cudaPitchedPtr(dev_out, rowb, rowb, ny);
  p.extent = make_cudaExtent(rowb, ny, nz);
  p.kind = cudaMemcpyHostToDevice;
  TE_CUDA(cudaMemcpy3DAsync(&p, static_cast<cudaStream_t>(stream)));
  return TE_OK;
}

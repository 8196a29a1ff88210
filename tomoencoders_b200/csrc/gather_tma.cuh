// TMA-staged gather / scatter fast paths (gather_tma.cu).  Each function returns false when the
// TMA path does not apply (row pitch or patch rows not a multiple of 16 bytes, oversized planes,
// unaligned buffers): the caller then uses the SIMT kernels of gather_scatter.cu.  When it returns
// true the kernel has been enqueued and *err holds the launch status.
#pragma once
#include "te_common.cuh"

namespace te {

bool gather_tma(const void* vol, int es, const int64_t vol_shape[3], const uint32_t* points, const uint32_t* widths,
                int64_t n, const int32_t patch[3], void* out, cudaStream_t st, cudaError_t* err);
// unbinned patches with short rows at arbitrary corners (row-group pattern); perm = optional processing order
bool gather_rows_tma(const void* vol, int es, const int64_t vol_shape[3], const uint32_t* points, const uint32_t* perm,
                     int64_t n, const int32_t patch[3], void* out, cudaStream_t st, cudaError_t* err);
bool gather_norm_tma(const void* vol, int vol_dtype, const int64_t vol_shape[3], const uint32_t* points,
                     const uint32_t* widths, int64_t n, const int32_t patch[3], int rescale, int do_clip, float lo,
                     float hi, void* out, int out_dtype, cudaStream_t st, cudaError_t* err);
bool scatter_tma(const void* sub_vols, int es, const uint32_t* points, int64_t n, const int32_t width[3],
                 void* vol_out, const int64_t vol_shape[3], cudaStream_t st, cudaError_t* err);

}  // namespace te

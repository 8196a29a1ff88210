// Write-only bandwidth probe: which store flavour fills 4 GiB fastest on B200?
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

template <int MODE>
__global__ void __launch_bounds__(256) fill_kernel(uint4* out, size_t n16) {
  const uint4 v = make_uint4(1, 2, 3, 4);
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) {
    if (MODE == 0) out[i] = v;
    else if (MODE == 1) asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(out + i), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
    else if (MODE == 2) asm volatile("st.global.wt.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(out + i), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
    else asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(out + i), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
  }
}
// contiguous per block: block b writes a contiguous span
__global__ void __launch_bounds__(256) fill_span_kernel(uint4* out, size_t n16) {
  const uint4 v = make_uint4(1, 2, 3, 4);
  const size_t per = (n16 + gridDim.x - 1) / gridDim.x;
  const size_t a = per * blockIdx.x, b = a + per < n16 ? a + per : n16;
  for (size_t i = a + threadIdx.x; i < b; i += 256) out[i] = v;
}
// TMA bulk store: each CTA stages CH bytes in smem once and bulk-stores it over and over
template <int CH>
__global__ void __launch_bounds__(128) fill_bulk_kernel(uint8_t* out, size_t nbytes) {
  extern __shared__ __align__(128) uint8_t sm[];
  for (int i = threadIdx.x; i < CH / 16; i += 128) reinterpret_cast<uint4*>(sm)[i] = make_uint4(1, 2, 3, 4);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    const size_t nch = nbytes / CH;
    int k = 0;
    for (size_t c = blockIdx.x; c < nch; c += gridDim.x, ++k) {
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out + c * CH), "r"((uint32_t)__cvta_generic_to_shared(sm)), "r"(CH) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      if ((k & 7) == 7) asm volatile("cp.async.bulk.wait_group.read 4;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

int main() {
  const size_t nbytes = 4ull << 30;
  uint8_t* buf; CK(cudaMalloc(&buf, nbytes));
  uint8_t* fl; CK(cudaMalloc(&fl, 512 << 20));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto run = [&](const char* name, auto fn) {
    float best = 1e9;
    for (int r = 0; r < 4; ++r) {
      cudaMemset(fl, r, 512 << 20);
      cudaEventRecord(e0); fn(); cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    cudaError_t e = cudaGetLastError();
    printf("%-44s %8.1f us  %7.0f GB/s %s\n", name, best * 1e3, nbytes / best / 1e6, e == cudaSuccess ? "" : cudaGetErrorString(e));
  };
  run("cudaMemsetAsync", [&] { cudaMemsetAsync(buf, 7, nbytes); });
  for (int g : {148 * 2, 148 * 4, 148 * 8, 148 * 16, 148 * 64}) {
    char nm[64];
    snprintf(nm, 64, "st.v4 grid-stride, %d blocks", g); run(nm, [&] { fill_kernel<0><<<g, 256>>>((uint4*)buf, nbytes / 16); });
    snprintf(nm, 64, "st.cs.v4 grid-stride, %d blocks", g); run(nm, [&] { fill_kernel<1><<<g, 256>>>((uint4*)buf, nbytes / 16); });
  }
  run("st.wt.v4, 1184 blocks", [&] { fill_kernel<2><<<1184, 256>>>((uint4*)buf, nbytes / 16); });
  run("st.L1::no_allocate.v4, 1184 blocks", [&] { fill_kernel<3><<<1184, 256>>>((uint4*)buf, nbytes / 16); });
  run("st.v4 contiguous span per block, 1184 blocks", [&] { fill_span_kernel<<<1184, 256>>>((uint4*)buf, nbytes / 16); });
  run("st.v4 contiguous span per block, 16384 blocks", [&] { fill_span_kernel<<<16384, 256>>>((uint4*)buf, nbytes / 16); });
  cudaFuncSetAttribute(fill_bulk_kernel<65536>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  for (int g : {148, 296, 444}) {
    char nm[64];
    snprintf(nm, 64, "bulk store 16 KB chunks, %d blocks", g); run(nm, [&] { fill_bulk_kernel<16384><<<g, 128, 16384>>>(buf, nbytes); });
    snprintf(nm, 64, "bulk store 64 KB chunks, %d blocks", g); run(nm, [&] { fill_bulk_kernel<65536><<<g, 128, 65536>>>(buf, nbytes); });
  }
  // copy for reference
  uint8_t* src; CK(cudaMalloc(&src, nbytes));
  run("cudaMemcpyAsync d2d (read+write = 2x)", [&] { cudaMemcpyAsync(buf, src, nbytes, cudaMemcpyDeviceToDevice); });
  return 0;
}

// CUDA-core layer kernels: the fp32 check mode of every layer, plus the layers that stay on CUDA
// cores in the tcgen05 modes (Cin = 1 first conv, max-pool, 1x1x1 head when not fused, dense).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace te {

// NDHWC activation tensor, possibly a channel slice [coff, coff + C) of a buffer with ctot channels.
struct TView {
  void* ptr;
  int ctot, coff, C;
  int N, D, H, W;
  __host__ __device__ long long voxels() const { return static_cast<long long>(N) * D * H * W; }
};

enum ActType { ACT_F32 = 0, ACT_BF16 = 1, ACT_F16 = 2 };

// 3x3x3 'same' conv, w = [27][Cin][Cout] fp32 (Keras layout, BN folded), fp32 accumulation.
// overflow (optional, fp16 activations): OR-ed with 1 when the tensor-core first layer produces a value outside the finite
// fp16 range (see ConvUmmaLaunch::overflow).
cudaError_t simt_conv3(const TView& in, const TView& out, const float* w, const float* bias, int lrelu, float alpha,
                       int act_type, cudaStream_t st, int* overflow = nullptr);
// Tensor-core version of the Cin = 1 first layer (conv1_umma.cu); returns false when it does not apply.
bool conv1_umma_try(const TView& in, const TView& out, const float* w, const float* bias, int lrelu, float alpha, int act_type,
                    cudaStream_t st, cudaError_t* err, int* overflow = nullptr);
// Toeplitz formulation of the same layer (conv1_tz.cu: raw input rows as the A operand, no im2col builder), tried first.
bool conv1_tz_applies(const TView& in, const TView& out, int act_type);
bool conv1_tz_try(const TView& in, const TView& out, const float* w, const float* bias, int lrelu, float alpha, int act_type,
                  cudaStream_t st, cudaError_t* err, int* overflow = nullptr);
// transposed conv k=2, stride s ('same': output = s x input; gaps get bias only), w = [8][Cout][Cin].
cudaError_t simt_upconv(const TView& in, const TView& out, const float* w, const float* bias, int stride, int lrelu,
                        float alpha, int act_type, cudaStream_t st);
// max-pool p^3, stride p (extents divisible by p).
cudaError_t simt_maxpool(const TView& in, const TView& out, int pool, int act_type, cudaStream_t st);
// 1x1x1 conv to one channel + sigmoid; any of labels / probs / logits may be null.
cudaError_t simt_head(const TView& in, const float* w, float b, uint8_t* labels, float* probs, float* logits,
                      int act_type, cudaStream_t st);
// dense: out[n][j] = act(sum_i x[n][i] w[i][j] + b[j]); x is the flattened NDHWC tensor of `in_type`;
// out_type selects the element type of out (dense row-major (N, nout)).  acc_scratch (optional, n * nout fp32):
// enables the split-K path for long input vectors.
cudaError_t simt_dense(const void* x, int in_type, long long n, int nin, int nout, const float* w, const float* b,
                       int lrelu, float alpha, void* out, int out_type, cudaStream_t st, float* acc_scratch = nullptr);
// view -> dense fp32 NDHWC copy (test hook).
cudaError_t simt_to_f32(const TView& in, float* out, int act_type, cudaStream_t st);
// dense fp32 -> activation type (decoder input etc.)
cudaError_t simt_from_f32(const float* in, void* out, long long count, int act_type, cudaStream_t st);

}  // namespace te

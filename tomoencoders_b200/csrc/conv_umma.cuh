// Launch interface of the tcgen05 implicit-GEMM convolution kernel (conv_umma.cu).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace te {

// One launch = one layer over a batch of patches.  Activations are NDHWC with 16-bit elements
// (bf16 or fp16); the output may be a channel slice [out_coff, out_coff + Cout) of a wider buffer
// (this is how the skip concatenation disappears: both producers write into one buffer).
struct ConvUmmaLaunch {
  CUtensorMap tmap_in;      // 5-D (C, W, H, D, N) map of the input tensor, box (CB, 8+2h, 16+2h, TZ+2h, 1)
  CUtensorMap tmap_in2;     // second input tensor holding the last Cin2 channels (planar concat); unused if Cin2 == 0
  int Cin2;                 // channels taken from tmap_in2 (multiple of CB; Cin - Cin2 from tmap_in)
  const void* w_packed;     // pre-swizzled weights, see pack_conv_weights() in net.cu
  const void* w_packed_pair;  // CTA-pair image of the same weights (conv_umma_pack_pair) or nullptr: single-CTA kernel
  const float* bias;        // Cout fp32 (conv bias with BN folded in)
  void* out;                // output activations (16-bit) or nullptr when the head is fused
  int out_ctot, out_coff;   // channel pitch / offset of the output buffer
  int N, D, H, W;           // batch and INPUT spatial extents
  int Cin, Cout;
  int ntaps;                // 27 (3x3x3 'same' conv) or 1 (one sub-position of a transposed conv)
  int up_stride;            // 0: conv.  s > 0: transposed conv k=2, stride s; 8 sub-positions
  int act_lrelu;            // 1: LeakyReLU(alpha) after bias
  float alpha;
  int fp16;                 // 0: bf16 operands / activations, 1: fp16
  // fused 1x1x1 head (Cout -> 1): logits = dot(act, head_w) + head_b
  const float* head_w;      // nullptr = no fused head
  float head_b;
  uint8_t* head_labels;     // optional (n,D,H,W) uint8: logit > 0
  float* head_probs;        // optional fp32 sigmoid(logit)
  float* head_logits;       // optional fp32 logits
  // fused MaxPool3D(2) of the activated output (conv only; D, H, W even): written next to `out`
  void* pool_out;           // nullptr = no fused pool
  int pool_ctot, pool_coff;
  // fp16 launches: *overflow is OR-ed with 1 when an activation (fp32, before rounding) exceeds the largest finite fp16
  // value, i.e. the stored tensor holds +-inf; nullptr = not tracked
  int* overflow;
  // kernel configuration chosen by conv_umma_config()
  int CB, NT, TZ;
  // un-haloed kernels (transposed / pointwise convs): output through TMA box stores.  tmap_out[sub] is a
  // 5-D map (C, W, H, D, N) of the output voxels this sub-position writes (for a transposed conv: every
  // second voxel, base shifted by the sub-position), box (min(NT, 64), 8, 16, 1, 1), swizzle = box row
  // bytes.  Built by conv_umma_make_out_maps(); ignored when use_tma_out == 0.
  CUtensorMap tmap_out[8];
  int use_tma_out;
};
// Fills L.tmap_out / L.use_tma_out from L.out, L.out_ctot, L.out_coff, L.N/D/H/W, L.Cout, L.up_stride, L.NT,
// L.fp16 (un-haloed launches only; leaves use_tma_out = 0 when the layout does not qualify).
void conv_umma_make_out_maps(ConvUmmaLaunch* L);

struct ConvUmmaConfig { int CB, NT, TZ; int pair; };   // pair: the layer may run on CTA pairs (cta_group::2) when its tile count is even
// Picks (channel chunk, Cout tile, z planes per tile) for a layer; returns false if unsupported.
bool conv_umma_config(int Cin, int Cout, ConvUmmaConfig* cfg);
// Bytes of one packed weight tensor for the given layer shape.
size_t conv_umma_packed_bytes(int Cin, int Cout, int ntaps, int nsub, const ConvUmmaConfig& cfg);
// Packs fp32 weights w[sub][tap][cout][cin] (already BN-folded) into the kernel's layout.
void conv_umma_pack(const float* w, int Cin, int Cout, int ntaps, int nsub, const ConvUmmaConfig& cfg, int fp16,
                    void* dst_host);
// CTA-pair image (cfg.pair only): bytes and packing of w[tap][cout][cin] (3x3x3 conv, BN folded).
size_t conv_umma_pair_packed_bytes(int Cin, int Cout, const ConvUmmaConfig& cfg);
void conv_umma_pack_pair(const float* w, int Cin, int Cout, const ConvUmmaConfig& cfg, int fp16, void* dst_host);
// Enqueues the layer.  Returns cudaSuccess or the launch error.
cudaError_t conv_umma_launch(const ConvUmmaLaunch& L, cudaStream_t stream);
// Number of CTAs the launch uses (for accounting / tests).
int conv_umma_grid(const ConvUmmaLaunch& L);

}  // namespace te

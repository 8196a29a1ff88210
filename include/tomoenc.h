/* tomoenc.h -- C ABI of libtomoenc.so, the B200 (sm_100a) data plane behind the TomoEncoders hot path.
 *
 * The reference (aniketkt/TomoEncoders) is pure Python and has no FFI layer of its own; its boundary for
 * this path is the Python API (Patches / Grid / SurfaceSegmenter / SelfSupervisedCAE).  The Python
 * host code in tomoencoders_b200/ keeps that API and routes every data-plane call to the entry points
 * below through ctypes.  Each entry point cites the reference interface it replaces (paths relative
 * to /root/reference/tomo_encoders/).
 *
 * Conventions
 *   - all data pointers are DEVICE pointers (e.g. torch.Tensor.data_ptr()) unless marked "host";
 *   - the caller allocates every output; nothing is retained across calls except opaque handles;
 *   - `stream` is a cudaStream_t passed as void*; calls enqueue work and return without synchronising
 *     unless documented otherwise;
 *   - return value 0 = success, < 0 = error (TE_ERR_*); te_last_error() returns a message for the
 *     calling thread; nothing throws; a handle may be used from one host thread at a time;
 *   - coordinates are (z, y, x) corner points, uint32, exactly like Patches.points / Grid.points
 *     (structures/patches.py:66-67, structures/grid.py:105).
 */
#ifndef TOMOENC_H_
#define TOMOENC_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TE_OK 0
#define TE_ERR_ARG (-1)      /* invalid argument (shape, dtype, null pointer, unsupported configuration) */
#define TE_ERR_CUDA (-2)     /* a CUDA runtime / driver call failed */
#define TE_ERR_STATE (-3)    /* handle used in the wrong state (e.g. weights not set) */
#define TE_ERR_UNSUPPORTED (-4)

/* element types of volumes / patch arrays */
#define TE_U8 0
#define TE_U16 1
#define TE_F16 2
#define TE_F32 3
#define TE_BF16 4

/* arithmetic modes of the network entry points */
#define TE_MODE_BF16 0       /* tcgen05 path: bf16 operands, fp32 TMEM accumulation, bf16 activations */
#define TE_MODE_FP16 1       /* tcgen05 path: fp16 operands, fp32 TMEM accumulation, fp16 activations */
#define TE_MODE_FP32_CHECK 2 /* CUDA-core fp32 check mode (slow; used to bound the reduced-precision error) */

int te_version(void);
const char* te_last_error(void);
/* Number of kernel launches this library has enqueued since load (process-wide, monotonic). */
int64_t te_launch_count(void);
/* A host layer that replays kernels of this library through a captured CUDA graph (the launches then bypass the entry
 * points that count) reports the number of kernel nodes it replayed here, so the counter stays "kernels of this library
 * enqueued".  Returns the new total. */
int64_t te_launch_count_add(int64_t n);

/* ------------------------------------------------------------------------------------------------
 * Patch gather.  Replaces Patches.extract (structures/patches.py:748-781, with _calc_binning
 * :734-746 and slices :436-453) and Grid.extract (structures/grid.py:313-336):
 *     out[i] = vol[z:z+wz:b, y:y+wy:b, x:x+wx:b],  b = widths[i] // patch  (isotropic, >= 1)
 * out is (n, patch[0], patch[1], patch[2]) C-contiguous with the element type of vol; the copy is
 * bit-exact.  widths is (n,3) uint32 (for a Grid pass the scalar width replicated).  The binning
 * rules (isotropy, >= 1, widths % patch == 0) are validated by the host code before the call; the
 * kernel re-derives b = widths[i][0] / patch[0].  elem_size in {1,2,4}.
 *
 * flags: TE_HINT_ALIGNED promises that every patch is unbinned and that its x corner is a multiple
 * of 16 bytes (true for every Grid / regular-grid tiling with width * elem_size % 16 == 0); the
 * call may then take the TMA-staged kernel (one box load + one bulk store per group of planes).
 * The hint only selects the kernel: a violated promise costs speed, never correctness.
 */
#define TE_HINT_ALIGNED 1
/* TE_HINT_UNBINNED: the caller promises widths[i] == patch for every i (no decimation).  Lets te_gather pick the row-group
 * kernel for short rows at arbitrary corners (random patches of BASELINE configs[0] / configs[2]).  Like every hint: a
 * violated promise is a caller error here (the widths are not re-read), so only set it from a checked table. */
#define TE_HINT_UNBINNED 2
int te_gather(const void* vol, int elem_size, const int64_t vol_shape[3], const uint32_t* points,
              const uint32_t* widths, int64_t n, const int32_t patch[3], void* out, int flags, void* stream);

/* te_gather with a processing order: work slot s handles patch perm[s] (perm = a permutation of 0..n-1, device uint32);
 * out is still written in patch-index order, so the result is identical to te_gather.  A list of heavily overlapping
 * random patches (BASELINE configs[0] / configs[2]: n P^3 is many times the volume) gathered in (z, y)-sorted order keeps
 * the slab the running blocks read resident in the 126 MB L2 instead of re-fetching it from HBM for every patch. */
int te_gather_perm(const void* vol, int elem_size, const int64_t vol_shape[3], const uint32_t* points,
                   const uint32_t* widths, int64_t n, const int32_t patch[3], void* out, int flags,
                   const uint32_t* perm, void* stream);

/* Gather fused with the pre-network clip / rescale.  Replaces extract followed by the per-chunk
 * np.clip + _rescale_data of SurfaceSegmenter.predict_patches (neural_nets/surface_segmenter.py:
 * 272-275, misc/voxel_processing.py:81-89):
 *     out[i] = act((clip(vol[...], lo, hi) - lo) / (hi - lo + 1e-12))   computed in fp32
 * do_clip = 0 skips the clip (predict_embeddings, neural_nets/autoencoders.py:753); rescale = 0
 * skips both (min_max=None).  out_dtype in {TE_BF16, TE_F16, TE_F32}.  vol_dtype in
 * {TE_U8, TE_U16, TE_F16, TE_F32}.
 */
int te_gather_norm(const void* vol, int vol_dtype, const int64_t vol_shape[3], const uint32_t* points,
                   const uint32_t* widths, int64_t n, const int32_t patch[3], int rescale, int do_clip,
                   float lo, float hi, void* out, int out_dtype, void* stream);
/* te_gather_norm with a processing order (see te_gather_perm). */
int te_gather_norm_perm(const void* vol, int vol_dtype, const int64_t vol_shape[3], const uint32_t* points,
                        const uint32_t* widths, int64_t n, const int32_t patch[3], int rescale, int do_clip,
                        float lo, float hi, void* out, int out_dtype, const uint32_t* perm, void* stream);

/* Gather fused with the in-graph input standardisation of the ICIP porosity encoder (build_CAE_3D(stdinput=True),
 * /root/reference/scratchpad/IEEE_ICIP_paper/porosity_encoders.py:89-98, 298-302): after the optional global rescale
 * (rescale = 1: (v - lo) / (hi - lo + 1e-12), no clip, as in predict_embeddings) every patch is mapped to
 *     (v - min_patch) / (max_patch - min_patch + 1e-12)        (fp32, IEEE divisions)
 * minmax_scratch: device float32 scratch of 2 * n values (receives the per-patch raw extrema). */
int te_gather_standardize(const void* vol, int vol_dtype, const int64_t vol_shape[3], const uint32_t* points,
                          const uint32_t* widths, int64_t n, const int32_t patch[3], int rescale, float lo, float hi,
                          float* minmax_scratch, void* out, int out_dtype, void* stream);

/* Patch scatter / stitch.  Replaces Patches.fill_patches_in_volume (structures/patches.py:811-821)
 * and Grid.fill_patches_in_volume (structures/grid.py:338-348): sequential
 *     vol_out[z:z+wz, y:y+wy, x:x+wx] = sub_vols[i]
 * with "highest patch index wins" on overlap.  sub_vols is (n, wz, wy, wx) of the same element
 * size as vol_out (the host code performs numpy's assignment cast beforehand); all patches share
 * one width[3].  If `owner` is NULL the patches must be pairwise disjoint (the host code proves it);
 * otherwise owner is a uint32 scratch volume of vol_shape elements and the call resolves overlaps
 * deterministically in two passes (atomicMax of the patch index, then owner-only writes).
 * flags: TE_HINT_ALIGNED as for te_gather (disjoint path only).
 */
int te_scatter(const void* sub_vols, int elem_size, const uint32_t* points, int64_t n,
               const int32_t width[3], void* vol_out, const int64_t vol_shape[3], uint32_t* owner,
               int flags, void* stream);

/* min / max over vol[::s, ::s, ::s].  Replaces _find_min_max (misc/voxel_processing.py:91-104).
 * out2 = device float[2] {min, max}; scratch = device float[2 * 1024]. */
int te_minmax_strided(const void* vol, int vol_dtype, const int64_t vol_shape[3], int stride,
                      float* out2, float* scratch, void* stream);

/* Host volume -> device copy of vol[::s, ::s, :] (dense (ceil(Z/s), ceil(Y/s), X) on the device) as one strided 3-D DMA:
 * the min / max pre-pass of a HOST volume (_find_min_max) moves 1 / s^2 of the bytes; the x stride is applied on the device
 * (te_minmax_nd).  host_vol should be pinned (the copy is asynchronous on `stream` then). */
int te_upload_sampled(const void* host_vol, int32_t elem_bytes, const int64_t vol_shape[3], int32_t s, void* dev_out,
                      void* stream);

/* ------------------------------------------------------------------------------------------------
 * Volume normalisation / contrast / masking -- the calls every user of the path makes right before
 * (and after) it (misc/voxel_processing.py).  Views are described by up to
 * four extents (outer -> inner) and strides in ELEMENTS, so the sub-sampled views the reference
 * builds by slicing (vol[::s, ::s, ::s], vol[:, ::s, ::s, ::s]) are read in place.
 */

/* min / max of a strided view (first half of np.histogram's range search, modified_autocontrast
 * misc/voxel_processing.py:160-204).  out2 = device float[2]; scratch = device float[2 * 1024]. */
int te_minmax_nd(const void* data, int dtype, const int64_t shape[4], const int64_t strides[4],
                 float* out2, float* scratch, void* stream);

/* np.histogram(view, bins = edges) (misc/voxel_processing.py:189): counts[i] = #{edges[i] <= a <
 * edges[i+1]}, last bin closed; edges = device double[nbins + 1], monotone (the host passes numpy's
 * own np.linspace edges in the bin dtype numpy would choose); counts = device uint64[nbins], zeroed
 * by the call.  nbins <= 1024.  Bit-exact against numpy. */
int te_histogram_nd(const void* data, int dtype, const int64_t shape[4], const int64_t strides[4],
                    const double* edges, int32_t nbins, unsigned long long* counts, void* stream);

/* out = (in - lo) / denom, fp32 round-to-nearest subtract and divide (bit-exact against numpy float32
 * arithmetic).  Replaces _rescale_data (misc/voxel_processing.py:81-89) on device arrays and the
 * per-chunk body of normalize_volume_gpu (:207-253); denom = max - min + 1e-12 is formed by the host
 * in float32.  out_dtype in {TE_F32, TE_F16}; in == out allowed for equal element sizes; both
 * buffers 16-byte aligned. */
int te_rescale(const void* in, int in_dtype, void* out, int out_dtype, int64_t n, float lo, float denom,
               void* stream);

/* edge_map (misc/voxel_processing.py:106-122): out[z,y,x] = 1 where any in-bounds 6-neighbour of
 * Y[z,y,x] holds a different value (bit pattern), else 0.  out = uint8 / bool volume. */
int te_edge_map(const void* Y, int elem_size, const int64_t shape[3], uint8_t* out, void* stream);

/* cylindrical_mask (misc/voxel_processing.py:124-140), in place: vol[z,y,x] = value wherever
 * NOT ((y - n//2)^2 + (x - n//2)^2 < rad^2), rad = int(mask_fac * n / 2) computed by the host, n =
 * shape[1] = shape[2].  value_bits = the mask value already cast to the element type. */
int te_cylindrical_mask(void* vol, int elem_size, const int64_t shape[3], int64_t rad, uint32_t value_bits,
                        void* stream);

/* get_values_cyl_mask (misc/voxel_processing.py:142-157): vol[cyl > 0] in C order.  The inside set of
 * row y is the interval [x_lo[y], x_hi[y]); row_off[y] = number of inside voxels of rows < y of one
 * slice, per_slice = total per slice (device int32[n], int32[n], int64[n]; built by the host from
 * rad).  strides in elements (sub-sampled views are read in place).  out holds shape[0] * per_slice
 * elements. */
int te_cyl_values(const void* vol, int elem_size, const int64_t shape[3], const int64_t strides[3],
                  const int32_t* x_lo, const int32_t* x_hi, const int64_t* row_off, int64_t per_slice,
                  void* out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Convolutional networks.  One handle = one Keras model of the reference:
 *   kind "unet"    build_Unet_3D            (neural_nets/Unet3D.py:199-291)       -> segmenter / enhancer
 *   kind "encoder" build_encoder_r          (neural_nets/autoencoders.py:246-347) -> latent vectors
 *   kind "decoder" build_decoder_r          (neural_nets/autoencoders.py:350-447) -> denoised patches
 * Weights are supplied once as a flat fp32 blob in Keras get_weights() order and layout
 * (Conv3D (kz,ky,kx,Cin,Cout), Conv3DTranspose (kz,ky,kx,Cout,Cin), Dense (in,out), BN gamma, beta,
 * moving_mean, moving_variance); BN (eps 1e-5) is folded at load.
 */
typedef struct te_net te_net;

typedef struct te_net_desc {
  int32_t kind;            /* 0 unet, 1 encoder, 2 decoder */
  int32_t n_blocks;        /* <= 8 */
  int32_t n_filters[8];
  int32_t pool_size[8];
  int32_t isconcat[8];     /* unet only */
  int32_t batch_norm;      /* 0/1 */
  int32_t kern_size;       /* 3 */
  int32_t kern_size_upconv;/* 2 */
  int32_t n_hidden;        /* encoder / decoder: len(hidden_units) (<= 8) */
  int32_t hidden_units[8];
  int32_t pool_flag;       /* encoder / decoder POOL_FLAG */
  int32_t patch[3];        /* model input size (z,y,x); the U-net is fully convolutional but a handle is planned for one size */
  int32_t max_batch;       /* patches per forward call (workspace is sized for it) */
  float lrelu_alpha;       /* 0.2 */
  float bn_eps;            /* 1e-5 */
  int32_t mode;            /* TE_MODE_* */
} te_net_desc;

/* Number of fp32 values the weight blob must hold for this descriptor (or < 0 on error). */
int64_t te_net_weight_count(const te_net_desc* desc);
/* Creates the handle on the current device: plans layers, allocates the activation workspace. */
int te_net_create(const te_net_desc* desc, te_net** out);
/* weights: HOST pointer to te_net_weight_count() floats.  Synchronous. */
int te_net_set_weights(te_net* net, const float* weights_host, int64_t count);
int te_net_destroy(te_net* net);
/* Bytes of device memory held by the handle. */
int64_t te_net_workspace_bytes(const te_net* net);

/* U-net forward.  Replaces models["segmenter"].predict + np.round (surface_segmenter.py:283-291)
 * and, with out_dtype TE_F32, the enhancer / raw-probability output.
 *   x        (n, pz, py, px) activations already rescaled, dtype TE_BF16 / TE_F16 (tcgen05 modes) or
 *            TE_F32 (check mode): the output layout of te_gather_norm
 *   out      (n, pz, py, px): TE_U8 -> labels, 1 iff sigmoid(logit) > 0.5 (np.round is half-to-even,
 *            so exactly 0.5 -> 0); TE_F32 -> sigmoid probabilities
 *   logits   optional (may be NULL) fp32 (n, pz, py, px) pre-sigmoid values
 * n <= max_batch. */
int te_unet_forward(te_net* net, const void* x, int64_t n, void* out, int out_dtype, float* logits,
                    void* stream);
/* Encoder forward.  Replaces models["encoder"].predict (neural_nets/autoencoders.py:760-767).
 * z is (n, latent) fp32. */
int te_encoder_forward(te_net* net, const void* x, int64_t n, float* z, void* stream);
/* Decoder forward.  Replaces models["decoder"].predict (neural_nets/keras_processor.py:117-131).
 * z (n, latent) fp32 -> out (n,pz,py,px) fp32 probabilities. */
int te_decoder_forward(te_net* net, const float* z, int64_t n, float* out, void* stream);

/* Debug / test hook: copies the activation tensor produced by layer `layer_index` during the last
 * forward call, converted to fp32 NDHWC, into out (device).  Returns the element count. */
int64_t te_net_layer_output(te_net* net, int layer_index, float* out, int64_t capacity, void* stream);
int te_net_num_layers(const te_net* net);
/* Algorithmic FLOPs (2 x MACs, no padding / halo) of one forward pass over one patch. */
double te_net_flops_per_patch(const te_net* net);
/* Name, algorithmic FLOPs and kernel class of layer i (host strings owned by the handle). */
int te_net_layer_info(const te_net* net, int layer_index, const char** name, double* flops, int* uses_tensor_cores);

/* ------------------------------------------------------------------------------------------------
 * Training data generator of the segmenter.
 *
 * te_patch_std replaces the np.std over every binned label patch in SurfaceSegmenter._find_blanks
 * (neural_nets/surface_segmenter.py:75-82): out_std[i] = population standard deviation (ddof = 0, float64 sums) of
 * vol[z:z+wz:b, y:y+wy:b, x:x+wx:b], b = widths[i][0] / patch[0].  out_std = device float[n].
 *
 * te_training_pairs replaces extract_training_patch_pairs (neural_nets/keras_processor.py:213-232) in ONE kernel:
 *   x_out[i] = rot90(X[binned patch i], nrots[i], axes[i]) + sigma[i] * N(0, 1)       float32 (n, P, P, P)
 *   y_out[i] = rot90(Y[binned patch i], nrots[i], axes[i])                             uint8
 * rot90 = numpy.rot90(m, k, axes=(a, b)) with k = nrots[i] & 3, (a, b) = axes[i] (ordered pair out of 0, 1, 2; cubic
 * patches).  The noise is Philox4x32-10 keyed by (seed, patch, voxel) + Box-Muller: reproducible for a given seed,
 * independent of the launch geometry.  sigma == NULL: no noise; nrots == NULL: no rotation.  All pointers are device
 * pointers; dtypes of X / Y in {TE_U8, TE_U16, TE_F16, TE_F32}. */
int te_patch_std(const void* vol, int vol_dtype, const int64_t vol_shape[3], const uint32_t* points,
                 const uint32_t* widths, int64_t n, const int32_t patch[3], float* out_std, void* stream);
int te_training_pairs(const void* X, int x_dtype, const void* Y, int y_dtype, const int64_t vol_shape[3],
                      const uint32_t* points, const uint32_t* widths, int64_t n, const int32_t patch[3],
                      const float* sigma, const int32_t* nrots, const int32_t* axes, uint64_t seed,
                      float* x_out, uint8_t* y_out, void* stream);

/* fp16 activations (TE_MODE_FP16) have a finite range of +-65504.  Random-init and BN-folded shipped weights keep
 * activations O(1), but a trained net with large gamma / sigma ratios can exceed it; Keras computes in float32
 * (Unet3D.py:44-81) and has no such limit.  Every conv epilogue therefore tracks the largest |activation| it produced
 * (fp32, before rounding) and raises a sticky per-handle flag when one leaves the finite fp16 range.  *flag = 1 if any
 * forward call since the last reset overflowed (always 0 for bf16 / fp32 handles); the call synchronises `stream`.
 * The host mirror re-runs such a call in bf16 (neural_nets/keras_processor.py). */
int te_net_overflow(te_net* net, int reset, int* flag, void* stream);

/* Per-layer device timing (bench.py's roofline leg): when enabled, every layer launch of a forward
 * call is bracketed by CUDA events on the launching stream.  te_net_layer_time synchronises on the
 * recorded events and returns the accumulated milliseconds and launch count since enabling. */
int te_net_set_profiling(te_net* net, int enable);
int te_net_layer_time(te_net* net, int layer_index, double* total_ms, int64_t* launches);
/* Name of the kernel (template instantiation) that executes layer i, written into buf. */
int te_net_layer_kernel(const te_net* net, int layer_index, char* buf, int buflen);

/* ------------------------------------------------------------------------------------------------
 * Latent PCA.  Replaces sklearn IncrementalPCA.fit / transform as called by fit_PCA / transform_PCA
 * (scratchpad/IEEE_ICIP_paper/latent_vis.py:94-143) with exact PCA from streamed moments.
 */
/* Accumulates moments[0] += n, moments[1..d] += sum_i h_i, moments[1+d ..] += sum_i h_i h_i^T
 * (row-major d x d), all float64, h (n,d) fp32, d <= 256.  moments must be zeroed by the caller
 * before the first call; after the last call the ranks allreduce(SUM) it. */
int te_pca_accumulate(const float* h, int64_t n, int32_t d, double* moments, void* stream);
/* Solves the d x d symmetric eigenproblem of the covariance (cyclic Jacobi, float64, one CTA) and
 * writes mean (d), the top-k components (k,d) with sklearn's sign convention (largest-|v| entry of
 * each component positive) and their explained variances (k), all float64 device arrays.
 * work: device float64 scratch of 2*d*d values (d > 128: the matrices; d <= 128: everything lives in shared memory and
 * work[0] receives the number of Jacobi sweeps of the solve). */
int te_pca_solve(const double* moments, int32_t d, int32_t k, double* mean, double* components,
                 double* explained_variance, double* work, void* stream);
/* z[i] = (h[i] - mean) @ components^T ; z (n,k) fp32. */
int te_pca_project(const float* h, int64_t n, int32_t d, const double* mean, const double* components,
                   int32_t k, float* z, void* stream);


/* ------------------------------------------------------------------------------------------------
 * Training operators.  Replace the Keras fit step of
 * SurfaceSegmenter.train (neural_nets/surface_segmenter.py:43-66; compile with Adam + BinaryCrossentropy
 * :322-323; BatchNormalization(momentum 0.9, eps 1e-5) in training mode, Unet3D.py:78).  One entry per
 * operator; the host walks the graph (tomoencoders_b200/neural_nets/_train.py).  Tensors are NDHWC
 * channel slices described by te_tensor; dtype TE_F32 (check mode) or TE_BF16 / TE_F16; weights,
 * gradients and reductions are fp32 / fp64 device arrays in Keras layout.
 */
typedef struct te_tensor {
  void* ptr;
  int32_t dtype;          /* TE_F32 / TE_F16 / TE_BF16 */
  int32_t ctot, coff, C;  /* channel pitch of the buffer, first channel and channel count of this view */
  int32_t N, D, H, W;
} te_tensor;

/* y = Conv3D(3, 'same')(concat(x, x2)) + bias, no activation (BN follows).  w: [27][Cin][Cout].  16-bit
 * dense inputs with channel counts % 16 run on the tcgen05 kernel; wpack = device scratch of
 * te_tr_conv3_wpack_bytes(Cin, Cout) bytes (weights are re-packed on the device at every call). */
int64_t te_tr_conv3_wpack_bytes(int32_t Cin, int32_t Cout);
int te_tr_conv3(const te_tensor* x, const te_tensor* x2, const float* w, const float* bias, const te_tensor* y,
                void* wpack, void* stream);
/* Which kernel te_tr_conv3 / the first layer of a network takes for a ONE-channel input x and output y (first Conv3D of
 * every graph, Unet3D.py:117-122): 2 = Toeplitz tcgen05 kernel (conv1_tz.cu), 1 = im2col tcgen05 kernel (conv1_umma.cu),
 * 0 = CUDA cores.  Host-only query (tests and the per-layer roofline table name the kernel with it). */
int te_conv1_kernel_kind(const te_tensor* x, const te_tensor* y);
/* wf[27][Cout][ci_n]: spatially flipped, channel-transposed slice [ci_lo, ci_lo + ci_n) of w[27][Cin][Cout];
 * te_tr_conv3(dy, NULL, wf, zero_bias, dx, ...) is then the data gradient of the convolution. */
int te_tr_flip_weights(const float* w, int32_t Cin, int32_t Cout, int32_t ci_lo, int32_t ci_n, float* wf, void* stream);
/* dw[27][dw_cin][Cout] += sum_v x[v + tap][ci] dy[v][co] for input channels [ci_off, ci_off + x.C);
 * dbias2c (double [2 * Cout], optional): [0, Cout) += sum_v dy.  16-bit tensors: mma.sync tensor-core tiles. */
int te_tr_conv3_wgrad(const te_tensor* x, const te_tensor* dy, float* dw, int32_t dw_cin, int32_t ci_off,
                      double* dbias2c, void* stream);
/* stats2c (double [2C]) += per-channel sum and sum of squares of y (BN batch statistics). */
int te_tr_bn_stats(const te_tensor* y, double* stats2c, void* stream);
/* a = LeakyReLU(scale[c] * y + shift[c]) with scale = gamma * rstd, shift = beta - mean * scale. */
int te_tr_bn_act_fwd(const te_tensor* y, const te_tensor* a, const float* scale, const float* shift, float alpha, void* stream);
/* red2c (double [2C]) += [sum dz, sum dz * xhat], dz = da * LeakyReLU'(z). */
int te_tr_bn_act_bwd_reduce(const te_tensor* y, const te_tensor* da, const float* scale, const float* shift,
                            const float* mean, const float* rstd, float alpha, double* red2c, void* stream);
/* dy = scale * (dz - dbeta_n - xhat * dgamma_n), dbeta_n = sum dz / n, dgamma_n = sum dz xhat / n (dy may be da).
 * dsum2c (double [2C], optional): [0, C) += sum_v dy -- the bias gradient of the convolution that feeds this
 * BatchNormalization, taken from the values in flight instead of by another pass over dy. */
int te_tr_bn_act_bwd_apply(const te_tensor* y, const te_tensor* da, const te_tensor* dy, const float* scale,
                           const float* shift, const float* mean, const float* rstd, const float* dbeta_n,
                           const float* dgamma_n, float alpha, double* dsum2c, void* stream);
int te_tr_maxpool_fwd(const te_tensor* x, const te_tensor* out, int32_t pool, void* stream);
/* dx (+)= gradient routed to the first maximum of every window. */
int te_tr_maxpool_bwd(const te_tensor* x, const te_tensor* dpool, const te_tensor* dx, int32_t pool, int32_t accumulate,
                      void* stream);
/* a = LeakyReLU(Conv3DTranspose(k 2, strides s)(x) + bias); w: (2,2,2,Cout,Cin). */
int te_tr_upconv_fwd(const te_tensor* x, const float* w, const float* bias, const te_tensor* a, int32_t stride,
                     float alpha, void* wpack, void* stream);
/* dz = da * LeakyReLU'(a); dbias2c[0..C) += sum dz; dw += weight gradient; dx = data gradient (optional).
 * 16-bit tensors: dz is written "space to depth" into dz2 (N, D/2, H/2, W/2, 8 C) and both gradients are
 * pointwise products on the tensor cores (wpack as above); fp32: da is overwritten with dz, CUDA cores. */
int te_tr_upconv_bwd(const te_tensor* x, const te_tensor* a, const te_tensor* da, const float* w, int32_t stride,
                     float alpha, float* dw, double* dbias2c, const te_tensor* dx, const te_tensor* dz2, void* wpack,
                     void* stream);
/* 1x1x1 head + sigmoid + Keras binary cross-entropy (mean over n_total voxels), forward and backward. */
int te_tr_head_loss(const te_tensor* a, const float* w, const float* b, const uint8_t* y, const te_tensor* da,
                    double* loss, double* dwb2c, double n_total, float* dl, float* probs, void* stream);
/* Dense layers and loss of the autoencoder training step (RegularizedAutoencoder.train_step,
 * neural_nets/autoencoders.py:498-526; hidden_layer :60-91).  Dense activations are fp32 (N, n) row-major. */
int te_tr_dense_fwd(const float* x, int64_t n, int32_t nin, int32_t nout, const float* w, const float* bias, float* y,
                    void* stream);
int te_tr_dense_bwd(const float* x, const float* dy, int64_t n, int32_t nin, int32_t nout, const float* w, float* dw,
                    double* dbias2c, float* dx, void* stream);
/* 1x1x1 head + sigmoid + mean squared error + weight * KL(target || p); loss3 += {total, mse, kl}. */
int te_tr_head_ae_loss(const te_tensor* a, const float* w, const float* b, const void* target, const te_tensor* da,
                       double* loss3, double* dwb2c, double n_total, float weight, float* dl, float* probs, void* stream);
/* BatchNormalization bookkeeping between the reduction kernels and the element-wise kernels, one launch each.
 * te_tr_bn_finalize: stats2c = [sum y (C), sum y^2 (C)] over n_total voxels (all ranks) -> batch mean / biased variance,
 * rstd = rsqrt(var + eps), scale = gamma * rstd, shift = beta - mean * scale; stats2c is zeroed.
 * te_tr_bn_bwd_finalize: stats2c = [sum dz (C), sum dz * xhat (C)] (global); the parameter gradients g_beta / g_gamma
 * take the LOCAL sums (stats2c_local, NULL = same buffer), dbeta_n / dgamma_n the global means; stats2c is zeroed.
 * te_tr_take_stats: dst[i] = (float) stats[i] for i < n (dst may be NULL), then stats[0, n_zero) = 0. */
int te_tr_bn_finalize(double* stats2c, int32_t C, double n_total, const float* gamma, const float* beta, float eps,
                      float* mean, float* var, float* rstd, float* scale, float* shift, void* stream);
int te_tr_bn_bwd_finalize(double* stats2c, const double* stats2c_local, int32_t C, double n_total, float* g_beta,
                          float* g_gamma, float* dbeta_n, float* dgamma_n, void* stream);
int te_tr_take_stats(double* stats, int32_t n, int32_t n_zero, float* dst, void* stream);
/* Keras Adam on a flat parameter buffer; g is multiplied by grad_scale first (1 / world size after an all-reduce). */
int te_tr_adam(float* w, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
               int64_t step, float grad_scale, void* stream);

/* ---- post-stitch labelling: replaces cupyx.scipy.ndimage.label and
 * scipy.ndimage.find_objects as used by /root/reference/tomo_encoders/tasks/digital_zoom.py:41-43 and
 * structures/voids.py:188-303, 356-373.  All pointers are device pointers.
 *
 * te_label: connected components of the non-zero voxels of vol (u8 / u16 / f16 / f32), connectivity 6 (scipy's default
 * structure) or 26 (np.ones((3,3,3))); labels (int32, same shape) = 0 for background, 1..n numbered in C order of each
 * component's first voxel (scipy.ndimage.label's numbering); *n_labels (device int64) = n.  workspace: device scratch of
 * te_label_workspace_bytes(nvox) bytes.  Up to 2^31 - 1 voxels. */
int64_t te_label_workspace_bytes(int64_t nvox);
int te_label(const void* vol, int vol_dtype, const int64_t vol_shape[3], int connectivity, int32_t* labels,
             int64_t* n_labels, void* workspace, void* stream);
/* find_objects + sizes: bbox6[i] = {z0, y0, x0, z1, y1, x1} (stops exclusive) and sizes[i] = voxel count of label i + 1. */
int te_label_stats(const int32_t* labels, const int64_t vol_shape[3], int64_t n_labels, int32_t* bbox6, int64_t* sizes,
                   void* stream);
/* out[offsets[i] + local C-order index] = (labels[voxel of box i] == ids[i]) for n boxes (the x_voids of
 * Voids.count_voids); max_box_voxels = the largest box (sizes the launch). */
int te_label_crop_masks(const int32_t* labels, const int64_t vol_shape[3], const int32_t* boxes6, const int32_t* ids,
                        int64_t n, const int64_t* offsets, int64_t max_box_voxels, uint8_t* out, void* stream);
/* vol[box i] = mask i for i = 0..n-1 in order (numpy slice assignment semantics: later boxes overwrite earlier ones);
 * owner: device scratch of 4 bytes per voxel of vol. */
int te_paint_boxes(const uint8_t* masks, const int64_t* offsets, const int32_t* boxes6, int64_t n, int64_t max_box_voxels,
                   uint8_t* vol, const int64_t vol_shape[3], uint32_t* owner, void* stream);
/* flags[tile] = 1 iff any voxel of the width^3 tile of Grid(vol_shape, width) is non-zero (tiles in Grid.points order). */
int te_tile_any(const void* vol, int elem_size, const int64_t vol_shape[3], int32_t width, uint8_t* flags, void* stream);

/* ---- filtered back-projection: replaces the CuPy RawKernels and cuFFT calls of
 * /root/reference/tomo_encoders/reconstruction/cuda_kernels.py:21-317, prep.py:21-77 and the device side of
 * recon.py:135-381.  All pointers are device pointers; projections are float32 (ntheta, nz, n), C order.
 *
 * te_fbp_padding: prep.py:21-31 (calc_padding).  te_fbp_filter: prep.py:33-60 in place over `rows` = ntheta * nz
 * detector rows of n samples: edge-pad to n_pad, rfft, multiply by rfftfreq(n_pad), irfft, crop (cuFFT, batched so
 * that a batch stays in L2); workspace of te_fbp_workspace_bytes(rows, n) bytes (a smaller one only lowers the batch).
 * te_preprocess: prep.py:63-77, data = -log(max((data - dark) / max(flat - dark, 1e-6), 1e-6)), dark / flat (nz, n).
 * te_make_mask: recon.py:369-381, mask (float32, shape) = 1 inside the wd^3 boxes at points (n, 3) uint32, else 0. */
void te_fbp_padding(int32_t n, int32_t* pad_left, int32_t* pad_right);
int64_t te_fbp_workspace_bytes(int64_t rows, int32_t n);
int te_fbp_filter(float* data, int64_t rows, int32_t n, void* workspace, int64_t workspace_bytes, void* stream);
int te_fbp_release_plans(void);
int te_preprocess(float* data, const float* dark, const float* flat, int32_t ntheta, int32_t nz, int32_t n, void* stream);
int te_make_mask(float* mask, const int64_t shape[3], const uint32_t* points, int64_t n, int32_t wd, void* stream);
/* Back-projection.  te_bp_prepare re-lays the (filtered) projections out for the kernels below and tabulates
 * cos / sin(theta) (evaluated in fp64, rounded to fp32) into `workspace` (te_bp_workspace_bytes); the projections may
 * be overwritten afterwards.  Per voxel (x, y, z): sum over k of g[k][z][s0] + (g[k][z][s0+1] - g[k][z][s0]) *
 * (sp - s0) / n with sp = (x - n/2) cos - (y - n/2) sin + center, s0 = roundf(sp), 0 <= s0 < n - 1
 * (cuda_kernels.py:73-147).  exact = 1: one IEEE rounding per reference operation (parity mode); 0: FMA-contracted.
 *   te_bp_box      rec_patch / rec (cuda_kernels.py:73-98, 269-302): out (pz, py, px) = voxels [stz:stz+pz, sty:sty+py,
 *                  stx:stx+px]; with stx = sty = stz = 0, px = py = n, pz = nz it is rec_all (:125-147); masked = 1 is
 *                  rec_mask (:100-123): voxels with out > 0 are reconstructed, all others set to 0.
 *   te_bp_patches  the fused recon_patches_3d + extract_segmented path (recon.py:199-203, 256-300): out (npatch, wd, wd,
 *                  wd) float32 / float16 / bfloat16 = the wd^3 boxes at points (z, y, x) uint32; normalize = 1 applies
 *                  clip(v, lo, hi) then (v - lo) / (hi - lo) (recon.py:283-285) before the store.
 *   te_bp_points   rec_pts (:23-43; pts (npts, 3) int32 z, y, x; out (npts)) or, xy_mode = 1, rec_pts_xy (:47-68;
 *                  pts (npts, 2) y, x; out (nz, npts)). */
int64_t te_bp_workspace_bytes(int32_t ntheta, int32_t nz, int32_t n);
int te_bp_prepare(const float* data, const float* theta, int32_t ntheta, int32_t nz, int32_t n, void* workspace, void* stream);
int te_bp_box(const void* workspace, int32_t ntheta, int32_t nz, int32_t n, float center, int32_t stx, int32_t px, int32_t sty,
              int32_t py, int32_t stz, int32_t pz, float* out, int32_t masked, int32_t exact, void* stream);
int te_bp_patches(const void* workspace, int32_t ntheta, int32_t nz, int32_t n, float center, const uint32_t* points,
                  int64_t npatch, int32_t wd, void* out, int32_t out_dtype, int32_t normalize, float clip_lo, float clip_hi,
                  int32_t exact, void* stream);
int te_bp_points(const void* workspace, int32_t ntheta, int32_t nz, int32_t n, float center, const int32_t* pts, int32_t npts,
                 int32_t xy_mode, float* out, int32_t exact, void* stream);

/* ------------------------------------------------------------------------------------------------
 * One-shot all-reduce of small fp64 vectors over NVLink peer memory (csrc/peer_comm.cu): the SyncBN reductions of the
 * data-parallel training step (per-channel sums of every BatchNormalization layer, forward and
 * backward).  One process per GPU on one box; every rank: te_peer_create, te_peer_export -> exchange the 64-byte handles
 * (all-gather of bytes, e.g. torch.distributed) -> te_peer_import, then the SAME sequence of te_peer_allreduce_f64 calls
 * on one stream.  A reduction is one single-CTA kernel per rank (publish + epoch flag, wait for the peers' flags, sum in
 * rank order: bit-identical on every rank); graph-capturable.  count <= slot_doubles.
 */
typedef struct te_peer te_peer;
int te_peer_create(int32_t rank, int32_t world, int32_t slot_doubles, int32_t nslots, te_peer** out);
int te_peer_export(te_peer* p, void* handle64);
int te_peer_import(te_peer* p, const void* handles);
int te_peer_allreduce_f64(te_peer* p, double* data, int32_t count, void* stream);
int te_peer_destroy(te_peer* p);

#ifdef __cplusplus
}
#endif
#endif /* TOMOENC_H_ */

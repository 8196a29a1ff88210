// Network runtime: plans a Keras graph of the reference as a list of fused layers, owns the weights
// (BN folded, packed for the tensor-core kernels) and the activation workspace, and runs the forward
// pass on a stream.  Graphs restated (paths relative to /root/reference/tomo_encoders/):
//   U-net    build_Unet_3D     neural_nets/Unet3D.py:199-291 (analysis_block :84-132, synthesis_block :134-194)
//   encoder  build_encoder_r   neural_nets/autoencoders.py:246-347 (analysis_block_small :146-189, hidden_layer :60-91)
//   decoder  build_decoder_r   neural_nets/autoencoders.py:350-447 (synthesis_block_small :192-243)
// Weight blob order = Keras get_weights() order = layer creation order: kernel, bias,
// [gamma, beta, moving_mean, moving_variance].
//
// Data layout in HBM: activations are NDHWC, 16-bit (bf16 / fp16) in the tensor-core modes and fp32
// in the check mode.  The skip concatenation is never materialised by a copy: the transposed conv
// and the encoder conv that produce the two halves write disjoint channel slices of one buffer.
#include "conv_umma.cuh"
#include "layers_simt.cuh"
#include "te_common.cuh"

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace te {
namespace {

enum Op { OP_CONV3 = 0, OP_UPCONV = 1, OP_POOL = 2, OP_HEAD = 3, OP_DENSE = 4 };

struct Layer {
  Op op;
  std::string name;
  TView in{}, out{};
  TView in2{};               // second input (planar skip concatenation): channels [in.C, Cin)
  bool has_in2 = false;
  int Cin = 0, Cout = 0;
  int stride = 0;            // pool size / transposed-conv stride
  bool bn = false, lrelu = false;
  // dense
  int nin = 0, nout = 0;
  const void* dense_in = nullptr;
  void* dense_out = nullptr;
  int dense_in_type = ACT_F32, dense_out_type = ACT_F32;
  // blob offsets (floats)
  long long off_k = -1, off_b = -1, off_bn = -1;
  // device weights
  float* d_w = nullptr;      // CUDA-core layout
  float* d_bias = nullptr;
  void* d_wpacked = nullptr; // tensor-core layout
  void* d_wpair = nullptr;   // CTA-pair image of the same weights (cfg.pair)
  bool use_umma = false;
  bool fused_head = false;   // this conv also evaluates the following 1x1x1 head
  bool fused_pool = false;   // this conv also writes the MaxPool3D(2) of its output
  bool store_out = true;     // false: only the pooled tensor is consumed downstream
  TView pool_view{};
  bool skip = false;         // layer folded into its predecessor (the fused head)
  ConvUmmaConfig cfg{};
  CUtensorMap tmap{}, tmap2{};
  CUtensorMap omaps[8]{};    // output tensor maps of the un-haloed kernels (TMA-store epilogue), built per batch size
  int omaps_n = -1, omaps_ok = 0;
  double flops = 0.0;
  float head_b = 0.f;
  // optional per-layer timing (te_net_set_profiling): event pairs recorded around every launch
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> events;
  double prof_ms = 0.0;
  long long prof_launches = 0;
};

}  // namespace
}  // namespace te

struct te_net {
  te_net_desc desc;
  int act_type = te::ACT_BF16;      // element type of activations
  int elem = 2;
  std::vector<te::Layer> layers;
  std::vector<void*> allocations;
  long long weight_count = 0;
  long long workspace_bytes = 0;
  bool weights_set = false;
  bool fuse_head = true;
  int latent = 0;
  int last_n = 0;
  bool profiling = false;
  float* d_dense_tmp[2] = {nullptr, nullptr};
  float* d_dense_acc = nullptr;     // split-K accumulation scratch of the dense layers
  int* d_overflow = nullptr;        // fp16 mode: sticky flag set by the conv epilogues when an activation overflows fp16
  te::TView input{};                // network input view (C = 1)
};

namespace te {
namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    cudaDriverEntryPointQueryResult q;
    void* p = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess) fn = (EncodeTiledFn)p;
  }
  return fn;
}

int make_tmap(CUtensorMap* tm, const TView& v, const ConvUmmaConfig& cfg, int halo, bool fp16) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled is not available"); return TE_ERR_CUDA; }
  cuuint64_t gdim[5] = {(cuuint64_t)v.ctot, (cuuint64_t)v.W, (cuuint64_t)v.H, (cuuint64_t)v.D, (cuuint64_t)v.N};
  cuuint64_t gstr[4] = {(cuuint64_t)v.ctot * 2, (cuuint64_t)v.W * v.ctot * 2, (cuuint64_t)v.H * v.W * v.ctot * 2,
                        (cuuint64_t)v.D * v.H * v.W * v.ctot * 2};
  cuuint32_t box[5] = {(cuuint32_t)cfg.CB, (cuuint32_t)(8 + 2 * halo), (cuuint32_t)(16 + 2 * halo),
                       (cuuint32_t)(cfg.TZ + 2 * halo), 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  const CUtensorMapSwizzle sw = cfg.CB == 16 ? CU_TENSOR_MAP_SWIZZLE_32B
                                : (cfg.CB == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B);
  CUresult r = enc(tm, fp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, v.ptr, gdim, gstr,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return TE_ERR_CUDA; }
  return TE_OK;
}

struct Planner {
  te_net* net;
  long long cursor = 0;      // running offset into the weight blob
  bool count_only;
  long long ws_bytes = 0;

  TView alloc_tensor(int C, int D, int H, int W) {
    TView v;
    v.ctot = C; v.coff = 0; v.C = C;
    v.N = net->desc.max_batch; v.D = D; v.H = H; v.W = W;
    v.ptr = nullptr;
    const size_t bytes = static_cast<size_t>(v.voxels()) * C * net->elem;
    ws_bytes += bytes;
    if (!count_only) {
      void* p = nullptr;
      if (cudaMalloc(&p, bytes + 256) != cudaSuccess) { failed = true; return v; }
      net->allocations.push_back(p);
      v.ptr = p;
    }
    return v;
  }
  static TView slice(const TView& buf, int coff, int C) {
    TView v = buf;
    v.coff = coff;
    v.C = C;
    return v;
  }
  bool failed = false;

  Layer& add(Op op, const char* name) {
    net->layers.emplace_back();
    Layer& L = net->layers.back();
    L.op = op;
    L.name = name;
    return L;
  }

  void conv3(const std::string& name, const TView& in, const TView& out, bool bn, const TView* in2 = nullptr) {
    Layer& L = add(OP_CONV3, name.c_str());
    L.in = in; L.out = out; L.Cin = in.C; L.Cout = out.C; L.bn = bn; L.lrelu = true;
    if (in2) { L.in2 = *in2; L.has_in2 = true; L.Cin += in2->C; }
    L.off_k = cursor; cursor += 27LL * L.Cin * L.Cout;
    L.off_b = cursor; cursor += L.Cout;
    if (bn) { L.off_bn = cursor; cursor += 4LL * L.Cout; }
    L.flops = 2.0 * 27.0 * L.Cin * L.Cout * in.D * in.H * in.W;
  }
  void upconv(const std::string& name, const TView& in, const TView& out, int stride) {
    Layer& L = add(OP_UPCONV, name.c_str());
    L.in = in; L.out = out; L.Cin = in.C; L.Cout = out.C; L.stride = stride; L.lrelu = true;
    L.off_k = cursor; cursor += 8LL * L.Cin * L.Cout;
    L.off_b = cursor; cursor += L.Cout;
    L.flops = 2.0 * 8.0 * L.Cin * L.Cout * in.D * in.H * in.W;
  }
  void pool(const std::string& name, const TView& in, const TView& out, int p) {
    Layer& L = add(OP_POOL, name.c_str());
    L.in = in; L.out = out; L.Cin = in.C; L.Cout = out.C; L.stride = p;
  }
  void head(const std::string& name, const TView& in) {
    Layer& L = add(OP_HEAD, name.c_str());
    L.in = in; L.Cin = in.C; L.Cout = 1;
    L.off_k = cursor; cursor += L.Cin;
    L.off_b = cursor; cursor += 1;
    L.flops = 2.0 * L.Cin * in.D * in.H * in.W;
  }
  void dense(const std::string& name, int nin, int nout, bool bn) {
    Layer& L = add(OP_DENSE, name.c_str());
    L.nin = nin; L.nout = nout; L.bn = bn; L.lrelu = true;
    L.off_k = cursor; cursor += static_cast<long long>(nin) * nout;
    L.off_b = cursor; cursor += nout;
    if (bn) { L.off_bn = cursor; cursor += 4LL * nout; }
    L.flops = 2.0 * nin * nout;
  }
};

int plan(te_net* net, bool count_only, long long* weight_count) {
  const te_net_desc& d = net->desc;
  TE_CHECK_ARG(d.kind >= 0 && d.kind <= 2, "te_net: kind must be 0 (unet), 1 (encoder) or 2 (decoder)");
  TE_CHECK_ARG(d.n_blocks >= 1 && d.n_blocks <= 8, "te_net: n_blocks must be in [1, 8]");
  TE_CHECK_ARG(d.max_batch >= 1, "te_net: max_batch must be >= 1");
  TE_CHECK_ARG(d.mode >= 0 && d.mode <= 2, "te_net: unknown mode");
  if (d.kern_size != 3 || d.kern_size_upconv != 2) {
    set_error("te_net: only kern_size = 3 and kern_size_upconv = 2 are implemented (the values of every shipped model)");
    return TE_ERR_UNSUPPORTED;
  }
  for (int i = 0; i < d.n_blocks; ++i) {
    TE_CHECK_ARG(d.n_filters[i] >= 1, "te_net: n_filters must be positive");
    TE_CHECK_ARG(d.pool_size[i] >= 1, "te_net: pool_size must be >= 1");
  }
  net->act_type = d.mode == TE_MODE_FP32_CHECK ? ACT_F32 : (d.mode == TE_MODE_FP16 ? ACT_F16 : ACT_BF16);
  net->elem = d.mode == TE_MODE_FP32_CHECK ? 4 : 2;
  net->layers.clear();
  Planner P{net};
  P.count_only = count_only;
  int D = d.patch[0], H = d.patch[1], W = d.patch[2];
  TE_CHECK_ARG(D >= 1 && H >= 1 && W >= 1, "te_net: patch extents must be positive");
  const bool bn = d.batch_norm != 0;
  char nm[64];

  if (d.kind == 0) {
    // ------------------------------------------------------------------ U-net
    const int nb = d.n_blocks;
    std::vector<TView> skips(nb);
    std::vector<TView> concat_bufs(nb);
    std::vector<int> c_up(nb);
    for (int i = 0; i < nb; ++i) c_up[i] = (i == nb - 1) ? 4 * d.n_filters[nb - 1] : 2 * d.n_filters[i + 1];
    // Skip concatenation.  Tensor-core modes: "planar" -- the up-sampled and the skip tensor stay two
    // dense tensors and the consuming conv reads its channel chunks through two TMA descriptors
    // (every voxel row is then a whole number of 128-byte lines for both producers).  Otherwise:
    // both producers write channel slices of one buffer.  Either way no concat copy exists.
    std::vector<char> planar(nb, 0);
    if (net->act_type != ACT_F32) {
      for (int i = 0; i < nb; ++i) {
        ConvUmmaConfig cfg;
        const int f2 = 2 * d.n_filters[i];
        planar[i] = d.pool_size[i] > 1 && d.isconcat[i] && conv_umma_config(c_up[i] + f2, f2, &cfg) &&
                    c_up[i] % cfg.CB == 0 && f2 % cfg.CB == 0;
      }
    }
    TView cur = P.alloc_tensor(1, D, H, W);
    net->input = cur;
    for (int i = 0; i < nb; ++i) {
      const int f = d.n_filters[i], p = d.pool_size[i];
      TE_CHECK_ARG(D % p == 0 && H % p == 0 && W % p == 0, "te_net: pool size must divide the spatial extent");
      TView t1 = P.alloc_tensor(f, D, H, W);
      snprintf(nm, sizeof nm, "enc%d.conv1", i); P.conv3(nm, cur, t1, bn);
      TView t2;
      if (p > 1 && d.isconcat[i] && !planar[i]) {
        concat_bufs[i] = P.alloc_tensor(c_up[i] + 2 * f, D, H, W);
        t2 = Planner::slice(concat_bufs[i], c_up[i], 2 * f);          // up first, skip second (Unet3D.py:180)
      } else {
        t2 = P.alloc_tensor(2 * f, D, H, W);
      }
      snprintf(nm, sizeof nm, "enc%d.conv2", i); P.conv3(nm, t1, t2, bn);
      skips[i] = t2;
      cur = t2;
      if (p > 1) {
        D /= p; H /= p; W /= p;
        TView t3 = P.alloc_tensor(2 * f, D, H, W);
        snprintf(nm, sizeof nm, "enc%d.pool", i); P.pool(nm, t2, t3, p);
        cur = t3;
      }
    }
    const int nf = cur.C;
    TView b1 = P.alloc_tensor(nf, D, H, W);
    P.conv3("bottleneck.conv1", cur, b1, bn);
    TView b2 = P.alloc_tensor(2 * nf, D, H, W);
    P.conv3("bottleneck.conv2", b1, b2, bn);
    cur = b2;
    for (int i = nb - 1; i >= 0; --i) {
      const int f = d.n_filters[i], p = d.pool_size[i];
      if (p > 1) {
        D *= p; H *= p; W *= p;
        TView up;
        if (d.isconcat[i] && !planar[i]) up = Planner::slice(concat_bufs[i], 0, cur.C);
        else up = P.alloc_tensor(cur.C, D, H, W);
        if (d.isconcat[i] && cur.C != c_up[i]) { set_error("te_net: internal channel bookkeeping error"); return TE_ERR_STATE; }
        snprintf(nm, sizeof nm, "dec%d.upconv", i); P.upconv(nm, cur, up, p);
        cur = (d.isconcat[i] && !planar[i]) ? concat_bufs[i] : up;
      }
      TView t1 = P.alloc_tensor(2 * f, D, H, W);
      snprintf(nm, sizeof nm, "dec%d.conv1", i);
      if (p > 1 && planar[i]) P.conv3(nm, cur, t1, bn, &skips[i]);
      else P.conv3(nm, cur, t1, bn);
      TView t2 = P.alloc_tensor(2 * f, D, H, W);
      snprintf(nm, sizeof nm, "dec%d.conv2", i); P.conv3(nm, t1, t2, bn);
      cur = t2;
    }
    P.head("head", cur);
  } else if (d.kind == 1) {
    // ------------------------------------------------------------------ encoder
    TE_CHECK_ARG(d.n_hidden >= 1 && d.n_hidden <= 8, "te_net: n_hidden must be in [1, 8]");
    TView cur = P.alloc_tensor(1, D, H, W);
    net->input = cur;
    for (int i = 0; i < d.n_blocks; ++i) {
      const int f = d.n_filters[i], p = d.pool_size[i];
      TE_CHECK_ARG(D % p == 0 && H % p == 0 && W % p == 0, "te_net: pool size must divide the spatial extent");
      TView t1 = P.alloc_tensor(f, D, H, W);
      snprintf(nm, sizeof nm, "enc%d.conv1", i); P.conv3(nm, cur, t1, bn);
      TView t2 = P.alloc_tensor(2 * f, D, H, W);
      snprintf(nm, sizeof nm, "enc%d.conv2", i); P.conv3(nm, t1, t2, bn);
      D /= p; H /= p; W /= p;
      TView t3 = P.alloc_tensor(2 * f, D, H, W);
      snprintf(nm, sizeof nm, "enc%d.pool", i); P.pool(nm, t2, t3, p);
      cur = t3;
    }
    if (d.pool_flag) {
      TE_CHECK_ARG(D % 2 == 0 && H % 2 == 0 && W % 2 == 0, "te_net: extra pool must divide the spatial extent");
      D /= 2; H /= 2; W /= 2;
      TView t = P.alloc_tensor(cur.C, D, H, W);
      P.pool("extra.pool", cur, t, 2);
      cur = t;
    }
    int nin = cur.C * D * H * W;
    const void* src = cur.ptr;
    int src_type = net->act_type;
    for (int i = 0; i < d.n_hidden; ++i) {
      const bool last = i == d.n_hidden - 1;
      snprintf(nm, sizeof nm, last ? "latent.dense" : "hidden%d.dense", i);
      P.dense(nm, nin, d.hidden_units[i], last ? true : bn);
      Layer& L = net->layers.back();
      L.dense_in = src; L.dense_in_type = src_type; L.dense_out_type = ACT_F32;
      nin = d.hidden_units[i];
      src = nullptr; src_type = ACT_F32;
    }
    net->latent = d.hidden_units[d.n_hidden - 1];
  } else {
    // ------------------------------------------------------------------ decoder
    TE_CHECK_ARG(d.n_hidden >= 1 && d.n_hidden <= 8, "te_net: n_hidden must be in [1, 8]");
    // Conv3DTranspose(k = 2, strides = 1, 'same') is a real overlapping 2x2x2 convolution (autoencoders.py:213); the
    // transposed-conv kernels here assume non-overlapping taps (k <= stride), so a stride of 1 is refused, not mis-computed
    for (int i = 0; i < d.n_blocks; ++i)
      if (d.pool_size[i] < 2) {
        set_error("te_net: decoder blocks need pool_size >= 2 (a stride-1 transposed conv with kernel 2 is not implemented)");
        return TE_ERR_UNSUPPORTED;
      }
    // pre-flatten shape of the matching encoder
    int pd = D, ph = H, pw = W;
    for (int i = 0; i < d.n_blocks; ++i) {
      const int p = d.pool_size[i];
      TE_CHECK_ARG(pd % p == 0 && ph % p == 0 && pw % p == 0, "te_net: pool size must divide the spatial extent");
      pd /= p; ph /= p; pw /= p;
    }
    if (d.pool_flag) { pd /= 2; ph /= 2; pw /= 2; }
    const int pc = 2 * d.n_filters[d.n_blocks - 1];
    const int flat = pd * ph * pw * pc;
    net->latent = d.hidden_units[d.n_hidden - 1];
    int nin = net->latent;
    for (int i = d.n_hidden - 2; i >= 0; --i) {
      snprintf(nm, sizeof nm, "hidden%d.dense", i);
      P.dense(nm, nin, d.hidden_units[i], bn);
      nin = d.hidden_units[i];
    }
    TView pre = P.alloc_tensor(pc, pd, ph, pw);
    P.dense("reshape.dense", nin, flat, bn);
    {
      Layer& L = net->layers.back();
      L.dense_out = pre.ptr; L.dense_out_type = net->act_type;
    }
    TView cur = pre;
    int cd = pd, ch = ph, cw = pw;
    if (d.pool_flag) {
      cd *= 2; ch *= 2; cw *= 2;
      TView up = P.alloc_tensor(cur.C, cd, ch, cw);
      P.upconv("extra.upconv", cur, up, 2);
      cur = up;
    }
    for (int i = d.n_blocks - 1; i >= 0; --i) {
      const int f = d.n_filters[i], p = d.pool_size[i];
      cd *= p; ch *= p; cw *= p;
      TView up = P.alloc_tensor(cur.C, cd, ch, cw);
      snprintf(nm, sizeof nm, "dec%d.upconv", i); P.upconv(nm, cur, up, p);
      TView t1 = P.alloc_tensor(f, cd, ch, cw);
      snprintf(nm, sizeof nm, "dec%d.conv1", i); P.conv3(nm, up, t1, bn);
      TView t2 = P.alloc_tensor(f, cd, ch, cw);
      snprintf(nm, sizeof nm, "dec%d.conv2", i); P.conv3(nm, t1, t2, bn);
      cur = t2;
    }
    P.head("head", cur);
  }
  if (P.failed) { set_error("te_net_create: out of device memory for the activation workspace"); return TE_ERR_CUDA; }
  *weight_count = P.cursor;
  net->workspace_bytes = P.ws_bytes;
  return TE_OK;
}

void free_net(te_net* net) {
  for (void* p : net->allocations) cudaFree(p);
  for (auto& L : net->layers) {
    for (auto& ev : L.events) { cudaEventDestroy(ev.first); cudaEventDestroy(ev.second); }
    if (L.d_w) cudaFree(L.d_w);
    if (L.d_bias) cudaFree(L.d_bias);
    if (L.d_wpacked) cudaFree(L.d_wpacked);
    if (L.d_wpair) cudaFree(L.d_wpair);
  }
  for (int i = 0; i < 2; ++i) if (net->d_dense_tmp[i]) cudaFree(net->d_dense_tmp[i]);
  if (net->d_dense_acc) cudaFree(net->d_dense_acc);
  if (net->d_overflow) cudaFree(net->d_overflow);
  delete net;
}

// BN folding: y = gamma (x - mean) / sqrt(var + eps) + beta applied to x = W a + b.
void bn_scale_shift(const float* bnp, int C, float eps, std::vector<double>& scale, std::vector<double>& shift) {
  scale.assign(C, 1.0);
  shift.assign(C, 0.0);
  if (!bnp) return;
  const float *g = bnp, *b = bnp + C, *m = bnp + 2 * C, *v = bnp + 3 * C;
  for (int c = 0; c < C; ++c) {
    scale[c] = static_cast<double>(g[c]) / std::sqrt(static_cast<double>(v[c]) + eps);
    shift[c] = static_cast<double>(b[c]) - static_cast<double>(m[c]) * scale[c];
  }
}

int upload(const void* host, size_t bytes, void** dev) {
  TE_CUDA(cudaMalloc(dev, bytes));
  TE_CUDA(cudaMemcpy(*dev, host, bytes, cudaMemcpyHostToDevice));
  return TE_OK;
}

}  // namespace
}  // namespace te

using namespace te;

extern "C" int64_t te_net_weight_count(const te_net_desc* desc) {
  if (!desc) { set_error("te_net_weight_count: null descriptor"); return TE_ERR_ARG; }
  te_net tmp;
  tmp.desc = *desc;
  long long count = 0;
  const int rc = plan(&tmp, true, &count);
  return rc == TE_OK ? count : rc;
}

extern "C" int te_net_create(const te_net_desc* desc, te_net** out) {
  TE_CHECK_ARG(desc && out, "te_net_create: null pointer");
  te_net* net = new te_net;
  net->desc = *desc;
  const char* nf = getenv("TOMOENC_NO_FUSE_HEAD");
  net->fuse_head = !(nf && nf[0] == '1');
  long long count = 0;
  int rc = plan(net, false, &count);
  if (rc != TE_OK) { free_net(net); return rc; }
  net->weight_count = count;
  const bool tc = net->act_type != ACT_F32;
  // decide the kernel of every layer, build TMA descriptors
  for (size_t i = 0; i < net->layers.size(); ++i) {
    Layer& L = net->layers[i];
    if (!tc) continue;
    if (L.op == OP_CONV3 || L.op == OP_UPCONV) {
      // stride > 2 (k = 2 < stride leaves bias-only gaps) stays on the CUDA-core kernel
      const bool k2s = L.op == OP_CONV3 || L.stride == 2;
      ConvUmmaConfig cfg;
      if (k2s && L.in.coff == 0 && L.in.C == L.in.ctot && conv_umma_config(L.Cin, L.Cout, &cfg)) {
        L.use_umma = true;
        L.cfg = cfg;
        rc = make_tmap(&L.tmap, L.in, cfg, L.op == OP_CONV3 ? 1 : 0, net->act_type == ACT_F16);
        if (rc == TE_OK && L.has_in2) rc = make_tmap(&L.tmap2, L.in2, cfg, 1, net->act_type == ACT_F16);
        if (rc != TE_OK) { free_net(net); return rc; }
      } else if (L.has_in2) {
        set_error("te_net_create: internal error: planar concat planned for a layer without a tensor-core kernel");
        free_net(net);
        return TE_ERR_STATE;
      }
    }
  }
  // fuse the 1x1x1 head into the last conv when that conv runs on the tensor cores with one N tile
  if (tc && net->fuse_head && net->layers.size() >= 2) {
    Layer& H = net->layers.back();
    Layer& Cv = net->layers[net->layers.size() - 2];
    if (H.op == OP_HEAD && Cv.op == OP_CONV3 && Cv.use_umma && Cv.cfg.NT == Cv.Cout) {
      Cv.fused_head = true;
      H.skip = true;
    }
  }
  // fuse MaxPool3D(2) into the epilogue of the tensor-core conv that feeds it; in the encoder graph
  // the un-pooled tensor has no other consumer, so it is not written at all
  const char* nfp = getenv("TOMOENC_NO_FUSE_POOL");
  if (tc && !(nfp && nfp[0] == '1')) {
    for (size_t i = 0; i + 1 < net->layers.size(); ++i) {
      Layer& Cv = net->layers[i];
      Layer& Pl = net->layers[i + 1];
      if (Cv.op == OP_CONV3 && Cv.use_umma && Pl.op == OP_POOL && Pl.stride == 2 && Cv.in.D % 2 == 0 &&
          Cv.in.H % 2 == 0 && Cv.in.W % 2 == 0 && Pl.in.ptr == Cv.out.ptr && Pl.in.coff == Cv.out.coff) {
        Cv.fused_pool = true;
        Cv.pool_view = Pl.out;
        Cv.store_out = desc->kind == 0;      // U-net: the skip connection still needs the full tensor
        Pl.skip = true;
      }
    }
  }
  if (net->act_type == ACT_F16) {
    if (cudaMalloc(&net->d_overflow, sizeof(int)) != cudaSuccess || cudaMemset(net->d_overflow, 0, sizeof(int)) != cudaSuccess) {
      set_error("te_net_create: out of device memory");
      free_net(net);
      return TE_ERR_CUDA;
    }
  }
  if (desc->kind != 0) {
    int maxh = 1;
    for (auto& L : net->layers) if (L.op == OP_DENSE) maxh = L.nout > maxh ? L.nout : maxh;
    for (int i = 0; i < 2; ++i) {
      if (cudaMalloc(&net->d_dense_tmp[i], static_cast<size_t>(desc->max_batch) * maxh * sizeof(float)) != cudaSuccess) {
        set_error("te_net_create: out of device memory");
        free_net(net);
        return TE_ERR_CUDA;
      }
      net->workspace_bytes += static_cast<long long>(desc->max_batch) * maxh * 4;
    }
    if (cudaMalloc(&net->d_dense_acc, static_cast<size_t>(desc->max_batch) * maxh * sizeof(float)) != cudaSuccess) {
      set_error("te_net_create: out of device memory");
      free_net(net);
      return TE_ERR_CUDA;
    }
  }
  *out = net;
  return TE_OK;
}

extern "C" int te_net_destroy(te_net* net) {
  if (net) free_net(net);
  return TE_OK;
}

extern "C" int64_t te_net_workspace_bytes(const te_net* net) { return net ? net->workspace_bytes : 0; }
extern "C" int te_net_num_layers(const te_net* net) { return net ? static_cast<int>(net->layers.size()) : 0; }

extern "C" double te_net_flops_per_patch(const te_net* net) {
  double f = 0.0;
  if (net) for (auto& L : net->layers) f += L.flops;
  return f;
}

extern "C" int te_net_layer_info(const te_net* net, int i, const char** name, double* flops, int* uses_tc) {
  TE_CHECK_ARG(net && i >= 0 && i < static_cast<int>(net->layers.size()), "te_net_layer_info: bad index");
  const Layer& L = net->layers[i];
  if (name) *name = L.name.c_str();
  if (flops) *flops = L.flops;
  if (uses_tc) *uses_tc = L.use_umma ? 1 : 0;
  return TE_OK;
}

extern "C" int te_net_set_weights(te_net* net, const float* w, int64_t count) {
  TE_CHECK_ARG(net && w, "te_net_set_weights: null pointer");
  TE_CHECK_ARG(count == net->weight_count, "te_net_set_weights: expected %lld floats, got %lld",
               (long long)net->weight_count, (long long)count);
  const float eps = net->desc.bn_eps;
  const int fp16 = net->act_type == ACT_F16;
  for (auto& L : net->layers) {
    if (L.d_w) { cudaFree(L.d_w); L.d_w = nullptr; }
    if (L.d_bias) { cudaFree(L.d_bias); L.d_bias = nullptr; }
    if (L.d_wpacked) { cudaFree(L.d_wpacked); L.d_wpacked = nullptr; }
    if (L.d_wpair) { cudaFree(L.d_wpair); L.d_wpair = nullptr; }
    std::vector<double> scale, shift;
    if (L.op == OP_CONV3) {
      const int Ci = L.Cin, Co = L.Cout;
      bn_scale_shift(L.bn ? w + L.off_bn : nullptr, Co, eps, scale, shift);
      std::vector<float> wd(27ull * Ci * Co), bias(Co);
      const float* k = w + L.off_k;                       // (kz,ky,kx,Cin,Cout)
      for (size_t e = 0; e < wd.size(); ++e) wd[e] = static_cast<float>(k[e] * scale[e % Co]);
      for (int c = 0; c < Co; ++c) bias[c] = static_cast<float>(w[L.off_b + c] * scale[c] + shift[c]);
      int rc = upload(bias.data(), bias.size() * 4, reinterpret_cast<void**>(&L.d_bias));
      if (rc != TE_OK) return rc;
      if (L.use_umma) {
        std::vector<float> canon(27ull * Ci * Co);        // [tap][cout][cin]
        for (int t = 0; t < 27; ++t)
          for (int ci = 0; ci < Ci; ++ci)
            for (int co = 0; co < Co; ++co)
              canon[(static_cast<size_t>(t) * Co + co) * Ci + ci] = wd[(static_cast<size_t>(t) * Ci + ci) * Co + co];
        std::vector<uint8_t> packed(conv_umma_packed_bytes(Ci, Co, 27, 1, L.cfg));
        conv_umma_pack(canon.data(), Ci, Co, 27, 1, L.cfg, fp16, packed.data());
        rc = upload(packed.data(), packed.size(), &L.d_wpacked);
        if (rc == TE_OK && L.cfg.pair) {
          std::vector<uint8_t> pp(conv_umma_pair_packed_bytes(Ci, Co, L.cfg));
          conv_umma_pack_pair(canon.data(), Ci, Co, L.cfg, fp16, pp.data());
          rc = upload(pp.data(), pp.size(), &L.d_wpair);
        }
      } else {
        rc = upload(wd.data(), wd.size() * 4, reinterpret_cast<void**>(&L.d_w));
      }
      if (rc != TE_OK) return rc;
    } else if (L.op == OP_UPCONV) {
      const int Ci = L.Cin, Co = L.Cout;
      const float* k = w + L.off_k;                       // (kz,ky,kx,Cout,Cin) == [sub][cout][cin]
      int rc = upload(w + L.off_b, Co * 4ull, reinterpret_cast<void**>(&L.d_bias));
      if (rc != TE_OK) return rc;
      if (L.use_umma) {
        std::vector<uint8_t> packed(conv_umma_packed_bytes(Ci, Co, 1, 8, L.cfg));
        conv_umma_pack(k, Ci, Co, 1, 8, L.cfg, fp16, packed.data());
        rc = upload(packed.data(), packed.size(), &L.d_wpacked);
      } else {
        rc = upload(k, 8ull * Ci * Co * 4, reinterpret_cast<void**>(&L.d_w));
      }
      if (rc != TE_OK) return rc;
    } else if (L.op == OP_HEAD) {
      int rc = upload(w + L.off_k, L.Cin * 4ull, reinterpret_cast<void**>(&L.d_w));
      if (rc != TE_OK) return rc;
      L.head_b = w[L.off_b];
    } else if (L.op == OP_DENSE) {
      bn_scale_shift(L.bn ? w + L.off_bn : nullptr, L.nout, eps, scale, shift);
      std::vector<float> wd(static_cast<size_t>(L.nin) * L.nout), bias(L.nout);
      const float* k = w + L.off_k;                       // (in, out)
      for (size_t e = 0; e < wd.size(); ++e) wd[e] = static_cast<float>(k[e] * scale[e % L.nout]);
      for (int c = 0; c < L.nout; ++c) bias[c] = static_cast<float>(w[L.off_b + c] * scale[c] + shift[c]);
      int rc = upload(wd.data(), wd.size() * 4, reinterpret_cast<void**>(&L.d_w));
      if (rc != TE_OK) return rc;
      rc = upload(bias.data(), bias.size() * 4, reinterpret_cast<void**>(&L.d_bias));
      if (rc != TE_OK) return rc;
    }
  }
  net->weights_set = true;
  return TE_OK;
}

namespace {

struct HeadOut { uint8_t* labels; float* probs; float* logits; };

TView with_n(TView v, int n) { v.N = n; return v; }

int run_layers(te_net* net, int n, const HeadOut& ho, cudaStream_t st) {
  const float alpha = net->desc.lrelu_alpha;
  const void* dense_src = nullptr;
  int dense_src_type = ACT_F32;
  int tmp_idx = 0;
  for (size_t i = 0; i < net->layers.size(); ++i) {
    Layer& L = net->layers[i];
    if (L.skip) continue;
    cudaError_t e = cudaSuccess;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    if (net->profiling) {
      cudaEventCreate(&ev0);
      cudaEventCreate(&ev1);
      cudaEventRecord(ev0, st);
    }
    switch (L.op) {
      case OP_CONV3:
      case OP_UPCONV: {
        if (L.use_umma) {
          ConvUmmaLaunch C{};
          C.tmap_in = L.tmap;
          C.tmap_in2 = L.has_in2 ? L.tmap2 : L.tmap;
          C.Cin2 = L.has_in2 ? L.in2.C : 0;
          C.w_packed = L.d_wpacked;
          C.w_packed_pair = L.op == OP_CONV3 ? L.d_wpair : nullptr;
          C.bias = L.d_bias;
          C.out = L.out.ptr;
          C.out_ctot = L.out.ctot;
          C.out_coff = L.out.coff;
          C.N = n; C.D = L.in.D; C.H = L.in.H; C.W = L.in.W;
          C.Cin = L.Cin; C.Cout = L.Cout;
          C.ntaps = L.op == OP_CONV3 ? 27 : 1;
          C.up_stride = L.op == OP_UPCONV ? L.stride : 0;
          C.act_lrelu = L.lrelu ? 1 : 0;
          C.alpha = alpha;
          C.fp16 = net->act_type == ACT_F16;
          C.CB = L.cfg.CB; C.NT = L.cfg.NT; C.TZ = L.cfg.TZ;
          C.overflow = net->d_overflow;
          if (L.fused_head) {
            const Layer& H = net->layers[i + 1];
            C.head_w = H.d_w;
            C.head_b = H.head_b;
            C.head_labels = ho.labels;
            C.head_probs = ho.probs;
            C.head_logits = ho.logits;
            C.out = nullptr;                      // the 32-channel activation is never written
          }
          if (L.fused_pool) {
            C.pool_out = L.pool_view.ptr;
            C.pool_ctot = L.pool_view.ctot;
            C.pool_coff = L.pool_view.coff;
            if (!L.store_out) C.out = nullptr;
          }
          {
            // timing experiments only (results are wrong): TOMOENC_DBG_SKIP bit 0 drops the activation
            // stores, bit 1 the pooled stores
            static const int dbg = getenv("TOMOENC_DBG_SKIP") ? atoi(getenv("TOMOENC_DBG_SKIP")) : 0;
            if (dbg & 1) C.out = nullptr;
            if (dbg & 2) C.pool_out = nullptr;
          }
          if (C.out != nullptr) {
            if (L.omaps_n != n) {
              conv_umma_make_out_maps(&C);
              for (int k = 0; k < 8; ++k) L.omaps[k] = C.tmap_out[k];
              L.omaps_ok = C.use_tma_out;
              L.omaps_n = n;
            } else {
              for (int k = 0; k < 8; ++k) C.tmap_out[k] = L.omaps[k];
              C.use_tma_out = L.omaps_ok;
            }
          }
          e = conv_umma_launch(C, st);
          count_launch();
        } else if (L.op == OP_CONV3) {
          e = simt_conv3(with_n(L.in, n), with_n(L.out, n), L.d_w, L.d_bias, L.lrelu, alpha, net->act_type, st, net->d_overflow);
        } else {
          e = simt_upconv(with_n(L.in, n), with_n(L.out, n), L.d_w, L.d_bias, L.stride, L.lrelu, alpha, net->act_type, st);
        }
      } break;
      case OP_POOL:
        e = simt_maxpool(with_n(L.in, n), with_n(L.out, n), L.stride, net->act_type, st);
        break;
      case OP_HEAD:
        e = simt_head(with_n(L.in, n), L.d_w, L.head_b, ho.labels, ho.probs, ho.logits, net->act_type, st);
        break;
      case OP_DENSE: {
        const void* src = L.dense_in ? L.dense_in : dense_src;
        const int src_type = L.dense_in ? L.dense_in_type : dense_src_type;
        void* dst = L.dense_out ? L.dense_out : static_cast<void*>(net->d_dense_tmp[tmp_idx]);
        e = simt_dense(src, src_type, n, L.nin, L.nout, L.d_w, L.d_bias, L.lrelu, alpha, dst, L.dense_out_type, st, net->d_dense_acc);
        dense_src = dst;
        dense_src_type = L.dense_out_type;
        tmp_idx ^= 1;
      } break;
    }
    if (net->profiling) {
      cudaEventRecord(ev1, st);
      L.events.emplace_back(ev0, ev1);
    }
    if (e != cudaSuccess) {
      set_error("layer %s failed to launch: %s", L.name.c_str(), cudaGetErrorString(e));
      return TE_ERR_CUDA;
    }
  }
  return TE_OK;
}

}  // namespace

extern "C" int te_unet_forward(te_net* net, const void* x, int64_t n, void* out, int out_dtype, float* logits,
                               void* stream) {
  TE_CHECK_ARG(net && net->desc.kind == 0, "te_unet_forward: handle is not a U-net");
  TE_CHECK_ARG(x && (out || logits), "te_unet_forward: null pointer");
  TE_CHECK_ARG(n >= 0 && n <= net->desc.max_batch, "te_unet_forward: n = %lld exceeds max_batch = %d", (long long)n,
               net->desc.max_batch);
  TE_CHECK_ARG(out_dtype == TE_U8 || out_dtype == TE_F32, "te_unet_forward: out_dtype must be TE_U8 or TE_F32");
  if (!net->weights_set) { set_error("te_unet_forward: weights have not been set"); return TE_ERR_STATE; }
  if (n == 0) return TE_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t in_bytes = static_cast<size_t>(n) * net->input.D * net->input.H * net->input.W * net->elem;
  TE_CUDA(cudaMemcpyAsync(net->input.ptr, x, in_bytes, cudaMemcpyDeviceToDevice, st));
  HeadOut ho{nullptr, nullptr, logits};
  if (out && out_dtype == TE_U8) ho.labels = static_cast<uint8_t*>(out);
  if (out && out_dtype == TE_F32) ho.probs = static_cast<float*>(out);
  net->last_n = static_cast<int>(n);
  return run_layers(net, static_cast<int>(n), ho, st);
}

extern "C" int te_encoder_forward(te_net* net, const void* x, int64_t n, float* z, void* stream) {
  TE_CHECK_ARG(net && net->desc.kind == 1, "te_encoder_forward: handle is not an encoder");
  TE_CHECK_ARG(x && z, "te_encoder_forward: null pointer");
  TE_CHECK_ARG(n >= 0 && n <= net->desc.max_batch, "te_encoder_forward: n = %lld exceeds max_batch = %d", (long long)n,
               net->desc.max_batch);
  if (!net->weights_set) { set_error("te_encoder_forward: weights have not been set"); return TE_ERR_STATE; }
  if (n == 0) return TE_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t in_bytes = static_cast<size_t>(n) * net->input.D * net->input.H * net->input.W * net->elem;
  TE_CUDA(cudaMemcpyAsync(net->input.ptr, x, in_bytes, cudaMemcpyDeviceToDevice, st));
  // the last dense layer writes straight into the caller's buffer
  Layer& last = net->layers.back();
  last.dense_out = z;
  net->last_n = static_cast<int>(n);
  HeadOut ho{nullptr, nullptr, nullptr};
  const int rc = run_layers(net, static_cast<int>(n), ho, st);
  last.dense_out = nullptr;
  return rc;
}

extern "C" int te_decoder_forward(te_net* net, const float* z, int64_t n, float* out, void* stream) {
  TE_CHECK_ARG(net && net->desc.kind == 2, "te_decoder_forward: handle is not a decoder");
  TE_CHECK_ARG(z && out, "te_decoder_forward: null pointer");
  TE_CHECK_ARG(n >= 0 && n <= net->desc.max_batch, "te_decoder_forward: n = %lld exceeds max_batch = %d", (long long)n,
               net->desc.max_batch);
  if (!net->weights_set) { set_error("te_decoder_forward: weights have not been set"); return TE_ERR_STATE; }
  if (n == 0) return TE_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Layer& first = net->layers.front();
  const void* saved = first.dense_in;
  first.dense_in = z;
  first.dense_in_type = ACT_F32;
  net->last_n = static_cast<int>(n);
  HeadOut ho{nullptr, out, nullptr};
  const int rc = run_layers(net, static_cast<int>(n), ho, st);
  first.dense_in = saved;
  return rc;
}

extern "C" int64_t te_net_layer_output(te_net* net, int i, float* out, int64_t capacity, void* stream) {
  if (!net || i < 0 || i >= static_cast<int>(net->layers.size())) { set_error("te_net_layer_output: bad index"); return TE_ERR_ARG; }
  const Layer& L = net->layers[i];
  if (L.op == OP_HEAD || L.op == OP_DENSE || L.fused_head || !L.store_out || L.out.ptr == nullptr) {
    set_error("te_net_layer_output: layer %s has no stored activation tensor", L.name.c_str());
    return TE_ERR_STATE;
  }
  TView v = L.out;
  v.N = net->last_n;
  const long long count = v.voxels() * v.C;
  if (count > capacity) { set_error("te_net_layer_output: capacity too small (%lld needed)", count); return TE_ERR_ARG; }
  cudaError_t e = simt_to_f32(v, out, net->act_type, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) { set_error("te_net_layer_output: %s", cudaGetErrorString(e)); return TE_ERR_CUDA; }
  return count;
}

extern "C" int te_net_overflow(te_net* net, int reset, int* flag, void* stream) {
  TE_CHECK_ARG(net && flag, "te_net_overflow: null pointer");
  *flag = 0;
  if (!net->d_overflow) return TE_OK;                     // bf16 / fp32: the activation range is fp32's
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TE_CUDA(cudaMemcpyAsync(flag, net->d_overflow, sizeof(int), cudaMemcpyDeviceToHost, st));
  TE_CUDA(cudaStreamSynchronize(st));
  if (reset && *flag) TE_CUDA(cudaMemsetAsync(net->d_overflow, 0, sizeof(int), st));
  return TE_OK;
}

extern "C" int te_net_set_profiling(te_net* net, int enable) {
  TE_CHECK_ARG(net, "te_net_set_profiling: null handle");
  net->profiling = enable != 0;
  for (auto& L : net->layers) {
    for (auto& ev : L.events) { cudaEventDestroy(ev.first); cudaEventDestroy(ev.second); }
    L.events.clear();
    L.prof_ms = 0.0;
    L.prof_launches = 0;
  }
  return TE_OK;
}

extern "C" int te_net_layer_time(te_net* net, int i, double* total_ms, int64_t* launches) {
  TE_CHECK_ARG(net && i >= 0 && i < static_cast<int>(net->layers.size()), "te_net_layer_time: bad index");
  Layer& L = net->layers[i];
  for (auto& ev : L.events) {
    TE_CUDA(cudaEventSynchronize(ev.second));
    float ms = 0.f;
    TE_CUDA(cudaEventElapsedTime(&ms, ev.first, ev.second));
    L.prof_ms += ms;
    L.prof_launches += 1;
    cudaEventDestroy(ev.first);
    cudaEventDestroy(ev.second);
  }
  L.events.clear();
  if (total_ms) *total_ms = L.prof_ms;
  if (launches) *launches = L.prof_launches;
  return TE_OK;
}

extern "C" int te_net_layer_kernel(const te_net* net, int i, char* buf, int buflen) {
  TE_CHECK_ARG(net && buf && buflen > 0 && i >= 0 && i < static_cast<int>(net->layers.size()), "te_net_layer_kernel: bad argument");
  const Layer& L = net->layers[i];
  if (L.skip) snprintf(buf, buflen, "fused into previous layer");
  else if (L.use_umma) snprintf(buf, buflen, "conv_umma_kernel<%d,%d,%d>%s", L.cfg.CB, L.cfg.NT, L.cfg.TZ, L.fused_head ? "+head" : "");
  else if (L.op == OP_CONV3 && L.Cin == 1 && conv1_tz_applies(L.in, L.out, net->act_type))
    snprintf(buf, buflen, "conv1_tz_kernel<%d>", L.Cout);
  else if (L.op == OP_CONV3 && L.Cin == 1 && net->act_type != ACT_F32 && (L.Cout == 16 || L.Cout == 32))
    snprintf(buf, buflen, "conv1_umma_kernel<%d>", L.Cout);
  else {
    static const char* names[] = {"conv3_kernel", "upconv_kernel", "maxpool_kernel", "head_kernel", "dense_kernel"};
    snprintf(buf, buflen, "%s", names[L.op]);
  }
  return TE_OK;
}

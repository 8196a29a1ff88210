// Latent PCA on the GPU: streamed moments -> covariance -> symmetric eigensolve -> projection.
//
// Replaces sklearn.decomposition.IncrementalPCA(n_components=2, batch_size=500).fit / .transform as
// called by fit_PCA / transform_PCA (/root/reference/scratchpad/IEEE_ICIP_paper/latent_vis.py:94-143)
// with exact PCA (BASELINE.json north_star: "covariance accumulation and PCA projection ... feeding a
// small eigensolve").  The moments buffer [count, sum h (d), sum h h^T (d x d)] is float64 so that
// it can be all-reduced across ranks and so that cov = E[hh^T] - mean mean^T keeps ~1e-9 relative
// accuracy for fp32 latents.
#include "te_common.cuh"

namespace te {
namespace {

constexpr int kRB = 32;   // rows staged per iteration

// grid (d/32 col tiles, d/32 row tiles, row splits), block (32, 8).
__global__ void __launch_bounds__(256) pca_moments_kernel(const float* __restrict__ h, long long n, int d,
                                                          double* __restrict__ moments) {
  __shared__ float hj[kRB][33], hk[kRB][33];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int j0 = blockIdx.y * 32, k0 = blockIdx.x * 32;
  const long long rows_per = (n + gridDim.z - 1) / gridDim.z;
  const long long r_begin = rows_per * blockIdx.z;
  const long long r_end = r_begin + rows_per < n ? r_begin + rows_per : n;
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  double colsum[4] = {0.0, 0.0, 0.0, 0.0};
  for (long long r0 = r_begin; r0 < r_end; r0 += kRB) {
    for (int rr = ty; rr < kRB; rr += 8) {
      const long long r = r0 + rr;
      const bool ok = r < r_end;
      hj[rr][tx] = (ok && j0 + tx < d) ? h[r * d + j0 + tx] : 0.f;
      hk[rr][tx] = (ok && k0 + tx < d) ? h[r * d + k0 + tx] : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int rr = 0; rr < kRB; ++rr) {
      const double a = static_cast<double>(hk[rr][tx]);
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[i] = fma(static_cast<double>(hj[rr][ty + 8 * i]), a, acc[i]);
    }
    if (blockIdx.x == 0) {
      // column sums of the j tile: lane tx takes row tx, reduced across the warp below
#pragma unroll
      for (int i = 0; i < 4; ++i) colsum[i] += static_cast<double>(hj[tx][ty + 8 * i]);
    }
    __syncthreads();
  }
  double* s1 = moments + 1;
  double* s2 = moments + 1 + d;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int j = j0 + ty + 8 * i, k = k0 + tx;
    if (j < d && k < d) atomicAdd(s2 + static_cast<long long>(j) * d + k, acc[i]);
  }
  if (blockIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      double v = colsum[i];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      const int j = j0 + ty + 8 * i;
      if (tx == 0 && j < d) atomicAdd(s1 + j, v);
    }
    if (blockIdx.y == 0 && blockIdx.z == 0 && tx == 0 && ty == 0) atomicAdd(moments, static_cast<double>(n));
  }
}

// One CTA: parallel-ordered cyclic Jacobi on the covariance matrix (float64).
// work = A (d x d) followed by V (d x d).
__global__ void __launch_bounds__(1024) pca_solve_kernel(const double* __restrict__ moments, int d, int k,
                                                          double* __restrict__ mean, double* __restrict__ comps,
                                                          double* __restrict__ evar, double* __restrict__ work) {
  double* A = work;
  double* V = work + static_cast<long long>(d) * d;
  __shared__ double cs[128], sn[128];
  __shared__ int pp[128], qq[128];
  __shared__ int order[256];
  __shared__ double off_norm;
  const int tid = threadIdx.x, nt = blockDim.x;
  const double cnt = moments[0];
  const double* s1 = moments + 1;
  const double* s2 = moments + 1 + d;
  for (int j = tid; j < d; j += nt) mean[j] = s1[j] / cnt;
  __syncthreads();
  for (int e = tid; e < d * d; e += nt) {
    const int i = e / d, j = e % d;
    // symmetrise the accumulated outer-product sum (atomic order makes it symmetric only up to rounding)
    const double s = 0.5 * (s2[e] + s2[j * d + i]);
    A[e] = (s - cnt * mean[i] * mean[j]) / (cnt - 1.0);
    V[e] = i == j ? 1.0 : 0.0;
  }
  __syncthreads();
  const int m = (d + 1) / 2 * 2;          // players of the round-robin tournament (one dummy when d is odd)
  const int half = m / 2;
  for (int sweep = 0; sweep < 30; ++sweep) {
    if (tid == 0) off_norm = 0.0;
    __syncthreads();
    {
      double loc = 0.0;
      for (int e = tid; e < d * d; e += nt) { const int i = e / d, j = e % d; if (i != j) loc += A[e] * A[e]; }
      for (int o = 16; o > 0; o >>= 1) loc += __shfl_xor_sync(0xffffffffu, loc, o);
      if ((tid & 31) == 0) atomicAdd(&off_norm, loc);
    }
    __syncthreads();
    double diag = 0.0;
    if (tid == 0) { for (int i = 0; i < d; ++i) diag += A[i * d + i] * A[i * d + i]; }
    __shared__ int done;
    if (tid == 0) done = off_norm <= 1e-18 * (diag + 1e-300) ? 1 : 0;
    __syncthreads();
    if (done) break;
    for (int round = 0; round < m - 1; ++round) {
      // circle method: player m-1 fixed, the others rotate
      for (int t = tid; t < half; t += nt) {
        int a = t == 0 ? m - 1 : (round + t) % (m - 1);
        int b = (round + (m - 1) - t) % (m - 1);
        if (t == 0) b = round % (m - 1);
        int p = a < b ? a : b, q = a < b ? b : a;
        pp[t] = p; qq[t] = q;
        double c = 1.0, s = 0.0;
        if (q < d) {
          const double apq = A[p * d + q];
          if (fabs(apq) > 1e-300) {
            const double tau = (A[q * d + q] - A[p * d + p]) / (2.0 * apq);
            const double tt = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
            c = 1.0 / sqrt(1.0 + tt * tt);
            s = tt * c;
          }
        }
        cs[t] = c; sn[t] = s;
      }
      __syncthreads();
      // columns: A <- A J, V <- V J
      for (int e = tid; e < half * d; e += nt) {
        const int t = e / d, i = e % d;
        const int p = pp[t], q = qq[t];
        if (q >= d) continue;
        const double c = cs[t], s = sn[t];
        const double aip = A[i * d + p], aiq = A[i * d + q];
        A[i * d + p] = c * aip - s * aiq;
        A[i * d + q] = s * aip + c * aiq;
        const double vip = V[i * d + p], viq = V[i * d + q];
        V[i * d + p] = c * vip - s * viq;
        V[i * d + q] = s * vip + c * viq;
      }
      __syncthreads();
      // rows: A <- J^T A
      for (int e = tid; e < half * d; e += nt) {
        const int t = e / d, j = e % d;
        const int p = pp[t], q = qq[t];
        if (q >= d) continue;
        const double c = cs[t], s = sn[t];
        const double apj = A[p * d + j], aqj = A[q * d + j];
        A[p * d + j] = c * apj - s * aqj;
        A[q * d + j] = s * apj + c * aqj;
      }
      __syncthreads();
    }
  }
  // top-k eigenpairs, descending; sign: largest-|v| entry positive (sklearn svd_flip, u_based_decision=False)
  if (tid == 0) {
    for (int i = 0; i < d; ++i) order[i] = i;
    for (int c = 0; c < k; ++c) {
      int best = c;
      for (int i = c + 1; i < d; ++i) if (A[order[i] * d + order[i]] > A[order[best] * d + order[best]]) best = i;
      const int tmp = order[c]; order[c] = order[best]; order[best] = tmp;
    }
  }
  __syncthreads();
  for (int c = 0; c < k; ++c) {
    const int col = order[c];
    __shared__ double flip;
    if (tid == 0) {
      int jm = 0;
      double vm = -1.0;
      for (int j = 0; j < d; ++j) { const double a = fabs(V[j * d + col]); if (a > vm) { vm = a; jm = j; } }
      flip = V[jm * d + col] < 0.0 ? -1.0 : 1.0;
      evar[c] = A[col * d + col];
    }
    __syncthreads();
    for (int j = tid; j < d; j += nt) comps[c * d + j] = flip * V[j * d + col];
    __syncthreads();
  }
}

// Fast path (d <= 128): the same parallel-ordered cyclic Jacobi with the matrix resident in shared
// memory (m = d rounded up to even; the dummy row / column of an odd d stays zero and only ever sees
// identity rotations).  A round is ONE phase: thread-owned 2 x 2 blocks (pair i) x (pair j) take the
// column rotation of pair j and the row rotation of pair i together (A <- J^T A J touches each block
// exactly once), and the eigenvector matrix is kept transposed so that its update is a two-row rotation.
// Two barriers per round.
// Storage (d <= 128, everything in shared memory, nothing in L2 during the sweeps): A is kept as its packed upper
// triangle (m (m + 1) / 2 doubles; a round touches each 2 x 2 block (pair i) x (pair j) with i <= j once, the mirrored
// block is implied), and the transposed eigenvector matrix VT (d x d doubles) sits next to it -- 66 + 131 KB at d = 128.
// With VT in global memory a round cost 6.3 us (8.0 ms per solve, 8 % of an encode step of 4096 patches).
__device__ __forceinline__ int sym_idx(int i, int j, int m) {          // i <= j
  return i * m - (i * (i - 1)) / 2 + (j - i);
}
__device__ __forceinline__ int sym_at(int i, int j, int m) { return i <= j ? sym_idx(i, j, m) : sym_idx(j, i, m); }

__global__ void __launch_bounds__(1024) pca_solve_smem_kernel(const double* __restrict__ moments, int d, int k,
                                                               double* __restrict__ mean, double* __restrict__ comps,
                                                               double* __restrict__ evar, double* __restrict__ work) {
  extern __shared__ double dyn[];
  const int tid = threadIdx.x, nt = blockDim.x;
  const int m = (d + 1) / 2 * 2, half = m / 2;
  const int ntri = m * (m + 1) / 2, nblk = half * (half - 1) / 2;          // off-diagonal 2 x 2 blocks, ti < tj
  double* As = dyn;                        // packed upper triangle of the m x m matrix
  double* VT = dyn + ntri;                 // d x d, row e = eigenvector e
  unsigned short* blkmap = reinterpret_cast<unsigned short*>(VT + d * d);   // nblk entries: ti | tj << 8, ti < tj
  __shared__ double cs[64], sn[64];
  __shared__ int pp[64], qq[64];
  __shared__ int order[128];
  __shared__ double red[32];
  __shared__ int done;
  __shared__ double flip;
  int sweeps_done = 0;
  const double cnt = moments[0];
  const double* s1 = moments + 1;
  const double* s2 = moments + 1 + d;
  for (int j = tid; j < d; j += nt) mean[j] = s1[j] / cnt;
  __syncthreads();
  for (int e = tid; e < m * m; e += nt) {
    const int i = e / m, j = e % m;
    if (i > j) continue;
    double a = 0.0;
    if (i < d && j < d) {
      const double s = 0.5 * (s2[i * d + j] + s2[j * d + i]);
      a = (s - cnt * mean[i] * mean[j]) / (cnt - 1.0);
    }
    As[sym_idx(i, j, m)] = a;
  }
  for (int e = tid; e < d * d; e += nt) VT[e] = (e / d == e % d) ? 1.0 : 0.0;
  for (int e = tid; e < half * half; e += nt) {
    const int ti = e / half, tj = e % half;
    if (ti < tj) blkmap[ti * (half - 1) - (ti * (ti - 1)) / 2 + (tj - ti - 1)] = static_cast<unsigned short>(ti | (tj << 8));
  }
  __syncthreads();
  for (int sweep = 0; sweep < 30; ++sweep) {
    double off = 0.0, dg = 0.0;
    for (int e = tid; e < m * m; e += nt) {
      const int i = e / m, j = e % m;
      if (i > j) continue;
      const double a = As[sym_idx(i, j, m)];
      if (i == j) dg += a * a; else off += 2.0 * a * a;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      off += __shfl_xor_sync(0xffffffffu, off, o);
      dg += __shfl_xor_sync(0xffffffffu, dg, o);
    }
    // deterministic two-level reduction (no atomics: every launch takes the same number of sweeps)
    if ((tid & 31) == 0) red[tid >> 5] = off;
    __syncthreads();
    if (tid == 0) { double t = 0.0; for (int w = 0; w < (nt >> 5); ++w) t += red[w]; red[0] = t; }
    __syncthreads();
    const double off_all = red[0];
    __syncthreads();
    if ((tid & 31) == 0) red[tid >> 5] = dg;
    __syncthreads();
    if (tid == 0) {
      double t = 0.0;
      for (int w = 0; w < (nt >> 5); ++w) t += red[w];
      // Converged: ||off||_F <= 1e-9 ||diag||_F (eigenvector angles ~1e-9 / relative gap, eigenvalues ~1e-18 relative:
      // far below the fp32 projections this feeds).  The former 1e-13 test sat at the rounding-noise floor of a sweep
      // (||off||^2 ~ 1e-26 ||diag||^2 at d = 128): graded spectra ran 19+ sweeps of 0.5 ms each.
      done = off_all <= 1e-18 * (t + 1e-300) ? 1 : 0;
    }
    __syncthreads();
    if (done) break;
    ++sweeps_done;
    for (int round = 0; round < m - 1; ++round) {
      if (tid < half) {
        // circle method: player m-1 fixed, the others rotate
        const int t = tid;
        int a = t == 0 ? m - 1 : (round + t) % (m - 1);
        int b = t == 0 ? round % (m - 1) : (round + (m - 1) - t) % (m - 1);
        const int p = a < b ? a : b, q = a < b ? b : a;
        double c = 1.0, s = 0.0;
        const int ipp = sym_idx(p, p, m), ipq = sym_idx(p, q, m), iqq = sym_idx(q, q, m);
        const double apq = As[ipq];
        if (q < d && fabs(apq) > 1e-300) {
          const double app = As[ipp], aqq = As[iqq];
          const double tau = (aqq - app) / (2.0 * apq);
          const double tt = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
          c = 1.0 / sqrt(1.0 + tt * tt);
          s = tt * c;
          // the pair's own (diagonal) block, J^T [[app, apq], [apq, aqq]] J -- nobody else touches these three entries
          const double bpp = c * app - s * apq, bpq = s * app + c * apq;
          const double bqp = c * apq - s * aqq, bqq = s * apq + c * aqq;
          As[ipp] = c * bpp - s * bqp;
          As[iqq] = s * bpq + c * bqq;
          As[ipq] = c * bpq - s * bqq;
        }
        pp[t] = p; qq[t] = q; cs[t] = c; sn[t] = s;
      }
      __syncthreads();
      for (int blk = tid; blk < nblk; blk += nt) {
        const int ti = blkmap[blk] & 255, tj = blkmap[blk] >> 8;
        const int p1 = pp[ti], q1 = qq[ti], p2 = pp[tj], q2 = qq[tj];
        const double c1 = cs[ti], s1r = sn[ti], c2 = cs[tj], s2r = sn[tj];
        const int ipp = sym_at(p1, p2, m), ipq = sym_at(p1, q2, m), iqp = sym_at(q1, p2, m), iqq = sym_at(q1, q2, m);
        const double app = As[ipp], apq = As[ipq], aqp = As[iqp], aqq = As[iqq];
        // columns (pair j): [x_p, x_q] <- [c x_p - s x_q, s x_p + c x_q]
        const double bpp = c2 * app - s2r * apq, bpq = s2r * app + c2 * apq;
        const double bqp = c2 * aqp - s2r * aqq, bqq = s2r * aqp + c2 * aqq;
        // rows (pair i)
        As[ipp] = c1 * bpp - s1r * bqp;
        As[iqq] = s1r * bpq + c1 * bqq;
        As[ipq] = c1 * bpq - s1r * bqq;
        As[iqp] = s1r * bpp + c1 * bqp;
      }
      if (nt % d == 0) {            // thread = one column j of VT, pairs t0, t0 + nt / d, ... (no division in the loop)
        const int j = tid % d, tstep = nt / d;
        for (int t = tid / d; t < half; t += tstep) {
          const int p = pp[t], q = qq[t];
          if (q >= d) continue;
          const double c = cs[t], s = sn[t];
          const double vp = VT[p * d + j], vq = VT[q * d + j];
          VT[p * d + j] = c * vp - s * vq;
          VT[q * d + j] = s * vp + c * vq;
        }
      } else {
        for (int e = tid; e < half * d; e += nt) {
          const int t = e / d, j = e - t * d;
          const int p = pp[t], q = qq[t];
          if (q >= d) continue;
          const double c = cs[t], s = sn[t];
          const double vp = VT[p * d + j], vq = VT[q * d + j];
          VT[p * d + j] = c * vp - s * vq;
          VT[q * d + j] = s * vp + c * vq;
        }
      }
      __syncthreads();
    }
  }
  if (tid == 0) work[0] = static_cast<double>(sweeps_done);       // diagnostics: Jacobi sweeps of this solve
  // top-k eigenpairs, descending; sign: largest-|v| entry positive (sklearn svd_flip, u_based_decision=False)
  if (tid == 0) {
    for (int i = 0; i < d; ++i) order[i] = i;
    for (int c = 0; c < k; ++c) {
      int best = c;
      for (int i = c + 1; i < d; ++i)
        if (As[sym_idx(order[i], order[i], m)] > As[sym_idx(order[best], order[best], m)]) best = i;
      const int tmp = order[c]; order[c] = order[best]; order[best] = tmp;
    }
  }
  __syncthreads();
  for (int c = 0; c < k; ++c) {
    const int col = order[c];
    if (tid == 0) {
      int jm = 0;
      double vm = -1.0;
      for (int j = 0; j < d; ++j) { const double a = fabs(VT[col * d + j]); if (a > vm) { vm = a; jm = j; } }
      flip = VT[col * d + jm] < 0.0 ? -1.0 : 1.0;
      evar[c] = As[sym_idx(col, col, m)];
    }
    __syncthreads();
    for (int j = tid; j < d; j += nt) comps[c * d + j] = flip * VT[col * d + j];
    __syncthreads();
  }
}

// one warp per row
__global__ void __launch_bounds__(256) pca_project_kernel(const float* __restrict__ h, long long n, int d,
                                                          const double* __restrict__ mean,
                                                          const double* __restrict__ comps, int k,
                                                          float* __restrict__ z) {
  const int lane = threadIdx.x & 31;
  const long long warp = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  for (long long r = warp; r < n; r += nwarps) {
    for (int c = 0; c < k; ++c) {
      double acc = 0.0;
      for (int j = lane; j < d; j += 32) acc = fma(static_cast<double>(h[r * d + j]) - mean[j], comps[c * d + j], acc);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (lane == 0) z[r * k + c] = static_cast<float>(acc);
    }
  }
}

}  // namespace
}  // namespace te

using namespace te;

extern "C" int te_pca_accumulate(const float* h, int64_t n, int32_t d, double* moments, void* stream) {
  TE_CHECK_ARG(n >= 0, "te_pca_accumulate: n < 0");
  if (n == 0) return TE_OK;
  TE_CHECK_ARG(h && moments, "te_pca_accumulate: null pointer");
  TE_CHECK_ARG(d >= 1 && d <= 256, "te_pca_accumulate: d must be in [1, 256]");
  const int tiles = (d + 31) / 32;
  long long splits = (n + 255) / 256;
  const long long cap = (4LL * kNumSMs + tiles * tiles - 1) / (tiles * tiles);
  if (splits > cap) splits = cap;
  if (splits < 1) splits = 1;
  dim3 grid(tiles, tiles, static_cast<unsigned>(splits)), block(32, 8);
  pca_moments_kernel<<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(h, n, d, moments);
  TE_LAUNCH_CHECK();
  return TE_OK;
}

extern "C" int te_pca_solve(const double* moments, int32_t d, int32_t k, double* mean, double* components,
                            double* explained_variance, double* work, void* stream) {
  TE_CHECK_ARG(moments && mean && components && explained_variance && work, "te_pca_solve: null pointer");
  TE_CHECK_ARG(d >= 1 && d <= 256, "te_pca_solve: d must be in [1, 256]");
  TE_CHECK_ARG(k >= 1 && k <= d, "te_pca_solve: k must be in [1, d]");
  if (d <= 128) {
    const int m = (d + 1) / 2 * 2;
    const size_t smem = (static_cast<size_t>(m) * (m + 1) / 2 + static_cast<size_t>(d) * d) * sizeof(double) +
                        static_cast<size_t>(m / 2) * (m / 2 + 1) / 2 * sizeof(unsigned short) + 16;
    static unsigned long long attr_set = 0;
    TE_CUDA(smem_attr_once(pca_solve_smem_kernel, (128 * 129 / 2 + 128 * 128) * 8 + 64 * 65 / 2 * 2 + 16, &attr_set));
    pca_solve_smem_kernel<<<1, 1024, smem, static_cast<cudaStream_t>(stream)>>>(moments, d, k, mean, components,
                                                                                 explained_variance, work);
  } else {
    pca_solve_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(moments, d, k, mean, components,
                                                                        explained_variance, work);
  }
  TE_LAUNCH_CHECK();
  return TE_OK;
}

extern "C" int te_pca_project(const float* h, int64_t n, int32_t d, const double* mean, const double* components,
                              int32_t k, float* z, void* stream) {
  TE_CHECK_ARG(n >= 0, "te_pca_project: n < 0");
  if (n == 0) return TE_OK;
  TE_CHECK_ARG(h && mean && components && z, "te_pca_project: null pointer");
  long long blocks = (n * 32 + 255) / 256;
  if (blocks > 8LL * kNumSMs) blocks = 8LL * kNumSMs;
  pca_project_kernel<<<static_cast<int>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(h, n, d, mean, components,
                                                                                            k, z);
  TE_LAUNCH_CHECK();
  return TE_OK;
}

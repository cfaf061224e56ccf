// General-weight forms of the score used by the drop-in NEM / NEMOrderMCMC methods (fp64, off the
// batched hot path): compute_ll_ratios (nem.py:72-80, nem_order_mcmc.py:79-87, utils.py:66-82) and
// calculate_ll / compute_ll (nem.py:82-86, nem_order_mcmc.py:89-93, utils.py:84-94).
#include "common.cuh"

namespace {

// cell[i][e] = U[i][e] + sum_j log(1 - w_ij + w_ij * exp(T[j][e]))  with T two-valued (nem.py:46):
// the log term is l1_ij where D[j][e]==1 and l0_ij where 0; U[i][e] = T[i][e] + N[e], U[S][e] = N[e].
// blockIdx.y selects a chunk of kCellRows child rows i: a CTA stages only its rows of the two S x S tables
// (2 * 48 * 128 * 8 B = 96 KB at S = 128; the whole tables would need 256 KB, more than a CTA can have)
constexpr int kCellRows = 48;
__global__ void __launch_bounds__(128) k_cell_ratios(const uint32_t *__restrict__ dcol, int S, int W, int E, double A, double B,
                                                     const double *__restrict__ weights, int apply_expit,
                                                     double *__restrict__ cell) {
  // blockIdx.z: chain of a batch (the device-resident fit loop): S x S weights and (S+1) x E cell ratios per chain
  weights += (size_t)blockIdx.z * S * S;
  cell += (size_t)blockIdx.z * (S + 1) * E;
  extern __shared__ double sh[];  // l0[rows*S], l1[rows*S]
  const int i_lo = blockIdx.y * kCellRows, i_hi = min(S, i_lo + kCellRows), rows = i_hi - i_lo;
  double *l0 = sh, *l1 = sh + rows * S;
  const double eA = exp(-A), eB = exp(B);
  for (int k = threadIdx.x; k < rows * S; k += blockDim.x) {
    const int i = i_lo + k / S, j = k % S;
    double w = weights[i * S + j];
    bool present = w != 0.0 && i != j;
    if (apply_expit && present) w = 1.0 / (1.0 + exp(-w));  // the reference's MCMC driver quirk
    l0[k] = present ? log(1.0 - w + w * eB) : 0.0;
    l1[k] = present ? log(1.0 - w + w * eA) : 0.0;
  }
  __syncthreads();
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  uint32_t d[4] = {0, 0, 0, 0};
  int ones = 0;
  for (int w = 0; w < W; ++w) { d[w] = dcol[(size_t)e * W + w]; ones += __popc(d[w]); }
  const double N = A * ones;
  for (int i = i_lo; i < i_hi; ++i) {
    double v = N + (((d[i >> 5] >> (i & 31)) & 1u) ? -A : B);
    const double *r0 = l0 + (i - i_lo) * S, *r1 = l1 + (i - i_lo) * S;
    for (int j = 0; j < S; ++j) v += ((d[j >> 5] >> (j & 31)) & 1u) ? r1[j] : r0[j];
    cell[(size_t)i * E + e] = v;
  }
  if (blockIdx.y == 0) cell[(size_t)S * E + e] = N;
}

// per-column logaddexp.reduce over `rows` rows, optional posterior weights exp(cell - colsum), block partial sums
__global__ void __launch_bounds__(256) k_cell_ll(const double *__restrict__ cell, int rows, int E, double *__restrict__ ow,
                                                 double *__restrict__ partial) {
  cell += (size_t)blockIdx.y * rows * E;                    // blockIdx.y: chain of a batch
  if (ow != nullptr) ow += (size_t)blockIdx.y * rows * E;
  partial += (size_t)blockIdx.y * gridDim.x;
  __shared__ double red[256];
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  double cs = 0.0;
  if (e < E) {
    double m = -INFINITY;
    for (int i = 0; i < rows; ++i) m = fmax(m, cell[(size_t)i * E + e]);
    double s = 0.0;
    for (int i = 0; i < rows; ++i) s += exp(cell[(size_t)i * E + e] - m);
    cs = m + log(s);
    if (ow != nullptr)
      for (int i = 0; i < rows; ++i) ow[(size_t)i * E + e] = exp(cell[(size_t)i * E + e] - cs);
  }
  red[threadIdx.x] = cs;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = red[0];
}

__global__ void k_sum_partials(const double *partial, int n, double *out) {  // blockIdx.x: chain of a batch
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += partial[(size_t)blockIdx.x * n + i];
    out[blockIdx.x] = s;
  }
}

}  // namespace

extern "C" int nemb200_cell_ratios(const nemb200_dataset *h, const double *d_weights, int apply_expit, double *d_cell,
                                   void *stream) {
  NEMB_REQUIRE(h && d_weights && d_cell, "null argument");
  const int rows = h->S < kCellRows ? h->S : kCellRows;
  size_t smem = sizeof(double) * 2 * rows * h->S;
  NEMB_CHECK_CUDA(cudaFuncSetAttribute(k_cell_ratios, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((h->E + 127) / 128, (h->S + kCellRows - 1) / kCellRows);
  k_cell_ratios<<<grid, 128, smem, (cudaStream_t)stream>>>(h->dcol, h->S, h->W, h->E, h->A, h->B, d_weights,
                                                          apply_expit, d_cell);
  NEMB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" size_t nemb200_cell_ll_workspace_bytes(int E) { return sizeof(double) * (size_t)((E + 255) / 256); }

extern "C" int nemb200_cell_ll(const double *d_cell, int rows, int E, double *d_order_weights, double *d_ll,
                               void *d_workspace, size_t workspace_bytes, void *stream) {
  NEMB_REQUIRE(d_cell && d_ll && d_workspace && rows >= 1 && E >= 1, "bad arguments");
  NEMB_REQUIRE(workspace_bytes >= nemb200_cell_ll_workspace_bytes(E), "cell_ll workspace too small");
  int blocks = (E + 255) / 256;
  k_cell_ll<<<blocks, 256, 0, (cudaStream_t)stream>>>(d_cell, rows, E, d_order_weights, (double *)d_workspace);
  k_sum_partials<<<1, 32, 0, (cudaStream_t)stream>>>((const double *)d_workspace, blocks, d_ll);
  NEMB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ---- device-side data synthesis: utils.create_real_knockdown_mat (utils.py:15-23) and
//      utils.create_observed_knockdown_mat (utils.py:25-35) for shapes where the reference's Python double
//      loop (S*E calls of random.random()) takes minutes.  The flip uniform of cell (i, e) is Philox4x32-10 with
//      key = seed and counter = (cell lo, cell hi, 0, 2), cell = i*E + e, u = (word0 + 0.5) / 2^32.
namespace {
__global__ void k_synth_knockdown(const uint32_t *__restrict__ adj_bits, const int32_t *__restrict__ end_nodes, int S, int W,
                                  int E, double alpha, double beta, uint64_t seed, uint8_t *__restrict__ real,
                                  uint8_t *__restrict__ obs) {
  const size_t cell = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= (size_t)S * E) return;
  const int i = (int)(cell / E), e = (int)(cell - (size_t)i * E);
  const int k = end_nodes[e];
  // E-gene e hangs on S-gene k: it reacts to the knock-down of k and of every i with an edge i -> k (utils.py:5-23)
  const int r = (i == k) || ((adj_bits[i * W + (k >> 5)] >> (k & 31)) & 1u);
  uint32_t ctr[4] = {(uint32_t)cell, (uint32_t)((unsigned long long)cell >> 32), 0u, 2u};
  philox4x32_10(ctr, (uint32_t)seed, (uint32_t)(seed >> 32));
  const double u = ((double)ctr[0] + 0.5) * (1.0 / 4294967296.0);
  int o = r;
  if (r == 0 && u < alpha) o = 1;        // false positive, utils.py:31-32
  else if (r == 1 && u < beta) o = 0;    // false negative, utils.py:33-34
  if (real != nullptr) real[cell] = (uint8_t)r;
  obs[cell] = (uint8_t)o;
}
}  // namespace

extern "C" int nemb200_synth_knockdown(const uint32_t *d_adj_bits, const int32_t *d_end_nodes, int S, int E, double alpha,
                                       double beta, uint64_t seed, uint8_t *d_real, uint8_t *d_observed, void *stream) {
  NEMB_REQUIRE(d_adj_bits && d_end_nodes && d_observed, "null argument");
  NEMB_REQUIRE(S >= 1 && S <= NEMB200_MAX_S && E >= 1, "bad shape S=%d E=%d", S, E);
  const size_t cells = (size_t)S * E;
  k_synth_knockdown<<<(unsigned)((cells + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_adj_bits, d_end_nodes, S, (S + 31) / 32, E,
                                                                                     alpha, beta, seed, d_real, d_observed);
  NEMB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ---- per-order edge-weight fit: the S(S-1)/2 one-dimensional fits of
//      NEMOrderMCMC.get_optimal_weights / calculate_local_optimum (nem_order_mcmc.py:160-208) as one launch.
//      One CTA per (child i, parent k) pair.  c[e] follows :161-164 (two-valued exp(T[k,e]), posterior attachment
//      weight of site k, expit of the current weight); the objective is local_ll_sum_penalized (:18-23) in the
//      variable s = expit(x) that the reference thresholds:
//          G(s) = -sum_e log(1 + c_e s) + |s - anc| + s (1 - s),        G'(s) = -H(s) + sign(s - anc) + 1 - 2 s,
//      H(s) = sum_e c_e / (1 + c_e s).  The reference runs scipy's L-BFGS-B stopped at tol = 0.01 (its result is not a
//      function of the inputs alone); this kernel returns the global minimiser over s in [0, 1]: end points, the kink
//      at anc and every stationary point bracketed by kFitGrid probes per side and refined by kFitBisect bisection
//      steps (the tests' CPU restatement runs the same procedure with the same constants).
namespace {
constexpr int kFitThreads = 256;
constexpr int kFitGrid = 16;
constexpr int kFitBisect = 40;

struct FitCtx {
  const uint32_t *dcol;
  const double *ow_k;  // posterior attachment weights of site k, E doubles
  int W, E, k;
  double w, u1, u0;    // expit(current weight), exp(-A) - 1, exp(B) - 1
  double *red;         // kFitThreads doubles of shared memory
};

// mode 0: H(s) = sum c/(1+cs);  mode 1: sum log(1+cs).  Every thread returns the block-wide sum (fixed tree: deterministic).
__device__ double fit_sum(const FitCtx &f, double s, int mode) {
  double acc = 0.0;
  for (int e = threadIdx.x; e < f.E; e += kFitThreads) {
    const uint32_t d = (f.dcol[(size_t)e * f.W + (f.k >> 5)] >> (f.k & 31)) & 1u;
    const double u = d ? f.u1 : f.u0;
    const double a = u * f.ow_k[e];
    const double b = 1.0 - f.w * a + f.w * u;
    const double c = a / b;
    acc += mode == 0 ? c / (1.0 + c * s) : log(1.0 + c * s);
  }
  __syncthreads();  // red may still be read by the previous call
  f.red[threadIdx.x] = acc;
  __syncthreads();
  for (int o = kFitThreads / 2; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) f.red[threadIdx.x] += f.red[threadIdx.x + o];
    __syncthreads();
  }
  return f.red[0];
}

__global__ void __launch_bounds__(kFitThreads) k_fit_edges(const uint32_t *__restrict__ dcol, int S, int W, int E, double A,
                                                           double B, const double *__restrict__ weights,
                                                           const double *__restrict__ ow, const double *__restrict__ anc_x,
                                                           const uint8_t *__restrict__ mask, double *__restrict__ out) {
  __shared__ double red[kFitThreads];
  {  // blockIdx.y: chain of a batch
    const size_t c = blockIdx.y;
    weights += c * S * S; anc_x += c * S * S; mask += c * S * S; out += c * S * S; ow += c * (size_t)(S + 1) * E;
  }
  const int pr = blockIdx.x, k = pr % S;
  if (!mask[pr]) {
    if (threadIdx.x == 0) out[pr] = weights[pr];
    return;
  }
  FitCtx f;
  f.dcol = dcol; f.ow_k = ow + (size_t)k * E; f.W = W; f.E = E; f.k = k;
  f.w = 1.0 / (1.0 + exp(-weights[pr]));
  f.u1 = exp(-A) - 1.0; f.u0 = exp(B) - 1.0;
  f.red = red;
  const double xa = anc_x[pr];
  const double anc = fmin(fmax(xa, 0.0), 1.0);
  // candidates in ascending order of s; ties keep the smaller s (strict < below)
  double cand[3 + 2 * kFitGrid];
  int nc = 0;
  cand[nc++] = 0.0;
  for (int side = 0; side < 2; ++side) {
    const double lo_end = side == 0 ? 0.0 : anc, hi_end = side == 0 ? anc : 1.0, sg = side == 0 ? -1.0 : 1.0;
    if (side == 1) cand[nc++] = anc;
    if (!(hi_end > lo_end)) continue;
    double prev_s = lo_end, prev_d = -fit_sum(f, lo_end, 0) + sg + 1.0 - 2.0 * lo_end;
    for (int g = 1; g <= kFitGrid; ++g) {
      const double s = lo_end + (hi_end - lo_end) * (double)g / (double)kFitGrid;
      const double d = -fit_sum(f, s, 0) + sg + 1.0 - 2.0 * s;
      if (prev_d < 0.0 && 0.0 <= d) {
        double lo = prev_s, hi = s;
        for (int it = 0; it < kFitBisect; ++it) {
          const double mid = 0.5 * (lo + hi);
          if (-fit_sum(f, mid, 0) + sg + 1.0 - 2.0 * mid < 0.0) lo = mid; else hi = mid;
        }
        cand[nc++] = 0.5 * (lo + hi);
      }
      prev_s = s; prev_d = d;
    }
  }
  cand[nc++] = 1.0;
  double best_s = 0.0, best_g = 0.0;
  for (int q = 0; q < nc; ++q) {
    const double s = cand[q];
    const double g = -fit_sum(f, s, 1) + fabs(s - xa) + s * (1.0 - s);
    if (q == 0 || g < best_g) { best_s = s; best_g = g; }
  }
  if (threadIdx.x == 0) out[pr] = best_s;
}
}  // namespace

// ---- ancestor_x of get_optimal_weights (nem_order_mcmc.py:185): clip((I - expit_parent_weights(W))^-1 - I, 0, 1).
// One CTA inverts the S x S matrix in shared memory: in-place Gauss-Jordan with partial (row) pivoting, the column
// permutation undone at the end - what scipy.linalg.inv's LU does for the reference, without leaving the device.
// (For weights that respect an order I - expit(W) is a permuted unit-triangular matrix: every pivot is exactly 1.)
namespace {
constexpr int kInvThreads = 256;
__global__ void __launch_bounds__(kInvThreads) k_ancestor_x(const double *__restrict__ weights, const uint8_t *__restrict__ perm,
                                                            int S, double *__restrict__ out, int *__restrict__ singular) {
  extern __shared__ double a[];  // S x S
  __shared__ int piv[NEMB200_MAX_S];
  __shared__ double red_v[kInvThreads];
  __shared__ int red_i[kInvThreads];
  const int tid = threadIdx.x;
  weights += (size_t)blockIdx.x * S * S; perm += (size_t)blockIdx.x * S * S; out += (size_t)blockIdx.x * S * S;  // chain of a batch
  for (int i = tid; i < S * S; i += kInvThreads) {
    const double w = weights[i];
    const double e = perm[i] ? 1.0 / (1.0 + exp(-w)) : w;  // expit on the permissible parents only (:152-157)
    a[i] = ((i / S) == (i % S) ? 1.0 : 0.0) - e;
  }
  __syncthreads();
  for (int c = 0; c < S; ++c) {
    // pivot: the row r >= c with the largest |a[r][c]|
    double bv = -1.0; int bi = c;
    for (int r = c + tid; r < S; r += kInvThreads) { const double v = fabs(a[r * S + c]); if (v > bv) { bv = v; bi = r; } }
    red_v[tid] = bv; red_i[tid] = bi;
    __syncthreads();
    for (int o = kInvThreads / 2; o > 0; o >>= 1) {
      if (tid < o && (red_v[tid + o] > red_v[tid] || (red_v[tid + o] == red_v[tid] && red_i[tid + o] < red_i[tid]))) {
        red_v[tid] = red_v[tid + o]; red_i[tid] = red_i[tid + o];
      }
      __syncthreads();
    }
    const int r = red_i[0];
    if (tid == 0) { piv[c] = r; if (!(red_v[0] > 0.0)) *singular = 1; }
    if (r != c) for (int j = tid; j < S; j += kInvThreads) { const double t = a[c * S + j]; a[c * S + j] = a[r * S + j]; a[r * S + j] = t; }
    __syncthreads();
    const double p = a[c * S + c];
    __syncthreads();
    for (int j = tid; j < S; j += kInvThreads) a[c * S + j] = (j == c ? 1.0 : a[c * S + j]) / p;
    __syncthreads();
    // eliminate column c from every other row (the column itself becomes -f / p, the in-place inverse update)
    for (int idx = tid; idx < S * S; idx += kInvThreads) {
      const int i = idx / S, j = idx - i * S;
      if (i == c) continue;
      const double f = a[i * S + c];
      if (j != c) a[idx] -= f * a[c * S + j];
    }
    __syncthreads();
    for (int i = tid; i < S; i += kInvThreads) if (i != c) a[i * S + c] = -a[i * S + c] * a[c * S + c];
    __syncthreads();
  }
  for (int c = S - 1; c >= 0; --c) {  // undo the row exchanges as column exchanges
    const int r = piv[c];
    if (r != c) for (int i = tid; i < S; i += kInvThreads) { const double t = a[i * S + c]; a[i * S + c] = a[i * S + r]; a[i * S + r] = t; }
    __syncthreads();
  }
  for (int i = tid; i < S * S; i += kInvThreads) {
    const double v = a[i] - ((i / S) == (i % S) ? 1.0 : 0.0);
    out[i] = fmin(fmax(v, 0.0), 1.0);
  }
}
}  // namespace

extern "C" int nemb200_ancestor_x(const double *d_weights, const uint8_t *d_permissible, int S, double *d_out,
                                  int *d_singular, void *stream) {
  NEMB_REQUIRE(d_weights && d_permissible && d_out && S >= 1 && S <= NEMB200_MAX_S, "bad arguments");
  const size_t smem = sizeof(double) * (size_t)S * S;
  NEMB_CHECK_CUDA(cudaFuncSetAttribute(k_ancestor_x, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_ancestor_x<<<1, kInvThreads, smem, (cudaStream_t)stream>>>(d_weights, d_permissible, S, d_out, d_singular);
  NEMB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int nemb200_fit_edges(const nemb200_dataset *h, const double *d_weights, const double *d_order_weights,
                                 const double *d_ancestor_x, const uint8_t *d_mask, double *d_out_weights, void *stream) {
  NEMB_REQUIRE(h && d_weights && d_order_weights && d_ancestor_x && d_mask && d_out_weights, "null argument");
  k_fit_edges<<<h->S * h->S, kFitThreads, 0, (cudaStream_t)stream>>>(h->dcol, h->S, h->W, h->E, h->A, h->B, d_weights,
                                                                    d_order_weights, d_ancestor_x, d_mask, d_out_weights);
  NEMB_CHECK_CUDA(cudaGetLastError());
  return 0;
}


// ---- the reference's order -> DAG step inside the device-resident loop (`rule = NEMB200_RULE_FIT`):
//      what NEMOrderMCMC.method does between proposal and acceptance (nem_order_mcmc.py:276-286), for ALL chains of a
//      batch at once and without a host round trip:
//        reset: every permissible pair gets weight w0 (:50-77; all pairs)            k_fit_prep
//        soft cell ratios of those weights, posterior attachment weights (:181-182)             k_cell_ratios, k_cell_ll
//        ancestor_x = clip((I - expit W)^-1 - I, 0, 1) (:185)                                   k_ancestor_x
//        one exact 1-D fit per permissible (node, parent) pair (:187-189, :160-170)            k_fit_edges
//        threshold at 0.5 -> DAG (:210-214), closed if use_nem (:216-221)                       k_fit_threshold (+ closure)
//        score of the thresholded graph through the same expit-on-weights cell ratios (:207)  k_fit_soft, k_cell_ratios, k_cell_ll
//      In: prop_bits = the permissibility graph of each proposed order (edge a -> b iff a comes before b);
//      out: prop_bits = the thresholded graph, prop_scores = its score.
namespace {
struct FitWs { uint8_t *perm; double *w, *cell, *ow, *anc, *fitted, *partial; int *sing; };
size_t fit_carve(int S, int E, int n, void *base, FitWs *f) {
  size_t off = 0;
  uint8_t *b = (uint8_t *)base;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
  const size_t ss = (size_t)n * S * S, ce = (size_t)n * (S + 1) * E;
  const size_t o_perm = take(ss), o_w = take(8 * ss), o_cell = take(8 * ce), o_ow = take(8 * ce), o_anc = take(8 * ss),
               o_fit = take(8 * ss), o_part = take(8 * (size_t)n * ((E + 255) / 256)), o_sing = take(16);
  if (f) {
    f->perm = b + o_perm; f->w = (double *)(b + o_w); f->cell = (double *)(b + o_cell); f->ow = (double *)(b + o_ow);
    f->anc = (double *)(b + o_anc); f->fitted = (double *)(b + o_fit); f->partial = (double *)(b + o_part);
    f->sing = (int *)(b + o_sing);
  }
  return off;
}
// perm[i][k] = 1 iff k is a permissible parent of i = edge k -> i of the permissibility graph; W[i][k] = w0 there
__global__ void k_fit_prep(const uint32_t *__restrict__ bits, int S, int W, double w0, uint8_t *__restrict__ perm,
                           double *__restrict__ w) {
  const size_t c = blockIdx.x;
  for (int idx = threadIdx.x; idx < S * S; idx += blockDim.x) {
    const int i = idx / S, k = idx - i * S;
    const uint32_t on = (bits[(c * S + k) * W + (i >> 5)] >> (i & 31)) & 1u;
    perm[c * S * S + idx] = (uint8_t)on;
    w[c * S * S + idx] = on ? w0 : 0.0;
  }
}
// fitted weights -> graph: edge k -> i iff W[i][k] > 0.5 (dag = dag_weights.T, :212-213)
__global__ void k_fit_threshold(const double *__restrict__ fitted, const uint8_t *__restrict__ perm, int S, int W,
                                uint32_t *__restrict__ bits) {
  const size_t c = blockIdx.x;
  for (int idx = threadIdx.x; idx < S * W; idx += blockDim.x) {
    const int k = idx / W, wd = idx - k * W;
    uint32_t m = 0;
    for (int b = 0; b < 32; ++b) {
      const int i = wd * 32 + b;
      if (i < S && perm[c * S * S + (size_t)i * S + k] && fitted[c * S * S + (size_t)i * S + k] > 0.5) m |= 1u << b;
    }
    bits[(c * S + k) * W + wd] = m;
  }
}
// weights of the thresholded (and possibly closed) graph for the soft score: 1 on its edges, "0 but present" on the
// other permissible pairs (the reference's loop over parents_list applies expit(0) = 0.5 to them), 0 elsewhere
__global__ void k_fit_soft(const uint32_t *__restrict__ bits, const uint8_t *__restrict__ perm, int S, int W,
                           double *__restrict__ w) {
  const size_t c = blockIdx.x;
  for (int idx = threadIdx.x; idx < S * S; idx += blockDim.x) {
    const int i = idx / S, k = idx - i * S;
    const uint32_t on = (bits[(c * S + k) * W + (i >> 5)] >> (i & 31)) & 1u;
    w[c * S * S + idx] = perm[c * S * S + idx] ? (on ? 1.0 : 1e-300) : 0.0;
  }
}
}  // namespace

size_t fit_pipeline_workspace_bytes(const nemb200_dataset *h, int n) { return fit_carve(h->S, h->E, n, nullptr, nullptr); }

int launch_fit_pipeline(const nemb200_dataset *h, int n, uint32_t *prop_bits, double *prop_scores, void *ws, double w0,
                        int use_nem, cudaStream_t st) {
  const int S = h->S, W = h->W, E = h->E;
  FitWs f;
  fit_carve(S, E, n, ws, &f);
  const int rows = S < kCellRows ? S : kCellRows;
  const size_t smem_cell = sizeof(double) * 2 * rows * S, smem_inv = sizeof(double) * (size_t)S * S;
  NEMB_CHECK_CUDA(cudaFuncSetAttribute(k_cell_ratios, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cell));
  NEMB_CHECK_CUDA(cudaFuncSetAttribute(k_ancestor_x, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_inv));
  const dim3 gcell((E + 127) / 128, (S + kCellRows - 1) / kCellRows, n), gll((E + 255) / 256, n);
  k_fit_prep<<<n, 256, 0, st>>>(prop_bits, S, W, w0, f.perm, f.w);
  k_cell_ratios<<<gcell, 128, smem_cell, st>>>(h->dcol, S, W, E, h->A, h->B, f.w, 1, f.cell);
  k_cell_ll<<<gll, 256, 0, st>>>(f.cell, S + 1, E, f.ow, f.partial);
  k_ancestor_x<<<n, kInvThreads, smem_inv, st>>>(f.w, f.perm, S, f.anc, f.sing);
  k_fit_edges<<<dim3(S * S, n), kFitThreads, 0, st>>>(h->dcol, S, W, E, h->A, h->B, f.w, f.ow, f.anc, f.perm, f.fitted);
  k_fit_threshold<<<n, 128, 0, st>>>(f.fitted, f.perm, S, W, prop_bits);
  if (use_nem) {
    int rc = launch_closure(prop_bits, n, S, prop_bits, st);
    if (rc) return rc;
  }
  k_fit_soft<<<n, 256, 0, st>>>(prop_bits, f.perm, S, W, f.w);
  k_cell_ratios<<<gcell, 128, smem_cell, st>>>(h->dcol, S, W, E, h->A, h->B, f.w, 1, f.cell);
  k_cell_ll<<<gll, 256, 0, st>>>(f.cell, S + 1, E, nullptr, f.partial);
  k_sum_partials<<<n, 32, 0, st>>>(f.partial, (E + 255) / 256, prop_scores);
  NEMB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

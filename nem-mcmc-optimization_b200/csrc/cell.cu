// General-weight forms of the score used by the drop-in NEM / NEMOrderMCMC methods (fp64, off the
// batched hot path): compute_ll_ratios (nem.py:72-80, nem_order_mcmc.py:79-87, utils.py:66-82) and
// calculate_ll / compute_ll (nem.py:82-86, nem_order_mcmc.py:89-93, utils.py:84-94).
#include "common.cuh"

namespace {

// cell[i][e] = U[i][e] + sum_j log(1 - w_ij + w_ij * exp(T[j][e]))  with T two-valued (nem.py:46):
// the log term is l1_ij where D[j][e]==1 and l0_ij where 0; U[i][e] = T[i][e] + N[e], U[S][e] = N[e].
__global__ void __launch_bounds__(128) k_cell_ratios(const uint32_t *__restrict__ dcol, int S, int W, int E, double A, double B,
                                                     const double *__restrict__ weights, int apply_expit,
                                                     double *__restrict__ cell) {
  extern __shared__ double sh[];  // l0[S*S], l1[S*S]
  double *l0 = sh, *l1 = sh + S * S;
  const double eA = exp(-A), eB = exp(B);
  for (int i = threadIdx.x; i < S * S; i += blockDim.x) {
    double w = weights[i];
    bool present = w != 0.0 && (i / S) != (i % S);
    if (apply_expit && present) w = 1.0 / (1.0 + exp(-w));  // the reference's MCMC driver quirk
    l0[i] = present ? log(1.0 - w + w * eB) : 0.0;
    l1[i] = present ? log(1.0 - w + w * eA) : 0.0;
  }
  __syncthreads();
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  uint32_t d[4] = {0, 0, 0, 0};
  int ones = 0;
  for (int w = 0; w < W; ++w) { d[w] = dcol[(size_t)e * W + w]; ones += __popc(d[w]); }
  const double N = A * ones;
  for (int i = 0; i < S; ++i) {
    double v = N + (((d[i >> 5] >> (i & 31)) & 1u) ? -A : B);
    for (int j = 0; j < S; ++j) v += ((d[j >> 5] >> (j & 31)) & 1u) ? l1[i * S + j] : l0[i * S + j];
    cell[(size_t)i * E + e] = v;
  }
  cell[(size_t)S * E + e] = N;
}

// per-column logaddexp.reduce over `rows` rows, optional posterior weights exp(cell - colsum), block partial sums
__global__ void __launch_bounds__(256) k_cell_ll(const double *__restrict__ cell, int rows, int E, double *__restrict__ ow,
                                                 double *__restrict__ partial) {
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

__global__ void k_sum_partials(const double *partial, int n, double *out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += partial[i];
    *out = s;
  }
}

}  // namespace

extern "C" int nemb200_cell_ratios(const nemb200_dataset *h, const double *d_weights, int apply_expit, double *d_cell,
                                   void *stream) {
  NEMB_REQUIRE(h && d_weights && d_cell, "null argument");
  size_t smem = sizeof(double) * 2 * h->S * h->S;
  NEMB_CHECK_CUDA(cudaFuncSetAttribute(k_cell_ratios, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_cell_ratios<<<(h->E + 127) / 128, 128, smem, (cudaStream_t)stream>>>(h->dcol, h->S, h->W, h->E, h->A, h->B, d_weights,
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

// Resident dataset: replaces the score tables of the reference (nem.py:25-64, Q11: the reference
// materialises S tables of S x E float64; everything the score needs is the 0/1 matrix D and A, B).
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <new>

#include "common.cuh"

static thread_local char g_err[512] = "";
void nemb200_set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
extern "C" const char *nemb200_last_error(void) { return g_err; }
extern "C" int nemb200_version(void) { return NEMB200_VERSION; }

namespace {

// pair-merged table kernel (score_tc.cu): 32 accumulator columns per graph, S + 9 K rows <= 64
bool pair_supported(int S) { return S >= 33 && S <= 52; }
// multi-level table kernel (score_tc.cu, EPI_LUTM): its own operand image (4*D + three constant rows, S + 3 K rows <= 128)
// and a compact table of (S+1)(S+2)/2 <= kLutmStride entries per level.  Meant for 64 <= S <= 125; smaller S use it when the
// error rates put the single-scale table out of range (e.g. 55 <= S <= 63 at alpha = 0.05, beta = 0.1)
bool lutm_supported(int S) { return S >= 16 && S <= NEMB200_LUTM_MAX_S; }

struct Layout { size_t dcol, a_lut, a_pair, a_lutm, lut, tabs, lutm, tile_first, sumN, counters, total; int W, n_mtiles; };
Layout make_layout(int S, int E) {
  Layout L;
  L.W = (S + 31) / 32;
  L.n_mtiles = (E + 127) / 128;
  size_t off = 0;
  L.dcol = off; off = align_up(off + (size_t)L.n_mtiles * 128 * L.W * 4, 1024);
  L.a_lut = off; off = align_up(off + (size_t)L.n_mtiles * 128 * 128, 1024);
  L.a_pair = off; off = align_up(off + (pair_supported(S) ? (size_t)L.n_mtiles * 128 * 128 : 0), 1024);
  L.a_lutm = off; off = align_up(off + (lutm_supported(S) ? (size_t)L.n_mtiles * 128 * 128 : 0), 1024);
  L.lut = off; off = align_up(off + 64 * 64 * 4, 1024);
  L.tabs = off; off = align_up(off + 65 * 66 / 2 * 4, 1024);
  L.lutm = off; off = align_up(off + (lutm_supported(S) ? (size_t)NEMB200_LUTM_LEVELS * NEMB200_LUTM_STRIDE * 4 : 0), 1024);
  L.tile_first = off; off = align_up(off + (lutm_supported(S) ? (size_t)L.n_mtiles : 0), 1024);
  L.sumN = off; off = align_up(off + 8, 1024);
  L.counters = off; off = align_up(off + 192 * 4, 1024);
  L.total = off;
  return L;
}

// bit i of dcol[e][w] = D[i][e]; rows e >= E are zero
__global__ void k_pack_dcol(const uint8_t *D, int S, int E, int W, int E_pad, uint32_t *dcol) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E_pad) return;
  for (int w = 0; w < W; ++w) {
    uint32_t m = 0;
    if (e < E)
      for (int b = 0; b < 32 && w * 32 + b < S; ++b) m |= (uint32_t)(D[(size_t)(w * 32 + b) * E + e] != 0) << b;
    dcol[(size_t)e * W + w] = m;
  }
}

// u8 A-operand image, one 16 KB SWIZZLE_128B atom per 128 E-genes:
//   a[e][k] = 2 if D[k][e] else 128 (k < S, e < E); a[e][S] = 128 for every row (exp-table base row); 0 elsewhere
// pair == 1: the image of the pair-merged kernel: a[e][k] = 4*D[k][e]; a[e][S..S+7] = 128, a[e][S+8] = 4 for every row
// (constant rows: the B operand's entries there add the byte address of the column's table)
// pair == 2: the image of the multi-level table kernel: the same with two rows of 128 (S, S+1) and the 4 in row S+2
__global__ void k_build_aimg(const uint32_t *dcol, int S, int E, int W, int E_pad, int pair, uint8_t *img) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;  // (row e, 16-byte chunk c)
  if (idx >= E_pad * 8) return;
  int e = idx >> 3, c = idx & 7;
  uint32_t bytes[4] = {0, 0, 0, 0};
  for (int b = 0; b < 16; ++b) {
    int k = c * 16 + b;
    uint32_t v = 0;
    if (pair) {
      if (k < S) v = (e < E && ((dcol[(size_t)e * W + (k >> 5)] >> (k & 31)) & 1u)) ? 4u : 0u;
      else if (k < S + (pair == 2 ? 2 : 8)) v = 128u;
      else if (k == S + (pair == 2 ? 2 : 8)) v = 4u;
    } else if (k < S && e < E) v = ((dcol[(size_t)e * W + (k >> 5)] >> (k & 31)) & 1u) ? 2u : 128u;
    else if (k == S) v = 128u;
    bytes[b >> 2] |= v << ((b & 3) * 8);
  }
  int tile = e >> 7, r = e & 127;
  size_t off = (size_t)tile * 16384 + (size_t)(r >> 3) * 1024 + (size_t)(r & 7) * 128 + (size_t)((c ^ (r & 7)) << 4);
  *reinterpret_cast<uint4 *>(img + off) = make_uint4(bytes[0], bytes[1], bytes[2], bytes[3]);
}

__global__ void k_sumN(const uint32_t *dcol, size_t n_words, double A, double *out) {
  __shared__ unsigned long long sh[256];
  unsigned long long c = 0;
  for (size_t i = threadIdx.x; i < n_words; i += blockDim.x) c += __popc(dcol[i]);
  sh[threadIdx.x] = c;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = A * (double)sh[0];  // nem.py:62 summed over all E-genes
}

// lut[q][p] = 2^(alpha2*p + beta2*q - s) for p+q <= S (0 elsewhere), computed in fp64
__global__ void k_build_lut(int S, double alpha2, double beta2, double s, float *lut, float *tabs) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 64 * 64) return;
  int q = i >> 6, p = i & 63, n = p + q;
  const float v = (n <= S) ? (float)exp2(alpha2 * p + beta2 * q - s) : 0.f;
  lut[i] = v;
  if (n <= S && n < 64) tabs[n * (n + 1) / 2 + p] = v;  // compact copy for the pair planner
}

// multi-level compact table: level l, entry n(n+1)/2 + p = 2^(alpha2*p + beta2*(n-p) - s_l) (fp64 -> fp32: +inf where a
// level overflows, 0 where it underflows; the epilogue never selects a level in either state, see score_tc.cu)
__global__ void k_build_lutm(int S, double alpha2, double beta2, double s0, double s1, double s2, int levels, float *lutm) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= NEMB200_LUTM_STRIDE) return;
  int n = 0;
  while ((n + 1) * (n + 2) / 2 <= i) ++n;
  const int p = i - n * (n + 1) / 2;
  for (int l = 0; l < NEMB200_LUTM_LEVELS; ++l) {
    const double s = l == 0 ? s0 : (l == 1 ? s1 : s2);
    float v = 0.f;
    if (n <= S && l < levels) {
      const double x = alpha2 * p + beta2 * (n - p) - s;
      v = x < -126.0 ? 0.f : (float)exp2(x);
    }
    lutm[(size_t)l * NEMB200_LUTM_STRIDE + i] = v;
  }
}

// Per E tile: the lowest-scale table level that cannot overflow for any of the tile's E-genes, whatever the graph.  A site's
// term is 2^(alpha2*p + beta2*q) with p <= ones(e), q <= S - ones(e), so x <= max(alpha2,0)*ones(e) + max(beta2,0)*(S-ones(e));
// levels above tile_first[m] are not needed for tile m (the epilogue skips their look-ups).  With the E-genes stored
// sorted by their number of ones the tiles are homogeneous and most of them need the last level only.
__global__ void k_tile_levels(const uint32_t *dcol, int S, int E, int W, int n_mtiles, double alpha2, double beta2, int levels,
                              double s0, double s1, double s2, uint8_t *tile_first) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= n_mtiles) return;
  const double ap = alpha2 > 0 ? alpha2 : 0.0, bp = beta2 > 0 ? beta2 : 0.0;
  double bound = 0.0;
  for (int e = m * 128; e < m * 128 + 128 && e < E; ++e) {
    int ones = 0;
    for (int w = 0; w < W; ++w) ones += __popc(dcol[(size_t)e * W + w]);
    bound = fmax(bound, ap * ones + bp * (S - ones));
  }
  const double top = bound + log2((double)S + 1.0);
  int first = 0;
  for (int l = 1; l < levels; ++l) {
    const double s = l == 1 ? s1 : s2;
    if (top < s + 127.0) first = l;
  }
  (void)s0;
  tile_first[m] = (uint8_t)first;
}

}  // namespace

extern "C" size_t nemb200_dataset_bytes(int S, int E) {
  if (S < 1 || S > NEMB200_MAX_S || E < 1) return 0;
  return make_layout(S, E).total;
}

extern "C" int nemb200_dataset_create(const uint8_t *d_D, int S, int E, double A, double B, void *d_workspace,
                                      size_t workspace_bytes, void *stream, nemb200_dataset **out) {
  NEMB_REQUIRE(out != nullptr && d_D != nullptr && d_workspace != nullptr, "null argument");
  NEMB_REQUIRE(S >= 1 && S <= NEMB200_MAX_S, "S must be in [1,%d] (got %d)", NEMB200_MAX_S, S);
  NEMB_REQUIRE(E >= 1, "E must be >= 1 (got %d)", E);
  NEMB_REQUIRE(isfinite(A) && isfinite(B), "A and B must be finite");
  Layout L = make_layout(S, E);
  NEMB_REQUIRE(workspace_bytes >= L.total, "dataset workspace too small: %zu < %zu", workspace_bytes, L.total);
  NEMB_REQUIRE(((uintptr_t)d_workspace & 1023) == 0, "dataset workspace must be 1024-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  nemb200_dataset *h = (nemb200_dataset *)calloc(1, sizeof(nemb200_dataset));
  NEMB_REQUIRE(h != nullptr, "out of host memory");
  h->cache = new (std::nothrow) nemb200_cache();
  if (h->cache == nullptr) { free(h); nemb200_set_error("out of host memory"); return -1; }
  uint8_t *base = (uint8_t *)d_workspace;
  h->S = S; h->E = E; h->W = L.W; h->KA = 1; h->n_mtiles = L.n_mtiles;
  h->CS = S;  // set below once the epilogue kind is known
  h->A = A; h->B = B;
  h->alpha2 = -A * NEMB200_LOG2E; h->beta2 = B * NEMB200_LOG2E;
  h->dcol = (uint32_t *)(base + L.dcol); h->a_lut = base + L.a_lut; h->a_pair = base + L.a_pair; h->a_lutm = base + L.a_lutm;
  h->lut = (float *)(base + L.lut); h->tabs = (float *)(base + L.tabs); h->sumN = (double *)(base + L.sumN);
  h->counters = (int *)(base + L.counters);
  // single-scale validity: largest log2-domain value over the lattice p+q<=S
  double vmax = 0.0;
  if (h->alpha2 > 0) vmax = fmax(vmax, h->alpha2 * S);
  if (h->beta2 > 0) vmax = fmax(vmax, h->beta2 * S);
  double s = ceil(vmax + log2((double)S + 1.0) - 126.0);
  if (s < 0) s = 0;
  h->scale = (int)s;
  h->fixed_valid = s <= 100.0;
  h->lut_valid = h->fixed_valid && S <= 63;
  // tile layout: the table kernels are compiled for a fixed number of accumulator columns per graph - S itself
  // for the headline shapes, otherwise S rounded up to a multiple of 8 (padding sites read a zero table entry)
  if (h->lut_valid && S >= 8 && !(S == 10 || S == 20 || S == 30 || S == 50)) h->CS = (S + 7) / 8 * 8;
  h->G = 256 / h->CS < 16 ? 256 / h->CS : 16;   // graphs per MMA N tile: a multiple of 4 (one quarter per
  h->G = h->G >= 4 ? (h->G & ~3) : 2;           // epilogue warpgroup), or 2 for S > 64
  h->G_lut = h->G;  // the specialised table kernel may use its own tile width (score_tc.cu: tile_graphs)
  int E_pad = L.n_mtiles * 128;
  if (cudaMemsetAsync(h->counters, 0, 192 * sizeof(int), st) != cudaSuccess) { delete h->cache; free(h); nemb200_set_error("cudaMemsetAsync failed"); return -2; }
  k_pack_dcol<<<(E_pad + 255) / 256, 256, 0, st>>>(d_D, S, E, L.W, E_pad, h->dcol);
  k_build_aimg<<<(E_pad * 8 + 255) / 256, 256, 0, st>>>(h->dcol, S, E, L.W, E_pad, 0, h->a_lut);
  h->pair_ok = h->lut_valid && pair_supported(S);
  // multi-level table epilogue (16 <= S <= 125): the null site bounds every E-gene's largest term from below by
  // 2^0, so the levels only have to cover [0, vmax]: the lowest level has scale T-1, each further one sits
  // `step` = 121 + T - 1 binary orders higher, and level 0 must hold vmax + log2(S+1) without overflow (T = 90 is
  // the selection threshold of the epilogue: a level is used iff its sum is >= 2^-T)
  h->lutm_levels = 0;
  h->CS_lutm = (S + 7) / 8 * 8;
  h->G_lutm = 256 / h->CS_lutm < 16 ? 256 / h->CS_lutm : 16;  // graphs per tile: a multiple of 4, or 2 (S > 64)
  h->G_lutm = h->G_lutm >= 4 ? (h->G_lutm & ~3) : 2;
  if (lutm_supported(S)) {
    const double T = 90.0, step = 121.0 + T - 1.0, top = vmax + log2((double)S + 1.0);
    for (int lv = 2; lv <= NEMB200_LUTM_LEVELS && h->lutm_levels == 0; ++lv)
      if (top < (T - 1.0) + (lv - 1) * step + 127.0) h->lutm_levels = lv;
    for (int l = 0; l < h->lutm_levels; ++l) h->lutm_scale[l] = (T - 1.0) + (h->lutm_levels - 1 - l) * step;
  }
  h->lutm = (float *)(base + L.lutm);
  if (h->lutm_levels > 0)
    k_build_lutm<<<(NEMB200_LUTM_STRIDE + 255) / 256, 256, 0, st>>>(S, h->alpha2, h->beta2, h->lutm_scale[0], h->lutm_scale[1],
                                                                  h->lutm_scale[2], h->lutm_levels, h->lutm);
  h->tile_first = base + L.tile_first;
  if (h->lutm_levels > 0)
    k_tile_levels<<<(L.n_mtiles + 127) / 128, 128, 0, st>>>(h->dcol, S, E, L.W, L.n_mtiles, h->alpha2, h->beta2, h->lutm_levels,
                                                             h->lutm_scale[0], h->lutm_scale[1], h->lutm_scale[2], h->tile_first);
  if (h->pair_ok) k_build_aimg<<<(E_pad * 8 + 255) / 256, 256, 0, st>>>(h->dcol, S, E, L.W, E_pad, 1, h->a_pair);
  if (h->lutm_levels > 0) k_build_aimg<<<(E_pad * 8 + 255) / 256, 256, 0, st>>>(h->dcol, S, E, L.W, E_pad, 2, h->a_lutm);
  k_sumN<<<1, 256, 0, st>>>(h->dcol, (size_t)E_pad * L.W, A, h->sumN);
  k_build_lut<<<16, 256, 0, st>>>(S, h->alpha2, h->beta2, (double)h->scale, h->lut, h->tabs);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    delete h->cache;
    free(h);
    nemb200_set_error("dataset kernels failed to launch: %s", cudaGetErrorString(e));
    return -2;
  }
  *out = h;
  return 0;
}

extern "C" int nemb200_dataset_destroy(nemb200_dataset *h) {
  if (h && h->cache) {
    nemb200_cache *c = h->cache;
    if (c->mcmc_graph_exec) cudaGraphExecDestroy((cudaGraphExec_t)c->mcmc_graph_exec);
    if (c->mcmc_ev_fork) cudaEventDestroy((cudaEvent_t)c->mcmc_ev_fork);
    if (c->mcmc_ev_join) cudaEventDestroy((cudaEvent_t)c->mcmc_ev_join);
    if (c->mcmc_side_stream) cudaStreamDestroy((cudaStream_t)c->mcmc_side_stream);
    delete c;
  }
  free(h);
  return 0;
}

extern "C" int nemb200_dataset_info(const nemb200_dataset *h, int64_t *info, int n) {
  NEMB_REQUIRE(h != nullptr && info != nullptr && n >= 8, "dataset_info needs a handle and 8 slots");
  info[0] = h->S; info[1] = h->E; info[2] = h->W; info[3] = 128; info[4] = h->lut_valid ? h->G_lut : h->G;
  info[5] = h->lut_valid; info[6] = h->scale; info[7] = h->n_mtiles;
  if (n >= 9) info[8] = h->lutm_levels;
  if (n >= 10) info[9] = h->pair_ok ? 1 : 0;
  return 0;
}

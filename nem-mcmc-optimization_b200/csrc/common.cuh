// Shared declarations of the nemb200 library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <mutex>

#include "../../include/nemb200.h"

#define NEMB200_LOG2E 1.4426950408889634074
#define NEMB200_LN2 0.69314718055994530942
// multi-level table epilogue (score_tc.cu, EPI_LUTM): floats per level (n(n+1)/2 + p, n <= 125), number of levels
#define NEMB200_LUTM_STRIDE 8192
#define NEMB200_LUTM_LEVELS 3
#define NEMB200_LUTM_MAX_S 125   // S + 3 K rows <= 128; (S+1)(S+2)/2 <= 8001 table entries per level
#define NEMB200_LUTM_ZERO 8190   // entry that is 0 at every level (n = 127 > S): what padding columns look up

// ---- host handle ---------------------------------------------------------------------------
// The handle itself is immutable after nemb200_dataset_create.  The few pieces of host state the launchers keep
// between calls (which score workspace owns which completion counter, the cached CUDA graph of the MCMC iteration
// body, the side stream of the two-group loop) live in this separate object behind a mutex, so a `const
// nemb200_dataset *` is never cast: several host threads may score through one handle; nemb200_mcmc_run calls on
// one handle are serialised by the mutex.
struct nemb200_cache {
  std::recursive_mutex mu;  // recursive: nemb200_mcmc_run holds it while its steps call the score launcher
  const void *ctr_ws[64] = {};  // which score workspace owns which completion counter (counters live in the dataset
  int n_ctr = 0;                // workspace: library-initialised memory, zero at rest, whatever the caller allocates)
  // one cached CUDA graph of the MCMC iteration body (re-used while the same buffers are passed)
  unsigned long long mcmc_graph_key = 0;
  void *mcmc_graph_exec = nullptr;  // cudaGraphExec_t
  // second capture branch of the two-group MCMC loop (created on first use, destroyed with the handle)
  void *mcmc_side_stream = nullptr;                      // cudaStream_t
  void *mcmc_ev_fork = nullptr, *mcmc_ev_join = nullptr; // cudaEvent_t
};

struct nemb200_dataset {
  int S, E, W;        // S-genes, E-genes, words per bit row
  int KA;             // 64-element K atoms (bf16, SWIZZLE_128B rows of 128 B)
  int G;              // graphs per MMA N tile (runtime-S kernels: a multiple of 4, or 2)
  int G_lut;          // graphs per MMA N tile of the specialised table kernel (5 for 48/50-column graphs)
  int CS;             // accumulator columns per graph (S, or S rounded up to 8 for the specialised table kernels)
  int n_mtiles;       // E tiles of 128 rows
  double A, B;        // nem.py:17-18
  double alpha2, beta2;  // -A*log2(e), B*log2(e): log2-domain increments per ancestor with D=1 / D=0
  int lut_valid;      // single-scale exp table usable (S <= 63 and range fits fp32)
  int fixed_valid;    // single fixed scale usable for the ex2 epilogue
  int scale;          // s: all terms are computed times 2^-s
  // device pointers (inside the caller's workspace)
  uint32_t *dcol;     // [E_pad][W]  bit i of row e = D[i][e]
  uint8_t *a_lut;     // [n_mtiles][KA][128 x 128 B] bf16 operand image, values {4,256}+bias row
  uint8_t *a_pair;    // operand image of the pair-merged kernel: 4*D, eight constant rows of 128 and one of 4
  uint8_t *a_lutm;    // operand image of the multi-level table kernel: 4*D, two constant rows of 128 and one of 4
  int pair_ok;        // pair-merged table kernel usable (table valid, 33 <= S <= 52)
  float *tabs;        // compact exp table of the pair planner: tabs[n(n+1)/2 + p] = lut[n-p][p], n <= S
  float *lutm;        // [levels][NEMB200_LUTM_STRIDE] multi-level compact exp table (16 <= S <= 125)
  int CS_lutm;        // accumulator columns per graph of the multi-level table kernel: S rounded up to a multiple of 8
  int G_lutm;         // graphs per MMA N tile of that kernel
  uint8_t *tile_first;  // [n_mtiles] first (highest-scale) table level E tile m needs (dataset.cu: k_tile_levels)
  int lutm_levels;    // 0 = not usable for this dataset, else 1..3 levels cover the dynamic range
  double lutm_scale[NEMB200_LUTM_LEVELS];  // scale of each level, highest first
  float *lut;         // [64][64] exp table, entry [q][p] = 2^(alpha2*p + beta2*q - s)
  double *sumN;       // [1] A * sum(D)   (nem.py:62 summed over e)
  int *counters;      // [3][64], zero at rest, slot = score workspace (nemb200_cache::ctr_ws): self-resetting completion counters of the
                      // score kernel's fused finalize, of the pair path's list-mode launch, and that path's list length
  nemb200_cache *cache;  // mutable launcher state (see above); owned by the handle
};

void nemb200_set_error(const char *fmt, ...);
#define NEMB_CHECK_CUDA(expr)                                                          \
  do {                                                                                 \
    cudaError_t _e = (expr);                                                           \
    if (_e != cudaSuccess) {                                                           \
      nemb200_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return -2;                                                                       \
    }                                                                                  \
  } while (0)
#define NEMB_REQUIRE(cond, ...)        \
  do {                                 \
    if (!(cond)) {                     \
      nemb200_set_error(__VA_ARGS__);  \
      return -1;                       \
    }                                  \
  } while (0)

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// internal launchers (defined in the per-kernel .cu files)
int launch_closure(const uint32_t *adj, int n, int S, uint32_t *out, cudaStream_t st);
int launch_transitive_reduction(const uint32_t *adj, int n, int S, uint32_t *out, cudaStream_t st);
int launch_order_to_graph(const int32_t *orders, int n, int S, int rule, const uint32_t *cand, int use_nem,
                          uint32_t *out, cudaStream_t st);
size_t score_witness_workspace_bytes(int n);
int launch_score_witness(const nemb200_dataset *h, const uint32_t *phi, int n, double *scores, void *ws,
                         cudaStream_t st);
size_t score_tc_workspace_bytes(const nemb200_dataset *h, int n, bool with_pair = false);
bool score_tc_has_mix(const nemb200_dataset *h);  // NEMB200_SCORE_TC_LUT_MIX available for this dataset
int launch_score_tc(const nemb200_dataset *h, const uint32_t *phi, int n, double *scores, void *ws, size_t ws_bytes,
                    int impl, cudaStream_t st, int leave_sms_free = 0, int close = 0, int pdl = 0);

size_t fit_pipeline_workspace_bytes(const nemb200_dataset *h, int n);
int launch_fit_pipeline(const nemb200_dataset *h, int n, uint32_t *prop_bits, double *prop_scores, void *ws, double w0,
                        int use_nem, cudaStream_t st);

bool mcmc_fused_supported(const nemb200_dataset *h);
bool mcmc_fused_preferred(const nemb200_dataset *h, int n_chains);
int launch_mcmc_fused(const nemb200_dataset *h, int n, int64_t n_iters, int64_t iter0, int64_t burn_in, const double *d_gamma,
                      unsigned long long sp, uint64_t seed, uint32_t chain_offset, int rule, const uint32_t *d_cand_bits,
                      int use_nem, int resume, int32_t *d_orders, double *d_curr_scores, uint32_t *d_curr_bits,
                      double *d_best_scores, int32_t *d_best_orders, uint32_t *d_best_bits, int64_t *d_accept,
                      uint32_t *d_edge_counts, double *d_score_trace, cudaStream_t st);

// ---- device helpers ------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Philox4x32-10 (Salmon et al., SC'11); the oracle uses the same function.
__device__ __forceinline__ void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t h0 = __umulhi(0xD2511F53u, c[0]), l0 = 0xD2511F53u * c[0];
    uint32_t h1 = __umulhi(0xCD9E8D57u, c[2]), l1 = 0xCD9E8D57u * c[2];
    uint32_t n0 = h1 ^ c[1] ^ k0, n2 = h0 ^ c[3] ^ k1;
    c[0] = n0; c[1] = l1; c[2] = n2; c[3] = l0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}

// Boolean Warshall on a graph held by one warp: lane l owns rows l, l+32, .. (TW rows of TW words).
template <int TW>
__device__ __forceinline__ void warp_closure(uint32_t (&r)[TW][TW], int S) {
#pragma unroll
  for (int slot = 0; slot < TW; ++slot) {
    for (int kk = 0; kk < 32; ++kk) {
      int k = slot * 32 + kk;
      if (k >= S) break;  // warp-uniform
      uint32_t rk[TW];
#pragma unroll
      for (int w = 0; w < TW; ++w) rk[w] = __shfl_sync(0xffffffffu, r[slot][w], kk);
#pragma unroll
      for (int t = 0; t < TW; ++t) {
        if ((r[t][slot] >> kk) & 1u) {
#pragma unroll
          for (int w = 0; w < TW; ++w) r[t][w] |= rk[w];
        }
      }
    }
  }
}
// Graph of an order held by one warp (nem_order_mcmc.py:54-77 + 210-221): labels[t] is the S-gene at
// position t*32+lane (-1 past the end).  row(label at position p) = labels at positions > p, i.e. the
// reference's `dag = W.T` orientation of "every earlier node is a permissible parent" (:63-64);
// MASKED rule ANDs the candidate-edge mask; use_nem closes the result (create_nem, :216-221).
// `rows` is S*TW words of warp-private shared memory; on return lane l holds rows l, l+32, ..
template <int TW>
__device__ __forceinline__ void warp_graph_from_labels(const int (&labels)[TW], int S, int rule,
                                                       const uint32_t *__restrict__ cand, int use_nem, uint32_t *rows,
                                                       uint32_t (&r)[TW][TW]) {
  const int lane = threadIdx.x & 31;
  uint32_t carry[TW];
#pragma unroll
  for (int w = 0; w < TW; ++w) carry[w] = 0u;
#pragma unroll
  for (int t = 0; t < TW; ++t) {
    const int label = labels[t];
    uint32_t m[TW];
#pragma unroll
    for (int w = 0; w < TW; ++w) m[w] = (label >= 0 && (label >> 5) == w) ? (1u << (label & 31)) : 0u;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {  // inclusive OR-scan over lanes = labels at positions <= p
#pragma unroll
      for (int w = 0; w < TW; ++w) {
        uint32_t v = __shfl_up_sync(0xffffffffu, m[w], o);
        if (lane >= o) m[w] |= v;
      }
    }
#pragma unroll
    for (int w = 0; w < TW; ++w) {
      m[w] |= carry[w];
      carry[w] = __shfl_sync(0xffffffffu, m[w], 31);
    }
    if (label >= 0) {
#pragma unroll
      for (int w = 0; w < TW; ++w) {
        int hi = S - w * 32;
        uint32_t full = hi >= 32 ? 0xffffffffu : (hi <= 0 ? 0u : ((1u << hi) - 1u));
        uint32_t v = ~m[w] & full;
        if (rule == NEMB200_RULE_MASKED) v &= cand[label * TW + w];
        rows[label * TW + w] = v;
      }
    }
  }
  __syncwarp();
#pragma unroll
  for (int t = 0; t < TW; ++t) {
    int row = t * 32 + lane;
#pragma unroll
    for (int w = 0; w < TW; ++w) r[t][w] = row < S ? rows[row * TW + w] : 0u;
  }
  if (use_nem && rule != NEMB200_RULE_FULL) warp_closure<TW>(r, S);  // FULL is already transitively closed
}
#endif

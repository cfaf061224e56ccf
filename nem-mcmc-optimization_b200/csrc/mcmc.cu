// Device-resident order-MCMC: replaces the loop of NEMOrderMCMC.method (nem_order_mcmc.py:272-302).
// Per iteration three launches, all captured into a CUDA graph so the host never synchronises:
//   propose (get_new_order :231-255 + order->graph :54-77,210-221)  ->  batched score  ->
//   accept (accepting :224-229 + best-so-far bookkeeping :290-302 + posterior edge counts).
// Randomness: Philox4x32-10, key = seed, counter = (iteration lo, iteration hi, global chain id, 0);
// word 0 -> swap vs adjacent, words 1,2 -> positions, word 3 -> accept uniform.
#include <math.h>
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int kWarpsPerBlock = 4;

struct McmcWs {
  long long *iter_a;      // iteration index read by propose (written by accept)
  long long *iter_b;      // iteration index read by accept (written by propose)
  long long *since;       // [n] iteration at which the current graph became current (edge-count bookkeeping)
  int32_t *prop_orders;   // [n][S]
  uint32_t *prop_bits;    // [n][S][W]
  double *prop_scores;    // [n]
  void *score_ws;
  size_t score_ws_bytes;
  void *fit_ws;           // NEMB200_RULE_FIT only: scratch of the per-order edge-weight fit (cell.cu: launch_fit_pipeline)
};

size_t carve(const nemb200_dataset *h, int n, void *base, McmcWs *w, bool fit = false) {
  size_t off = 0;
  uint8_t *b = (uint8_t *)base;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
  size_t o_iter = take(16), o_since = take(8 * (size_t)n), o_po = take(4 * (size_t)n * h->S),
         o_pb = take(4 * (size_t)n * h->S * h->W), o_ps = take(8 * (size_t)n);
  size_t sws = score_tc_workspace_bytes(h, n, /*with_pair=*/true), wws = score_witness_workspace_bytes(n);  // chains are few: room for impl = PAIR
  if (wws > sws) sws = wws;
  size_t o_sw = take(sws);
  size_t o_fit = fit ? take(fit_pipeline_workspace_bytes(h, n)) : 0;
  if (w) {
    w->fit_ws = fit ? b + o_fit : nullptr;
    w->iter_a = (long long *)(b + o_iter); w->iter_b = w->iter_a + 1;
    w->since = (long long *)(b + o_since); w->prop_orders = (int32_t *)(b + o_po);
    w->prop_bits = (uint32_t *)(b + o_pb); w->prop_scores = (double *)(b + o_ps);
    w->score_ws = b + o_sw; w->score_ws_bytes = sws;
  }
  return off;
}

__global__ void k_mcmc_init(long long *iter_a, long long *iter_b, long long *since, int n, long long iter0) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) { *iter_a = iter0; *iter_b = iter0; }
  if (i < n) since[i] = iter0;
}

template <int TW>
__global__ void __launch_bounds__(kWarpsPerBlock * 32) k_mcmc_propose(
    const int32_t *__restrict__ orders, int n, int S, const long long *iter_a, long long *iter_b, uint64_t seed,
    uint32_t chain_offset, unsigned long long swap_prob_u33, int rule, const uint32_t *__restrict__ cand, int use_nem,
    int init, int32_t *__restrict__ prop_orders, uint32_t *__restrict__ prop_bits) {
  __shared__ uint32_t sh[kWarpsPerBlock][128 * 4];
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = blockIdx.x * kWarpsPerBlock + wib;
  const long long t = *iter_a;
  if (blockIdx.x == 0 && threadIdx.x == 0) *iter_b = t;
  if (c >= n) return;
  int si = -1, sj = -1;
  if (!init) {
    uint32_t ctr[4] = {(uint32_t)t, (uint32_t)((unsigned long long)t >> 32), chain_offset + (uint32_t)c, 0u};
    philox4x32_10(ctr, (uint32_t)seed, (uint32_t)(seed >> 32));
    if ((unsigned long long)ctr[0] < swap_prob_u33) {  // swap two distinct uniform positions (:238-244)
      si = (int)__umulhi(ctr[1], (uint32_t)S);
      sj = (int)__umulhi(ctr[2], (uint32_t)(S - 1));
      if (sj >= si) ++sj;
    } else {                                            // swap adjacent positions (:245-251)
      si = (int)__umulhi(ctr[1], (uint32_t)(S - 1));
      sj = si + 1;
    }
  }
  const int32_t *ord = orders + (size_t)c * S;
  int labels[TW];
#pragma unroll
  for (int tt = 0; tt < TW; ++tt) {
    int p = tt * 32 + lane;
    int src = p == si ? sj : (p == sj ? si : p);
    labels[tt] = p < S ? ord[src] : -1;
    if (p < S) prop_orders[(size_t)c * S + p] = labels[tt];
  }
  uint32_t r[TW][TW];
  warp_graph_from_labels<TW>(labels, S, rule, cand, use_nem, sh[wib], r);
  uint32_t *dst = prop_bits + (size_t)c * S * TW;
#pragma unroll
  for (int tt = 0; tt < TW; ++tt) {
    int row = tt * 32 + lane;
    if (row < S) {
#pragma unroll
      for (int w = 0; w < TW; ++w) dst[row * TW + w] = r[tt][w];
    }
  }
}

// counts[a*S+b] += k for every edge a->b of `bits`; eight independent read-modify-writes per lane are in
// flight at a time so the warp is bandwidth- rather than latency-bound
__device__ __forceinline__ void warp_add_edges(uint32_t *counts, const uint32_t *bits, int S, int W, uint32_t k, int lane) {
  const int cells = S * S;
  for (int base = 0; base < cells; base += 32 * 8) {
    uint32_t v[8];
    bool on[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int cell = base + u * 32 + lane;
      on[u] = false;
      if (cell < cells) {
        const int a = cell / S, b = cell - a * S;
        on[u] = (bits[a * W + (b >> 5)] >> (b & 31)) & 1u;
      }
      v[u] = on[u] ? counts[cell] : 0u;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u)
      if (on[u]) counts[base + u * 32 + lane] = v[u] + k;
  }
}

constexpr int kAcceptThreads = 128;

// one CTA per chain: every thread evaluates the (cheap) accept test, then the whole CTA applies it
__global__ void __launch_bounds__(kAcceptThreads) k_mcmc_accept(
    int n, int S, int W, long long *iter_a, const long long *iter_b, long long iter0, long long burn_in, uint64_t seed,
    uint32_t chain_offset, const double *__restrict__ gamma, int init, const int32_t *__restrict__ prop_orders,
    const uint32_t *__restrict__ prop_bits, const double *__restrict__ prop_scores, int32_t *__restrict__ orders,
    double *__restrict__ curr_scores, uint32_t *__restrict__ curr_bits, double *__restrict__ best_scores,
    int32_t *__restrict__ best_orders, uint32_t *__restrict__ best_bits, long long *__restrict__ accept,
    uint32_t *__restrict__ edge_counts, long long *__restrict__ since, double *__restrict__ trace) {
  const int c = blockIdx.x, tid = threadIdx.x;
  const long long t = *iter_b;
  if (!init && c == 0 && tid == 0) *iter_a = t + 1;
  const double ll = prop_scores[c];
  const double curr = curr_scores[c], best = best_scores[c];
  const long long since_c = since[c];
  bool acc, better;
  if (init) {
    acc = true; better = true;
  } else {
    uint32_t ctr[4] = {(uint32_t)t, (uint32_t)((unsigned long long)t >> 32), chain_offset + (uint32_t)c, 0u};
    philox4x32_10(ctr, (uint32_t)seed, (uint32_t)(seed >> 32));
    const double u = ((double)ctr[3] + 0.5) * (1.0 / 4294967296.0);
    acc = u < exp(gamma[c] * (ll - curr));            // accepting(), :224-229
    better = acc && ll > best;                        // :290-292
    if (trace != nullptr && tid == 0) {  // [iteration][0: proposed score, 1: current score after the decision][chain]
      trace[((size_t)(t - iter0) * 2 + 0) * n + c] = ll;
      trace[((size_t)(t - iter0) * 2 + 1) * n + c] = acc ? ll : curr;
    }
  }
  if (!acc) return;
  const int nw = S * W;
  __shared__ uint32_t old_bits[128 * 4];
  if (!init && edge_counts != nullptr) {
    // the graph being replaced was current for iterations [since, t-1]; count the post-burn-in ones
    const long long from = since_c > burn_in ? since_c : burn_in;
    if (t > from) {
      for (int i = tid; i < nw; i += kAcceptThreads) old_bits[i] = curr_bits[(size_t)c * nw + i];
      __syncthreads();
      const uint32_t k = (uint32_t)(t - from);
      uint32_t *counts = edge_counts + (size_t)c * S * S;
      for (int cell = tid; cell < S * S; cell += kAcceptThreads) {
        const int a = cell / S, b = cell - a * S;
        if ((old_bits[a * W + (b >> 5)] >> (b & 31)) & 1u) counts[cell] += k;
      }
    }
  }
  __syncthreads();  // every thread has read curr/best/since of this chain before they are overwritten
  for (int i = tid; i < S; i += kAcceptThreads) {
    const int v = prop_orders[(size_t)c * S + i];
    orders[(size_t)c * S + i] = v;
    if (better) best_orders[(size_t)c * S + i] = v;
  }
  for (int i = tid; i < nw; i += kAcceptThreads) {
    const uint32_t v = prop_bits[(size_t)c * nw + i];
    curr_bits[(size_t)c * nw + i] = v;
    if (better) best_bits[(size_t)c * nw + i] = v;
  }
  if (tid == 0) {
    curr_scores[c] = ll;
    if (better) best_scores[c] = ll;
    if (!init) {
      accept[c] += 1;
      if (edge_counts != nullptr) since[c] = t;
    }
  }
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32) k_mcmc_flush(int n, int S, int W, long long t_end, long long burn_in,
                                                                    const uint32_t *__restrict__ curr_bits,
                                                                    uint32_t *__restrict__ edge_counts,
                                                                    long long *__restrict__ since) {
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = blockIdx.x * kWarpsPerBlock + wib;
  if (c >= n) return;
  long long from = since[c] > burn_in ? since[c] : burn_in;
  if (t_end > from)
    warp_add_edges(edge_counts + (size_t)c * S * S, curr_bits + (size_t)c * S * W, S, W, (uint32_t)(t_end - from), lane);
  __syncwarp();
  if (lane == 0) since[c] = t_end;
}

// ---- replica exchange (nem_order_mcmc.py:326-340): one warp per neighbouring pair ----------------
__global__ void __launch_bounds__(kWarpsPerBlock * 32) k_replica_exchange(
    int n_chains, int ladder, int phase, long long round, uint64_t seed, int S, int W, const double *__restrict__ gamma,
    int32_t *__restrict__ orders, double *__restrict__ curr_scores, uint32_t *__restrict__ curr_bits,
    long long *__restrict__ n_swaps, double *__restrict__ best_scores, int32_t *__restrict__ best_orders,
    uint32_t *__restrict__ best_bits) {
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pairs_per_ladder = ladder / 2;  // upper bound; odd phases may have one fewer
  const int pid = blockIdx.x * kWarpsPerBlock + wib;
  const int lad = pid / pairs_per_ladder, k = pid - lad * pairs_per_ladder;
  const int i = lad * ladder + 2 * k + (phase ? 1 : 0), j = i + 1;
  if (lad * ladder >= n_chains || j >= (lad + 1) * ladder || j >= n_chains) return;
  uint32_t ctr[4] = {(uint32_t)round, (uint32_t)((unsigned long long)round >> 32), (uint32_t)pid, 1u};
  philox4x32_10(ctr, (uint32_t)seed, (uint32_t)(seed >> 32));
  const double u = ((double)ctr[3] + 0.5) * (1.0 / 4294967296.0);
  // default: the chains' CURRENT scores (standard parallel tempering); faithful mode (best_scores != NULL): the
  // reference's scores[i] = replicas[i].best_score (:321), and whole replicas change places (:335), best state included
  const double si = best_scores ? best_scores[i] : curr_scores[i], sj = best_scores ? best_scores[j] : curr_scores[j];
  const double delta = gamma[i] * sj - gamma[i] * si + gamma[j] * si - gamma[j] * sj;  // :333
  if (!(u < exp(-delta))) return;
  if (best_scores != nullptr) {
    for (int q = lane; q < S; q += 32) {
      int a = best_orders[(size_t)i * S + q], b = best_orders[(size_t)j * S + q];
      best_orders[(size_t)i * S + q] = b; best_orders[(size_t)j * S + q] = a;
    }
    for (int q = lane; q < S * W; q += 32) {
      uint32_t a = best_bits[(size_t)i * S * W + q], b = best_bits[(size_t)j * S * W + q];
      best_bits[(size_t)i * S * W + q] = b; best_bits[(size_t)j * S * W + q] = a;
    }
  }
  for (int q = lane; q < S; q += 32) {
    int a = orders[(size_t)i * S + q], b = orders[(size_t)j * S + q];
    orders[(size_t)i * S + q] = b; orders[(size_t)j * S + q] = a;
  }
  for (int q = lane; q < S * W; q += 32) {
    uint32_t a = curr_bits[(size_t)i * S * W + q], b = curr_bits[(size_t)j * S * W + q];
    curr_bits[(size_t)i * S * W + q] = b; curr_bits[(size_t)j * S * W + q] = a;
  }
  __syncwarp();
  if (lane == 0) {
    const double ci = curr_scores[i], cj = curr_scores[j];
    curr_scores[i] = cj; curr_scores[j] = ci;
    if (best_scores != nullptr) { best_scores[i] = sj; best_scores[j] = si; }
    if (n_swaps) atomicAdd((unsigned long long *)n_swaps, 1ull);
  }
}

__global__ void k_edge_reduce(const uint32_t *__restrict__ counts, int n, int cells, long long *__restrict__ out) {
  int cell = blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= cells) return;
  long long s = 0;
  for (int c = 0; c < n; ++c) s += counts[(size_t)c * cells + cell];
  out[cell] = s;
}

}  // namespace

// Large runs are split into two chain groups whose iterations are captured as two parallel branches of one
// CUDA graph: while the score kernel of one group owns the SMs, the (small, latency-bound) propose / accept /
// finalize kernels of the other group run beside it, so the step time approaches the score time alone.
// Chains keep their global Philox id, so the result does not depend on the grouping.
constexpr int kMinChainsForTwoGroups = 512;
constexpr int kSmsLeftFree = 20;  // per score launch in two-group mode (measured: 108 vs 126 us/iteration at 1024 chains)
static int first_group_size(const nemb200_dataset *h, int n) {
  if (n < kMinChainsForTwoGroups || mcmc_fused_preferred(h, n)) return n;
  const int G = h->lut_valid ? h->G_lut : (!h->fixed_valid && h->lutm_levels > 0 ? h->G_lutm : (h->G > 0 ? h->G : 1));
  return ((n / 2 + G - 1) / G) * G;  // whole MMA N tiles in the first group
}

extern "C" size_t nemb200_mcmc_workspace_bytes_rule(const nemb200_dataset *h, int n_chains, int rule) {
  if (!h || n_chains < 0) return 0;
  if (rule != NEMB200_RULE_FIT) return nemb200_mcmc_workspace_bytes(h, n_chains);
  return carve(h, n_chains, nullptr, nullptr, true);
}

extern "C" size_t nemb200_mcmc_workspace_bytes(const nemb200_dataset *h, int n_chains) {
  if (!h || n_chains < 0) return 0;
  const int nA = first_group_size(h, n_chains);
  size_t one = carve(h, n_chains, nullptr, nullptr);
  size_t two = nA < n_chains ? carve(h, nA, nullptr, nullptr) + carve(h, n_chains - nA, nullptr, nullptr) : 0;
  return one > two ? one : two;
}

// AUTO of the two sweep entry points (nemb200_score_batch, nemb200_closure_score_batch): the pair-merged path where it is
// the fastest (measured at E = 10 000, 1024 graphs: S = 40 one-site 78 us / pair 83, S = 44 87 / 83, 47 87 / 83, 50 92 / 85),
// else the dataset's best one-site kernel.  nemb200_mcmc_run's AUTO does not come through here.
static bool sweep_auto_is_pair(const nemb200_dataset *h) { return h->pair_ok && h->S >= 44 && h->S <= 50; }
static int sweep_auto(const nemb200_dataset *h, int impl) {
  if (impl != NEMB200_SCORE_AUTO || h->S < 8) return impl;
  return sweep_auto_is_pair(h) ? NEMB200_SCORE_TC_PAIR : impl;
}

// a launch outside stream capture may carry the programmatic-dependent-launch attribute (score_tc.cu: launch_ex)
static int pdl_ok(cudaStream_t st) {
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  return cudaStreamIsCapturing(st, &cs) == cudaSuccess && cs == cudaStreamCaptureStatusNone ? 1 : 0;
}

static int score_dispatch(const nemb200_dataset *h, const uint32_t *phi, int n, double *scores, void *ws, size_t ws_bytes,
                          int impl, cudaStream_t st, int leave_sms_free = 0, int pdl = 0) {
  if (impl == NEMB200_SCORE_WITNESS || (impl == NEMB200_SCORE_AUTO && h->S < 8)) {
    NEMB_REQUIRE(ws_bytes >= score_witness_workspace_bytes(n), "score workspace too small");
    return launch_score_witness(h, phi, n, scores, ws, st);
  }
  return launch_score_tc(h, phi, n, scores, ws, ws_bytes, impl, st, leave_sms_free, 0, pdl);
}

extern "C" size_t nemb200_score_workspace_bytes_impl(const nemb200_dataset *h, int n_graphs, int impl) {
  if (!h || n_graphs < 0) return 0;
  size_t a = score_tc_workspace_bytes(h, n_graphs, impl == NEMB200_SCORE_TC_PAIR || (impl == NEMB200_SCORE_AUTO && sweep_auto_is_pair(h))),
         b = score_witness_workspace_bytes(n_graphs);
  return a > b ? a : b;
}
extern "C" size_t nemb200_score_workspace_bytes(const nemb200_dataset *h, int n_graphs) {
  return nemb200_score_workspace_bytes_impl(h, n_graphs, NEMB200_SCORE_AUTO);
}

extern "C" int nemb200_score_batch(const nemb200_dataset *h, const uint32_t *d_phi_bits, int n_graphs, double *d_scores,
                                   void *d_workspace, size_t workspace_bytes, int impl, void *stream) {
  NEMB_REQUIRE(h != nullptr && n_graphs >= 0, "bad handle / n_graphs");
  NEMB_REQUIRE(n_graphs == 0 || (d_phi_bits && d_scores && d_workspace), "null device pointer");
  const int ready = (impl & NEMB200_SCORE_INPUTS_READY) ? 1 : 0;
  impl = sweep_auto(h, impl & ~NEMB200_SCORE_INPUTS_READY);
  return score_dispatch(h, d_phi_bits, n_graphs, d_scores, d_workspace, workspace_bytes, impl, (cudaStream_t)stream, 0,
                        pdl_ok((cudaStream_t)stream) * (1 + ready));
}

extern "C" int nemb200_closure_score_batch(const nemb200_dataset *h, const uint32_t *d_adj_bits, int n_graphs, double *d_scores,
                                           void *d_workspace, size_t workspace_bytes, int impl, void *stream) {
  NEMB_REQUIRE(h != nullptr && n_graphs >= 0, "bad handle / n_graphs");
  NEMB_REQUIRE(n_graphs == 0 || (d_adj_bits && d_scores && d_workspace), "null device pointer");
  const int ready = (impl & NEMB200_SCORE_INPUTS_READY) ? 1 : 0;
  impl &= ~NEMB200_SCORE_INPUTS_READY;
  NEMB_REQUIRE(impl != NEMB200_SCORE_WITNESS && h->S >= 8, "closure fused into the score needs a tensor-core score kernel (S >= 8)");
  impl = sweep_auto(h, impl);
  return launch_score_tc(h, d_adj_bits, n_graphs, d_scores, d_workspace, workspace_bytes, impl, (cudaStream_t)stream, 0, 1,
                         pdl_ok((cudaStream_t)stream) * (1 + ready));
}

extern "C" int nemb200_closure_u32(const uint32_t *d_adj_bits, int n_graphs, int S, uint32_t *d_out_bits, void *stream) {
  NEMB_REQUIRE(S >= 1 && S <= NEMB200_MAX_S, "S must be in [1,%d] (got %d)", NEMB200_MAX_S, S);
  NEMB_REQUIRE(n_graphs >= 0 && (n_graphs == 0 || (d_adj_bits && d_out_bits)), "bad arguments");
  return launch_closure(d_adj_bits, n_graphs, S, d_out_bits, (cudaStream_t)stream);
}

extern "C" int nemb200_transitive_reduction_u32(const uint32_t *d_adj_bits, int n_graphs, int S, uint32_t *d_out_bits,
                                                void *stream) {
  NEMB_REQUIRE(S >= 1 && S <= NEMB200_MAX_S, "S must be in [1,%d] (got %d)", NEMB200_MAX_S, S);
  NEMB_REQUIRE(n_graphs >= 0 && (n_graphs == 0 || (d_adj_bits && d_out_bits)), "bad arguments");
  return launch_transitive_reduction(d_adj_bits, n_graphs, S, d_out_bits, (cudaStream_t)stream);
}

extern "C" int nemb200_order_to_graph(const int32_t *d_orders, int n_chains, int S, int rule, const uint32_t *d_cand_bits,
                                      int use_nem, uint32_t *d_out_bits, void *stream) {
  NEMB_REQUIRE(S >= 1 && S <= NEMB200_MAX_S, "S must be in [1,%d] (got %d)", NEMB200_MAX_S, S);
  NEMB_REQUIRE(rule == NEMB200_RULE_FULL || (rule == NEMB200_RULE_MASKED && d_cand_bits), "MASKED rule needs a candidate mask");
  NEMB_REQUIRE(n_chains >= 0 && (n_chains == 0 || (d_orders && d_out_bits)), "bad arguments");
  return launch_order_to_graph(d_orders, n_chains, S, rule, d_cand_bits, use_nem, d_out_bits, (cudaStream_t)stream);
}

extern "C" int nemb200_mcmc_run(const nemb200_dataset *h, int n, int64_t n_iters, int64_t iter0, int64_t burn_in,
                                const double *d_gamma, double swap_prob, uint64_t seed, uint32_t chain_offset, int rule_arg,
                                const uint32_t *d_cand_bits, int use_nem_arg, int resume, int score_impl, int32_t *d_orders,
                                double *d_curr_scores, uint32_t *d_curr_bits, double *d_best_scores, int32_t *d_best_orders,
                                uint32_t *d_best_bits, int64_t *d_accept, uint32_t *d_edge_counts, double *d_score_trace,
                                void *d_workspace, size_t workspace_bytes, void *stream) {
  NEMB_REQUIRE(h != nullptr, "null dataset");
  NEMB_REQUIRE(n >= 0 && n_iters >= 0, "negative sizes");
  if (n == 0) return 0;
  NEMB_REQUIRE(h->S >= 2, "order-MCMC needs S >= 2");
  NEMB_REQUIRE(d_gamma && d_orders && d_curr_scores && d_curr_bits && d_best_scores && d_best_orders && d_best_bits &&
                   d_accept && d_workspace,
               "null device pointer");
  const bool fit = rule_arg == NEMB200_RULE_FIT;
  NEMB_REQUIRE(rule_arg == NEMB200_RULE_FULL || fit || (rule_arg == NEMB200_RULE_MASKED && d_cand_bits), "MASKED rule needs a candidate mask");
  NEMB_REQUIRE(workspace_bytes >= nemb200_mcmc_workspace_bytes_rule(h, n, rule_arg),
               fit ? "mcmc workspace too small for rule FIT: size it with nemb200_mcmc_workspace_bytes_rule" : "mcmc workspace too small");
  NEMB_REQUIRE(((uintptr_t)d_workspace & 255) == 0, "mcmc workspace must be 256-byte aligned");
  NEMB_REQUIRE(swap_prob >= 0.0 && swap_prob <= 1.0, "swap_prob must be in [0,1]");
  cudaStream_t st = (cudaStream_t)stream;
  const int S = h->S, W = h->W;
  const unsigned long long sp = swap_prob >= 1.0 ? 0x100000000ull : (unsigned long long)(swap_prob * 4294967296.0 + 0.5);
  NEMB_REQUIRE(!(fit && score_impl == NEMB200_MCMC_FUSED), "rule FIT runs in the kernel loop, not in the fused persistent kernel");
  if (!fit && (score_impl == NEMB200_MCMC_FUSED || (score_impl == NEMB200_SCORE_AUTO && mcmc_fused_preferred(h, n))))
    return launch_mcmc_fused(h, n, n_iters, iter0, burn_in, d_gamma, sp, seed, chain_offset, rule_arg, d_cand_bits, use_nem_arg,
                             resume, d_orders, d_curr_scores, d_curr_bits, d_best_scores, d_best_orders, d_best_bits,
                             d_accept, d_edge_counts, d_score_trace, st);
  // AUTO of the kernel loop at the column stride 50: the one-site table kernel with a quarter of the sites through ex2.approx
  // (NEMB200_SCORE_TC_LUT_MIX: 94 -> 87 us per iteration of 1024 chains at S=50, E=10k; the pair-merged path does not pay
  // here, its planner would run per group and iteration: 95 us).  Scores then carry a relative error of <= 5e-7 instead
  // of <= 5e-8 - still 200 times inside the 1e-4 the accept step is specified to.
  if (score_impl == NEMB200_SCORE_AUTO && S >= 8 && score_tc_has_mix(h)) score_impl = NEMB200_SCORE_TC_LUT_MIX;
  constexpr int kUnroll = 16;  // iterations per captured CUDA graph
  const int thr = kWarpsPerBlock * 32;
  nemb200_cache *hm = h->cache;  // graph cache + side stream: mutable launcher state, serialised per handle
  std::lock_guard<std::recursive_mutex> lock(hm->mu);

  // ---- chain groups (see first_group_size)
  struct Group { int c0, n; McmcWs w; };
  Group grp[2];
  int n_groups = 1;
  const int nA = first_group_size(h, n);
  if (st != nullptr && nA < n && n_iters >= 2 * kUnroll && !fit) n_groups = 2;
  if (n_groups == 2) {
    grp[0].c0 = 0; grp[0].n = nA; grp[1].c0 = nA; grp[1].n = n - nA;
    size_t off = carve(h, nA, d_workspace, &grp[0].w);
    carve(h, n - nA, (uint8_t *)d_workspace + off, &grp[1].w);
    if (hm->mcmc_side_stream == nullptr) {
      cudaStream_t s2; cudaEvent_t e1, e2;
      NEMB_CHECK_CUDA(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
      hm->mcmc_side_stream = s2;
      NEMB_CHECK_CUDA(cudaEventCreateWithFlags(&e1, cudaEventDisableTiming));
      hm->mcmc_ev_fork = e1;
      NEMB_CHECK_CUDA(cudaEventCreateWithFlags(&e2, cudaEventDisableTiming));
      hm->mcmc_ev_join = e2;
    }
  } else {
    grp[0].c0 = 0; grp[0].n = n;
    carve(h, n, d_workspace, &grp[0].w, fit);
  }

  static const char *sms_env = getenv("NEMB200_SMS_FREE");  // tuning knob (diagnostics)
  const int sms_free = sms_env ? atoi(sms_env) : kSmsLeftFree;
  auto step = [&](const Group &g, int init, cudaStream_t s) -> int {
    const McmcWs &w = g.w;
    const size_t c0 = (size_t)g.c0;
    const int grid = (g.n + kWarpsPerBlock - 1) / kWarpsPerBlock;
    int32_t *orders = d_orders + c0 * S;
    const uint32_t coff = chain_offset + (uint32_t)g.c0;
    // rule FIT: the proposal kernel emits the permissibility graph of the proposed order (FULL rule, not closed -
    // it is closed by construction); the fit pipeline turns it into the thresholded graph and its score
    const int rule = fit ? NEMB200_RULE_FULL : rule_arg, use_nem = fit ? 0 : use_nem_arg;
    switch (W) {
      case 1: k_mcmc_propose<1><<<grid, thr, 0, s>>>(orders, g.n, S, w.iter_a, w.iter_b, seed, coff, sp, rule, d_cand_bits, use_nem, init, w.prop_orders, w.prop_bits); break;
      case 2: k_mcmc_propose<2><<<grid, thr, 0, s>>>(orders, g.n, S, w.iter_a, w.iter_b, seed, coff, sp, rule, d_cand_bits, use_nem, init, w.prop_orders, w.prop_bits); break;
      case 3: k_mcmc_propose<3><<<grid, thr, 0, s>>>(orders, g.n, S, w.iter_a, w.iter_b, seed, coff, sp, rule, d_cand_bits, use_nem, init, w.prop_orders, w.prop_bits); break;
      default: k_mcmc_propose<4><<<grid, thr, 0, s>>>(orders, g.n, S, w.iter_a, w.iter_b, seed, coff, sp, rule, d_cand_bits, use_nem, init, w.prop_orders, w.prop_bits); break;
    }
    NEMB_CHECK_CUDA(cudaGetLastError());
    // (method() scores the start order with the constructor's weights, 1.0 on every permissible pair; every later
    //  iteration resets them to 0.5: nem_order_mcmc.py:47, :57, :276)
    int rc = fit ? launch_fit_pipeline(h, g.n, w.prop_bits, w.prop_scores, w.fit_ws, init ? 1.0 : 0.5, use_nem_arg, s)
                 : score_dispatch(h, w.prop_bits, g.n, w.prop_scores, w.score_ws, w.score_ws_bytes, score_impl, s,
                                  n_groups == 2 ? sms_free : 0);
    if (rc) return rc;
    // `n` (all chains) is the row stride of the score trace; every other array is passed at the group's offset
    k_mcmc_accept<<<g.n, kAcceptThreads, 0, s>>>(
        n, S, W, w.iter_a, w.iter_b, iter0, burn_in, seed, coff, d_gamma + c0, init, w.prop_orders, w.prop_bits,
        w.prop_scores, orders, d_curr_scores + c0, d_curr_bits + c0 * S * W, d_best_scores + c0, d_best_orders + c0 * S,
        d_best_bits + c0 * S * W, (long long *)d_accept + c0, d_edge_counts ? d_edge_counts + c0 * S * S : nullptr, w.since,
        d_score_trace ? d_score_trace + c0 : nullptr);
    NEMB_CHECK_CUDA(cudaGetLastError());
    return 0;
  };

  for (int gi = 0; gi < n_groups; ++gi) {
    k_mcmc_init<<<(grp[gi].n + 255) / 256, 256, 0, st>>>(grp[gi].w.iter_a, grp[gi].w.iter_b, grp[gi].w.since, grp[gi].n, iter0);
    NEMB_CHECK_CUDA(cudaGetLastError());
    if (!resume) {
      int rc = step(grp[gi], 1, st);  // method(): score of the start order (:258), best = current
      if (rc) return rc;
    }
  }
  // capture kUnroll iterations into a CUDA graph (no host sync per iteration); the legacy NULL stream
  // cannot be captured, so it falls back to direct launches
  int64_t done = 0;
  if (st != nullptr && n_iters >= 2 * kUnroll) {
    // the captured body depends only on these arguments (iteration indices live in device memory), so the
    // instantiated graph is cached in the handle and re-used by later calls with the same buffers
    unsigned long long key = 1469598103934665603ull;
    auto mix = [&](unsigned long long v) { key = (key ^ v) * 1099511628211ull; };
    mix((unsigned long long)n); mix(17ull + (unsigned long long)n_groups); mix((unsigned long long)burn_in); mix((unsigned long long)(uintptr_t)d_gamma);
    mix(sp); mix(seed); mix(chain_offset); mix((unsigned long long)rule_arg); mix((unsigned long long)(uintptr_t)d_cand_bits);
    mix((unsigned long long)use_nem_arg); mix((unsigned long long)score_impl); mix((unsigned long long)(uintptr_t)d_orders);
    mix((unsigned long long)(uintptr_t)d_curr_scores); mix((unsigned long long)(uintptr_t)d_curr_bits);
    mix((unsigned long long)(uintptr_t)d_best_scores); mix((unsigned long long)(uintptr_t)d_best_orders);
    mix((unsigned long long)(uintptr_t)d_best_bits); mix((unsigned long long)(uintptr_t)d_accept);
    mix((unsigned long long)(uintptr_t)d_edge_counts); mix((unsigned long long)(uintptr_t)d_score_trace);
    mix((unsigned long long)(uintptr_t)d_workspace); mix((unsigned long long)(uintptr_t)st);
    mix((unsigned long long)iter0 * (d_score_trace != nullptr));  // the trace offset is baked into the kernels
    cudaGraphExec_t exec = (cudaGraphExec_t)hm->mcmc_graph_exec;
    if (exec == nullptr || hm->mcmc_graph_key != key) {
      if (exec) { cudaGraphExecDestroy(exec); hm->mcmc_graph_exec = nullptr; exec = nullptr; }
      cudaGraph_t graph = nullptr;
      cudaStream_t s2 = (cudaStream_t)hm->mcmc_side_stream;
      NEMB_CHECK_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
      int rc = 0;
      cudaError_t fe = cudaSuccess;
      if (n_groups == 2) {  // fork: the side stream joins the capture
        fe = cudaEventRecord((cudaEvent_t)hm->mcmc_ev_fork, st);
        if (fe == cudaSuccess) fe = cudaStreamWaitEvent(s2, (cudaEvent_t)hm->mcmc_ev_fork, 0);
      }
      for (int u = 0; u < kUnroll && rc == 0 && fe == cudaSuccess; ++u) {
        rc = step(grp[0], 0, st);
        if (rc == 0 && n_groups == 2) rc = step(grp[1], 0, s2);
      }
      if (n_groups == 2 && fe == cudaSuccess) {  // join
        fe = cudaEventRecord((cudaEvent_t)hm->mcmc_ev_join, s2);
        if (fe == cudaSuccess) fe = cudaStreamWaitEvent(st, (cudaEvent_t)hm->mcmc_ev_join, 0);
      }
      cudaError_t ce = cudaStreamEndCapture(st, &graph);
      if (rc != 0 || ce != cudaSuccess || fe != cudaSuccess) {
        if (graph) cudaGraphDestroy(graph);
        if (rc == 0) nemb200_set_error("CUDA graph capture failed: %s", cudaGetErrorString(ce != cudaSuccess ? ce : fe));
        return rc ? rc : -2;
      }
      cudaError_t ie = cudaGraphInstantiate(&exec, graph, 0);
      cudaGraphDestroy(graph);
      if (ie != cudaSuccess) { nemb200_set_error("cudaGraphInstantiate failed: %s", cudaGetErrorString(ie)); return -2; }
      hm->mcmc_graph_exec = exec;
      hm->mcmc_graph_key = key;
    }
    for (; done + kUnroll <= n_iters; done += kUnroll) {
      cudaError_t le = cudaGraphLaunch(exec, st);
      if (le != cudaSuccess) {
        nemb200_set_error("cudaGraphLaunch failed: %s", cudaGetErrorString(le));
        return -2;
      }
    }
  }
  for (; done < n_iters; ++done) {
    for (int gi = 0; gi < n_groups; ++gi) {
      int rc = step(grp[gi], 0, st);
      if (rc) return rc;
    }
  }
  if (d_edge_counts != nullptr) {
    for (int gi = 0; gi < n_groups; ++gi) {
      const size_t c0 = (size_t)grp[gi].c0;
      k_mcmc_flush<<<(grp[gi].n + kWarpsPerBlock - 1) / kWarpsPerBlock, thr, 0, st>>>(
          grp[gi].n, S, W, iter0 + n_iters, burn_in, d_curr_bits + c0 * S * W, d_edge_counts + c0 * S * S, grp[gi].w.since);
      NEMB_CHECK_CUDA(cudaGetLastError());
    }
  }
  return 0;
}

static int replica_exchange_launch(int n_chains, int ladder, int phase, int64_t round, uint64_t seed, int S,
                                   const double *d_gamma, int32_t *d_orders, double *d_curr_scores, uint32_t *d_curr_bits,
                                   int64_t *d_n_swaps, double *d_best_scores, int32_t *d_best_orders,
                                   uint32_t *d_best_bits, void *stream) {
  NEMB_REQUIRE(n_chains >= 0 && ladder >= 2 && S >= 1 && S <= NEMB200_MAX_S, "bad arguments");
  if (n_chains < 2) return 0;
  NEMB_REQUIRE(d_gamma && d_orders && d_curr_scores && d_curr_bits, "null device pointer");
  int ladders = (n_chains + ladder - 1) / ladder, pairs = ladders * (ladder / 2);
  k_replica_exchange<<<(pairs + kWarpsPerBlock - 1) / kWarpsPerBlock, kWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(
      n_chains, ladder, phase & 1, round, seed, S, (S + 31) / 32, d_gamma, d_orders, d_curr_scores, d_curr_bits,
      (long long *)d_n_swaps, d_best_scores, d_best_orders, d_best_bits);
  NEMB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int nemb200_replica_exchange(int n_chains, int ladder, int phase, int64_t round, uint64_t seed, int S,
                                        const double *d_gamma, int32_t *d_orders, double *d_curr_scores,
                                        uint32_t *d_curr_bits, int64_t *d_n_swaps, void *stream) {
  return replica_exchange_launch(n_chains, ladder, phase, round, seed, S, d_gamma, d_orders, d_curr_scores, d_curr_bits,
                                 d_n_swaps, nullptr, nullptr, nullptr, stream);
}

extern "C" int nemb200_replica_exchange_best(int n_chains, int ladder, int phase, int64_t round, uint64_t seed, int S,
                                             const double *d_gamma, int32_t *d_orders, double *d_curr_scores,
                                             uint32_t *d_curr_bits, double *d_best_scores, int32_t *d_best_orders,
                                             uint32_t *d_best_bits, int64_t *d_n_swaps, void *stream) {
  NEMB_REQUIRE(d_best_scores && d_best_orders && d_best_bits, "null device pointer");
  return replica_exchange_launch(n_chains, ladder, phase, round, seed, S, d_gamma, d_orders, d_curr_scores, d_curr_bits,
                                 d_n_swaps, d_best_scores, d_best_orders, d_best_bits, stream);
}

extern "C" int nemb200_edge_counts_reduce(const uint32_t *d_edge_counts, int n_chains, int S, int64_t *d_out, void *stream) {
  NEMB_REQUIRE(d_edge_counts && d_out && n_chains >= 0 && S >= 1 && S <= NEMB200_MAX_S, "bad arguments");
  int cells = S * S;
  k_edge_reduce<<<(cells + 255) / 256, 256, 0, (cudaStream_t)stream>>>(d_edge_counts, n_chains, cells, (long long *)d_out);
  NEMB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

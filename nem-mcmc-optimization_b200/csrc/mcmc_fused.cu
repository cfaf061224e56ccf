// Persistent fused order-MCMC for small problems: ONE launch runs every iteration of every chain.
//
// Same algorithm, draws and bookkeeping as the three-kernel loop of mcmc.cu (NEMOrderMCMC.method,
// nem_order_mcmc.py:272-302), but one CTA owns one chain for the whole run: warp 0 proposes
// (get_new_order :231-255), builds the graph (:54-77, 210-221) and closes it (utils.py:37-54) in registers,
// all threads score it against the shared-memory copy of D (AND + popcount gives the integer counts p, q;
// the same exp table as the tensor-core epilogue gives 2^(alpha2*p + beta2*q - s)), a block reduction gives
// the log-likelihood and thread 0 accepts (:224-229).  No kernel launch, no HBM round trip and no host
// synchronisation per iteration - the tensor-core path pays ~25 us of launches per iteration, which
// dominates at S=10..20 / E<=1000 (BASELINE configs 1 and 2).
#include <math.h>

#include "common.cuh"

namespace {

// threads per CTA (template parameter): two proposer warps + 192 / 256 scorer threads with one E-gene each for
// E <= 192 / 256, else + 512 scorer threads with two E-genes per thread and pass (1024 E-genes per pass: config 2 in one)
// E-genes per thread and pass (template parameter EPT): 4 amortises the site record over four E-genes; with
// E <= 256 one E-gene per thread keeps four times as many threads busy (S=10 x E=100: 2.42 -> 2.09 us/iteration)

struct FusedParams {
  const uint32_t *dcol;   // [E_pad][W]
  const float *lut;       // [64][64]
  const double *sumN;
  int n, S, E;
  long long n_iters, iter0, burn_in;
  const double *gamma;
  unsigned long long swap_prob_u33;
  uint64_t seed;
  uint32_t chain_offset;
  int rule, use_nem, resume;
  const uint32_t *cand;
  int32_t *orders;
  double *curr_scores;
  uint32_t *curr_bits;
  double *best_scores;
  int32_t *best_orders;
  uint32_t *best_bits;
  long long *accept;
  uint32_t *edge_counts;
  double *trace;
  float scale;
  int dcol_in_smem;
};

__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ float log2_split(float x) {
  uint32_t b = __float_as_uint(x);
  int e = (int)(b >> 23) - 127;
  float m = __uint_as_float((b & 0x007FFFFFu) | 0x3F800000u), l;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(m));
  return (float)e + l;
}

// One proposal slot: what a proposer warp prepares and the scorer warps consume.
constexpr int kSlotBytes = 64 * 16 + 64 * 2 * 4 + 64 * 4;  // site records, graph rows, proposed order
constexpr int kFixedSmem = 16384 + 3 * kSlotBytes + 512 + 256 + 256 + 64;

// Proposal of iteration t (or, with is_init, the unswapped start order) built by ONE warp into a slot: Philox draw ->
// swapped order (get_new_order :231-255) -> graph rows (:54-77, 210-221, closure utils.py:37-54) -> per attachment site
// the ancestor mask (column j of I|Phi) and the byte address of its exp-table row.
template <int TW>
__device__ __forceinline__ void build_proposal(const FusedParams &p, int c, long long t, bool is_init, const int *base_order,
                                               uint4 *sSite, uint32_t *sRows, int *sProp, int *sU, uint32_t lutbase) {
  const int lane = threadIdx.x & 31, S = p.S;
  uint32_t ctr[4] = {(uint32_t)t, (uint32_t)((unsigned long long)t >> 32), p.chain_offset + (uint32_t)c, 0u};
  int si = -1, sj = -1;
  if (!is_init) {
    philox4x32_10(ctr, (uint32_t)p.seed, (uint32_t)(p.seed >> 32));
    if ((unsigned long long)ctr[0] < p.swap_prob_u33) {
      si = (int)__umulhi(ctr[1], (uint32_t)S);
      sj = (int)__umulhi(ctr[2], (uint32_t)(S - 1));
      if (sj >= si) ++sj;
    } else {
      si = (int)__umulhi(ctr[1], (uint32_t)(S - 1));
      sj = si + 1;
    }
  }
  int labels[TW];
#pragma unroll
  for (int tt = 0; tt < TW; ++tt) {
    const int pos = tt * 32 + lane;
    const int src = pos == si ? sj : (pos == sj ? si : pos);
    labels[tt] = pos < S ? base_order[src] : -1;
  }
#pragma unroll
  for (int tt = 0; tt < TW; ++tt)
    if (tt * 32 + lane < S) sProp[tt * 32 + lane] = labels[tt];
  uint32_t r[TW][TW];
  warp_graph_from_labels<TW>(labels, S, p.rule, p.cand, p.use_nem, sRows, r);
  __syncwarp();
#pragma unroll
  for (int tt = 0; tt < TW; ++tt) {
    const int row = tt * 32 + lane;
    if (row < S) {
#pragma unroll
      for (int w = 0; w < TW; ++w) sRows[row * TW + w] = r[tt][w];
    }
  }
  if (lane == 0) *sU = (int)ctr[3];  // accept uniform word (unused for the start order)
  __syncwarp();
  for (int j = lane; j < S; j += 32) {
    uint32_t m[2] = {0u, 0u};
    int cnt = 0;
    for (int a = 0; a < S; ++a) {
      const uint32_t bit = ((sRows[a * TW + (j >> 5)] >> (j & 31)) & 1u) | (a == j ? 1u : 0u);
      m[a >> 5] |= bit << (a & 31);
      cnt += (int)bit;
    }
    sSite[j] = make_uint4(m[0], m[1], lutbase + 4u * 64u * (uint32_t)cnt, 0u);
  }
}

// Warps 0 and 1 are proposers, the others score.  While the scorers work on the proposal of iteration t, warp 0
// prepares the proposal of iteration t+1 for the case that t is accepted (base order = the proposal being scored)
// and warp 1 for the case that it is rejected (base = the current order): the Philox draw does not depend on the
// state, so after the accept decision the right slot is simply selected and no iteration ever waits for a proposal.
template <int TW, int EPT, int kThreads>
__global__ void __launch_bounds__(kThreads) k_mcmc_fused(const FusedParams p) {
  extern __shared__ __align__(16) uint8_t smem[];
  float *sLut = reinterpret_cast<float *>(smem);                                   // 16 KB
  uint8_t *slots = smem + 16384;                                                   // 3 proposal slots
  uint32_t *sCurr = reinterpret_cast<uint32_t *>(slots + 3 * kSlotBytes);          // [S][TW] current graph rows
  int *sOrder = reinterpret_cast<int *>(sCurr + 64 * 2);                           // [S] current order
  double *sRed = reinterpret_cast<double *>(sOrder + 64);                          // [32]: one partial sum per warp
  int *sFlag = reinterpret_cast<int *>(sRed + 32);                                 // [0] accept flag, [4+s] uniform of slot s
  uint32_t *sD = reinterpret_cast<uint32_t *>(smem + kFixedSmem);
  auto slot_site = [&](int s) { return reinterpret_cast<uint4 *>(slots + s * kSlotBytes); };
  auto slot_rows = [&](int s) { return reinterpret_cast<uint32_t *>(slots + s * kSlotBytes + 64 * 16); };
  auto slot_prop = [&](int s) { return reinterpret_cast<int *>(slots + s * kSlotBytes + 64 * 16 + 64 * 2 * 4); };

  const int c = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int S = p.S, E = p.E;
  for (int i = tid; i < 4096; i += kThreads) sLut[i] = p.lut[i];
  if (p.dcol_in_smem)
    for (int i = tid; i < E * TW; i += kThreads) sD[i] = p.dcol[i];
  for (int i = tid; i < S; i += kThreads) sOrder[i] = p.orders[(size_t)c * S + i];
  if (p.resume)
    for (int i = tid; i < S * TW; i += kThreads) sCurr[i] = p.curr_bits[(size_t)c * S * TW + i];
  __syncthreads();
  const uint32_t *dsrc = p.dcol_in_smem ? sD : p.dcol;
  const uint32_t lutbase = smem_u32(sLut);
  const float null_term = exp2f(-p.scale);
  const double gamma = p.gamma[c], sumN = *p.sumN;
  double curr = p.resume ? p.curr_scores[c] : 0.0, best = p.resume ? p.best_scores[c] : 0.0;
  long long n_acc = 0, since = p.iter0;
  uint32_t *counts = p.edge_counts ? p.edge_counts + (size_t)c * S * S : nullptr;

  const long long it0 = p.resume ? 0 : -1;  // -1: score the start order first (method(), :258)
  int x = 0, sa = 1, sr = 2;                // slot being scored, slots of the two speculative successors
  if (it0 < p.n_iters && warp == 0)
    build_proposal<TW>(p, c, p.iter0 + (it0 < 0 ? 0 : it0), it0 < 0, sOrder, slot_site(x), slot_rows(x), slot_prop(x), sFlag + 4 + x, lutbase);
  __syncthreads();
  for (long long it = it0; it < p.n_iters; ++it) {
    const bool init = it < 0;
    const long long t = p.iter0 + (init ? 0 : it);
    const uint4 *sSite = slot_site(x);
    if (warp < 2) {
      // ---- proposers: the two possible proposals of iteration it + 1 -------------------------------------
      if (it + 1 < p.n_iters) {
        const int s = warp == 0 ? sa : sr;
        build_proposal<TW>(p, c, p.iter0 + it + 1, false, warp == 0 ? slot_prop(x) : sOrder, slot_site(s), slot_rows(s), slot_prop(s),
                           sFlag + 4 + s, lutbase);
      }
    } else {
      // ---- scorers: every thread takes EPT E-genes per pass ----------------------------------------------
      double sum = 0.0;
      for (int e0 = (tid - 64) * EPT; e0 < E; e0 += (kThreads - 64) * EPT) {
        uint32_t d[EPT][TW];
        float acc[EPT][4];
#pragma unroll
        for (int u = 0; u < EPT; ++u) {
#pragma unroll
          for (int w = 0; w < TW; ++w) d[u][w] = e0 + u < E ? dsrc[(size_t)(e0 + u) * TW + w] : 0u;
          acc[u][0] = null_term; acc[u][1] = 0.f; acc[u][2] = 0.f; acc[u][3] = 0.f;
        }
        for (int j0 = 0; j0 < S; j0 += 4) {
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            const int j = j0 + jj;
            if (j < S) {
              const uint4 site = sSite[j];
#pragma unroll
              for (int u = 0; u < EPT; ++u) {
                int pc = __popc(site.x & d[u][0]);
                if (TW > 1) pc += __popc(site.y & d[u][1]);
                // entry [q][p] with q = cnt - p: byte offset 4*(64*cnt - 63*p)
                acc[u][jj] += lds_f32(site.z - 252u * (uint32_t)pc);
              }
            }
          }
        }
#pragma unroll
        for (int u = 0; u < EPT; ++u)
          if (e0 + u < E) sum += (double)(log2_split((acc[u][0] + acc[u][1]) + (acc[u][2] + acc[u][3])) + p.scale);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      if (lane == 0) sRed[warp] = sum;
    }
    __syncthreads();
    // ---- accept.  Small CTAs: every thread takes the same decision from the same numbers (no flag broadcast, no
    // second barrier: 759 k -> 813 k iterations/s at config 1); large CTAs: thread 0 decides and a barrier publishes
    // the flag (576 threads evaluating the fp64 exponential cost more than the barrier: 28.6 vs 27.2 M at config 2)
    int flag = 0;
    if (kThreads <= 320 || tid == 0) {
      double v = 0.0;
      for (int w8 = 2; w8 < kThreads / 32; ++w8) v += sRed[w8];
      const double ll = __fma_rn(NEMB200_LN2, v, sumN);  // (explicit, like the score kernel's finalize: identical roundings on both MCMC paths)
      bool acc_ok, better;
      if (init) {
        acc_ok = true; better = true;
      } else {
        const double u = ((double)(uint32_t)sFlag[4 + x] + 0.5) * (1.0 / 4294967296.0);
        acc_ok = u < exp(gamma * (ll - curr));
        better = acc_ok && ll > best;
        if (tid == 0 && p.trace != nullptr) {
          p.trace[((size_t)it * 2 + 0) * p.n + c] = ll;
          p.trace[((size_t)it * 2 + 1) * p.n + c] = acc_ok ? ll : curr;
        }
      }
      if (acc_ok) curr = ll;
      if (better) best = ll;
      flag = (acc_ok ? 1 : 0) | (better ? 2 : 0);
      if (kThreads > 320) sFlag[0] = flag;
    }
    if (kThreads > 320) {
      __syncthreads();
      flag = sFlag[0];
    }
    if (flag & 1) {
      if (!init) {
        if (counts != nullptr) {
          const long long from = since > p.burn_in ? since : p.burn_in;
          if (t > from) {
            const uint32_t k = (uint32_t)(t - from);
            for (int cell = tid; cell < S * S; cell += kThreads) {
              const int a = cell / S, b = cell - a * S;
              if ((sCurr[a * TW + (b >> 5)] >> (b & 31)) & 1u) counts[cell] += k;
            }
          }
        }
        since = t;
        ++n_acc;
      }
      __syncthreads();  // edge counts read sCurr before it is overwritten
      const int *sProp = slot_prop(x);
      const uint32_t *sRows = slot_rows(x);
      for (int i = tid; i < S; i += kThreads) {
        const int v = sProp[i];
        sOrder[i] = v;
        if (flag & 2) p.best_orders[(size_t)c * S + i] = v;
      }
      for (int i = tid; i < S * TW; i += kThreads) {
        const uint32_t v = sRows[i];
        sCurr[i] = v;
        if (flag & 2) p.best_bits[(size_t)c * S * TW + i] = v;
      }
    }
    __syncthreads();
    // the successor that matches the decision becomes the proposal being scored; the other two slots are free
    if (flag & 1) { const int o = x; x = sa; sa = o; } else { const int o = x; x = sr; sr = o; }
  }
  // ---- write the chain state back -----------------------------------------------------------------------
  if (counts != nullptr) {
    const long long t_end = p.iter0 + p.n_iters, from = since > p.burn_in ? since : p.burn_in;
    if (t_end > from) {
      const uint32_t k = (uint32_t)(t_end - from);
      for (int cell = tid; cell < S * S; cell += kThreads) {
        const int a = cell / S, b = cell - a * S;
        if ((sCurr[a * TW + (b >> 5)] >> (b & 31)) & 1u) counts[cell] += k;
      }
    }
  }
  for (int i = tid; i < S; i += kThreads) p.orders[(size_t)c * S + i] = sOrder[i];
  for (int i = tid; i < S * TW; i += kThreads) p.curr_bits[(size_t)c * S * TW + i] = sCurr[i];
  if (tid == 0) {
    p.curr_scores[c] = curr;
    p.best_scores[c] = best;
    p.accept[c] += n_acc;
  }
}


}  // namespace

// Can the fused kernel run this problem at all / is it expected to beat the tensor-core loop?
bool mcmc_fused_supported(const nemb200_dataset *h) { return h->lut_valid && h->S >= 2 && h->S <= 63; }

bool mcmc_fused_preferred(const nemb200_dataset *h, int n_chains) {
  if (!mcmc_fused_supported(h)) return false;
  if ((size_t)h->E * h->W * 4 > 96 * 1024) return false;  // the bit columns of D must fit in shared memory
  // Measured model (tools/mcmc_crossover.py, us per iteration of all chains): a CTA-iteration of the persistent kernel costs
  // 1.5 + 0.085 per 1000 (site, E-gene) pairs and the 148 SMs take the chains in waves (two CTAs per SM share it); the
  // three-kernel loop pays ~25 us of launches per iteration plus the tensor-core score of the batch.  Chains are
  // independent, so the persistent kernel does not need all of them resident at once.
  const double lookups = (double)h->S * h->E;
  const double share = h->E <= 256 ? 3.0 : 1.0;  // the small (256/320-thread) CTAs overlap about three deep on an SM
  const double waves = ceil(n_chains / 148.0) / share;
  const double fused = (1.5 + 0.085e-3 * lookups) * (waves > 1.0 ? waves : 1.0);
  const double loop = 25.0 + (double)n_chains * lookups * 1.5e-7;
  return fused < loop;
}

int launch_mcmc_fused(const nemb200_dataset *h, int n, int64_t n_iters, int64_t iter0, int64_t burn_in, const double *d_gamma,
                      unsigned long long sp, uint64_t seed, uint32_t chain_offset, int rule, const uint32_t *d_cand_bits,
                      int use_nem, int resume, int32_t *d_orders, double *d_curr_scores, uint32_t *d_curr_bits,
                      double *d_best_scores, int32_t *d_best_orders, uint32_t *d_best_bits, int64_t *d_accept,
                      uint32_t *d_edge_counts, double *d_score_trace, cudaStream_t st) {
  NEMB_REQUIRE(mcmc_fused_supported(h), "fused MCMC kernel needs a valid exp table and 2 <= S <= 63 (S=%d)", h->S);
  FusedParams p;
  p.dcol = h->dcol; p.lut = h->lut; p.sumN = h->sumN; p.n = n; p.S = h->S; p.E = h->E;
  p.n_iters = n_iters; p.iter0 = iter0; p.burn_in = burn_in; p.gamma = d_gamma; p.swap_prob_u33 = sp; p.seed = seed;
  p.chain_offset = chain_offset; p.rule = rule; p.use_nem = use_nem; p.resume = resume; p.cand = d_cand_bits;
  p.orders = d_orders; p.curr_scores = d_curr_scores; p.curr_bits = d_curr_bits; p.best_scores = d_best_scores;
  p.best_orders = d_best_orders; p.best_bits = d_best_bits; p.accept = (long long *)d_accept; p.edge_counts = d_edge_counts;
  p.trace = d_score_trace; p.scale = (float)h->scale;
  const size_t dbytes = (size_t)h->E * h->W * 4;
  p.dcol_in_smem = dbytes <= 96 * 1024;
  const size_t smem = kFixedSmem + (p.dcol_in_smem ? dbytes : 0);
  const bool one = h->E <= 256;  // one E-gene per thread when a single pass of 256 threads covers them all
#define NEMB_FUSED(TW_, EPT_, THR_)                                                                                            \
  do {                                                                                                                         \
    NEMB_CHECK_CUDA(cudaFuncSetAttribute(k_mcmc_fused<TW_, EPT_, THR_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    k_mcmc_fused<TW_, EPT_, THR_><<<n, THR_, smem, st>>>(p);                                                                   \
  } while (0)
  const bool tiny = h->E <= 192;  // 192 scorer threads cover it: the smallest CTA has the cheapest barriers
  if (h->W == 1) { if (tiny) NEMB_FUSED(1, 1, 256); else if (one) NEMB_FUSED(1, 1, 320); else NEMB_FUSED(1, 2, 576); }
  else { if (tiny) NEMB_FUSED(2, 1, 256); else if (one) NEMB_FUSED(2, 1, 320); else NEMB_FUSED(2, 2, 576); }
#undef NEMB_FUSED
  NEMB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

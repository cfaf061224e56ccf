/* nemb200.h - C ABI of the B200-native NEM scoring / order-MCMC hot path.
 *
 * The reference (MrGreyPanda/NEM-MCMC-optimization) is pure Python and has no FFI of its own;
 * each entry point below names the reference method(s) (file:line under /root/reference) whose
 * work it replaces.  INTEGRATION.md shows the ctypes binding a reference maintainer would add.
 *
 * Conventions
 *  - every pointer named d_* is DEVICE memory owned by the caller (torch); the library never
 *    allocates device memory, never synchronises the device and launches only on `stream`
 *    (a cudaStream_t passed as void*; NULL = legacy default stream);
 *  - every function returns 0 on success, <0 on error; nemb200_last_error() gives the
 *    thread-local message of the last failure;
 *  - graphs are bit-packed: W = (S+31)/32 uint32 words per row, rows[g][a][w] bit b set iff
 *    edge a -> (32*w+b).  This is the reference's `dag`/`nem` orientation
 *    (dag[a][b]==1 iff a->b, nem_order_mcmc.py:213,220); the score ignores the diagonal (it uses I|Phi),
 *    the closure treats it like any other entry;
 *  - orders are int32[S] position -> S-gene label (the reference's perm_order, :29-48);
 *  - S <= 128 (tensor-core path), E arbitrary;
 *  - threads: a dataset handle is immutable after nemb200_dataset_create; the little host state the launchers keep
 *    (completion-counter assignment of score workspaces, the cached CUDA graph of the MCMC loop) sits behind a mutex
 *    inside the handle, so several host threads may call nemb200_score_batch on one handle as long as each uses its
 *    OWN score workspace and stream (launches that share a workspace must be stream-ordered); nemb200_mcmc_run
 *    calls on one handle are serialised.  nemb200_dataset_destroy must not race with any other call.
 */
#ifndef NEMB200_H
#define NEMB200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NEMB200_VERSION 1
#define NEMB200_MAX_S 128

enum {
  NEMB200_RULE_FULL = 0,   /* W = permissibility mask of the order */
  NEMB200_RULE_MASKED = 1, /* W = permissibility & candidate-edge mask */
  NEMB200_RULE_FIT = 2     /* nemb200_mcmc_run only: the reference's own order -> DAG step - per-order edge-weight fit
                              (get_optimal_weights, nem_order_mcmc.py:172-208) + threshold, scored with the driver's
                              expit-on-weights cell ratios (:207), all chains at once on the device */
};
/* score implementations (nemb200_score_batch `impl`) */
enum {
  NEMB200_SCORE_AUTO = 0,    /* tensor-core path, best epilogue for the dataset */
  NEMB200_SCORE_WITNESS = 1, /* popcount + fp64 logsumexp CUDA kernel (independent witness) */
  NEMB200_SCORE_TC_LUT = 2,  /* tcgen05 GEMM + shared-memory exp-table epilogue */
  NEMB200_SCORE_TC_EX2 = 3,  /* tcgen05 GEMM + ex2 epilogue, fixed scale */
  NEMB200_SCORE_TC_EX2_2PASS = 4, /* tcgen05 GEMM + max-subtracted ex2 epilogue (any range) */
  NEMB200_MCMC_FUSED = 5, /* nemb200_mcmc_run only: persistent one-CTA-per-chain kernel (small S*E) */
  NEMB200_SCORE_TC_PAIR = 6, /* tcgen05 GEMM + pair-merged table epilogue: two attachment sites per look-up (33 <= S <= 52): a planner kernel
                                pairs the sites of every graph (and closes the graph, for nemb200_closure_score_batch), graphs it
                                cannot plan go through the one-site kernel in the same call.  Three launches, no host round trip.
                                AUTO of nemb200_score_batch / nemb200_closure_score_batch picks it for 44 <= S <= 50. */
  NEMB200_SCORE_TC_LUTM = 7, /* tcgen05 GEMM + multi-level exp-table epilogue: dynamic ranges of up to ~630 binary orders, 16 <= S <= 125 */
  NEMB200_SCORE_TC_LUT_MIX = 8 /* NEMB200_SCORE_TC_LUT with every fourth group of four attachment sites through ex2.approx
                                  instead of the table (the look-ups bound the kernel on the shared-memory pipe, the MUFU
                                  pipe is idle): about 5 % faster, relative error of a score <= 5e-7 instead of <= 5e-8.
                                  Column stride 50 (S = 50) only; AUTO of nemb200_mcmc_run's kernel loop at that stride. */
};

/* flag, OR-ed into `impl` of nemb200_score_batch / nemb200_closure_score_batch: the caller promises that the input graphs
 * are NOT written by the kernel that precedes this call on `stream` (they are static, or they were produced earlier, or
 * they arrive from another stream and are joined by an event).  The launches of a call are dependent launches
 * (cudaLaunchAttributeProgrammaticStreamSerialization, never under stream capture); with the promise the first kernel
 * reads its graphs and works ahead while the previous kernel of the stream drains, and waits for it (griddepcontrol.wait)
 * only in front of its first write.  Without the flag it waits before its first read, as a plain launch would. */
enum { NEMB200_SCORE_INPUTS_READY = 0x100 };

typedef struct nemb200_dataset nemb200_dataset; /* opaque host handle */

int nemb200_version(void);
const char *nemb200_last_error(void);

/* ---- resident data: replaces NEM.get_score_tables / build_score_table / get_node_lr_table
 *      (nem.py:25-64) and the copies NEMOrderMCMC.__init__ keeps (nem_order_mcmc.py:40-44).
 *      d_D: S x E row-major uint8 0/1 = NEM.observed_knockdown_mat (nem.py:19).
 *      A, B: nem.py:17-18.  d_workspace: nemb200_dataset_bytes(S,E) bytes, 1024-byte aligned,
 *      must outlive the handle. */
size_t nemb200_dataset_bytes(int S, int E);
int nemb200_dataset_create(const uint8_t *d_D, int S, int E, double A, double B,
                           void *d_workspace, size_t workspace_bytes, void *stream,
                           nemb200_dataset **out);
int nemb200_dataset_destroy(nemb200_dataset *h);
/* host-side facts about the dataset: fills info[0..7] =
 * {S, E, W, K_pad, graphs_per_tile, lut_valid, scale_log2, n_mtiles}; with n >= 9 also info[8] = number of levels
 * of the multi-level exp table (NEMB200_SCORE_TC_LUTM; 0 = not usable for this dataset); with n >= 10 also info[9] = 1 if
 * the pair-merged path (NEMB200_SCORE_TC_PAIR) is available for this dataset */
int nemb200_dataset_info(const nemb200_dataset *h, int64_t *info, int n);

/* ---- transitive closure: replaces utils.ancestor (utils.py:37-54) as called by
 *      NEMOrderMCMC.create_nem (nem_order_mcmc.py:216-221).  Boolean (mathematical) closure of the
 *      relation as given (a DAG without self loops stays irreflexive; nodes on a cycle get their
 *      diagonal bit, like the reference's path counts).  In-place allowed (d_out == d_adj). */
int nemb200_closure_u32(const uint32_t *d_adj_bits, int n_graphs, int S, uint32_t *d_out_bits,
                        void *stream);

/* ---- transitive reduction: replaces utils.transitive_reduction (utils.py:120-129) as used by
 *      main.output_handling (main.py:50) on the best graph of a run, for a batch of graphs (e.g. the best graph of
 *      every chain).  Reproduces the reference's in-place update order exactly: for a closed DAG the result is its
 *      Hasse diagram; graphs with cycles or self loops get what the reference's triple loop leaves.  In-place
 *      allowed (d_out == d_adj). */
int nemb200_transitive_reduction_u32(const uint32_t *d_adj_bits, int n_graphs, int S,
                                     uint32_t *d_out_bits, void *stream);

/* ---- batched NEM score: replaces compute_ll_ratios + calculate_ll for binary graphs
 *      (nem.py:72-86 driven as nem.py:132-142; utils.py:66-94; nem_order_mcmc.py:79-93 without
 *      the expit quirk).  d_phi_bits: n_graphs closed (or not) graphs; score is
 *      sum_e N[e] + sum_e LSE(0, ((I|Phi)^T T)[:,e]).  d_scores: n_graphs float64.
 *      d_workspace: nemb200_score_workspace_bytes(h, n_graphs) bytes. */
size_t nemb200_score_workspace_bytes(const nemb200_dataset *h, int n_graphs);
/* the same for one specific `impl`: NEMB200_SCORE_TC_PAIR needs room for its plan and pair tables
 * (about 10 KB per graph); every other implementation fits nemb200_score_workspace_bytes(), which is the size for
 * NEMB200_SCORE_AUTO (it includes the pair-merged path's room where AUTO resolves to it, 44 <= S <= 50). */
size_t nemb200_score_workspace_bytes_impl(const nemb200_dataset *h, int n_graphs, int impl);
int nemb200_score_batch(const nemb200_dataset *h, const uint32_t *d_phi_bits, int n_graphs,
                        double *d_scores, void *d_workspace, size_t workspace_bytes, int impl,
                        void *stream);

/* diagnostics: number of graphs of the last NEMB200_SCORE_TC_PAIR launch on `d_workspace` (a batch of n_graphs) that the
 * pair planner could not plan and that were scored by the one-site kernel instead; synchronises the device; < 0 on error */
int nemb200_pair_fallback_count(const nemb200_dataset *h, const void *d_workspace, int n_graphs);

/* ---- closure + score in ONE launch: create_nem's utils.ancestor (nem_order_mcmc.py:216-221, utils.py:37-54) followed by
 *      the score above, for candidate graphs that are not transitively closed yet (what a sweep over proposals or
 *      `Dataset.score_graphs(close=True)` needs).  The score kernel's operand-expansion warps close each graph in
 *      registers (the same warp Warshall as nemb200_closure_u32) before they build the MMA operand, so no closed
 *      copy is written to HBM and no second kernel is launched.  Same workspace and `impl` rules as
 *      nemb200_score_batch; NEMB200_SCORE_WITNESS is not available here (S >= 8).  With NEMB200_SCORE_TC_PAIR the closure
 *      runs in the pair planner (which has every graph in shared memory anyway) instead of the score kernel. */
int nemb200_closure_score_batch(const nemb200_dataset *h, const uint32_t *d_adj_bits, int n_graphs,
                                double *d_scores, void *d_workspace, size_t workspace_bytes, int impl,
                                void *stream);

/* ---- general-weight cell ratios: replaces compute_ll_ratios / compute_cell_ratios for real-valued
 *      parent weights (nem.py:72-80, utils.py:66-82, nem_order_mcmc.py:79-87).  d_weights: S x S
 *      float64, w[i][j] = weight of parent j for node i (0 = no edge, diagonal ignored);
 *      apply_expit != 0 reproduces the MCMC driver's expit-on-weights (nem_order_mcmc.py:83-86).
 *      d_cell: (S+1) x E float64 out, row S = the null row N (nem.py:62). */
int nemb200_cell_ratios(const nemb200_dataset *h, const double *d_weights, int apply_expit,
                        double *d_cell, void *stream);

/* ---- column logsumexp + total: replaces calculate_ll / compute_ll (nem.py:82-86,
 *      nem_order_mcmc.py:89-93, utils.py:84-94).  d_cell: rows x E float64; d_order_weights
 *      (rows x E, may be NULL) receives exp(cell - colsum); d_ll[0] the log-likelihood. */
size_t nemb200_cell_ll_workspace_bytes(int E);
int nemb200_cell_ll(const double *d_cell, int rows, int E, double *d_order_weights, double *d_ll,
                    void *d_workspace, size_t workspace_bytes, void *stream);

/* ---- per-order edge-weight fit: replaces the loop of one-dimensional scipy fits in
 *      NEMOrderMCMC.get_optimal_weights (nem_order_mcmc.py:183-194) with calculate_local_optimum (:160-170) and
 *      the objective local_ll_sum_penalized (:18-23).  All matrices are S x S float64 row-major, [i][k] = node i,
 *      parent k.  d_weights: current parent weights (raw, the kernel applies expit as :163 does);
 *      d_order_weights: (S+1) x E posterior attachment weights from nemb200_cell_ll (:89-93); d_ancestor_x: the
 *      clipped (I - expit(W))^-1 - I of :185; d_mask: uint8, 1 = fit this pair (permissible parents, or only those
 *      touching the swapped nodes as in :190-194), 0 = copy the weight through.  d_out_weights[i][k] = the
 *      minimiser of the objective in s = expit(x) over [0,1] (what the reference thresholds at 0.5, :212).  The
 *      reference stops L-BFGS-B at tol=0.01; this returns the exact global minimiser. */
/* ---- ancestor_x of get_optimal_weights (nem_order_mcmc.py:185): clip((I - expit_parent_weights(W))^-1 - I, 0, 1), the
 *      reference's scipy.linalg.inv call, on the device (one CTA: in-place Gauss-Jordan with partial pivoting in shared
 *      memory, S <= 128).  d_weights: S x S float64 raw weights; d_permissible: uint8 S x S, 1 where k is a permissible
 *      parent of i (expit is applied to those entries only, :152-157); d_out: S x S float64; *d_singular (int, device) is set to 1 if a pivot vanishes (never for weights that respect an order). */
int nemb200_ancestor_x(const double *d_weights, const uint8_t *d_permissible, int S, double *d_out,
                       int *d_singular, void *stream);

int nemb200_fit_edges(const nemb200_dataset *h, const double *d_weights, const double *d_order_weights,
                      const double *d_ancestor_x, const uint8_t *d_mask, double *d_out_weights,
                      void *stream);

/* ---- order -> graph: replaces get_permissible_parents (nem_order_mcmc.py:54-77) +
 *      create_dag / create_nem (:210-221) with the deterministic rules of DESIGN.md
 *      (FULL: every earlier node is a parent; MASKED: & d_cand_bits, S x W words shared by all
 *      chains).  use_nem != 0 closes the graph. */
int nemb200_order_to_graph(const int32_t *d_orders, int n_chains, int S, int rule,
                           const uint32_t *d_cand_bits, int use_nem, uint32_t *d_out_bits,
                           void *stream);

/* ---- device-resident order-MCMC: replaces the loop of NEMOrderMCMC.method
 *      (nem_order_mcmc.py:272-302): get_new_order (:231-255), order->graph, score, accepting
 *      (:224-229), best-so-far bookkeeping (:290-309).  Philox4x32-10 keyed by `seed`, counter
 *      (iteration, global chain id = chain_offset + local index, 0).
 *      State (all device, caller-owned, n_chains leading dimension):
 *        d_orders      int32 [n][S]      in: start orders, out: current orders
 *        d_curr_scores f64   [n]         in (if resume) / out
 *        d_curr_bits   u32   [n][S][W]   in (if resume) / out: current graphs
 *        d_best_scores f64   [n], d_best_orders int32 [n][S], d_best_bits u32 [n][S][W]
 *        d_accept      i64   [n]         accepted proposals (accumulated)
 *        d_edge_counts u32   [n][S*S]    per-chain sum over post-burn-in iterations of the current
 *                                        graph (may be NULL); edge a->b at [a*S+b]
 *        d_score_trace f64   [n_iters][2][n] per iteration: proposed score, current score after the
 *                                        accept decision (may be NULL)
 *        d_gamma       f64   [n]         per-chain inverse temperature (gamma of :224-229)
 *      iter0: global index of the first iteration (RNG counter / burn-in bookkeeping);
 *      resume == 0 scores the start orders first (method(), :258).
 *      score_impl: NEMB200_SCORE_AUTO picks the persistent one-CTA-per-chain kernel where it is expected to be
 *      faster (small S*E, from a measured cost model of both paths; the chains run in waves when they do not all
 *      fit on the GPU at once) and the CUDA-graph loop around the tensor-core score otherwise.
 *      d_workspace: nemb200_mcmc_workspace_bytes(h, n_chains) bytes. */
size_t nemb200_mcmc_workspace_bytes(const nemb200_dataset *h, int n_chains);
/* the same for one specific `rule`: NEMB200_RULE_FIT needs two (S+1) x E float64 tables and five S x S matrices per chain */
size_t nemb200_mcmc_workspace_bytes_rule(const nemb200_dataset *h, int n_chains, int rule);
int nemb200_mcmc_run(const nemb200_dataset *h, int n_chains, int64_t n_iters, int64_t iter0,
                     int64_t burn_in, const double *d_gamma, double swap_prob, uint64_t seed,
                     uint32_t chain_offset, int rule, const uint32_t *d_cand_bits, int use_nem,
                     int resume, int score_impl, int32_t *d_orders, double *d_curr_scores,
                     uint32_t *d_curr_bits, double *d_best_scores, int32_t *d_best_orders,
                     uint32_t *d_best_bits, int64_t *d_accept, uint32_t *d_edge_counts,
                     double *d_score_trace, void *d_workspace, size_t workspace_bytes,
                     void *stream);

/* ---- replica exchange between neighbouring chains of one ladder: replaces the swap half of
 *      replica_exchange_step (nem_order_mcmc.py:326-340).  Chains are grouped in ladders of
 *      `ladder` consecutive chains; phase 0 pairs (0,1),(2,3).., phase 1 pairs (1,2),(3,4)..;
 *      swap iff u < exp(-delta), delta as :333.  States (order, graph, score) are exchanged,
 *      gammas stay.  The uniform of pair k of ladder l is (r3 + 0.5) / 2^32 of Philox4x32-10 with key = seed and
 *      counter (round lo, round hi, l * (ladder / 2) + k, 1). */
int nemb200_replica_exchange(int n_chains, int ladder, int phase, int64_t round, uint64_t seed,
                             int S, const double *d_gamma, int32_t *d_orders,
                             double *d_curr_scores, uint32_t *d_curr_bits, int64_t *d_n_swaps,
                             void *stream);
/* the same with the reference's own swap criterion: replica_exchange_step takes scores[i] =
 * replicas[i].best_score (nem_order_mcmc.py:321), i.e. the best score each replica reached in the round that
 * just ended (method() restarts best_score from the current score, :259-260), puts them into delta (:333) and
 * swaps WHOLE replica objects (:335): current state and best state change places, scores[] with them (:336).
 * Same pairing, same Philox counter as above. */
int nemb200_replica_exchange_best(int n_chains, int ladder, int phase, int64_t round, uint64_t seed,
                                  int S, const double *d_gamma, int32_t *d_orders,
                                  double *d_curr_scores, uint32_t *d_curr_bits, double *d_best_scores,
                                  int32_t *d_best_orders, uint32_t *d_best_bits, int64_t *d_n_swaps,
                                  void *stream);

/* ---- device-side data synthesis: replaces utils.create_real_knockdown_mat (utils.py:15-23) and
 *      utils.create_observed_knockdown_mat (utils.py:25-35) for large shapes.  d_adj_bits: the (closed) true
 *      graph as S x W bit rows; d_end_nodes: int32 [E] attachment of every E-gene; d_real (may be NULL) and
 *      d_observed: S x E row-major uint8 out.  The flip uniform of cell i*E+e is Philox4x32-10(key = seed,
 *      counter = (cell lo, cell hi, 0, 2)) word 0 (the reference uses MT19937, so only the law is the same). */
int nemb200_synth_knockdown(const uint32_t *d_adj_bits, const int32_t *d_end_nodes, int S, int E,
                            double alpha, double beta, uint64_t seed, uint8_t *d_real,
                            uint8_t *d_observed, void *stream);

/* ---- posterior edge frequencies: sum per-chain counts into int64 [S*S] (what the NCCL
 *      all-reduce then sums across GPUs).  New in this build. */
int nemb200_edge_counts_reduce(const uint32_t *d_edge_counts, int n_chains, int S,
                               int64_t *d_out, void *stream);

/* ---- host-only self-test of the score kernel's work partition (no device work; used by the CPU tests):
 *      0 if every (N tile, E tile) pair is owned by exactly one CTA and every N tile has a dense set of at most
 *      24 partial-sum slots for this problem size. */
int nemb200_selftest_partition(int n_ntiles, int n_mtiles, int max_ctas);

#ifdef __cplusplus
}
#endif
#endif

/* CPU oracle (C restatement) for NEM scoring inside order-MCMC.
 *
 * TEST INFRASTRUCTURE ONLY - never linked into or called from the product library.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * load this.  Each function cites the reference file:line it follows
 * (reference: MrGreyPanda/NEM-MCMC-optimization, pure Python/numpy).
 *
 * Parity: score + closure are PINNED by the reference's golden values (tests/golden/),
 * the MCMC trajectory is UNPINNED in the reference (see oracle/nem_oracle.py header).
 *
 * Conventions: adj[a*S+b]==1 iff edge a->b (the reference's `dag = W.T`,
 * nem_order_mcmc.py:213); D is S x E row-major 0/1 (observed_knockdown_mat).
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* utils.py:37-54 (boolean semantics: SURVEY Q8) - Warshall. */
void orc_closure(const uint8_t *adj, int S, uint8_t *out) {
  memcpy(out, adj, (size_t)S * S);
  for (int k = 0; k < S; ++k)
    for (int i = 0; i < S; ++i)
      if (out[i * S + k])
        for (int j = 0; j < S; ++j) out[i * S + j] |= out[k * S + j];
}

/* nem.py:72-86 driven as nem.py:132-142 (weights 1):  cell[j] = U[j] + sum_{i in pa(j)} T[i],
 * ll = sum_e logaddexp.reduce(cell[:,e]) over S+1 rows; written as
 * ll = sum_e N[e] + sum_e LSE(0, (Phi'^T T)[:,e])  (SURVEY section 8). */
double orc_score(const uint8_t *adj, const uint8_t *D, int S, int E, double A, double B) {
  double *v = (double *)malloc(sizeof(double) * (size_t)S * E);
  double total = 0.0;
  for (int j = 0; j < S; ++j) {
    double *vj = v + (size_t)j * E;
    for (int e = 0; e < E; ++e) vj[e] = 0.0;
    for (int i = 0; i < S; ++i) {
      if (i == j || adj[i * S + j]) {
        const uint8_t *di = D + (size_t)i * E;
        for (int e = 0; e < E; ++e) vj[e] += di[e] ? -A : B; /* T[i,e], nem.py:46 */
      }
    }
  }
  for (int e = 0; e < E; ++e) {
    double m = 0.0, n = 0.0;
    for (int j = 0; j < S; ++j) {
      double x = v[(size_t)j * E + e];
      if (x > m) m = x;
    }
    double s = exp(-m);
    for (int j = 0; j < S; ++j) s += exp(v[(size_t)j * E + e] - m);
    for (int i = 0; i < S; ++i) n += D[(size_t)i * E + e] ? A : 0.0; /* nem.py:62 */
    total += n + m + log(s);
  }
  free(v);
  return total;
}

typedef struct {
  const uint8_t *adj, *D; int n_graphs, S, E, tid, nthreads; double A, B; double *out;
} batch_arg;

static void *batch_worker(void *p) {
  batch_arg *a = (batch_arg *)p;
  for (int g = a->tid; g < a->n_graphs; g += a->nthreads)
    a->out[g] = orc_score(a->adj + (size_t)g * a->S * a->S, a->D, a->S, a->E, a->A, a->B);
  return 0;
}

/* Same score for a batch of graphs, fanned out over nthreads host threads (pthreads). */
void orc_score_batch(const uint8_t *adj, int n_graphs, const uint8_t *D, int S, int E, double A,
                     double B, double *out, int nthreads) {
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 256) nthreads = 256;
  pthread_t th[256]; batch_arg args[256];
  for (int t = 0; t < nthreads; ++t) {
    args[t] = (batch_arg){adj, D, n_graphs, S, E, t, nthreads, A, B, out};
    pthread_create(&th[t], 0, batch_worker, &args[t]);
  }
  for (int t = 0; t < nthreads; ++t) pthread_join(th[t], 0);
}

/* Philox4x32-10 (Salmon et al., SC'11) - the counter-based RNG of the CUDA path. */
static void philox(uint32_t c[4], uint32_t k0, uint32_t k1) {
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1, n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}

/* nem_order_mcmc.py:54-77 + :210-221: graph of an order.  rule 0: every earlier node is a
 * parent; rule 1: & cand (S*S candidate edges).  use_nem closes it. */
static void order_graph(const int32_t *order, int S, int rule, const uint8_t *cand, int use_nem,
                        uint8_t *g, uint8_t *tmp) {
  int pos[256];
  for (int p = 0; p < S; ++p) pos[order[p]] = p;
  for (int a = 0; a < S; ++a)
    for (int b = 0; b < S; ++b)
      tmp[a * S + b] = (pos[a] < pos[b]) && (rule == 0 || cand[a * S + b]);
  if (use_nem) orc_closure(tmp, S, g); else memcpy(g, tmp, (size_t)S * S);
}

/* nem_order_mcmc.py:272-302 with proposal :231-255 and accept :224-229.
 * outputs: best_score, best_order[S], best_graph[S*S], curr_score, curr_order[S],
 * n_accept, edge_counts[S*S] (int64), optional all_scores[n_iter+1]. */
void orc_mcmc_chain(const uint8_t *D, int S, int E, double A, double B, const int32_t *order0,
                    int64_t n_iter, double gamma, double swap_prob, uint64_t seed, uint32_t chain,
                    int rule, const uint8_t *cand, int use_nem, int64_t burn_in,
                    double *best_score, int32_t *best_order, uint8_t *best_graph,
                    double *curr_score_out, int32_t *curr_order_out, int64_t *n_accept,
                    int64_t *edge_counts, double *all_scores) {
  size_t SS = (size_t)S * S;
  uint8_t *cg = malloc(SS), *pg = malloc(SS), *tmp = malloc(SS);
  int32_t *co = malloc(sizeof(int32_t) * S), *po = malloc(sizeof(int32_t) * S);
  memcpy(co, order0, sizeof(int32_t) * S);
  order_graph(co, S, rule, cand, use_nem, cg, tmp);
  double curr = orc_score(cg, D, S, E, A, B), best = curr;
  memcpy(best_order, co, sizeof(int32_t) * S);
  memcpy(best_graph, cg, SS);
  memset(edge_counts, 0, sizeof(int64_t) * SS);
  if (all_scores) all_scores[0] = curr;
  double spd = swap_prob < 0 ? 0 : swap_prob;
  uint64_t sp = spd >= 1.0 ? 0x100000000ull : (uint64_t)(spd * 4294967296.0 + 0.5);
  int64_t acc = 0;
  for (int64_t t = 0; t < n_iter; ++t) {
    uint32_t c[4] = {(uint32_t)t, (uint32_t)((uint64_t)t >> 32), chain, 0};
    philox(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    memcpy(po, co, sizeof(int32_t) * S);
    int i, j;
    if ((uint64_t)c[0] < sp) {
      i = (int)(((uint64_t)c[1] * (uint64_t)S) >> 32);
      j = (int)(((uint64_t)c[2] * (uint64_t)(S - 1)) >> 32);
      if (j >= i) ++j;
    } else {
      i = (int)(((uint64_t)c[1] * (uint64_t)(S - 1)) >> 32);
      j = i + 1;
    }
    int32_t t0 = po[i]; po[i] = po[j]; po[j] = t0;
    order_graph(po, S, rule, cand, use_nem, pg, tmp);
    double ll = orc_score(pg, D, S, E, A, B);
    if (all_scores) all_scores[t + 1] = ll;
    double u = ((double)c[3] + 0.5) / 4294967296.0;
    if (u < exp(gamma * (ll - curr))) {
      ++acc; curr = ll;
      memcpy(co, po, sizeof(int32_t) * S);
      memcpy(cg, pg, SS);
      if (curr > best) {
        best = curr;
        memcpy(best_order, co, sizeof(int32_t) * S);
        memcpy(best_graph, cg, SS);
      }
    }
    if (t >= burn_in)
      for (size_t q = 0; q < SS; ++q) edge_counts[q] += cg[q];
  }
  *best_score = best; *curr_score_out = curr; *n_accept = acc;
  memcpy(curr_order_out, co, sizeof(int32_t) * S);
  free(cg); free(pg); free(tmp); free(co); free(po);
}

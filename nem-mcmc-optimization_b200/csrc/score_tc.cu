// Batched NEM score on tcgen05 tensor cores with the logsumexp fused into the epilogue.
//
// Replaces, for a batch of binary graphs, compute_ll_ratios + calculate_ll of the reference
// (nem.py:72-86 as driven by nem.py:132-142; nem_order_mcmc.py:79-93 without the expit quirk):
//     ll(g) = sum_e N[e] + sum_e log( 1 + sum_j exp( ((I|Phi_g)^T T)[j,e] ) )
//
// Data model: T[i,e] takes two values (B if D[i,e]==0, -A if 1), so
//     ((I|Phi)^T T)[j,e] = -A * p + B * q,  p = #{ancestors-or-self i of j with D[i,e]=1}, q likewise D=0.
// GEMM orientation: M = E-genes (128 TMEM lanes = 128 E-genes per tile), N = (graph, attachment site j)
// stacked, K = S-genes.  Operands are unsigned 8-bit (kind::i8, s32 accumulate, exact):
//     A[e][k] = 2 if D[k,e]==1 else 128   (resident image, streamed by TMA bulk copy)
//     B[(g,j)][k] = 2 if k is j or an ancestor of j in graph g (expanded in-kernel from bit rows)
// so one accumulator element is 4p + 256q: exactly the byte offset of entry [q][p] of a float table
// lut[q][p] = 2^(alpha2*p + beta2*q - s).  Epilogue "LUT": one LDS + one FADD per (site, E-gene); an extra
// K row adds the table's shared-memory base so the accumulator IS the address.  Epilogue "EX2": the same
// accumulator is mapped linearly to the log2-domain value and exponentiated with ex2.approx (any S<=128);
// "EX2_2PASS" subtracts the running maximum first (any dynamic range).  For the headline shapes the
// epilogue is compiled for a fixed S and splits each graph's sites between the two pipes (LSU table
// look-ups and MUFU ex2) so that neither is the single limiter.
// No B x S x E intermediate ever reaches HBM: per (graph, E-chunk, warpgroup) one double is written.
//
// Roles (640 threads, no CTA-wide barrier in the main loop): warp 0 = TMA producer (A tiles, 4-stage
// mbarrier ring), warp 1 = MMA issuer + TMEM owner (2 x 256-column accumulators), warps 2-3 = B-operand
// expanders (double-buffered, run one work item ahead), warps 4-19 = four epilogue warpgroups.  All four
// drain the SAME accumulator buffer (each takes a quarter of the tile's graphs) while the MMA of the next
// tile fills the other buffer, so the MMA + mbarrier hand-shake latency is hidden behind epilogue work and
// four warps per scheduler hide the LDS / MUFU / TMEM latencies.  Work items (N tile of G graphs x chunk
// of E tiles) are statically strided over CTAs.
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int kThreads = 640;
constexpr int kStages = 4;
constexpr int kTileBytes = 128 * 128;  // one operand atom: 128 rows x 128 B (SWIZZLE_128B)
constexpr int kLutBytes = 64 * 256;
constexpr int kMaxG = 16;
constexpr int kBiasStride = 64;        // floats per graph in the bias array (S <= 128 uses two slots)

// shared memory carve-up (offsets from the 1024-aligned base)
constexpr int OFF_LUT = 0;                                  // first: its base must be < 32 KB and 128-aligned
constexpr int OFF_B = OFF_LUT + kLutBytes;                  // 2 buffers x (256 x 128 B)
constexpr int OFF_A = OFF_B + 2 * 2 * kTileBytes;           // kStages x 16 KB
constexpr int OFF_BIAS = OFF_A + kStages * kTileBytes;      // 2 buffers x 1024 floats
constexpr int OFF_ROWS = OFF_BIAS + 2 * 1024 * 4;           // 1024 words of bit rows (expander scratch)
constexpr int OFF_RED = OFF_ROWS + 1024 * 4;                // 16 warps x 8 doubles
constexpr int OFF_BAR = OFF_RED + 16 * 8 * 8;               // barriers + scalars
constexpr int kSmemBytes = OFF_BAR + 256;

struct Params {
  const uint32_t *phi;   // [n_graphs][S][W]
  const uint8_t *a_img;  // [n_mtiles][16 KB]
  const float *lut;      // [64][64]
  double *partials;      // [n_ntiles*G][n_chunks]
  int n_graphs, S, W, E, G, n_ntiles, n_mtiles, tpc, n_chunks, n_items, nmma, ksteps;
  int CS;                // accumulator columns per graph (>= S; the specialised kernels pad S up to a multiple of 8)
  float k1;              // EX2: x = k1 * acc + bias_j
  float kc;              // EX2: bias_j = kc * c_j - s
  float scale;           // s
  long long *trace;      // diagnostics only: clock64 time line of CTA 0 (NULL normally)
  int debug;             // diagnostics only (NEMB200_DEBUG): 1 skip epilogue math, 2 skip MMA, 4 skip TMA, 8 skip expansion
};

// ---- PTX wrappers ---------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n.reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
        "selp.u32 %0, 1, 0, p;\n}"
        : "=r"(done)
        : "r"(bar), "r"(parity), "r"(20000u)  // suspend-time hint (ns): sleep in hardware instead of spinning
        : "memory");
  } while (!done);
}
// non-blocking probe: the predicate can be consumed much later, hiding the ~150-cycle mbarrier latency
__device__ __forceinline__ uint32_t mbar_test(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n.reg .pred p;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n}"
      : "=r"(done)
      : "r"(bar), "r"(parity)
      : "memory");
  return done;
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major SWIZZLE_128B shared-memory descriptor (cute::UMMA::SmemDescriptor): start>>4 | LBO=1<<16 |
// SBO=(1024>>4)<<32 | version=1<<46 | layout SWIZZLE_128B=2<<61
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ uint32_t sw128(uint32_t row, uint32_t byte_in_row) {
  return (row >> 3) * 1024u + (row & 7u) * 128u + ((((byte_in_row >> 4) ^ (row & 7u)) << 4) | (byte_in_row & 15u));
}

#define TMEM_LD32(taddr, r)                                                                                            \
  asm volatile(                                                                                                        \
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"                   \
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"                                        \
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),     \
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),          \
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),         \
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])                       \
      : "r"(taddr))
#define TMEM_LD16(taddr, r)                                                                                         \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];" \
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),   \
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) \
               : "r"(taddr))
#define TMEM_LD8(taddr, r)                                                                   \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"      \
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) \
               : "r"(taddr))
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ float ex2f(float x) {
  float v;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(v) : "f"(x));
  return v;
}
// log2 of a positive normal float: exponent split off in integer arithmetic, lg2.approx on [1,2)
__device__ __forceinline__ float log2_split(float x) {
  uint32_t b = __float_as_uint(x);
  int e = (int)(b >> 23) - 127;
  float m = __uint_as_float((b & 0x007FFFFFu) | 0x3F800000u), l;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(m));
  return (float)e + l;
}

enum { EPI_LUT = 0, EPI_EX2 = 1, EPI_EX2_2PASS = 2 };

// diagnostics (compile with -DNEMB200_TRACE): time line slots [role 0..3][tile 0..255][event 0..3]
#ifdef NEMB200_TRACE
#define DBG(flag) (p.debug & (flag))
#else
#define DBG(flag) false
#endif
#ifndef NEMB200_TRACE
#define TRACE(role, tile, ev)
#else
#define TRACE(role, tile, ev)                                                                   \
  do {                                                                                           \
    if (p.trace != nullptr && blockIdx.x == 0 && (tile) < 256) p.trace[((role) * 256 + (tile)) * 4 + (ev)] = clock64(); \
  } while (0)
#endif

// ---- generic epilogue (runtime S): one graph's columns [0, S) of one accumulator tile ---------------
// Returns log2( 1 + sum_j 2^x_j ) for this thread's E-gene.
template <int EPI, int CH>
__device__ __forceinline__ float graph_lse2(uint32_t taddr, int S, const float *sBias, const Params &p) {
  static_assert(CH == 16 || CH == 8, "chunk");
  float acc0 = 0.f, acc1 = 0.f, m = 0.f;
  if (EPI == EPI_EX2_2PASS) {
    for (int c = 0; c < S; c += CH) {  // pass 1: running max of x_j (and 0 for the null site)
      int cs = min(c, S - CH), skip = c - cs;
      uint32_t r[CH];
      if (CH == 16) TMEM_LD16(taddr + cs, r); else TMEM_LD8(taddr + cs, r);
      tmem_wait_ld();
#pragma unroll
      for (int i = 0; i < CH; ++i)
        if (i >= skip) m = fmaxf(m, fmaf((float)(int)r[i], p.k1, sBias[cs + i]));
    }
  }
  for (int c = 0; c < S; c += CH) {
    int cs = min(c, S - CH), skip = c - cs;  // the last chunk is shifted back so it never over-reads
    uint32_t r[CH];
    if (CH == 16) TMEM_LD16(taddr + cs, r); else TMEM_LD8(taddr + cs, r);
    tmem_wait_ld();
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      float v;
      if (EPI == EPI_LUT) v = lds_f32(r[i]);
      else if (EPI == EPI_EX2) v = ex2f(fmaf((float)(int)r[i], p.k1, sBias[cs + i]));
      else v = ex2f(fmaf((float)(int)r[i], p.k1, sBias[cs + i] - m));
      if (i >= skip) { if (i & 1) acc1 += v; else acc0 += v; }
    }
  }
  if (EPI == EPI_EX2_2PASS) return m + log2_split(ex2f(-m) + acc0 + acc1);
  return log2_split(ex2f(-p.scale) + acc0 + acc1) + p.scale;
}

template <int EPI>
__device__ __forceinline__ void tile_epilogue_generic(uint32_t taddr, const float *sBias, const Params &p, bool valid,
                                                      int ng_mine, double (&part)[kMaxG / 4]) {
  const int S = p.S, bstride = kBiasStride * (S > 64 ? 2 : 1);
#pragma unroll
  for (int g = 0; g < kMaxG / 4; ++g) {
    if (g < ng_mine) {
      float l = S >= 16 ? graph_lse2<EPI, 16>(taddr + g * p.CS, S, sBias + g * bstride, p)
                        : graph_lse2<EPI, 8>(taddr + g * p.CS, S, sBias + g * bstride, p);
      if (valid) part[g] += (double)l;
    }
  }
}

// ---- specialised epilogue (compile-time S): straight-line code, software-pipelined over tiles ----------
// A warpgroup owns NCOL = (G/4)*S_T accumulator columns of every tile, split in chunk A = the first 32
// columns and chunk B = the rest.  Sites j < NLUT of every graph go through the shared-memory exp table
// (LSU pipe), the others through I2F + FFMA + ex2.approx (ALU/FMA/MUFU pipes).  Per tile:
//     [A(t) in registers]  issue TMEM load of B(t)  | math on A(t)
//     [B(t) in registers]  release the accumulator buffer, wait for tile t+1, issue load of A(t+1) | math on B(t)
// so every TMEM load and the MMA + mbarrier round trip of the next tile overlap this warp's own math.
#define TMEM_LDN(n, taddr, r)                        \
  do {                                               \
    if ((n) > 16) TMEM_LD32(taddr, r);               \
    else if ((n) > 8) TMEM_LD16(taddr, r);           \
    else TMEM_LD8(taddr, r);                         \
  } while (0)

// Which attachment sites of a graph go through the exp table (true) and which through ex2.approx (false).
// All warps of a CTA run in near lockstep, so the two kinds are interleaved in groups of four sites (a
// float4 of biases) - otherwise the LSU and MUFU pipes would be used one after the other, not together.
// NLUT == S_T: table only.  Otherwise every third group of four sites is an ex2 group.
template <int S_T, int NLUT>
__host__ __device__ constexpr bool is_lut_site(int j) { return NLUT >= S_T ? true : ((j >> 2) % 3) != 2; }

// The logarithm + float64 accumulation of a finished graph is deferred (pend_*) into the next graph's
// look-up stream so that its MUFU / F2F / DADD latency chain is off the per-tile critical path.
struct Pending { float sum; int g; bool valid; };
__device__ __forceinline__ void finalize_pending(Pending &pd, const Params &p, double (&part)[kMaxG / 4]) {
  if (pd.valid) {
    const float l = log2_split(pd.sum) + p.scale;
#pragma unroll
    for (int g = 0; g < kMaxG / 4; ++g)
      if (g == pd.g) part[g] += (double)l;
  }
  pd.valid = false;
}

template <int S_T, int NLUT, int C0, int C1>
__device__ __forceinline__ void static_cols(const uint32_t (&r)[32], const float *sBias, const Params &p, bool valid,
                                            int ng_mine, float (&acc)[4], float4 &b4, float null_term, Pending &pd,
                                            double (&part)[kMaxG / 4]) {
#pragma unroll
  for (int col = C0; col < C1; ++col) {
    const int g = col / S_T, j = col % S_T;
    if (j == 0) { acc[0] = null_term; acc[1] = 0.f; acc[2] = 0.f; acc[3] = 0.f; }
    float v;
    if (is_lut_site<S_T, NLUT>(j)) {
      v = lds_f32(r[col - C0]);
    } else {
      // (re)load the float4 of biases that holds site j (ex2 groups start on a 16-byte boundary)
      if ((j & 3) == 0 || col == C0) b4 = *reinterpret_cast<const float4 *>(sBias + g * kBiasStride + (j & ~3));
      const int w = j & 3;
      const float b = w == 0 ? b4.x : (w == 1 ? b4.y : (w == 2 ? b4.z : b4.w));
      v = ex2f(fmaf((float)(int)r[col - C0], p.k1, b));
    }
    if (j == 8 || (S_T <= 8 && j == S_T - 1)) finalize_pending(pd, p, part);  // previous graph, mid-stream
    acc[j & 3] += v;
    if (j == S_T - 1) {
      if (S_T <= 8) finalize_pending(pd, p, part);
      pd.sum = (acc[0] + acc[1]) + (acc[2] + acc[3]);
      pd.g = g;
      pd.valid = valid && g < ng_mine;
    }
  }
}

template <int S_T, int NLUT>
__device__ __forceinline__ void item_tiles_static(const Params &p, uint32_t tmem_cols, const float *sBias, int ng_mine,
                                                  int row0, int nt, uint32_t it0, uint32_t bar_tfull, uint32_t bar_tempty,
                                                  double (&part)[kMaxG / 4]) {
  constexpr int G = ((256 / S_T) < kMaxG ? (256 / S_T) : kMaxG) & ~3;
  constexpr int NCOL = (G / 4) * S_T;  // this warpgroup's quarter of the tile's graphs
  constexpr int NA = NCOL < 32 ? NCOL : 32, NB = NCOL - NA;
  static_assert(NCOL <= 64, "a warpgroup's share must fit two 32-column TMEM loads");
  const int lane = threadIdx.x & 31;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
  const float null_term = ex2f(-p.scale);
  uint32_t ra[32], rb[32];
  uint32_t buf = it0 & 1u;
  mbar_wait(bar_tfull + 8 * buf, (it0 >> 1) & 1u);
  tc_fence_after();
  TMEM_LDN(NA, tmem_cols + buf * 256u, ra);
#ifdef NEMB200_TRACE
  const int trole = (threadIdx.x == 128) ? 0 : (threadIdx.x == 128 + 256 + 32 ? 1 : -1);
#define ETRACE(ev)                                                                                     \
  do {                                                                                                 \
    if (p.trace != nullptr && blockIdx.x == 0 && trole >= 0 && it0 + t < 256)                          \
      p.trace[4096 + ((trole * 256 + (int)(it0 + t)) * 8) + (ev)] = clock64();                          \
  } while (0)
#else
#define ETRACE(ev)
#endif
  Pending pd = {0.f, 0, false};
  for (int t = 0; t < nt; ++t) {
    const bool valid = row0 + t * 128 < p.E;
    ETRACE(0);
    tmem_wait_ld();  // A(t) is in registers
    ETRACE(1);
    if (NB > 0) TMEM_LDN(NB, tmem_cols + buf * 256u + 32u, rb);
    static_cols<S_T, NLUT, 0, NA>(ra, sBias, p, valid, ng_mine, acc, b4, null_term, pd, part);
    ETRACE(2);
    const uint32_t itn = it0 + (uint32_t)t + 1u, nbuf = itn & 1u;
    uint32_t ready = 0;
    if (t + 1 < nt) ready = mbar_test(bar_tfull + 8 * nbuf, (itn >> 1) & 1u);  // probe early, consume after the release
    if (NB > 0) tmem_wait_ld();  // B(t) is in registers: this warp no longer needs the accumulator buffer
    ETRACE(3);
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_tempty + 8 * buf);
    ETRACE(4);
    if (t + 1 < nt) {
      buf = nbuf;
      if (!ready) mbar_wait(bar_tfull + 8 * buf, (itn >> 1) & 1u);
      ETRACE(5);
      tc_fence_after();
      TMEM_LDN(NA, tmem_cols + buf * 256u, ra);
    }
    ETRACE(6);
    if (NB > 0) static_cols<S_T, NLUT, NA, NCOL>(rb, sBias, p, valid, ng_mine, acc, b4, null_term, pd, part);
    ETRACE(7);
  }
  finalize_pending(pd, p, part);
#undef ETRACE
}

template <int EPI, int S_T, int NLUT>
__global__ void __launch_bounds__(kThreads, 1) k_score_tc(const Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t sbase = smem_u32(smem);
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + OFF_BAR);
  // barrier map: a_full[4] a_empty[4] t_full[2] t_empty[2] b_full[2] e_done[2]
  const uint32_t bar_afull = sbase + OFF_BAR, bar_aempty = bar_afull + 8 * kStages, bar_tfull = bar_aempty + 8 * kStages,
                 bar_tempty = bar_tfull + 16, bar_bfull = bar_tempty + 16, bar_edone = bar_bfull + 16;
  volatile uint32_t *s_tmem = reinterpret_cast<volatile uint32_t *>(bars + 2 * kStages + 8);
  uint32_t *sRows = reinterpret_cast<uint32_t *>(smem + OFF_ROWS);
  double *sRed = reinterpret_cast<double *>(smem + OFF_RED);

  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) { mbar_init(bar_afull + 8 * i, 1); mbar_init(bar_aempty + 8 * i, 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_tfull + 8 * i, 1); mbar_init(bar_tempty + 8 * i, 16);
      mbar_init(bar_bfull + 8 * i, 2); mbar_init(bar_edone + 8 * i, 16);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32((const void *)s_tmem)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (EPI == EPI_LUT)
    for (int i = threadIdx.x; i < kLutBytes / 4; i += kThreads) reinterpret_cast<float *>(smem + OFF_LUT)[i] = p.lut[i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  const int S = p.S, W = p.W, G = p.G;
  // instruction descriptor (cute::UMMA::InstrDescriptor): D=S32 (2<<4), A=B=U8 (0), K-major, N>>3 at 17, M>>4 at 24
  const uint32_t idesc = (2u << 4) | ((uint32_t)(p.nmma >> 3) << 17) | ((128u >> 4) << 24);
  const int n_my_items = p.n_items > (int)blockIdx.x ? (p.n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  if (warp == 0) {
    // =============================== TMA producer: A tiles of every item, in order ===================
    if (lane == 0) {
      uint32_t it = 0;
      for (int k = 0; k < n_my_items; ++k) {
        const int item = (int)blockIdx.x + k * (int)gridDim.x;
        const int chunk = item / p.n_ntiles;
        const int m0 = chunk * p.tpc, nt = min(p.tpc, p.n_mtiles - m0);
        for (int t = 0; t < nt; ++t, ++it) {
          const uint32_t st = it % kStages, ph = (it / kStages) & 1u;
          mbar_wait(bar_aempty + 8 * st, ph ^ 1u);
          TRACE(0, it, 0);
          if (DBG(4)) { mbar_arrive(bar_afull + 8 * st); continue; }
          mbar_expect_tx(bar_afull + 8 * st, kTileBytes);
          bulk_g2s(sbase + OFF_A + st * kTileBytes, p.a_img + (size_t)(m0 + t) * kTileBytes, kTileBytes, bar_afull + 8 * st);
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer =======================================================
    if (lane == 0) {
      uint32_t it = 0;
      for (int k = 0; k < n_my_items; ++k) {
        const int item = (int)blockIdx.x + k * (int)gridDim.x;
        const int chunk = item / p.n_ntiles;
        const int m0 = chunk * p.tpc, nt = min(p.tpc, p.n_mtiles - m0);
        const uint32_t bb = k & 1u, bbph = (k >> 1) & 1u;
        mbar_wait(bar_bfull + 8 * bb, bbph);  // this item's B operand is expanded
        const uint32_t b0 = sbase + OFF_B + bb * 2 * kTileBytes;
        for (int t = 0; t < nt; ++t, ++it) {
          const uint32_t st = it % kStages, ph = (it / kStages) & 1u, buf = it & 1u, bph = (it >> 1) & 1u;
          TRACE(1, it, 0);
          mbar_wait(bar_afull + 8 * st, ph);
          TRACE(1, it, 1);
          mbar_wait(bar_tempty + 8 * buf, bph ^ 1u);
          TRACE(1, it, 2);
          tc_fence_after();
          const uint32_t a0 = sbase + OFF_A + st * kTileBytes;
          for (int ks = 0; ks < (DBG(2) ? 0 : p.ksteps); ++ks)
            umma_i8(tmem_base + buf * 256u, umma_desc(a0 + ks * 32), umma_desc(b0 + ks * 32), idesc, ks > 0 ? 1u : 0u);
          umma_commit(bar_aempty + 8 * st);  // A slot reusable once these MMAs retire
          umma_commit(bar_tfull + 8 * buf);  // accumulator ready for the epilogue
          TRACE(1, it, 3);
        }
      }
    }
  } else if (warp < 4) {
    // =============================== B expanders (64 threads, one item ahead) ==========================
    const int te = threadIdx.x - 64;
    for (int k = 0; k < n_my_items; ++k) {
      const int item = (int)blockIdx.x + k * (int)gridDim.x;
      const int ntile = item % p.n_ntiles;
      const int g0 = ntile * G, ng = min(G, p.n_graphs - g0);
      const uint32_t bb = k & 1u;
      if (k >= 2) mbar_wait(bar_edone + 8 * bb, ((k - 2) >> 1) & 1u);  // epilogue finished the item that used this buffer
      const int nwords = ng * S * W;
      for (int i = te; i < nwords; i += 64) sRows[i] = p.phi[(size_t)g0 * S * W + i];
      asm volatile("bar.sync 5, 64;" ::: "memory");
      uint8_t *brow = smem + OFF_B + bb * 2 * kTileBytes;
      float *bias = reinterpret_cast<float *>(smem + OFF_BIAS) + bb * 1024;
      const int bstride = kBiasStride * (S > 64 ? 2 : 1);
      const int CS = p.CS;
      // padding sites (S <= j < CS) of the specialised kernels look up a table entry that is zero:
      // offset 128*m with m odd is entry [q=(m-1)/2][p=32], zero as soon as 32 + (m-1)/2 > S
      const uint32_t zero_m = S < 32 ? 1u : (uint32_t)(2 * (S - 32) + 3);
      for (int col = te; col < (DBG(8) ? 0 : 256); col += 64) {
        const int gi = col / CS, j = col - gi * CS;
        uint32_t anc[4] = {0u, 0u, 0u, 0u};
        int cnt = 0;
        const bool used = gi < ng && j < S;
        if (used) {
          const uint32_t *rows = sRows + gi * S * W;
          const int jw = j >> 5, jb = j & 31;
          for (int a = 0; a < S; ++a) {
            uint32_t bit = ((rows[a * W + jw] >> jb) & 1u) | (a == j ? 1u : 0u);
            anc[a >> 5] |= bit << (a & 31);
            cnt += (int)bit;
          }
        }
        // table look-up sites get the extra K row S: B = lut_base/128 (A holds 128 there), so that the
        // accumulator is the shared-memory address of lut[q][p]; unused columns point at lut[0][0]
        bool lut_site = EPI == EPI_LUT && gi < G;
        if constexpr (S_T > 0) lut_site = lut_site && is_lut_site<S_T, NLUT>(j);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          uint32_t half = (anc[c >> 1] >> ((c & 1) * 16)) & 0xFFFFu;
          uint4 v;
          v.x = (((half & 0xFu) * 0x00204081u) & 0x01010101u) << 1;
          v.y = ((((half >> 4) & 0xFu) * 0x00204081u) & 0x01010101u) << 1;
          v.z = ((((half >> 8) & 0xFu) * 0x00204081u) & 0x01010101u) << 1;
          v.w = ((((half >> 12) & 0xFu) * 0x00204081u) & 0x01010101u) << 1;
          if (lut_site && c == (S >> 4)) {
            uint32_t extv = ((sbase + OFF_LUT) >> 7) + (j >= S ? zero_m : 0u);
            if (extv > 255u) asm volatile("trap;");  // the table must sit in the first 32 KB of shared memory
            uint32_t ext = extv << ((S & 3) * 8);
            int wsel = (S >> 2) & 3;
            if (wsel == 0) v.x |= ext; else if (wsel == 1) v.y |= ext; else if (wsel == 2) v.z |= ext; else v.w |= ext;
          }
          *reinterpret_cast<uint4 *>(brow + sw128((uint32_t)col, (uint32_t)c * 16u)) = v;
        }
        if (gi < G && j < kBiasStride * 2) bias[gi * bstride + j] = p.kc * (float)cnt - p.scale;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_bfull + 8 * bb);
      asm volatile("bar.sync 5, 64;" ::: "memory");  // sRows is reused by the next item
    }
  } else {
    // =============================== epilogue: 4 warpgroups share every tile, a quarter of its graphs each ===
    const int wg = (warp - 4) >> 2, quarter = warp & 3, tw = threadIdx.x - 128 - wg * 128;
    const int gq = G >= 4 ? (G >> 2) : 1;                        // graphs per warpgroup (G == 2: warpgroups 2,3 idle)
    const int bstride = kBiasStride * (S > 64 ? 2 : 1);
    uint32_t it = 0;
    for (int k = 0; k < n_my_items; ++k) {
      const int item = (int)blockIdx.x + k * (int)gridDim.x;
      const int chunk = item / p.n_ntiles, ntile = item % p.n_ntiles;
      const int m0 = chunk * p.tpc, nt = min(p.tpc, p.n_mtiles - m0);
      const int gl = wg * gq;                                     // first graph (within the tile) of this warpgroup
      const int g0 = ntile * G + gl;
      const int ng_mine = gl < G ? max(0, min(gq, p.n_graphs - g0)) : 0;
      const uint32_t bb = k & 1u, bbph = (k >> 1) & 1u;
      const float *bias = reinterpret_cast<const float *>(smem + OFF_BIAS) + bb * 1024 + gl * bstride;
      mbar_wait(bar_bfull + 8 * bb, bbph);  // bias values of this item are written
      double part[kMaxG / 4];
#pragma unroll
      for (int g = 0; g < kMaxG / 4; ++g) part[g] = 0.0;
      if (S_T > 0 && ng_mine > 0) {
        if constexpr (S_T > 0)
          item_tiles_static<S_T, NLUT>(p, tmem_base + (uint32_t)(gl * p.CS) + ((uint32_t)(quarter * 32) << 16), bias, ng_mine,
                                       m0 * 128 + quarter * 32 + lane, nt, it, bar_tfull, bar_tempty, part);
        it += (uint32_t)nt;
      } else {
        for (int t = 0; t < nt; ++t, ++it) {
          const uint32_t buf = it & 1u, bph = (it >> 1) & 1u;
          mbar_wait(bar_tfull + 8 * buf, bph);
          tc_fence_after();
          if (ng_mine > 0 && !DBG(1)) {
            const uint32_t taddr = tmem_base + buf * 256u + (uint32_t)(gl * p.CS) + ((uint32_t)(quarter * 32) << 16);
            const bool valid = (m0 + t) * 128 + quarter * 32 + lane < p.E;
            tile_epilogue_generic<EPI>(taddr, bias, p, valid, ng_mine, part);
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_tempty + 8 * buf);
        }
      }
      // ---- per-item reduction inside the warpgroup -> one double per (graph, chunk)
#pragma unroll
      for (int g = 0; g < kMaxG / 4; ++g) {
        if (g < ng_mine) {
          double v = part[g];
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
          if (lane == 0) sRed[(warp - 4) * 8 + g] = v;
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_edone + 8 * bb);  // bias / B buffer of this item may be recycled
      asm volatile("bar.sync %0, 128;" ::"r"(wg + 1) : "memory");
      if (tw < ng_mine) {
        double v = 0.0;
        for (int w4 = 0; w4 < 4; ++w4) v += sRed[(wg * 4 + w4) * 8 + tw];
        p.partials[(size_t)(g0 + tw) * p.n_chunks + chunk] = v;
      }
      asm volatile("bar.sync %0, 128;" ::"r"(wg + 1) : "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base));
}

__global__ void k_finalize(const double *partials, int n_slots, const double *sumN, int n_graphs, double *scores) {
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n_graphs) return;
  double v = 0.0;
  for (int c = 0; c < n_slots; ++c) v += partials[(size_t)g * n_slots + c];
  scores[g] = *sumN + NEMB200_LN2 * v;
}

struct Plan { int n_ntiles, tpc, n_chunks, n_items; };
Plan make_plan(const nemb200_dataset *h, int n, int nsm) {
  Plan pl;
  pl.n_ntiles = (n + h->G - 1) / h->G;
  // Items are statically strided over the CTAs, so pick the E chunking whose item count wastes the least
  // of the last round (items are equal-sized); prefer >= 6 tiles per item to amortise the B expansion.
  int best_chunks = 1;
  double best_cost = 1e30;
  for (int nc = 1; nc <= h->n_mtiles; ++nc) {
    int tpc = (h->n_mtiles + nc - 1) / nc;
    int chunks = (h->n_mtiles + tpc - 1) / tpc;
    if (chunks != nc) continue;
    long items = (long)pl.n_ntiles * chunks;
    long rounds = (items + nsm - 1) / nsm;
    double cost = (double)rounds * (tpc + 1.5);  // +1.5 tiles of per-item overhead
    if (cost < best_cost - 1e-9) { best_cost = cost; best_chunks = chunks; }
  }
  pl.n_chunks = best_chunks;
  pl.tpc = (h->n_mtiles + pl.n_chunks - 1) / pl.n_chunks;
  pl.n_chunks = (h->n_mtiles + pl.tpc - 1) / pl.tpc;
  pl.n_items = pl.n_ntiles * pl.n_chunks;
  return pl;
}

int num_sms() {
  static int nsm = 0;
  if (nsm == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || nsm <= 0)
      nsm = 148;
  }
  return nsm;
}

template <typename K>
int launch_kernel(K kern, const Params &p, int grid, cudaStream_t st) {
  NEMB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
  kern<<<grid, kThreads, kSmemBytes, st>>>(p);
  NEMB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace

size_t score_tc_workspace_bytes(const nemb200_dataset *h, int n) {
  Plan pl = make_plan(h, n, num_sms());
  return 256 + sizeof(double) * (size_t)pl.n_ntiles * h->G * pl.n_chunks + 8 * 8192;  // + diagnostics time line
}

int launch_score_tc(const nemb200_dataset *h, const uint32_t *phi, int n, double *scores, void *ws, size_t ws_bytes,
                    int impl, cudaStream_t st) {
  NEMB_REQUIRE(h->S >= 8 && h->S <= 128, "tensor-core score needs 8 <= S <= 128 (S=%d)", h->S);
  if (impl == NEMB200_SCORE_AUTO)
    impl = h->lut_valid ? NEMB200_SCORE_TC_LUT : (h->fixed_valid ? NEMB200_SCORE_TC_EX2 : NEMB200_SCORE_TC_EX2_2PASS);
  NEMB_REQUIRE(impl != NEMB200_SCORE_TC_LUT || h->lut_valid, "exp-table epilogue not valid for this dataset (S=%d)", h->S);
  NEMB_REQUIRE(impl != NEMB200_SCORE_TC_EX2 || h->fixed_valid, "fixed-scale ex2 epilogue not valid for this dataset");
  NEMB_REQUIRE(ws_bytes >= score_tc_workspace_bytes(h, n), "score workspace too small");
  if (n == 0) return 0;
  const int nsm = num_sms();
  Plan pl = make_plan(h, n, nsm);
  Params p;
  p.phi = phi; p.a_img = h->a_lut; p.lut = h->lut;
  p.partials = reinterpret_cast<double *>(reinterpret_cast<uint8_t *>(ws) + 256);
  p.n_graphs = n; p.S = h->S; p.W = h->W; p.E = h->E; p.G = h->G;
  p.n_ntiles = pl.n_ntiles; p.n_mtiles = h->n_mtiles; p.tpc = pl.tpc; p.n_chunks = pl.n_chunks; p.n_items = pl.n_items;
  p.CS = h->CS;
  p.nmma = (h->G * h->CS + 15) / 16 * 16;
  p.ksteps = (h->S + (impl == NEMB200_SCORE_TC_LUT ? 1 : 0) + 31) / 32;
  // accumulator a = 4p + 256q = 256c - 252p  ->  x = alpha2*p + beta2*q = k1*a + kc*c
  const double kappa2 = h->alpha2 - h->beta2;
  p.k1 = (float)(-kappa2 / 252.0);
  p.kc = (float)(256.0 * kappa2 / 252.0 + h->beta2);
  p.scale = impl == NEMB200_SCORE_TC_EX2_2PASS ? 0.f : (float)h->scale;
  const int grid = pl.n_items < nsm ? pl.n_items : nsm;
  int rc;
  static const char *dbg_env = getenv("NEMB200_DEBUG");
  p.debug = dbg_env ? atoi(dbg_env) : 0;
  p.trace = (p.debug & 32) ? reinterpret_cast<long long *>(reinterpret_cast<uint8_t *>(ws) + 256 + sizeof(double) * (size_t)pl.n_ntiles * h->G * pl.n_chunks) : nullptr;
  if (impl == NEMB200_SCORE_TC_LUT) {
    // the dataset fixed the column stride: S itself for the four headline shapes, S rounded up to 8 otherwise
    switch (h->CS == h->S ? h->S : -h->CS) {
      case 10: rc = launch_kernel(k_score_tc<EPI_LUT, 10, 10>, p, grid, st); break;
      case 20: rc = launch_kernel(k_score_tc<EPI_LUT, 20, 20>, p, grid, st); break;
      case 30: rc = launch_kernel(k_score_tc<EPI_LUT, 30, 30>, p, grid, st); break;
      case 50: rc = launch_kernel(k_score_tc<EPI_LUT, 50, 50>, p, grid, st); break;
      case 8: case -8: rc = launch_kernel(k_score_tc<EPI_LUT, 8, 8>, p, grid, st); break;
      case 16: case -16: rc = launch_kernel(k_score_tc<EPI_LUT, 16, 16>, p, grid, st); break;
      case 24: case -24: rc = launch_kernel(k_score_tc<EPI_LUT, 24, 24>, p, grid, st); break;
      case 32: case -32: rc = launch_kernel(k_score_tc<EPI_LUT, 32, 32>, p, grid, st); break;
      case 40: case -40: rc = launch_kernel(k_score_tc<EPI_LUT, 40, 40>, p, grid, st); break;
      case 48: case -48: rc = launch_kernel(k_score_tc<EPI_LUT, 48, 48>, p, grid, st); break;
      case 56: case -56: rc = launch_kernel(k_score_tc<EPI_LUT, 56, 56>, p, grid, st); break;
      case -64: rc = launch_kernel(k_score_tc<EPI_LUT, 64, 64>, p, grid, st); break;
      default: rc = launch_kernel(k_score_tc<EPI_LUT, 0, 0>, p, grid, st); break;
    }
  } else if (impl == NEMB200_SCORE_TC_EX2) {
    rc = launch_kernel(k_score_tc<EPI_EX2, 0, 0>, p, grid, st);
  } else if (impl == NEMB200_SCORE_TC_EX2_2PASS) {
    rc = launch_kernel(k_score_tc<EPI_EX2_2PASS, 0, 0>, p, grid, st);
  } else {
    nemb200_set_error("unknown score impl %d", impl);
    return -1;
  }
  if (rc) return rc;
  k_finalize<<<(n + 255) / 256, 256, 0, st>>>(p.partials, p.n_chunks, h->sumN, n, scores);
  NEMB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

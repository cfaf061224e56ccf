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
// "EX2_2PASS" subtracts the running maximum first (any dynamic range).  For every S <= 63 the table epilogue is
// compiled for a fixed column stride (straight-line, software-pipelined over tiles).  "PAIR" (AUTO of the sweep entries for 44 <= S <= 50): a planner
// merges pairs of sites so that one look-up serves two of them (k_pair_plan below).
// No B x S x E intermediate ever reaches HBM: per (graph, CTA piece of its N tile) one double is written.
//
// Roles (640 threads, no CTA-wide barrier in the main loop): warp 0 = TMA producer (A tiles, 4-stage
// mbarrier ring), warp 1 = MMA issuer + TMEM owner (2 x 256-column accumulators), warps 2-3 = B-operand
// expanders (double-buffered, run one work item ahead; the first item is expanded by all warps >= 2),
// warps 4-19 = four epilogue warpgroups.  All four drain the SAME accumulator buffer (each takes a quarter of
// the tile's graphs) while the MMA of the next tile fills the other buffer.  The flattened (N tile, E tile) grid
// is cut into one contiguous range per CTA (Walker); for small batches the last CTA to finish reduces the
// partial sums (fused finalize).
#include <stdlib.h>

#include <type_traits>

#include "common.cuh"
#include <vector>

namespace {

// Epilogue warpgroups of the specialised table kernel.  Five warpgroups (five 50-column graphs in the 256-column
// accumulator, 20 epilogue warps, 80 registers) were measured SLOWER than four (111 vs 105 us at S=50, E=10k,
// 1024 graphs): the kernel is bound by shared-memory bandwidth (look-ups + MMA operand reads + TMA writes), not
// by the number of warps in flight, so the knob stays at four.
__host__ __device__ constexpr int epi_warpgroups(int S_T) { return (void)S_T, 4; }
// graphs per MMA N tile of the specialised kernels
__host__ __device__ constexpr int tile_graphs(int S_T) {
  return epi_warpgroups(S_T) == 5 ? 5 : (((256 / S_T) < 16 ? (256 / S_T) : 16) & ~3);
}
__host__ __device__ constexpr int cta_threads(int S_T) { return 128 + 128 * (S_T > 0 ? epi_warpgroups(S_T) : 4); }
constexpr int kTileBytes = 128 * 128;  // one operand atom: 128 rows x 128 B (SWIZZLE_128B)
constexpr int kLutBytes = 64 * 256;
constexpr int kMaxG = 16;
constexpr int kBiasStride = 64;        // floats per graph in the bias array (S <= 128 uses two slots)

// shared memory carve-up (offsets from the 1024-aligned base).  PAIR = the pair-merged table kernel (below):
// three A stages, no bias array, a small single-site table and two buffers of per-item pair tables.
constexpr int kPairCS = 32;                 // accumulator columns per graph of the pair kernel
constexpr int kPairG = 256 / kPairCS;       // graphs per MMA N tile
constexpr int kPairCap = 1568;              // table entries per graph (planner cap)
constexpr int kPairLim = 256;               // table entries per pair (keeps every B weight <= 128)
constexpr int kPairZero = 2040;             // float index (in the single-site table region) of a permanent 0
constexpr int kPlanColBytes = 80;           // planner output per accumulator column: 64 B-operand bytes + 16 bytes of meta
constexpr int kLutmStride = NEMB200_LUTM_STRIDE;  // floats per level of the multi-level table (n(n+1)/2 + p, n <= 125)
template <bool PAIR, bool LUT3 = false>
struct Lay {
  static constexpr int STAGES = (PAIR || LUT3) ? 3 : 4;
  static constexpr int LUT = 0;                                   // standard kernel: base must be < 32 KB and 128-aligned
  static constexpr int LUT_BYTES = LUT3 ? NEMB200_LUTM_LEVELS * kLutmStride * 4 : (PAIR ? 8192 : kLutBytes);
  static constexpr int B = LUT + LUT_BYTES;                       // 2 buffers x (256 x 128 B)
  static constexpr int A = B + 2 * 2 * kTileBytes;                // STAGES x 16 KB
  static constexpr int BIAS = A + STAGES * kTileBytes;            // 2 buffers x 1024 floats (ex2 epilogues)
  // (multi-level table kernel: the same region holds the level sums two warpgroups exchange, [slot 2][pair 2][level 3][half 2][128] floats)
  static constexpr int ROWS = BIAS + (PAIR ? 0 : (LUT3 ? 2 * 2 * 3 * 2 * 128 * 4 : 2 * 1024 * 4));   // 1024 words of bit rows (+ 512 words of scan scratch)
  static constexpr int RED = ROWS + 1024 * 4 + (PAIR ? 2048 : 0); // 20 warps x 8 doubles
  static constexpr int BAR = RED + 20 * 8 * 8;                    // barriers + scalars
  static constexpr int TAB = BAR + 256;                           // PAIR: 2 buffers of pair tables
  static constexpr int TAB_BYTES = PAIR ? kPairG * kPairCap * 4 : 0;
  static constexpr int BYTES = TAB + 2 * TAB_BYTES;
};
static_assert(Lay<true>::BYTES <= 232448 && Lay<false>::BYTES <= 232448 && Lay<false, true>::BYTES <= 232448, "shared memory budget");

struct Params {
  const uint32_t *phi;   // [n_graphs][S][W]
  const uint8_t *a_img;  // [n_mtiles][16 KB]
  const float *lut;      // [64][64]
  double *partials;      // [n_ntiles*G][n_slots]
  int n_graphs, S, W, E, G, n_mtiles, n_slots, nmma, ksteps;
  int CS;                // accumulator columns per graph (>= S; the specialised kernels pad S up to a multiple of 8)
  float k1;              // EX2: x = k1 * acc + bias_j
  float kc;              // EX2: bias_j = kc * c_j - s
  float scale;           // s
  float lvl_scale[3];    // EPI_LUTM: scale of each table level, highest first
  const uint8_t *tile_first;  // EPI_LUTM: per E tile, the first level its E-genes can need
  // pair kernel: per graph kPairCS columns of (site, partner or 255); 255,255 = padding column
  const uint8_t *plan;
  const float *tables;   // pair kernel: per graph kPairCap floats (the graph's pair tables, built by the planner)
  // list mode (fallback pass of the pair kernel): graph ids and their count live in device memory
  const int *glist;
  int *n_dev;          // list mode: number of graphs in glist (reset to 0 by the fused finalize: the planner counts from 0)
  const int *fbslot;   // pair kernel with fused finalize: graphs with fbslot[g] >= 0 belong to the list-mode launch that follows
  // fused finalize: the last CTA to finish sums the partials of every graph (self-resetting counter); NULL = separate k_finalize
  int late_wait;       // dependent launch with the caller's NEMB200_SCORE_INPUTS_READY promise: wait for the previous grid only before the first write
  int *done_counter;
  const double *sumN;
  double *scores;
  long long *trace;      // diagnostics only: clock64 time line of CTA 0 (NULL normally)
  int close;             // 1: the input graphs are NOT transitively closed - the expanders close them (warp Warshall) before use
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
// one lane of a converged warp (elect.sync): the idiom ptxas turns into `ELECT` + predicated issue
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
  return pred != 0;
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
template <int OFF>
__device__ __forceinline__ float lds_f32_off(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(addr), "n"(OFF));
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

enum { EPI_LUT = 0, EPI_EX2 = 1, EPI_EX2_2PASS = 2, EPI_LUTM = 3 };
#ifndef NEMB200_OLD_EPILOGUE
#define NEMB200_OLD_EPILOGUE 0  // 1: the first (scalar FADD, branchy) specialised epilogue, kept for A/B timing
#endif
#ifndef NEMB200_LDS_BATCH
#define NEMB200_LDS_BATCH 1  // 1: a chunk's look-ups all issue before the first is consumed (0: ptxas' own schedule, for A/B timing)
#endif
#ifndef NEMB200_CA
#define NEMB200_CA 32   // lean epilogue: columns of chunk A for shares wider than 32 columns
#endif
#ifndef NEMB200_NA1
#define NEMB200_NA1 12  // lean epilogue: look-ups of chunk A before the accumulator buffer is handed back
#endif
#ifndef NEMB200_EXP
#define NEMB200_EXP 0  // development skip-experiments (wrong results): 1 no LDS, 2 no TMEM loads in the tile loop, 4 no log/fp64, 16 no TMA copies, 32 one MMA k-step, 64 no MMA
#endif

// Named barrier of one epilogue warpgroup (ids 1..5).  The id is an immediate on purpose: with a register operand
// ptxas reserves all 16 hardware barriers for the kernel, and then no CTA that uses __syncthreads() (the MCMC accept
// kernel of the other chain group) can be co-resident with a score CTA.
__device__ __forceinline__ void wg_bar_sync(int wg) {
  switch (wg) {
    case 0: asm volatile("bar.sync 1, 128;" ::: "memory"); break;
    case 1: asm volatile("bar.sync 2, 128;" ::: "memory"); break;
    case 2: asm volatile("bar.sync 3, 128;" ::: "memory"); break;
    case 3: asm volatile("bar.sync 4, 128;" ::: "memory"); break;
    default: asm volatile("bar.sync 5, 128;" ::: "memory"); break;
  }
}

// ---- work distribution ---------------------------------------------------------------------------
// The (N tile, E tile) grid is flattened N-tile-major and cut into `ctas` equal contiguous ranges, one per
// CTA, so every CTA gets the same number of tiles (+-1) whatever the batch size.  A CTA's range is a
// sequence of "items" (runs of E tiles inside one N tile: one B-operand expansion each); an N tile cut by CTA
// boundaries gives one partial sum per piece ("slot" = CTA index - first CTA of that N tile).
constexpr int kShortFirstItem = 16;  // tiles: below this the CTA's second operand is expanded in the prologue as well
constexpr int kMinTilesPerCta = 20;  // a CTA gets at least ceil(n_mtiles / 20) tiles: at most 22 slots per N tile
constexpr int kMaxSlots = 24;
__host__ __device__ inline int effective_ctas(unsigned T, int n_mtiles, int max_ctas) {
  const unsigned pmin = (unsigned)(n_mtiles + kMinTilesPerCta - 1) / kMinTilesPerCta;
  unsigned c = T / (pmin > 0 ? pmin : 1u);
  if (c < 1) c = 1;
  return (int)(c < (unsigned)max_ctas ? c : (unsigned)max_ctas);
}
// first flat tile of CTA c: ranges [c*T/ctas, (c+1)*T/ctas); T * ctas < 2^32 is checked by the host
// (32-bit arithmetic: the host guarantees T * ctas + T < 2^32; 64-bit divisions are emulated and slow on the device)
__host__ __device__ inline unsigned cta_first_tile(unsigned c, unsigned T, unsigned ctas) { return (c * T) / ctas; }
__host__ __device__ inline int cta_of_tile(unsigned x, unsigned T, unsigned ctas) {  // owner of flat tile x
  return (int)(((x + 1u) * ctas + T - 1u) / T) - 1;
}
__host__ __device__ inline int slots_of_ntile(int ntile, int n_mtiles, unsigned T, int ctas) {
  const unsigned lo = (unsigned)ntile * (unsigned)n_mtiles;
  return cta_of_tile(lo + n_mtiles - 1, T, ctas) - cta_of_tile(lo, T, ctas) + 1;
}
struct Item { int ntile, m0, nt; };
struct Walker {  // two registers of state: every role walks the same item sequence
  unsigned pos, end;
  __device__ __forceinline__ Walker(int n_ntiles, int n_mtiles, int max_ctas, int cta) {
    const unsigned T = (unsigned)n_ntiles * (unsigned)n_mtiles;
    const int ctas = effective_ctas(T, n_mtiles, max_ctas);
    pos = cta < ctas ? cta_first_tile(cta, T, ctas) : 0u;
    end = cta < ctas ? cta_first_tile(cta + 1, T, ctas) : 0u;
  }
  __device__ __forceinline__ bool next(Item &it, int n_mtiles) {
    if (pos >= end) return false;
    it.ntile = (int)(pos / (unsigned)n_mtiles);
    it.m0 = (int)(pos - (unsigned)it.ntile * (unsigned)n_mtiles);
    const unsigned rem = end - pos;
    it.nt = (unsigned)(n_mtiles - it.m0) < rem ? n_mtiles - it.m0 : (int)rem;
    pos += (unsigned)it.nt;
    return true;
  }
};
// partial-sum slot of this CTA inside N tile `ntile`
__device__ __forceinline__ int item_slot(int ntile, int n_ntiles, int n_mtiles, int max_ctas, int cta) {
  const unsigned T = (unsigned)n_ntiles * (unsigned)n_mtiles;
  return cta - cta_of_tile((unsigned)ntile * (unsigned)n_mtiles, T, effective_ctas(T, n_mtiles, max_ctas));
}

// diagnostics (compile with -DNEMB200_TRACE): time line slots [role 0..3][tile 0..255][event 0..3]
#ifdef NEMB200_TRACE
#define DBG(flag) (p.debug & (flag))
#else
#define DBG(flag) false
#endif
// development: -DNEMB200_PHASES + NEMB200_DEBUG=32 -> %globaltimer stamps of CTA 0's phases at trace[8000 + k]
#ifdef NEMB200_PHASES
#define PHASE(k)                                                                                          \
  do {                                                                                                    \
    if (p.trace != nullptr && blockIdx.x == 0) {                                                          \
      unsigned long long t_;                                                                              \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                                              \
      p.trace[8000 + (k)] = (long long)t_;                                                                \
    }                                                                                                     \
  } while (0)
#else
#define PHASE(k)
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

// ---- multi-level table epilogue (EPI_LUTM, 16 <= S <= 125, NLEV = 2 or 3 levels): a site's term is
// V(n_j, p) = 2^(x - s_l) with n_j = |ancestors-or-self of j| known per column and p the count of ones among them, so
// the accumulator is made the shared-memory address of entry n_j(n_j+1)/2 + p of the level-0 table; the same entry at
// the lower scales sits kLutmStride floats further each.  One site costs NLEV look-ups and no exponential.  A single
// scale cannot hold the terms (x spans > 700 binary orders at S = 100), but the null site bounds every E-gene's
// largest term from below by 2^0, and the scales (dataset.cu) are spaced so that the FIRST level whose sum reaches
// 2^-90 has neither overflowed nor lost a term within 2^-22 of the sum; the last level always qualifies (it holds
// the null term 2^-89).  Returns log2(1 + sum_j 2^x_j).
template <int MASK, int N>  // MASK: which levels are looked up (bit l = level l)
__device__ __forceinline__ void level_add(const uint32_t (&r)[N], float &a0, float &a1, float &a2) {
#pragma unroll
  for (int i = 0; i < N; ++i) {
    if (MASK & 1) a0 += lds_f32(r[i]);
    if (MASK & 2) a1 += lds_f32_off<kLutmStride * 4>(r[i]);
    if (MASK & 4) a2 += lds_f32_off<2 * kLutmStride * 4>(r[i]);
  }
}
// columns [0, n), n a multiple of 8 (padding columns look up an entry that is zero at every level, so nothing in
// here is predicated: the first version spent more issue slots on ISETP/FSEL than on look-ups).  Keeping the TMEM
// load of the next chunk in flight during this chunk's look-ups (two register buffers) was measured much slower
// (2890 vs 1736 us at S=100, E=100k, 512 graphs).
template <int MASK>
__device__ __forceinline__ void level_sums_mask(uint32_t taddr, int n, float &a0, float &a1, float &a2) {
  int c = 0;
  for (; c + 16 <= n; c += 16) {
    uint32_t r[16];
    TMEM_LD16(taddr + c, r);
    tmem_wait_ld();
    level_add<MASK, 16>(r, a0, a1, a2);
  }
  if (c < n) {  // warp-uniform
    uint32_t r[8];
    TMEM_LD8(taddr + c, r);
    tmem_wait_ld();
    level_add<MASK, 8>(r, a0, a1, a2);
  }
}
// `first` (warp-uniform, from the dataset's per-E-tile range analysis): levels above it cannot be needed for this
// tile and are not looked up - with sorted E-genes most tiles need the last level only (one look-up per site).
// Where the analysis allows a higher level, it is still looked up LAZILY: the last level first, and the next higher one
// only if some E-gene of the warp has a sum of 2^119 or more there - below that, the higher level's sum (2^210 times
// smaller, term by term) is under 2^-91 and the selection rule of levels_log (first level, from the highest scale down,
// whose sum reaches 2^-90) rejects it anyway, so the result is bit-identical to looking everything up.  At config 4
// (S=100, E=100 000) 36 % of the tiles allow level 0, ~40 % of those (per graph) really need it: 1.36 -> 1.15 look-ups
// per site.
constexpr float kLevelHi = 6.646139978924579e35f;  // 2^119 = 2^(step - T - 1), dataset.cu: step = 210, T = 90
template <int NLEV>
__device__ __forceinline__ void level_sums(uint32_t taddr, int n, int first, float &a0, float &a1, float &a2, const Params &p) {
  a0 = 0.f; a1 = 0.f; a2 = 0.f;
  level_sums_mask<1 << (NLEV - 1)>(taddr, n, a0, a1, a2);
  if (first >= NLEV - 1) return;
  const float last = (NLEV == 2 ? a1 : a2) + ex2f(-p.lvl_scale[NLEV - 1]);
  if (!__any_sync(0xffffffffu, !(last < kLevelHi))) return;
  level_sums_mask<1 << (NLEV - 2)>(taddr, n, a0, a1, a2);
  if (NLEV == 3 && first == 0 && __any_sync(0xffffffffu, !(a1 + ex2f(-p.lvl_scale[1]) < kLevelHi)))
    level_sums_mask<1>(taddr, n, a0, a1, a2);
}
template <int NLEV>
__device__ __forceinline__ float levels_log(float a0, float a1, float a2, const Params &p) {
  // null site x = 0 at each level (flushes to zero where it is far below the level's range)
  a0 += ex2f(-p.lvl_scale[0]); a1 += ex2f(-p.lvl_scale[1]);
  if (NLEV > 2) a2 += ex2f(-p.lvl_scale[2]);
  const float lo = 8.077935669463161e-28f;  // 2^-90
  if (a0 >= lo) return log2_split(a0) + p.lvl_scale[0];
  if (NLEV == 2 || a1 >= lo) return log2_split(a1) + p.lvl_scale[1];
  return log2_split(a2) + p.lvl_scale[2];
}
template <int NLEV>
__device__ __forceinline__ float graph_lse_levels(uint32_t taddr, int n_cols, int first, const Params &p) {
  float a0, a1, a2;
  level_sums<NLEV>(taddr, n_cols, first, a0, a1, a2, p);
  return levels_log<NLEV>(a0, a1, a2, p);
}
// Two graphs per tile (S > 64) would leave two of the four epilogue warpgroups idle: instead warpgroups 2h and 2h+1
// share graph h, each sums the table entries of one half of its sites, and they hand each other their level sums through
// shared memory (block = tile parity; a 256-thread named barrier per tile for the last level, which every tile looks up).
// The accumulator hand-back (t_empty counts all 16 warps) keeps a warpgroup from running two tiles ahead, so two blocks
// suffice.  A higher level (see level_sums) is looked up by a PAIR OF WARPS - warp w of both warpgroups holds the same 32
// E-genes and reaches the same decision from the same total - which then meet on an mbarrier of their own (64 arrivals;
// named barriers are scarce: the MCMC accept kernel must stay co-resident).  xch: [level 3][half 2][128] floats.
template <int NLEV>
__device__ __forceinline__ void tile_epilogue_split(uint32_t taddr_graph, const Params &p, bool valid, int first, int half, int pair,
                                                    float *xch, int tw, uint32_t bar_x, uint32_t &ph_x, double &part0) {
  const int Sh = ((p.CS + 15) >> 4) << 3;  // both halves are multiples of 8 columns
  const uint32_t ta = taddr_graph + (half ? Sh : 0);
  const int nc = half ? p.CS - Sh : Sh;
  float tot[3] = {0.f, 0.f, 0.f};
  // one level: my half's sum, the partner's through shared memory, total = even half + odd half (the same bits in both)
  auto level = [&](auto lv_c, bool first_exchange) {
    constexpr int LV = decltype(lv_c)::value;
    float m0 = 0.f, m1 = 0.f, m2 = 0.f;
    level_sums_mask<1 << LV>(ta, nc, m0, m1, m2);
    const float mine = LV == 0 ? m0 : (LV == 1 ? m1 : m2);
    xch[(LV * 2 + half) * 128 + tw] = mine;
    if (first_exchange) {
      if (pair == 0) asm volatile("bar.sync 8, 256;" ::: "memory"); else asm volatile("bar.sync 9, 256;" ::: "memory");
    } else {
      mbar_arrive(bar_x);
      mbar_wait(bar_x, ph_x);
      ph_x ^= 1u;
    }
    const float other = xch[(LV * 2 + (half ^ 1)) * 128 + tw];
    tot[LV] = half ? other + mine : mine + other;
  };
  level(std::integral_constant<int, NLEV - 1>{}, true);
  if (first < NLEV - 1 && __any_sync(0xffffffffu, !(tot[NLEV - 1] + ex2f(-p.lvl_scale[NLEV - 1]) < kLevelHi))) {
    level(std::integral_constant<int, NLEV - 2>{}, false);
    if constexpr (NLEV == 3) {
      if (first == 0 && __any_sync(0xffffffffu, !(tot[1] + ex2f(-p.lvl_scale[1]) < kLevelHi))) level(std::integral_constant<int, 0>{}, false);
    }
  }
  if (!half) {
    const float l = levels_log<NLEV>(tot[0], tot[1], tot[2], p);
    if (valid) part0 += (double)l;
  }
}

template <int EPI, int NLEV>
__device__ __forceinline__ void tile_epilogue_generic(uint32_t taddr, const float *sBias, const Params &p, bool valid,
                                                      int ng_mine, double (&part)[kMaxG / 4], int first = 0) {
  const int S = p.S, bstride = kBiasStride * (S > 64 ? 2 : 1);
#pragma unroll
  for (int g = 0; g < kMaxG / 4; ++g) {
    if (g < ng_mine) {
      float l;
      if constexpr (EPI == EPI_LUTM) l = graph_lse_levels<NLEV>(taddr + g * p.CS, p.CS, first, p);
      else l = S >= 16 ? graph_lse2<EPI, 16>(taddr + g * p.CS, S, sBias + g * bstride, p)
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
// NLUT >= S_T: table only.  Otherwise NLUT is a period P: of every P groups of four sites the last one is an ex2 group
// (a fraction 1/P of the sites leaves the LSU pipe for the MUFU pipe).
constexpr int kMixPeriod = 4;  // the shipped mix (measured at S=50, E=10k, 1024 graphs: P=3 90.1 us, 4 87.1, 5 87.3, table only 92.0)
template <int S_T, int NLUT>
__host__ __device__ constexpr bool is_lut_site(int j) { return NLUT >= S_T ? true : ((j >> 2) % NLUT) != NLUT - 1; }
// int -> float without I2F (I2F and ex2 share the XU pipe): for 0 <= a < 2^14, (a << 9) + 0x46800000 is the bit pattern
// of 2^14 + a; the expander folds -k1 * 2^14 into the site's bias
constexpr uint32_t kMagic14 = 0x46800000u;

// The logarithm + float64 accumulation of a finished graph is deferred (pend_*) into the next graph's
// look-up stream so that its MUFU / F2F / DADD latency chain is off the per-tile critical path.
struct Pending { float sum; int g; bool valid; };
__device__ __forceinline__ void finalize_pending(Pending &pd, const Params &p, double (&part)[kMaxG / 4]) {
  if ((NEMB200_EXP & 4) && pd.sum != 123.f) { pd.valid = false; return; }
  if (pd.valid) {
    const float l = log2_split(pd.sum) + p.scale;
#pragma unroll
    for (int g = 0; g < kMaxG / 4; ++g)
      if (g == pd.g) part[g] += (double)l;
  }
  pd.valid = false;
}

template <int S_T, int NLUT, int C0, int C1, int SHIFT = 0>
__device__ __forceinline__ void static_cols(const uint32_t (&r)[32], const float *sBias, const Params &p, bool valid,
                                            int ng_mine, float (&acc)[4], float4 &b4, float null_term, Pending &pd,
                                            double (&part)[kMaxG / 4]) {
#pragma unroll
  for (int col = C0; col < C1; ++col) {
    const int g = col / S_T, j = col % S_T;
    if (j == 0) { acc[0] = null_term; acc[1] = 0.f; acc[2] = 0.f; acc[3] = 0.f; }
    float v;
    if (is_lut_site<S_T, NLUT>(j)) {
      v = (NEMB200_EXP & 1) ? __uint_as_float(r[col - C0 + SHIFT]) : lds_f32(r[col - C0 + SHIFT]);
    } else {
      // (re)load the float4 of biases that holds site j (ex2 groups start on a 16-byte boundary)
      if ((j & 3) == 0 || col == C0) b4 = *reinterpret_cast<const float4 *>(sBias + g * kBiasStride + (j & ~3));
      const int w = j & 3;
      const float b = w == 0 ? b4.x : (w == 1 ? b4.y : (w == 2 ? b4.z : b4.w));
      v = ex2f(fmaf(__uint_as_float((r[col - C0 + SHIFT] << 9) + kMagic14), p.k1, b));
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
  constexpr int G = tile_graphs(S_T);
  constexpr int NCOL = (G / epi_warpgroups(S_T)) * S_T;  // this warpgroup's share of the tile's graphs
  constexpr int NA = NCOL < 32 ? NCOL : 32, NB = NCOL - NA;
  constexpr int NA1 = NB > 0 ? (NA < 12 ? NA : 12) : 0;  // look-ups of chunk A done before the early hand-back of the buffer
  static_assert(NCOL <= 64, "a warpgroup's share must fit two 32-column TMEM loads");
  // The second load is 8, 16 or 32 columns wide and ENDS at the warpgroup's last column (it never reads past
  // the warpgroup's share, so the last warpgroup stays inside the 256-column buffer); its first BSH registers
  // repeat columns of chunk A and are ignored.
  constexpr int NBL = NB > 16 ? 32 : (NB > 8 ? 16 : 8);
  constexpr int BSTART = NB > 0 ? NCOL - NBL : 0, BSH = NB > 0 ? NA - BSTART : 0;
  static_assert(NB == 0 || BSTART >= 0, "second TMEM load would start before the warpgroup's share");
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
    if (NB > 0 && (!(NEMB200_EXP & 2) || t == 0)) TMEM_LDN(NB, tmem_cols + buf * 256u + (uint32_t)BSTART, rb);
    // The accumulator buffer goes back to the MMA issuer as soon as chunk B(t) has landed in registers, which is a
    // few look-ups into chunk A (the load was issued just above): the MMA of tile t+2 then has most of a tile
    // period to complete, and nothing this warp waits for later in the tile depends on the other 15 warps.
    static_cols<S_T, NLUT, 0, NA1>(ra, sBias, p, valid, ng_mine, acc, b4, null_term, pd, part);
    ETRACE(2);
    if (NB > 0) tmem_wait_ld();  // B(t) is in registers: this warp no longer needs the accumulator buffer
    ETRACE(3);
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_tempty + 8 * buf);
    ETRACE(4);
    static_cols<S_T, NLUT, NA1, NA, NA1>(ra, sBias, p, valid, ng_mine, acc, b4, null_term, pd, part);
    const uint32_t itn = it0 + (uint32_t)t + 1u, nbuf = itn & 1u;
    if (t + 1 < nt) {
      buf = nbuf;
      mbar_wait(bar_tfull + 8 * buf, (itn >> 1) & 1u);  // (complete long ago in steady state: ~90 clk)
      ETRACE(5);
      tc_fence_after();
      if (!(NEMB200_EXP & 2)) TMEM_LDN(NA, tmem_cols + buf * 256u, ra);  // lands while chunk B is being looked up
    }
    ETRACE(6);
    if (NB > 0) static_cols<S_T, NLUT, NA, NCOL, BSH>(rb, sBias, p, valid, ng_mine, acc, b4, null_term, pd, part);
    ETRACE(7);
  }
  finalize_pending(pd, p, part);
#undef ETRACE
}

// ---- lean table epilogue (every site through the table): the kernel is bound by the instructions its 16 epilogue
// warps issue (4 warps per scheduler, each in order), so this path counts them: look-ups are consumed in PAIRS by one
// packed FADD2 (add.f32x2), the tile loop is unrolled over the two accumulator buffers so that every barrier address
// and TMEM column is an immediate, the per-graph logarithm is branch-free (FSEL instead of a divergent block) and
// takes the float exponent by a magic-number add instead of I2F.  Per (warp, tile): S_T LDS + S_T/2 FADD2 + ~35.
__device__ __forceinline__ uint64_t pack2(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void unpack2(uint64_t v, float &a, float &b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) { uint64_t d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

struct LeanPending { float sum; int g; bool ok; };
// log2 of a positive normal float + cexp (the scale): biased exponent taken as 2^23 + e by one shift-add (no I2F),
// mantissa through lg2.approx
__device__ __forceinline__ void lean_finalize(LeanPending &pd, float cexp, double (&part)[kMaxG / 4]) {
  if ((NEMB200_EXP & 4) && pd.sum != 123.f) return;
  const uint32_t b = __float_as_uint(pd.sum);
  const float ef = __uint_as_float((b >> 23) + 0x4B000000u);  // 2^23 + biased exponent, exact
  float m = __uint_as_float((b & 0x007FFFFFu) | 0x3F800000u), l;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(m));
  // same roundings as log2_split(x) + scale (the fused MCMC kernel computes exactly that: both order-MCMC paths
  // then take bit-identical accept decisions): (e + lg2 m) + scale, with e = ef - (2^23 + 127) exact
  l = ((ef - 8388735.f) + l) + cexp;
  l = pd.ok ? l : 0.f;
#pragma unroll
  for (int g = 0; g < kMaxG / 4; ++g)
    if (g == pd.g) part[g] += (double)l;                     // pd.g is a compile-time constant at every call site
}

// all look-ups of the chunk are ISSUED before the first one is consumed: the kernel lives on the number of shared-memory
// loads in flight (16 in-order warps; ptxas on its own batches them in eights, which caps the LSU at ~0.6 loads/clk)
template <int N>
__device__ __forceinline__ void pin_regs(float (&v)[N]) {
  if constexpr (N >= 16) {
    float(&a)[16] = reinterpret_cast<float(&)[16]>(v[0]);
    asm volatile("" : "+f"(a[0]), "+f"(a[1]), "+f"(a[2]), "+f"(a[3]), "+f"(a[4]), "+f"(a[5]), "+f"(a[6]), "+f"(a[7]), "+f"(a[8]),
                      "+f"(a[9]), "+f"(a[10]), "+f"(a[11]), "+f"(a[12]), "+f"(a[13]), "+f"(a[14]), "+f"(a[15]));
    if constexpr (N > 16) pin_regs<N - 16>(reinterpret_cast<float(&)[N - 16]>(v[16]));
  } else if constexpr (N >= 4) {
    asm volatile("" : "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3]));
    if constexpr (N > 4) pin_regs<N - 4>(reinterpret_cast<float(&)[N - 4]>(v[4]));
  } else if constexpr (N >= 2) {
    asm volatile("" : "+f"(v[0]), "+f"(v[1]));
  }
}

// number of ex2 sites per graph of a mixed instantiation (NLUT < S_T: of every NLUT groups of four sites the last one), and
// the rank of site j among them
template <int S_T, int NLUT>
__host__ __device__ constexpr int n_ex2_sites() { int n = 0; for (int j = 0; j < S_T; ++j) n += is_lut_site<S_T, NLUT>(j) ? 0 : 1; return n; }
template <int S_T, int NLUT>
__host__ __device__ constexpr int ex2_rank(int j) { int n = 0; for (int q = 0; q < j; ++q) n += is_lut_site<S_T, NLUT>(q) ? 0 : 1; return n; }

template <int S_T, int NLUT, int C0, int C1, int SHIFT, int NB>
__device__ __forceinline__ void lean_cols(const uint32_t (&r)[32], float null_term, float cexp, bool ok_tile, int ng_mine,
                                          uint64_t &a0, uint64_t &a1, LeanPending &pd, double (&part)[kMaxG / 4],
                                          const float (&bq)[NB], float k1) {
  static_assert(S_T % 2 == 0 && C0 % 2 == 0 && C1 % 2 == 0, "column pairs must not straddle graphs or chunks");
  if constexpr (C1 > C0) {
    float v[C1 - C0];
#pragma unroll
    for (int i = 0; i < C1 - C0; ++i) {
      constexpr int kMixP = (NEMB200_EXP & 128) ? 4 : ((NEMB200_EXP & 256) ? 3 : ((NEMB200_EXP & 512) ? 2 : 0));  // timing probe: 1/P of the pairs through ex2 (wrong bias)
      const int jj = (C0 + i) % S_T;
      if (kMixP > 0 && ((jj >> 1) % kMixP) == kMixP - 1) v[i] = ex2f(fmaf(__uint_as_float((r[i + SHIFT] << 9) + kMagic14), -1e-4f, 1.0f));
      else if (!is_lut_site<S_T, NLUT>(jj))  // mixed instantiation: this site goes through the MUFU pipe (x = k1 * acc + bias_j)
        v[i] = ex2f(fmaf(__uint_as_float((r[i + SHIFT] << 9) + kMagic14), k1, bq[ex2_rank<S_T, NLUT>(jj) < NB ? ex2_rank<S_T, NLUT>(jj) : 0]));
      else v[i] = (NEMB200_EXP & 1) ? __uint_as_float((r[i + SHIFT] & 0xFFFFu) | 0x3F800000u) : lds_f32(r[i + SHIFT]);
    }
    if constexpr ((NEMB200_EXP & 1024) != 0) {  // sensitivity probe: every look-up issued twice
      float w[C1 - C0];
#pragma unroll
      for (int i = 0; i < C1 - C0; ++i) w[i] = lds_f32_off<4>(r[i + SHIFT]);
      pin_regs<C1 - C0>(w);
    }
#if NEMB200_LDS_BATCH
    pin_regs<C1 - C0>(v);
#endif
    if constexpr ((NEMB200_EXP & 4096) != 0) {  // sensitivity probe: twice the packed adds
#pragma unroll
      for (int i = 0; i + 1 < C1 - C0; i += 2) { if (i & 2) a1 = add2(a1, pack2(v[i], v[i + 1])); else a0 = add2(a0, pack2(v[i], v[i + 1])); }
    }
#pragma unroll
    for (int col = C0; col < C1; col += 2) {
      const int g = col / S_T, j = col % S_T;
      if (j == 0) { a0 = pack2(null_term, 0.f); a1 = pack2(0.f, 0.f); }
      const uint64_t v2 = pack2(v[col - C0], v[col + 1 - C0]);
      if (j == (S_T > 8 ? 8 : S_T - 2)) lean_finalize(pd, cexp, part);  // the previous graph's logarithm, mid-stream
      if (j & 2) a1 = add2(a1, v2); else a0 = add2(a0, v2);
      if (j == S_T - 2) {
        float x0, y0, x1, y1;  // (acc0 + acc1) + (acc2 + acc3): the association of every other epilogue of this file
        unpack2(a0, x0, y0); unpack2(a1, x1, y1);
        pd.sum = (x0 + y0) + (x1 + y1); pd.g = g; pd.ok = ok_tile && g < ng_mine;
      }
    }
  }
}

template <int S_T, int NLUT>
__device__ __forceinline__ void item_tiles_lean(const Params &p, uint32_t tmem_cols, const float *sBias, int ng_mine, int row0, int nt,
                                                uint32_t it0, uint32_t bar_tfull, uint32_t bar_tempty, double (&part)[kMaxG / 4]) {
  constexpr int G = tile_graphs(S_T);
  constexpr int NCOL = (G / epi_warpgroups(S_T)) * S_T;
  // chunk A = columns [0, NA), chunk B = [NA, NCOL): two TMEM loads of up to 32 columns per tile, each issued one
  // chunk of look-ups before it is needed (NEMB200_CA: columns of chunk A when the share is wider than 32)
  constexpr int NA = NCOL <= 32 ? NCOL : ((NEMB200_CA >= NCOL - 32 && NEMB200_CA <= 32) ? NEMB200_CA : 32), NB = NCOL - NA;
  // look-ups of chunk A before the accumulator buffer is handed back: 12 at the stride 50 (tuned there), the whole chunk at
  // the stride 32 (the pair-merged kernel: 72.3 us per 1024 graphs against 73.9 with 12, 73.4 with 4, 73.2 with 20)
  constexpr int NA1 = NB > 0 ? (S_T == 32 ? NA : (NA < NEMB200_NA1 ? NA : NEMB200_NA1)) : 0;
  static_assert(NCOL <= 64 && NA <= 32 && NB <= 32 && NA % 2 == 0, "a warpgroup's share must fit two 32-column TMEM loads");
  constexpr int NAL = NA > 16 ? 32 : (NA > 8 ? 16 : 8);
  constexpr int NBL = NB > 16 ? 32 : (NB > 8 ? 16 : 8);
  constexpr int BSTART = NB > 0 ? NCOL - NBL : 0, BSH = NB > 0 ? NA - BSTART : 0;
  const int lane = threadIdx.x & 31;
  const float null_term = ex2f(-p.scale), cexp = p.scale;
  // opaque copies: ptxas otherwise rematerialises the shared-window base (S2R + LEA) in front of every barrier op
  uint32_t bfull = bar_tfull, bempty = bar_tempty, tcols = tmem_cols;
  asm volatile("" : "+r"(bfull), "+r"(bempty), "+r"(tcols));
  uint32_t ra[32], rb[32];
  uint64_t a0 = 0, a1 = 0;
  LeanPending pd = {1.f, 0, false};
  // mixed instantiation (one graph per warpgroup): the biases of the graph's ex2 sites stay in registers for the item
  constexpr int NBQ = NLUT < S_T ? n_ex2_sites<S_T, NLUT>() : 1;
  static_assert(NLUT >= S_T || NCOL == S_T, "the ex2 mix is written for one graph per warpgroup");
  float bq[NBQ];
  bq[0] = 0.f;
  if constexpr (NLUT < S_T) {
#pragma unroll
    for (int j = 0; j < S_T; ++j)
      if (!is_lut_site<S_T, NLUT>(j)) bq[ex2_rank<S_T, NLUT>(j)] = sBias[j];
  }
  const float k1 = p.k1;
  uint32_t ph = (it0 >> 1) & 1u;  // parity of the t_full barriers for the current PAIR of tiles
  // one tile out of accumulator buffer BUF; `more`: another tile follows (its chunk A is fetched before chunk B is used)
  auto tile = [&](auto buf_c, int t, bool more) {
    constexpr uint32_t BUF = decltype(buf_c)::value;
    const bool ok_tile = row0 + t * 128 < p.E;
    tmem_wait_ld();  // A(t) is in registers
    if constexpr (NB > 0) { if (!(NEMB200_EXP & 2)) TMEM_LDN(NB, tcols + BUF * 256u + (uint32_t)BSTART, rb); }
    if constexpr (NB > 0 && (NEMB200_EXP & 2048) != 0) TMEM_LDN(NB, tcols + BUF * 256u + (uint32_t)BSTART, rb);  // sensitivity probe: twice the TMEM loads
    lean_cols<S_T, NLUT, 0, NA1, 0>(ra, null_term, cexp, ok_tile, ng_mine, a0, a1, pd, part, bq, k1);
    if constexpr (NB > 0) tmem_wait_ld();  // B(t) is in registers: this warp no longer needs the accumulator buffer
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(bempty + 8 * BUF);
    lean_cols<S_T, NLUT, NA1, NA, NA1>(ra, null_term, cexp, ok_tile, ng_mine, a0, a1, pd, part, bq, k1);
    if (more) {
      mbar_wait(bfull + 8 * (BUF ^ 1u), BUF ? (ph ^ 1u) : ph);
      tc_fence_after();
      if (!(NEMB200_EXP & 2)) TMEM_LDN(NAL, tcols + (BUF ^ 1u) * 256u, ra);  // lands while chunk B is being looked up
      if constexpr ((NEMB200_EXP & 2048) != 0) TMEM_LDN(NAL, tcols + (BUF ^ 1u) * 256u, ra);
    }
    if constexpr (NB > 0) lean_cols<S_T, NLUT, NA, NCOL, BSH>(rb, null_term, cexp, ok_tile, ng_mine, a0, a1, pd, part, bq, k1);
  };
  int t = 0;
  if (it0 & 1u) {  // the item starts on buffer 1
    mbar_wait(bfull + 8, ph);
    tc_fence_after();
    TMEM_LDN(NAL, tcols + 256u, ra);
    tile(std::integral_constant<uint32_t, 1>{}, 0, nt > 1);
    ph ^= 1u;
    t = 1;
  } else {
    mbar_wait(bfull, ph);
    tc_fence_after();
    TMEM_LDN(NAL, tcols, ra);
  }
  for (; t + 1 < nt; t += 2) {
    tile(std::integral_constant<uint32_t, 0>{}, t, true);
    tile(std::integral_constant<uint32_t, 1>{}, t + 1, t + 2 < nt);
    ph ^= 1u;
  }
  if (t < nt) tile(std::integral_constant<uint32_t, 0>{}, t, false);
  lean_finalize(pd, cexp, part);
}

template <int EPI, int S_T, int NLUT, bool PAIR = false>
__global__ void __launch_bounds__(cta_threads(S_T), 1) k_score_tc(const Params p) {
  constexpr int NWG = S_T > 0 ? epi_warpgroups(S_T) : 4;
  using L = Lay<PAIR, EPI == EPI_LUTM>;
  constexpr int kStages = L::STAGES;
  static_assert(!PAIR || (EPI == EPI_LUT && S_T == kPairCS), "the pair kernel is the 32-column table kernel");
  extern __shared__ __align__(1024) uint8_t smem[];
  // Programmatic dependent launch (the sweep entry points launch with cudaLaunchAttributeProgrammaticStreamSerialization):
  // the next kernel of the stream may be set up while this grid runs, and this grid runs its prologue - barriers, TMEM,
  // the exp table (dataset constants) - while its predecessor drains; griddepcontrol.wait (a no-op for a plain launch)
  // stands before the first access to anything the predecessor may produce or still read: graphs, plan, workspace.
  asm volatile("griddepcontrol.launch_dependents;");
  if (p.n_dev != nullptr) {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (*p.n_dev == 0) return;  // list mode with an empty list: nothing to do (uniform exit)
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) PHASE(0);
  const uint32_t sbase = smem_u32(smem);
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + L::BAR);
  // barrier map: a_full[4] a_empty[4] t_full[2] t_empty[2] b_full[2] e_done[2]
  const uint32_t bar_afull = sbase + L::BAR, bar_aempty = bar_afull + 8 * 4, bar_tfull = bar_aempty + 8 * 4,
                 bar_tempty = bar_tfull + 16, bar_bfull = bar_tempty + 16, bar_edone = bar_bfull + 16;
  [[maybe_unused]] const uint32_t bar_xch = bar_afull + 160;  // EPI_LUTM, split epilogue: [pair 2][warp 4] mbarriers of 64 arrivals (after the 16 above + s_tmem)
  volatile uint32_t *s_tmem = reinterpret_cast<volatile uint32_t *>(bars + 2 * 4 + 8);
  uint32_t *sRows = reinterpret_cast<uint32_t *>(smem + L::ROWS);
  double *sRed = reinterpret_cast<double *>(smem + L::RED);

  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) { mbar_init(bar_afull + 8 * i, 1); mbar_init(bar_aempty + 8 * i, 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_tfull + 8 * i, 1); mbar_init(bar_tempty + 8 * i, 4 * NWG);
      mbar_init(bar_bfull + 8 * i, 2); mbar_init(bar_edone + 8 * i, 4 * NWG);
    }
    if constexpr (EPI == EPI_LUTM) {
      for (int i = 0; i < 8; ++i) mbar_init(bar_xch + 8 * i, 64);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32((const void *)s_tmem)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if constexpr (PAIR) {
    // single-site table, compact: tabS[n(n+1)/2 + p] = 2^(alpha2*p + beta2*(n-p) - s) = lut[n-p][p]; the rest of the region is 0
    float *tabS = reinterpret_cast<float *>(smem + L::LUT);
    for (int i = threadIdx.x; i < L::LUT_BYTES / 4; i += cta_threads(S_T)) tabS[i] = 0.f;
    __syncthreads();
    for (int i = threadIdx.x; i < 64 * 64; i += cta_threads(S_T)) {
      const int q = i >> 6, pp = i & 63, n = pp + q;
      if (n <= p.S) tabS[n * (n + 1) / 2 + pp] = p.lut[i];
    }
  } else if (EPI == EPI_LUT) {
    for (int i = threadIdx.x; i < kLutBytes / 4; i += cta_threads(S_T)) reinterpret_cast<float *>(smem + L::LUT)[i] = p.lut[i];
  } else if (EPI == EPI_LUTM) {  // NLUT = number of table levels
    for (int i = threadIdx.x; i < NLUT * kLutmStride; i += cta_threads(S_T)) reinterpret_cast<float *>(smem + L::LUT)[i] = p.lut[i];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // (one-site kernel with the caller's promise that the graphs were complete before the previous kernel of the stream
  //  began: nothing it reads can still be in flight, so it works on while that kernel drains and waits only in front of
  //  its first write to the workspace; the pair kernel needs the planner's output right away)
  const bool late_wait = !PAIR && p.n_dev == nullptr && p.late_wait != 0;
  if (p.n_dev == nullptr && !late_wait) asm volatile("griddepcontrol.wait;" ::: "memory");
  if (threadIdx.x == 0) PHASE(1);
  const uint32_t tmem_base = *s_tmem;
  const int S = p.S, W = p.W, G = p.G;
  // list mode: the graph count (and with it the item count) is only known on the device
  const int n_graphs = p.n_dev != nullptr ? *p.n_dev : p.n_graphs;
  const int n_ntiles = (n_graphs + G - 1) / G;
  // instruction descriptor (cute::UMMA::InstrDescriptor): D=S32 (2<<4), A=B=U8 (0), K-major, N>>3 at 17, M>>4 at 24
  const uint32_t idesc = (2u << 4) | ((uint32_t)(p.nmma >> 3) << 17) | ((128u >> 4) << 24);

  // ---- B-operand expansion of one item by `nthr` cooperating threads (thread index te): bit rows -> u8 columns.
  // The expander warps use it with 64 threads; the FIRST item of a CTA is expanded by all warps >= 2 (nothing
  // else can run before the first operand exists, and 576 threads finish it four times sooner).
  auto expand_std = [&](int g0, int ng, uint32_t bb, int te, int nthr, auto &&sync) {
    const int gw = S * W;  // words per graph
    for (int i = te; i < ng * gw; i += nthr) {
      const int gi = i / gw;
      const int gid = p.glist != nullptr ? p.glist[g0 + gi] : g0 + gi;
      sRows[i] = __ldcg(p.phi + (size_t)gid * gw + (i - gi * gw));  // (L2: a dependent launch may start on an SM whose L1 predates the graphs)
    }
    sync();
    if (p.close && !((NEMB200_EXP & 8192) && nthr > 64)) {
      // fused transitive closure (utils.ancestor, utils.py:37-54, as create_nem calls it, nem_order_mcmc.py:216-221): one
      // warp per staged graph, rows in registers, boolean Warshall with row k broadcast by shuffle - the same
      // warp_closure the closure kernel and the MCMC propose kernel run.  Every CTA that works on an N tile closes
      // its own copy (a few hundred instructions per graph, hidden behind the previous item's tiles).
      const int wi = te >> 5, nw = nthr >> 5, ln = te & 31;
      for (int gi = wi; gi < ng; gi += nw) {
        uint32_t *rows = sRows + gi * gw;
        auto run = [&](auto tw_c) {
          constexpr int TW = decltype(tw_c)::value;
          uint32_t r[TW][TW];
#pragma unroll
          for (int t = 0; t < TW; ++t)
#pragma unroll
            for (int w = 0; w < TW; ++w) r[t][w] = (t * 32 + ln < S) ? rows[(t * 32 + ln) * TW + w] : 0u;
          warp_closure<TW>(r, S);
#pragma unroll
          for (int t = 0; t < TW; ++t)
#pragma unroll
            for (int w = 0; w < TW; ++w)
              if (t * 32 + ln < S) rows[(t * 32 + ln) * TW + w] = r[t][w];
        };
        if (W == 1) run(std::integral_constant<int, 1>{});
        else if (W == 2) run(std::integral_constant<int, 2>{});
        else if (W == 3) run(std::integral_constant<int, 3>{});
        else run(std::integral_constant<int, 4>{});
      }
      sync();
    }
    uint8_t *brow = smem + L::B + bb * 2 * kTileBytes;
    [[maybe_unused]] float *bias = reinterpret_cast<float *>(smem + L::BIAS) + bb * 1024;
    [[maybe_unused]] const int bstride = kBiasStride * (S > 64 ? 2 : 1);
    const int CS = p.CS;
    // ---- phase 1: rows -> columns.  A warp holds a staged graph one row (or W rows) per lane and transposes it as
    // 32 x 32 bit blocks (five shuffle-xor butterfly steps per block): afterwards lane j holds, for column j, j+32, ..,
    // the mask of the rows that have an edge into it - the ancestor mask of site j.  The masks go back in place
    // (sRows[gi][j][w]; the rows are not needed any more).  ~40 instructions per block instead of a 50-iteration
    // bit-gather loop per column.
    {
      const int wi = te >> 5, nw = nthr >> 5, ln = te & 31;
      for (int gi = wi; gi < ng; gi += nw) {
        uint32_t *rows = sRows + gi * gw;
        auto run = [&](auto tw_c) {
          constexpr int TW = decltype(tw_c)::value;
          uint32_t x[TW][TW];  // x[t][w]: word w of row t*32 + lane
#pragma unroll
          for (int t = 0; t < TW; ++t)
#pragma unroll
            for (int w = 0; w < TW; ++w) x[t][w] = (t * 32 + ln < S) ? rows[(t * 32 + ln) * TW + w] : 0u;
          __syncwarp();
#pragma unroll
          for (int t = 0; t < TW; ++t)
#pragma unroll
            for (int w = 0; w < TW; ++w) {
              uint32_t v = x[t][w];
#pragma unroll
              for (int sft = 16; sft > 0; sft >>= 1) {
                const uint32_t m = sft == 16 ? 0x0000FFFFu : (sft == 8 ? 0x00FF00FFu : (sft == 4 ? 0x0F0F0F0Fu : (sft == 2 ? 0x33333333u : 0x55555555u)));
                const uint32_t y = __shfl_xor_sync(0xffffffffu, v, sft);
                v = (ln & sft) ? ((v & ~m) | ((y >> sft) & m)) : ((v & m) | ((y << sft) & ~m));
              }
              x[t][w] = v;  // now: rows t*32.. of column w*32 + lane
            }
          // column j = w*32 + lane: word t of its mask = x[t][w]; the site itself belongs to its ancestor-or-self set
#pragma unroll
          for (int w = 0; w < TW; ++w) {
            const int j = w * 32 + ln;
            if (j < S) {
#pragma unroll
              for (int t = 0; t < TW; ++t) rows[j * TW + t] = x[t][w] | ((t == w) ? (1u << ln) : 0u);
            }
          }
        };
        if (W == 1) run(std::integral_constant<int, 1>{});
        else if (W == 2) run(std::integral_constant<int, 2>{});
        else if (W == 3) run(std::integral_constant<int, 3>{});
        else run(std::integral_constant<int, 4>{});
      }
    }
    sync();
    // ---- phase 2: masks -> operand bytes, one (column, 16-byte chunk of K) per step.  Only the K bytes the MMA reads
    // are written (ksteps x 32 per row).  Padding sites (S <= j < CS) of the specialised kernels look up a table entry
    // that is zero: offset 128*m with m odd is entry [q=(m-1)/2][p=32], zero as soon as 32 + (m-1)/2 > S.
    const uint32_t zero_m = S < 32 ? 1u : (uint32_t)(2 * (S - 32) + 3);
    const int nch = 2 * p.ksteps;
    for (int idx = te; idx < (DBG(8) ? 0 : 256 * nch); idx += nthr) {
      const int col = idx / nch, c = idx - col * nch;
      const int gi = col / CS, j = col - gi * CS;
      const bool used = gi < ng && j < S;
      int cnt = 0;
      uint32_t half = 0;
      if (used) {
        const uint32_t *mk = sRows + gi * gw + j * W;
        for (int w = 0; w < W; ++w) cnt += __popc(mk[w]);
        half = (c >> 1) < W ? ((mk[c >> 1] >> ((c & 1) * 16)) & 0xFFFFu) : 0u;
      }
      // table look-up sites get the extra K row S: B = lut_base/128 (A holds 128 there), so that the
      // accumulator is the shared-memory address of lut[q][p]; unused columns point at lut[0][0]
      bool lut_site = EPI == EPI_LUT && gi < G;
      if constexpr (S_T > 0) lut_site = lut_site && is_lut_site<S_T, NLUT>(j);
      uint4 v;
      constexpr int sh = EPI == EPI_LUTM ? 0 : 1;  // operand value 2 (the A operand tells D=1 from D=0); LUTM: 1 (A = 4*D)
      v.x = (((half & 0xFu) * 0x00204081u) & 0x01010101u) << sh;
      v.y = ((((half >> 4) & 0xFu) * 0x00204081u) & 0x01010101u) << sh;
      v.z = ((((half >> 8) & 0xFu) * 0x00204081u) & 0x01010101u) << sh;
      v.w = ((((half >> 12) & 0xFu) * 0x00204081u) & 0x01010101u) << sh;
      auto put_byte = [&](int k, uint32_t val) {  // byte k of the row, if it lies in this chunk
        if ((k >> 4) == c) {
          const uint32_t e = val << ((k & 3) * 8);
          const int wsel = (k >> 2) & 3;
          if (wsel == 0) v.x |= e; else if (wsel == 1) v.y |= e; else if (wsel == 2) v.z |= e; else v.w |= e;
        }
      };
      if (lut_site) {
        const uint32_t extv = ((sbase + L::LUT) >> 7) + (j >= S ? zero_m : 0u);
        if (extv > 255u) asm volatile("trap;");  // the table must sit in the first 32 KB of shared memory
        put_byte(S, extv);
      }
      if constexpr (EPI == EPI_LUTM) {
        // constant K rows (A operand: 128 in rows S and S+1, 4 in row S+2): the byte address of the column's first
        // table entry, V(n_j, 0) of level 0; padding columns and absent graphs look up an entry that is zero at every level
        const uint32_t addr = sbase + L::LUT + 4u * (uint32_t)(used ? cnt * (cnt + 1) / 2 : NEMB200_LUTM_ZERO);
        const uint32_t q = addr >> 7, b0 = q < 255u ? q : 255u;
        if (q - b0 > 255u) asm volatile("trap;");  // level 0 must sit in the first 64 KB of shared memory
        put_byte(S, b0);
        put_byte(S + 1, q - b0);
        put_byte(S + 2, (addr & 127u) >> 2);
      } else {
        // (specialised mixed kernels convert the accumulator with the 2^14 magic number, see kMagic14)
        if (c == 0 && gi < G && j < kBiasStride * 2)
          bias[gi * bstride + j] = p.kc * (float)cnt - p.scale - ((S_T > 0 && NLUT < S_T) ? p.k1 * 16384.f : 0.f);
      }
      *reinterpret_cast<uint4 *>(brow + sw128((uint32_t)col, (uint32_t)c * 16u)) = v;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  };
  constexpr int kWideThreads = cta_threads(S_T) - 64;  // warps >= 2
  if constexpr (!PAIR) {
    if (warp >= 2) {
      Walker wk0(n_ntiles, p.n_mtiles, (int)gridDim.x, (int)blockIdx.x);
      Item i0;
      if (wk0.next(i0, p.n_mtiles)) {  // CTA-uniform
        const int g0 = i0.ntile * G;
        expand_std(g0, min(G, n_graphs - g0), 0u, (int)threadIdx.x - 64, kWideThreads,
                   [] { asm volatile("bar.sync 7, %0;" ::"n"(kWideThreads) : "memory"); });
        asm volatile("bar.sync 7, %0;" ::"n"(kWideThreads) : "memory");  // every column is written and fenced
        if (warp < 4 && lane == 0) mbar_arrive(bar_bfull);               // the two arrivals the barrier expects
        if (threadIdx.x == 64) PHASE(2);
        // A CTA whose range starts just before an N-tile boundary has a first item of a few tiles only: the two
        // expander warps (lowest scheduling priority of the CTA) cannot finish the second operand behind it and the
        // whole CTA would wait for them (measured: such CTAs finished 6-9 us late).  Expand that one here too.
        Item i1;
        if (i0.nt < kShortFirstItem && wk0.next(i1, p.n_mtiles)) {
          const int g1 = i1.ntile * G;
          expand_std(g1, min(G, n_graphs - g1), 1u, (int)threadIdx.x - 64, kWideThreads,
                     [] { asm volatile("bar.sync 7, %0;" ::"n"(kWideThreads) : "memory"); });
          asm volatile("bar.sync 7, %0;" ::"n"(kWideThreads) : "memory");
          if (warp < 4 && lane == 0) mbar_arrive(bar_bfull + 8);
        }
      }
    }
  }

  constexpr bool kRegSplit = S_T > 0 && NLUT < S_T && !PAIR;
  if (warp < 4) {
  if constexpr (kRegSplit) {
    // the mixed epilogue keeps its ex2 biases in registers: the four non-epilogue warps (one warpgroup: they meet here
    // after the prologue expansion) give registers to the sixteen epilogue warps (CTA pool: 128 x 32 released = 4 x 128 x 8 taken)
    asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
  }
  if (warp == 0) {
    // =============================== TMA producer: A tiles of every item, in order ===================
    // The whole warp walks the loop (uniform control flow: addresses and barrier parities stay in uniform
    // registers) and ONE elected lane issues; a loop under `if (lane == 0)` makes ptxas wrap every bulk copy / MMA
    // in an ELECT + R2UR.BROADCAST retry loop (five register moves per instruction through the vector pipe).
    uint32_t it = 0;
    Walker wk(n_ntiles, p.n_mtiles, (int)gridDim.x, (int)blockIdx.x);
    Item im;
    while (wk.next(im, p.n_mtiles)) {
      const int m0 = im.m0, nt = im.nt;
      for (int t = 0; t < nt; ++t, ++it) {
        const uint32_t st = it % kStages, ph = (it / kStages) & 1u;
        mbar_wait(bar_aempty + 8 * st, ph ^ 1u);
        TRACE(0, it, 0);
        if (elect_one()) {
          if (DBG(4) || (NEMB200_EXP & 16)) {
            mbar_arrive(bar_afull + 8 * st);
          } else {
            const uint32_t tma_bytes = DBG(16) ? kTileBytes / 2 : kTileBytes;  // diagnostics: half the operand traffic (wrong results)
            mbar_expect_tx(bar_afull + 8 * st, tma_bytes);
            bulk_g2s(sbase + L::A + st * kTileBytes, p.a_img + (size_t)(m0 + t) * kTileBytes, tma_bytes, bar_afull + 8 * st);
          }
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer (warp-uniform loop, one elected lane issues) ==========
    uint32_t it = 0;
    Walker wk(n_ntiles, p.n_mtiles, (int)gridDim.x, (int)blockIdx.x);
    Item im;
    for (int k = 0; wk.next(im, p.n_mtiles); ++k) {
      const int nt = im.nt;
      const uint32_t bb = k & 1u, bbph = (k >> 1) & 1u;
      mbar_wait(bar_bfull + 8 * bb, bbph);  // this item's B operand is expanded
      const uint32_t b0 = sbase + L::B + bb * 2 * kTileBytes;
      for (int t = 0; t < nt; ++t, ++it) {
        const uint32_t st = it % kStages, ph = (it / kStages) & 1u, buf = it & 1u, bph = (it >> 1) & 1u;
        TRACE(1, it, 0);
        mbar_wait(bar_afull + 8 * st, ph);
        TRACE(1, it, 1);
        mbar_wait(bar_tempty + 8 * buf, bph ^ 1u);
        TRACE(1, it, 2);
        tc_fence_after();
        const uint32_t a0 = sbase + L::A + st * kTileBytes;
        if (elect_one()) {
          for (int ks = 0; ks < (DBG(2) ? 0 : ((NEMB200_EXP & 32) ? 1 : ((NEMB200_EXP & 64) ? 0 : p.ksteps))); ++ks)
            umma_i8(tmem_base + buf * 256u, umma_desc(a0 + ks * 32), umma_desc(b0 + ks * 32), idesc, ks > 0 ? 1u : 0u);
          umma_commit(bar_aempty + 8 * st);  // A slot reusable once these MMAs retire
          umma_commit(bar_tfull + 8 * buf);  // accumulator ready for the epilogue
        }
        __syncwarp();
        TRACE(1, it, 3);
      }
    }
  } else if (warp < 4) {
    // =============================== B expanders (64 threads, one item ahead) ==========================
    const int te = threadIdx.x - 64;
    Walker wk(n_ntiles, p.n_mtiles, (int)gridDim.x, (int)blockIdx.x);
    Item im;
    [[maybe_unused]] int first_nt = 0;
    for (int k = 0; wk.next(im, p.n_mtiles); ++k) {
      const int ntile = im.ntile;
      const int g0 = ntile * G, ng = min(G, n_graphs - g0);
      const uint32_t bb = k & 1u;
      if (k >= 2) mbar_wait(bar_edone + 8 * bb, ((k - 2) >> 1) & 1u);  // epilogue finished the item that used this buffer
      if (!PAIR && k == 0) { first_nt = im.nt; continue; }        // expanded by all warps above
      if (!PAIR && k == 1 && first_nt < kShortFirstItem) continue;  // ... and so was this one (short first item)
      if constexpr (!PAIR) expand_std(g0, ng, bb, te, 64, [] { asm volatile("bar.sync 6, 64;" ::: "memory"); });
      else asm volatile("bar.sync 6, 64;" ::: "memory");
      uint8_t *brow = smem + L::B + bb * 2 * kTileBytes;
      if constexpr (PAIR) {
        // ---- pair-merged columns: the planner (k_pair_plan) wrote every column's 64 operand bytes and the shape
        // of its table; here the constant K rows get the table's shared-memory address and the pair tables of
        // this item are built (graph gi owns a fixed slot of kPairCap entries in table buffer bb)
        const uint32_t tab_addr = sbase + L::TAB + bb * L::TAB_BYTES, tabs_addr = sbase + L::LUT;
        uint32_t *sTot = sRows;  // table entries per graph of this item
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int col = te + 64 * u, gi = col / kPairCS, c = col - gi * kPairCS;
          uint4 v[4] = {make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0)};
          uint4 meta = make_uint4(0, 0, 0, 0);
          if (gi < ng) {
            const int gid = p.glist != nullptr ? p.glist[g0 + gi] : g0 + gi;
            const uint4 *src = reinterpret_cast<const uint4 *>(p.plan + ((size_t)gid * kPairCS + c) * kPlanColBytes);
            v[0] = src[0]; v[1] = src[1]; v[2] = src[2]; v[3] = src[3]; meta = src[4];
          }
          const int kind = (int)(meta.x & 255u), n1 = (int)((meta.x >> 8) & 255u), nC = (int)(meta.x >> 24),
                    d1 = (int)(meta.y & 255u), d2 = (int)((meta.y >> 8) & 255u);
          const uint32_t toff = (uint32_t)gi * kPairCap + (meta.y >> 16);
          if (kind == 2 && (meta.y >> 16) + (uint32_t)((nC + 1) * (d1 + 1) * (d2 + 1)) > (uint32_t)kPairCap) asm volatile("trap;");
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) *reinterpret_cast<uint4 *>(brow + sw128((uint32_t)col, (uint32_t)cc * 16u)) = v[cc];
          // byte address of the column's table entry 0 -> K rows S..S+7 (x128) and S+8 (x4)
          const uint32_t addr = kind == 2 ? tab_addr + 4u * toff
                                          : (kind == 1 ? tabs_addr + 4u * (uint32_t)(n1 * (n1 + 1) / 2) : tabs_addr + 4u * kPairZero);
          uint32_t q = addr >> 7;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const uint32_t b = q < 255u ? q : 255u;
            q -= b;
            brow[sw128((uint32_t)col, (uint32_t)(S + i))] = (uint8_t)b;
          }
          brow[sw128((uint32_t)col, (uint32_t)(S + 8))] = (uint8_t)((addr & 127u) >> 2);
          if (c == 0) sTot[gi] = meta.z;  // table entries of graph gi (0 for absent / unplanned graphs)
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if constexpr (PAIR) {
        // the planner left every graph's pair tables in global memory: one bulk copy per graph into its slot of
        // table buffer bb, completing on the same barrier as the operand (the epilogue waits on it)
        asm volatile("bar.sync 6, 64;" ::: "memory");  // sTot is complete
        if (warp == 2 && lane == 0) {
          uint32_t bytes = 0;
          for (int gi = 0; gi < ng; ++gi) bytes += (sRows[gi] * 4u + 15u) & ~15u;
          mbar_expect_tx(bar_bfull + 8 * bb, bytes);
          for (int gi = 0; gi < ng; ++gi) {
            const uint32_t nb = (sRows[gi] * 4u + 15u) & ~15u;
            const int gid = p.glist != nullptr ? p.glist[g0 + gi] : g0 + gi;
            if (nb) bulk_g2s(sbase + L::TAB + bb * L::TAB_BYTES + (uint32_t)gi * kPairCap * 4u, p.tables + (size_t)gid * kPairCap, nb,
                             bar_bfull + 8 * bb);
          }
        } else if (lane == 0) {
          mbar_arrive(bar_bfull + 8 * bb);
        }
      } else {
        if (lane == 0) mbar_arrive(bar_bfull + 8 * bb);
      }
      asm volatile("bar.sync 6, 64;" ::: "memory");  // sRows is reused by the next item
    }
  }
  } else {
    // =============================== epilogue: 4 warpgroups share every tile, a quarter of its graphs each ===
    if constexpr (kRegSplit) asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    const int wg = (warp - 4) >> 2, quarter = warp & 3, tw = threadIdx.x - 128 - wg * 128;
    const int gq = G >= NWG ? G / NWG : 1;                       // graphs per warpgroup (G == 2: warpgroups 2,3 idle)
    const bool split = EPI == EPI_LUTM && G == 2;                // ... unless two warpgroups share a graph (tile_epilogue_split)
    [[maybe_unused]] uint32_t ph_xch = 0;                        // phase of this warp pair's exchange mbarrier (split epilogue)
    const int bstride = kBiasStride * (S > 64 ? 2 : 1);
    uint32_t it = 0;
    Walker wk(n_ntiles, p.n_mtiles, (int)gridDim.x, (int)blockIdx.x);
    Item im;
    for (int k = 0; wk.next(im, p.n_mtiles); ++k) {
      const int ntile = im.ntile, m0 = im.m0, nt = im.nt;
      const int gl = split ? (wg >> 1) : wg * gq;                 // first graph (within the tile) of this warpgroup
      const int g0 = ntile * G + gl;
      const int ng_mine = gl < G ? max(0, min(gq, n_graphs - g0)) : 0;
      const uint32_t bb = k & 1u, bbph = (k >> 1) & 1u;
      const float *bias = reinterpret_cast<const float *>(smem + L::BIAS) + bb * 1024 + gl * bstride;
      // partial-sum slot of this CTA inside the item's N tile (two emulated divisions: computed up front by the few
      // threads that store partial sums, not between the item's last tile and its hand-over)
      int slot = 0;
      if (tw < kMaxG / 4) slot = item_slot(ntile, n_ntiles, p.n_mtiles, (int)gridDim.x, (int)blockIdx.x);
      mbar_wait(bar_bfull + 8 * bb, bbph);  // bias values (pair kernel: tables) of this item are written
      if (threadIdx.x == 128 && k == 0) PHASE(3);
      double part[kMaxG / 4];
#pragma unroll
      for (int g = 0; g < kMaxG / 4; ++g) part[g] = 0.0;
      if (S_T > 0 && ng_mine > 0) {
        if constexpr (S_T > 0 && S_T % 2 == 0 && !NEMB200_OLD_EPILOGUE)
          item_tiles_lean<S_T, NLUT>(p, tmem_base + (uint32_t)(gl * p.CS) + ((uint32_t)(quarter * 32) << 16), bias, ng_mine,
                                     m0 * 128 + quarter * 32 + lane, nt, it, bar_tfull, bar_tempty, part);
        else if constexpr (S_T > 0)
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
            if constexpr (EPI == EPI_LUTM) {
              const int first = (int)p.tile_first[m0 + t];
              if (split)
                tile_epilogue_split<NLUT>(taddr, p, valid, first, wg & 1, wg >> 1,
                                          reinterpret_cast<float *>(smem + L::BIAS) + ((buf * 2 + (wg >> 1)) * 6) * 128, tw,
                                          bar_xch + 8u * (uint32_t)((wg >> 1) * 4 + quarter), ph_xch, part[0]);
              else
                tile_epilogue_generic<EPI, NLUT>(taddr, bias, p, valid, ng_mine, part, first);
            } else {
              tile_epilogue_generic<EPI, NLUT>(taddr, bias, p, valid, ng_mine, part);
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_tempty + 8 * buf);
        }
      }
      // ---- per-item reduction inside the warpgroup -> one double per (graph, chunk)
      const int ng_out = (split && (wg & 1)) ? 0 : ng_mine;  // the odd warpgroup of a pair only helps
#pragma unroll
      for (int g = 0; g < kMaxG / 4; ++g) {
        if (g < ng_out) {
          double v = part[g];
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
          if (lane == 0) sRed[(warp - 4) * 8 + g] = v;
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_edone + 8 * bb);  // bias / B buffer of this item may be recycled
      wg_bar_sync(wg);
      if (tw < ng_out) {
        double v = 0.0;
        for (int w4 = 0; w4 < 4; ++w4) v += sRed[(wg * 4 + w4) * 8 + tw];
        if (late_wait) asm volatile("griddepcontrol.wait;" ::: "memory");  // the previous launch's finalize may still read the slots
        p.partials[(size_t)(g0 + tw) * p.n_slots + slot] = v;  // list mode: indexed by list position
        // (no gpu-scope fence per item - it would stall the warpgroup; every thread fences ONCE, before the barrier at
        //  the end of the kernel, see below)
      }
      wg_bar_sync(wg);
    }
  }
  if (threadIdx.x == 128) PHASE(4);
  // fused finalize: the partial sums this CTA stored must be visible gpu-wide before thread 0 reports in (a fence by
  // thread 0 alone does not cover the other threads' stores: each storing thread fences, then the CTA barrier)
  if (p.done_counter != nullptr) __threadfence();
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) PHASE(5);
#ifdef NEMB200_PHASES
  if (threadIdx.x == 0 && p.trace != nullptr) {  // every CTA: when it finished, when it started, on which SM
    unsigned long long t_; unsigned sm_;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));
    asm volatile("mov.u32 %0, %%smid;" : "=r"(sm_));
    p.trace[7600 + blockIdx.x] = (long long)t_;
    p.trace[7800 + blockIdx.x] = (long long)sm_;
  }
#endif
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base));
  if (p.done_counter != nullptr) {
    // ---- fused finalize ("last block" pattern): score = sum_e N[e] + ln2 * sum of the graph's partial sums
    volatile int *s_last = reinterpret_cast<volatile int *>(sRows);
    if (threadIdx.x == 0) {
      if (late_wait) asm volatile("griddepcontrol.wait;" ::: "memory");  // the counter is the previous launch's until that grid is complete
      __threadfence();
      *s_last = atomicAdd(p.done_counter, 1) == (int)gridDim.x - 1;
    }
    __syncthreads();
    if (*s_last) {
      __threadfence();
      const unsigned T = (unsigned)n_ntiles * (unsigned)p.n_mtiles;
      const double sumN = *p.sumN;
      const int ctas_eff = effective_ctas(T, p.n_mtiles, (int)gridDim.x);  // (list mode launches a full grid for a list it cannot see)
      for (int g = threadIdx.x; g < n_graphs; g += cta_threads(S_T)) {
        const int fb = p.fbslot != nullptr ? __ldcg(p.fbslot + g) : -1;  // (in flight together with the partial sums: -0.5 us per call)
        const int cnt = slots_of_ntile(g / G, p.n_mtiles, T, ctas_eff);
        const double *src = p.partials + (size_t)g * p.n_slots;
        // (the first four slots - all there are unless E is huge - are fetched together: one L2 round trip, same order of additions)
        const double s0 = __ldcg(src), s1 = cnt > 1 ? __ldcg(src + 1) : 0.0, s2 = cnt > 2 ? __ldcg(src + 2) : 0.0,
                     s3 = cnt > 3 ? __ldcg(src + 3) : 0.0;
        if (fb >= 0) continue;  // scored by the one-site kernel (list mode)
        double v = s0;
        if (cnt > 1) v += s1;
        if (cnt > 2) v += s2;
        if (cnt > 3) v += s3;
        for (int c = 4; c < cnt && c < p.n_slots; ++c) v += __ldcg(src + c);
        p.scores[p.glist != nullptr ? p.glist[g] : g] = __fma_rn(NEMB200_LN2, v, sumN);  // (explicit: both finalize paths must round alike)
      }
      if (threadIdx.x == 0) {
        *p.done_counter = 0;  // ready for the next launch on this workspace
        if (p.n_dev != nullptr) *p.n_dev = 0;  // every CTA has read the list length; the next planner run appends from 0
      }
#ifdef NEMB200_PHASES
      if (threadIdx.x == 0 && p.trace != nullptr) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); p.trace[8007] = (long long)t_; p.trace[8008] = blockIdx.x; }
#endif
    }
  }
  if (threadIdx.x == 0) PHASE(6);
}

// ---- pair planner --------------------------------------------------------------------------------------
// The epilogue is bound by shared-memory look-ups: one per (attachment site, E-gene).  Two sites j, j' whose
// ancestor sets A, A' overlap can share ONE look-up: with C = A & A', D1 = A \ C, D2 = A' \ C the pair's term
//     2^x_j + 2^x_j' = V(|A|, pc + p1) + V(|A'|, pc + p2),   V(n, p) = 2^(alpha2*p + beta2*(n-p) - s)
// depends on the E-gene only through the three counts pc, p1, p2 of ones among C, D1, D2.  A B-operand column
// with weights 1 on C, w1 = |C|+1 on D1 and w2 = w1*(|D1|+1) on D2 makes the accumulator the mixed-radix index
// pc + w1*p1 + w2*p2 into a per-pair table of (|C|+1)(|D1|+1)(|D2|+1) entries that the expander warps build in
// shared memory (constant K rows add the table's address, as in the single-site kernel).  The planner (one
// warp per graph) pairs sites greedily in order of ancestor count, keeps the S - 32 cheapest pairs (tables of
// at most kPairLim entries, kPairCap per graph) and leaves the other sites single: a 50-site graph becomes 32
// accumulator columns.  Graphs that cannot be planned are appended to a list and scored by the one-site kernel.
constexpr int kPlanWindow = 8;     // partner window in sorted order
constexpr int kPlanThreads = 128;  // one CTA per graph, all graphs of a 1024-graph batch resident at once (7 CTAs per SM)
__global__ void __launch_bounds__(kPlanThreads, 7) k_pair_plan(const uint32_t *__restrict__ phi, int n, int S, int W,
                                                              const float *__restrict__ tabs, uint8_t *__restrict__ plan,
                                                              float *__restrict__ tables, int *__restrict__ fbslot,
                                                              int *__restrict__ flist, int *__restrict__ fcount,
                                                              long long *__restrict__ trace, uint32_t *__restrict__ closed, int late_wait) {
#define PTRACE(k) do { if (trace != nullptr && blockIdx.x == 500 && threadIdx.x == 0) trace[k] = clock64(); } while (0)
  __shared__ uint32_t s_rows[64 * 2];
  __shared__ uint4 s_site[64];      // by site j:            {mask word 0, mask word 1, ancestor count, j}
  __shared__ uint4 s_sorted[64];    // by sorted position r: the same record of the r-th smallest site
  __shared__ uint4 s_col[kPairCS];  // by accumulator column: {kind | ceil(2^16 / w2) << 2, n1 | n2 << 8 | nC << 16, table offset, ceil(2^16 / w1)}
  __shared__ uint4 s_mask[kPairCS][2];  // by accumulator column: {weight-1 mask (2 words), weight-w1 mask}, {weight-w2 mask, w1, w2}
  __shared__ uint8_t s_plan[2 * kPairCS];
  __shared__ __align__(16) uint32_t s_key[64];  // by site: ancestor count << 6 | site
  __shared__ uint32_t s_cost[32][kPlanWindow];  // pair cost of sorted positions (a, a+1+w), 0xFFFF = not allowed: rows 2k | 2k+1 << 16
  __shared__ __align__(16) float s_tab[1440];   // V(n, p) = lut[n-p][p] at n(n+1)/2 + p, n <= S <= 52
  const int g = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
  // (rotating the single-warp roles with the CTA's residency slot was measured slower: the hardware already spreads them)
  const int warp = tid >> 5;
  PTRACE(0);
  // (the table of V is only needed by the last phase: asynchronous copy, waited for there)
  for (int i = tid; i * 4 < (S + 1) * (S + 2) / 2; i += kPlanThreads)
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(s_tab + 4 * i)), "l"(tabs + 4 * i) : "memory");
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;");
  // (see k_score_tc.  Without the caller's promise the graphs may come from the previous kernel of the stream: wait here.
  //  With it, everything up to the first global write - the whole planning - runs while the previous call's pair kernel
  //  drains; that kernel still reads the plan and the tables this one is about to overwrite.)
  if (!late_wait) asm volatile("griddepcontrol.wait;" ::: "memory");
  if (tid < S * W) s_rows[tid] = __ldcg(phi + (size_t)g * S * W + tid);  // 33 <= S <= 52, W = 2 (L2: see k_score_tc)
  __syncthreads();
  if (closed != nullptr) {
    // the batch is not transitively closed yet (nemb200_closure_score_batch): close it here (utils.ancestor, utils.py:37-54,
    // the same warp Warshall as nemb200_closure_u32) and leave the closed rows for the one-site kernel, should the
    // graph end up on its list
    if (warp == 0) {
      uint32_t r[2][2];
#pragma unroll
      for (int t = 0; t < 2; ++t)
#pragma unroll
        for (int w = 0; w < 2; ++w) r[t][w] = (t * 32 + lane < S) ? s_rows[(t * 32 + lane) * 2 + w] : 0u;
      warp_closure<2>(r, S);
#pragma unroll
      for (int t = 0; t < 2; ++t)
        if (t * 32 + lane < S) *reinterpret_cast<uint2 *>(s_rows + (t * 32 + lane) * 2) = make_uint2(r[t][0], r[t][1]);
    }
    __syncthreads();
  }
  PTRACE(1);
  // ancestor-or-self masks (columns of I | Phi) by 32 x 32 bit-block transposes (five shuffle-xor butterfly steps each,
  // like the score kernel's operand expansion): warp w < 2 transposes the two row blocks of column block w, lane l then
  // holds the mask of site 32 w + l
  if (warp < 2) {
    uint32_t x[2];
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      uint32_t v = (t * 32 + lane < S) ? s_rows[(t * 32 + lane) * W + warp] : 0u;
#pragma unroll
      for (int sft = 16; sft > 0; sft >>= 1) {
        const uint32_t m = sft == 16 ? 0x0000FFFFu : (sft == 8 ? 0x00FF00FFu : (sft == 4 ? 0x0F0F0F0Fu : (sft == 2 ? 0x33333333u : 0x55555555u)));
        const uint32_t y = __shfl_xor_sync(0xffffffffu, v, sft);
        v = (lane & sft) ? ((v & ~m) | ((y >> sft) & m)) : ((v & m) | ((y << sft) & ~m));
      }
      x[t] = v;
    }
    const int j = warp * 32 + lane;
    x[warp] |= 1u << lane;  // the site itself
    s_site[j] = make_uint4(j < S ? x[0] : 0u, j < S ? x[1] : 0u, j < S ? (uint32_t)(__popc(x[0]) + __popc(x[1])) : (1u << 20), (uint32_t)j);
    s_sorted[j] = make_uint4(0u, 0u, 1u << 20, 255u);
    s_key[j] = j < S ? ((uint32_t)(__popc(x[0]) + __popc(x[1])) << 6) | (uint32_t)j : 0xFFFFFFFFu;
  }
  __syncthreads();
  PTRACE(2);
  // sorted by (ancestor count, index): rank = number of smaller keys, four keys per shared-memory load
  if (warp * 32 + lane < S) {
    const int j = warp * 32 + lane;
    const uint32_t key = s_key[j];
    int r = 0;
    for (int k = 0; k < S; k += 4) {  // (keys of positions >= S are larger than any site's)
      const uint4 q = *reinterpret_cast<const uint4 *>(s_key + k);
      r += (q.x < key ? 1 : 0) + (q.y < key ? 1 : 0) + (q.z < key ? 1 : 0) + (q.w < key ? 1 : 0);
    }
    s_sorted[r] = s_site[j];
  }
  __syncthreads();
  PTRACE(3);
  // all pair costs inside the window, in parallel: thread (a, h) takes partners a+1+4h .. a+4+4h.  The window is the
  // next kPlanWindow sites in sorted order (measured on the benchmark's graph families: 8 finds as many cheap pairs
  // as 32 - no graph falls back - at a quarter of the evaluations)
  {
    const int a = tid >> 1, h = tid & 1;
    const uint4 si = s_sorted[a];
    uint32_t cost[4];
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const int idx = a + 1 + 4 * h + w;
      cost[w] = 0xFFFFu;
      if (a < S && idx < S) {
        const uint4 sj = s_sorted[idx];
        const int nC = __popc(si.x & sj.x) + __popc(si.y & sj.y);
        const uint32_t c = (uint32_t)((nC + 1) * ((int)si.z - nC + 1) * ((int)sj.z - nC + 1));
        if (c <= (uint32_t)kPairLim) cost[w] = c;
      }
    }
    uint16_t *dst = reinterpret_cast<uint16_t *>(&s_cost[a >> 1][4 * h]) + (a & 1);
#pragma unroll
    for (int w = 0; w < 4; ++w) dst[2 * w] = (uint16_t)cost[w];
  }
  __syncthreads();
  PTRACE(4);
  const int need = S - kPairCS;
  if (warp == 0) {
    // greedy matching over sorted positions: every position that is still free takes the cheapest free partner of its
    // window.  Straight-line code over the positions (compile-time shifts, the window costs of lane w preloaded in
    // registers): the chain per pairing round is a funnel shift, one warp reduction and the mask update.
    uint32_t u0 = 0, u1 = ~((1u << (S - 32)) - 1u);  // positions >= S are "used" (33 <= S <= 52)
    int npairs = 0;
    uint32_t my_cost = 0xFFFFFFFFu, my_a = 0, my_b = 0;  // lane k keeps pair k
    const int wl = lane & (kPlanWindow - 1);
    uint32_t cst[26];
#pragma unroll
    for (int k = 0; k < 26; ++k) cst[k] = s_cost[k][wl];
#pragma unroll
    for (int a = 0; a < 51; ++a) {
      const bool used = a < 32 ? ((u0 >> a) & 1u) != 0u : ((u1 >> (a - 32)) & 1u) != 0u;
      if (a < S - 1 && !used) {  // warp-uniform
        // bit l of `win`: position a + 1 + l is taken
        const uint32_t win = a + 1 < 32 ? __funnelshift_r(u0, u1, a + 1) : (u1 >> (a + 1 - 32));
        const uint32_t c = (a & 1) ? (cst[a >> 1] >> 16) : (cst[a >> 1] & 0xFFFFu);
        const uint32_t packed = (lane >= kPlanWindow || c == 0xFFFFu || ((win >> lane) & 1u)) ? 0xFFFFFFFFu : ((c << 8) | (uint32_t)lane);
        const uint32_t best = __reduce_min_sync(0xffffffffu, packed);
        if (best != 0xFFFFFFFFu) {
          const int b = a + 1 + (int)(best & 255u);
          if (b < 32) u0 |= 1u << b; else u1 |= 1u << (b - 32);
          if (lane == npairs) { my_cost = best >> 8; my_a = (uint32_t)a; my_b = (uint32_t)b; }
          ++npairs;  // at most S/2 <= 32
        }
      }
    }
    PTRACE(5);
    // keep the `need` cheapest pairs
    int rank = 0;
#pragma unroll 8
    for (int m = 0; m < 32; ++m) {
      const uint32_t cm = __shfl_sync(0xffffffffu, my_cost, m);
      rank += (cm < my_cost || (cm == my_cost && m < lane)) ? 1 : 0;
    }
    const bool sel = lane < npairs && rank < need;
    const uint32_t total = __reduce_add_sync(0xffffffffu, sel ? my_cost : 0u);
    const bool ok = npairs >= need && total <= (uint32_t)kPairCap;
    s_plan[lane * 2] = 255; s_plan[lane * 2 + 1] = 255;  // kPairCS == 32 columns: padding unless overwritten
    __syncwarp();
    if (late_wait) asm volatile("griddepcontrol.wait;" ::: "memory");  // first global writes of this warp
    if (!ok) {
      if (lane == 0) { const int slot = atomicAdd(fcount, 1); flist[slot] = g; fbslot[g] = slot; }
    } else {
      if (lane == 0) fbslot[g] = -1;
      uint32_t sel0 = 0, sel1 = 0;
      if (sel) {
        const uint32_t ji = s_sorted[my_a].w, jj = s_sorted[my_b].w;
        s_plan[rank * 2] = (uint8_t)ji; s_plan[rank * 2 + 1] = (uint8_t)jj;
        if (ji < 32) sel0 |= 1u << ji; else sel1 |= 1u << (ji - 32);
        if (jj < 32) sel0 |= 1u << jj; else sel1 |= 1u << (jj - 32);
      }
      sel0 = __reduce_or_sync(0xffffffffu, sel0);
      sel1 = __reduce_or_sync(0xffffffffu, sel1);
      // single sites fill the remaining columns (need + (S - 2*need) == kPairCS exactly)
      const uint32_t full0 = 0xffffffffu, full1 = (1u << (S - 32)) - 1u;
      const uint32_t free0 = ~sel0 & full0, free1 = ~sel1 & full1;
      const uint32_t below = (1u << lane) - 1u;
      if ((free0 >> lane) & 1u) { const int c = need + __popc(free0 & below); s_plan[c * 2] = (uint8_t)lane; s_plan[c * 2 + 1] = 255; }
      if ((free1 >> lane) & 1u) { const int c = need + __popc(free0) + __popc(free1 & below); s_plan[c * 2] = (uint8_t)(lane + 32); s_plan[c * 2 + 1] = 255; }
    }
    __syncwarp();
    PTRACE(6);
    // lane c: shape of column c - the masks of weight 1 / w1 / w2, the table's size and offset
    const int j1 = s_plan[lane * 2], j2 = s_plan[lane * 2 + 1];
    uint32_t a0 = 0, a1m = 0, b0 = 0, b1 = 0;
    int kind = 0;
    if (j1 < S) { kind = 1; a0 = s_site[j1].x; a1m = s_site[j1].y; }
    if (kind == 1 && j2 < S) { kind = 2; b0 = s_site[j2].x; b1 = s_site[j2].y; }
    const uint32_t c0 = a0 & b0, c1w = a1m & b1;
    const int n1 = __popc(a0) + __popc(a1m), n2 = __popc(b0) + __popc(b1), nC = __popc(c0) + __popc(c1w);
    const int d1 = n1 - nC, d2 = n2 - nC;
    const uint32_t w1 = (uint32_t)(nC + 1), w2 = w1 * (uint32_t)(d1 + 1);
    const uint32_t tsz = kind == 2 ? w2 * (uint32_t)(d2 + 1) : 0u;
    uint32_t incl = tsz;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t2 = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t2;
    }
    const uint32_t toff = incl - tsz;
    const uint32_t total_g = __shfl_sync(0xffffffffu, incl, 31);
    // weight 1 on the common ancestors (single site: all of them), w1 / w2 on the two differences
    s_mask[lane][0] = make_uint4(kind == 2 ? c0 : a0, kind == 2 ? c1w : a1m, kind == 2 ? (a0 & ~c0) : 0u, kind == 2 ? (a1m & ~c1w) : 0u);
    s_mask[lane][1] = make_uint4(kind == 2 ? (b0 & ~c0) : 0u, kind == 2 ? (b1 & ~c1w) : 0u, w1, w2);
    s_col[lane] = make_uint4((uint32_t)kind | (((65536u + w2 - 1u) / w2) << 2), (uint32_t)n1 | ((uint32_t)n2 << 8) | ((uint32_t)nC << 16), toff, (65536u + w1 - 1u) / w1);
    // meta: kind | n1<<8 | n2<<16 | nC<<24 ; d1 | d2<<8 | toff<<16 ; table entries of the whole graph
    reinterpret_cast<uint4 *>(plan + ((size_t)g * kPairCS + lane) * kPlanColBytes)[4] =
        make_uint4((uint32_t)kind | ((uint32_t)n1 << 8) | ((uint32_t)n2 << 16) | ((uint32_t)nC << 24),
                   (uint32_t)d1 | ((uint32_t)d2 << 8) | (toff << 16), total_g, 0u);
  }
  __syncthreads();
  PTRACE(7);
  if (late_wait) asm volatile("griddepcontrol.wait;" ::: "memory");  // first global writes of the other warps
  if (closed != nullptr && tid < S * W) closed[(size_t)g * S * W + tid] = s_rows[tid];
  // thread (c, quarter): 16 of column c's 64 B-operand bytes (constant rows left 0 for the expander)
  {
    const int c = tid >> 2, qd = tid & 3;
    const uint4 m0 = s_mask[c][0], m1 = s_mask[c][1];
    const uint32_t one = qd < 2 ? m0.x : m0.y, xx = qd < 2 ? m0.z : m0.w, yy = qd < 2 ? m1.x : m1.y, w1 = m1.z, w2 = m1.w;
    uint32_t words[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int sh = ((qd & 1) * 4 + q) * 4;
      const uint32_t o = (one >> sh) & 0xFu, x = (xx >> sh) & 0xFu, y = (yy >> sh) & 0xFu;
      words[q] = ((o * 0x00204081u) & 0x01010101u) + ((x * 0x00204081u) & 0x01010101u) * w1 + ((y * 0x00204081u) & 0x01010101u) * w2;
    }
    reinterpret_cast<uint4 *>(plan + ((size_t)g * kPairCS + c) * kPlanColBytes)[qd] = make_uint4(words[0], words[1], words[2], words[3]);
  }
  // the pair tables of this graph, entry pc + w1*p1 + w2*p2 = V(n1, pc+p1) + V(n2, pc+p2), V(n, p) = lut[n-p][p]
  // (the score kernel copies the block into shared memory with one bulk copy per graph).  One warp per pair column
  // (pairs are the first `need` columns, cheapest first), a lane per entry (mixed-radix digits by two multiply-shifts).
  asm volatile("cp.async.wait_all;" ::: "memory");
  __syncthreads();
  float *tg = tables + (size_t)g * kPairCap;
  for (int c = warp; c < need; c += kPlanThreads / 32) {
    const uint4 cm = s_col[c];
    if ((cm.x & 3u) != 2u) continue;  // warp-uniform
    const int cn1 = (int)(cm.y & 255u), cn2 = (int)((cm.y >> 8) & 255u), cnC = (int)(cm.y >> 16);
    const int cw1 = cnC + 1, cw2 = cw1 * (cn1 - cnC + 1), tsz = cw2 * (cn2 - cnC + 1);
    const uint32_t m1 = cm.w, m2 = cm.x >> 2;  // ceil(2^16 / cw1), ceil(2^16 / cw2): x / cw = (x * m) >> 16 for x * cw < 2^16
    float *td = tg + cm.z;
    const float *t1 = s_tab + cn1 * (cn1 + 1) / 2, *t2 = s_tab + cn2 * (cn2 + 1) / 2;
    for (int i = lane; i < tsz; i += 32) {
      const int p2 = (int)(((uint32_t)i * m2) >> 16), e = i - p2 * cw2;
      const int p1 = (int)(((uint32_t)e * m1) >> 16), pc = e - p1 * cw1;
      td[i] = t1[pc + p1] + t2[pc + p2];
    }
  }
  PTRACE(8);
#undef PTRACE
}

// Sum of a graph's partial log2-likelihoods (one per CTA that worked on its N tile, see Walker) + the null term.
// Graphs the pair planner gave up on (fbslot >= 0) take theirs from the fallback pass, whose tile count depends on
// the device-side list length.
__global__ void k_finalize(const double *partials, int n_slots, int G, int n_mtiles, int max_ctas, const double *sumN,
                           int n_graphs, double *scores, const int *fbslot, const int *fcount, const double *fb_partials,
                           int fb_G) {
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n_graphs) return;
  const int fs = fbslot != nullptr ? fbslot[g] : -1;
  const double *src = partials + (size_t)g * n_slots;
  int idx = g, tiles_g = G, n = n_graphs;
  if (fs >= 0) { src = fb_partials + (size_t)fs * kMaxSlots; idx = fs; tiles_g = fb_G; n = *fcount; n_slots = kMaxSlots; }
  const unsigned T = (unsigned)((n + tiles_g - 1) / tiles_g) * (unsigned)n_mtiles;
  const int cnt = slots_of_ntile(idx / tiles_g, n_mtiles, T, effective_ctas(T, n_mtiles, max_ctas));
  double v = 0.0;
  for (int c = 0; c < cnt && c < n_slots; ++c) v += src[c];
  scores[g] = __fma_rn(NEMB200_LN2, v, *sumN);  // (explicit: both finalize paths must round alike)
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

// launches of the current launch_score_tc call carry the programmatic-dependent-launch attribute (sweep entry points on
// a stream that is not being captured; the MCMC loop's graph capture launches plainly)
thread_local bool t_pdl = false;
template <typename K, typename... Args>
int launch_ex(K kern, int grid, int threads, int smem_bytes, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)threads); cfg.dynamicSmemBytes = (size_t)smem_bytes; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = t_pdl ? 1 : 0;
  NEMB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, args...));
  return 0;
}
template <typename K>
int launch_kernel(K kern, const Params &p, int grid, int threads, cudaStream_t st, int smem_bytes = Lay<false>::BYTES) {
  NEMB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  return launch_ex(kern, grid, threads, smem_bytes, st, p);
}

// the specialised one-site table kernel for the dataset's column stride
int launch_lut(const nemb200_dataset *h, const Params &p, int grid, cudaStream_t st, bool mix = false) {
  // the dataset fixed the column stride: S itself for the four headline shapes, S rounded up to 8 otherwise
  switch (h->CS == h->S ? h->S : -h->CS) {
#ifndef NEMB200_DEV_S50  // (development builds compile the S=50 instantiations only)
    case 10: return launch_kernel(k_score_tc<EPI_LUT, 10, 10>, p, grid, cta_threads(10), st);
    case 20: return launch_kernel(k_score_tc<EPI_LUT, 20, 20>, p, grid, cta_threads(20), st);
    case 30: return launch_kernel(k_score_tc<EPI_LUT, 30, 30>, p, grid, cta_threads(30), st);
#endif
    case 50:
      // mix: of every kMixPeriod groups of four sites the last one goes through ex2.approx (NEMB200_SCORE_TC_LUT_MIX)
      if (mix) return launch_kernel(k_score_tc<EPI_LUT, 50, kMixPeriod>, p, grid, cta_threads(50), st);
      return launch_kernel(k_score_tc<EPI_LUT, 50, 50>, p, grid, cta_threads(50), st);
#ifndef NEMB200_DEV_S50
    case 8: case -8: return launch_kernel(k_score_tc<EPI_LUT, 8, 8>, p, grid, cta_threads(8), st);
    case 16: case -16: return launch_kernel(k_score_tc<EPI_LUT, 16, 16>, p, grid, cta_threads(16), st);
    case 24: case -24: return launch_kernel(k_score_tc<EPI_LUT, 24, 24>, p, grid, cta_threads(24), st);
    case 32: case -32: return launch_kernel(k_score_tc<EPI_LUT, 32, 32>, p, grid, cta_threads(32), st);
    case 40: case -40: return launch_kernel(k_score_tc<EPI_LUT, 40, 40>, p, grid, cta_threads(40), st);
    case 48: case -48: return launch_kernel(k_score_tc<EPI_LUT, 48, 48>, p, grid, cta_threads(48), st);
    case 56: case -56: return launch_kernel(k_score_tc<EPI_LUT, 56, 56>, p, grid, cta_threads(56), st);
    case -64: return launch_kernel(k_score_tc<EPI_LUT, 64, 64>, p, grid, cta_threads(64), st);
#endif
    default: return launch_kernel(k_score_tc<EPI_LUT, 0, 0>, p, grid, cta_threads(0), st);
  }
}

// workspace carve-up (bytes from the base)
struct WsLay { size_t fcount, partials, fb_partials, plan, tables, fbslot, flist, closed, trace, total; };
WsLay ws_layout(const nemb200_dataset *h, int n, bool with_pair) {
  WsLay w;
  const bool pr = with_pair && h->pair_ok;
  size_t off = 0;
  w.fcount = off; off += 256;
  w.partials = off; off = align_up(off + sizeof(double) * (size_t)(n + 16) * kMaxSlots, 256);
  w.fb_partials = off; off = align_up(off + (pr ? sizeof(double) * (size_t)(n + 16) * kMaxSlots : 0), 256);
  w.plan = off; off = align_up(off + (pr ? (size_t)n * kPairCS * kPlanColBytes : 0), 256);
  w.tables = off; off = align_up(off + (pr ? (size_t)n * kPairCap * 4 : 0), 256);
  w.fbslot = off; off = align_up(off + (pr ? (size_t)n * 4 : 0), 256);
  w.flist = off; off = align_up(off + (pr ? (size_t)(n + 16) * 4 : 0), 256);
  w.closed = off; off = align_up(off + (pr ? (size_t)n * h->S * h->W * 4 : 0), 256);  // closed rows (closure fused into the planner)
  w.trace = off; off += 8 * 8192;  // diagnostics time line
  w.total = off;
  return w;
}

}  // namespace

size_t score_tc_workspace_bytes(const nemb200_dataset *h, int n, bool with_pair) { return ws_layout(h, n, with_pair).total; }
bool score_tc_has_mix(const nemb200_dataset *h) { return h->lut_valid && h->CS == 50 && h->S == 50; }

int launch_score_tc(const nemb200_dataset *h, const uint32_t *phi, int n, double *scores, void *ws, size_t ws_bytes,
                    int impl, cudaStream_t st, int leave_sms_free, int close, int pdl) {
  struct PdlScope { bool old; explicit PdlScope(bool v) : old(t_pdl) { t_pdl = v; } ~PdlScope() { t_pdl = old; } } pdl_scope(pdl != 0);
  NEMB_REQUIRE(h->S >= 8 && h->S <= 128, "tensor-core score needs 8 <= S <= 128 (S=%d)", h->S);
  // AUTO at THIS level (what the MCMC loop reaches) is a one-site kernel: the pair-merged path would run its planner per
  // chain group and iteration (95 vs 94 us).  The sweep entry points resolve AUTO themselves (mcmc.cu: sweep_auto) and pick
  // the pair-merged path for 44 <= S <= 50.
  if (impl == NEMB200_SCORE_AUTO)
    impl = h->lut_valid ? NEMB200_SCORE_TC_LUT
                        : (h->fixed_valid ? NEMB200_SCORE_TC_EX2 : (h->lutm_levels > 0 ? NEMB200_SCORE_TC_LUTM : NEMB200_SCORE_TC_EX2_2PASS));
  NEMB_REQUIRE(impl != NEMB200_SCORE_TC_LUTM || h->lutm_levels > 0,
               "multi-level table epilogue not available for this dataset (needs 16 <= S <= %d and a range of at most three levels, S=%d)", NEMB200_LUTM_MAX_S, h->S);
  NEMB_REQUIRE(impl != NEMB200_SCORE_TC_PAIR || h->pair_ok, "pair-merged table epilogue not available for this dataset (S=%d)", h->S);
  NEMB_REQUIRE((impl != NEMB200_SCORE_TC_LUT && impl != NEMB200_SCORE_TC_LUT_MIX) || h->lut_valid, "exp-table epilogue not valid for this dataset (S=%d)", h->S);
  NEMB_REQUIRE(impl != NEMB200_SCORE_TC_LUT_MIX || score_tc_has_mix(h), "table + ex2 mix needs the column stride 50 (S=%d)", h->S);
  NEMB_REQUIRE(impl != NEMB200_SCORE_TC_EX2 || h->fixed_valid, "fixed-scale ex2 epilogue not valid for this dataset");
  NEMB_REQUIRE(ws_bytes >= score_tc_workspace_bytes(h, n, impl == NEMB200_SCORE_TC_PAIR),
               impl == NEMB200_SCORE_TC_PAIR ? "score workspace too small for the pair-merged kernel: size it with nemb200_score_workspace_bytes_impl"
                                             : "score workspace too small");
  if (n == 0) return 0;
  // leave_sms_free: the two-group MCMC loop keeps a few SMs out of each score launch so that the other group's
  // propose / accept / finalize kernels (and the first CTAs of its score kernel) never queue behind a full wave
  const int nsm = num_sms() - leave_sms_free > 8 ? num_sms() - leave_sms_free : num_sms();
  const bool pair = impl == NEMB200_SCORE_TC_PAIR;
  const bool lut = impl == NEMB200_SCORE_TC_LUT || impl == NEMB200_SCORE_TC_LUT_MIX;
  const bool lutm = impl == NEMB200_SCORE_TC_LUTM;
  const int G = pair ? kPairG : (lut ? h->G_lut : (lutm ? h->G_lutm : h->G));  // graphs per MMA N tile of the kernel that will run
  const int CS = pair ? kPairCS : (lutm ? h->CS_lutm : h->CS);
  const WsLay wl = ws_layout(h, n, pair);
  uint8_t *wb = reinterpret_cast<uint8_t *>(ws);
  const int n_ntiles = (n + G - 1) / G;
  NEMB_REQUIRE((double)n_ntiles * h->n_mtiles * nsm < 4.0e9 && (double)((n + 1) / 2 + 1) * h->n_mtiles * nsm < 4.0e9,
               "batch too large for one launch (%d graphs x %d E tiles): split it", n, h->n_mtiles);
  const unsigned T = (unsigned)n_ntiles * (unsigned)h->n_mtiles;
  Params p;
  p.phi = phi; p.a_img = pair ? h->a_pair : (lutm ? h->a_lutm : h->a_lut); p.lut = lutm ? h->lutm : h->lut;
  for (int l = 0; l < 3; ++l) p.lvl_scale[l] = (float)h->lutm_scale[l];
  p.tile_first = h->tile_first;
  p.partials = reinterpret_cast<double *>(wb + wl.partials);
  p.n_graphs = n; p.S = h->S; p.W = h->W; p.E = h->E; p.G = G;
  p.n_mtiles = h->n_mtiles; p.n_slots = kMaxSlots;
  p.CS = CS;
  p.nmma = (G * CS + 15) / 16 * 16;
  p.ksteps = pair ? (h->S + 9 + 31) / 32 : (lutm ? (h->S + 3 + 31) / 32 : (h->S + (lut ? 1 : 0) + 31) / 32);
  // accumulator a = 4p + 256q = 256c - 252p  ->  x = alpha2*p + beta2*q = k1*a + kc*c
  const double kappa2 = h->alpha2 - h->beta2;
  p.k1 = (float)(-kappa2 / 252.0);
  p.kc = (float)(256.0 * kappa2 / 252.0 + h->beta2);
  p.scale = impl == NEMB200_SCORE_TC_EX2_2PASS ? 0.f : (float)h->scale;
  p.plan = nullptr; p.tables = nullptr; p.glist = nullptr; p.n_dev = nullptr; p.fbslot = nullptr;
  p.late_wait = pdl == 2 ? 1 : 0;
  p.close = close ? 1 : 0;
  p.done_counter = nullptr; p.sumN = h->sumN; p.scores = scores;
  // (one CTA finalizing a large batch is slower than the separate, fully parallel kernel; in the two-group MCMC loop
  //  the separate kernel overlaps with the other group's score kernel and the tail of a fused one does not)
  int ctr_slot = -1;
  if (n <= 2048 && leave_sms_free == 0) {
    // one self-resetting completion counter per distinct workspace (launches that share a workspace are ordered by
    // the caller anyway); when the pool is exhausted the separate finalize kernel is used
    // (the counters themselves sit in the dataset workspace, which the library zeroed: caller memory is not zero at rest)
    nemb200_cache *hm = h->cache;
    int slot = -1;
    {
      std::lock_guard<std::recursive_mutex> lock(hm->mu);
      for (int i = 0; i < hm->n_ctr; ++i) if (hm->ctr_ws[i] == ws) { slot = i; break; }
      if (slot < 0 && hm->n_ctr < 64) { slot = hm->n_ctr++; hm->ctr_ws[slot] = ws; }
    }
    if (slot >= 0) p.done_counter = h->counters + slot;
    ctr_slot = slot;
  }
  const int grid = effective_ctas(T, h->n_mtiles, nsm);
  int rc;
  static const char *dbg_env = getenv("NEMB200_DEBUG");
  p.debug = dbg_env ? atoi(dbg_env) : 0;
  p.trace = (p.debug & 32) ? reinterpret_cast<long long *>(wb + wl.trace) : nullptr;
  const int *fbslot = nullptr;
  if (pair) {
    // plan -> pair-merged kernel -> one-site kernel over the (device-side) list of graphs the planner gave up on
    // Fused finalize (a counter slot was found above): the last CTA of the pair kernel writes the scores of the planned
    // graphs, the last CTA of the list-mode launch those of the listed ones and puts the list length back to zero -
    // three launches (plan, score, list), no memset and no finalize kernel.  The list length then lives beside the
    // completion counters in the dataset's workspace (zero at rest; caller memory is not).
    const bool fused = p.done_counter != nullptr;
    int *fcount = fused ? h->counters + 128 + ctr_slot : reinterpret_cast<int *>(wb + wl.fcount);
    int *flist = reinterpret_cast<int *>(wb + wl.flist);
    int *fslot = reinterpret_cast<int *>(wb + wl.fbslot);
    uint8_t *plan = wb + wl.plan;
    static bool planned_once = false;  // (diagnostics, bit 128: after the first call, time the score kernels alone on the plan it left in the workspace)
    if (!(p.debug & 128) || !planned_once) {
      planned_once = true;
      if (!fused) NEMB_CHECK_CUDA(cudaMemsetAsync(fcount, 0, sizeof(int), st));
      const uint32_t *phi_c = phi; const float *tabs_c = h->tabs;
      long long *trace_p = (p.debug & 64) ? reinterpret_cast<long long *>(wb + wl.trace) : nullptr;
      uint32_t *closed_p = close ? reinterpret_cast<uint32_t *>(wb + wl.closed) : nullptr;
      if (launch_ex(k_pair_plan, n, kPlanThreads, 0, st, phi_c, n, h->S, h->W, tabs_c, plan, reinterpret_cast<float *>(wb + wl.tables), fslot, flist,
                    fcount, trace_p, closed_p, (int)p.late_wait)) return -1;
    }
    Params f = p;
    f.a_img = h->a_lut; f.plan = nullptr; f.glist = flist; f.n_dev = fcount; f.fbslot = nullptr; f.close = 0;
    if (close) f.phi = reinterpret_cast<const uint32_t *>(wb + wl.closed);
    if (fused) f.done_counter = h->counters + 64 + ctr_slot;
    f.partials = reinterpret_cast<double *>(wb + wl.fb_partials);
    f.G = h->G_lut; f.CS = h->CS; f.nmma = (f.G * f.CS + 15) / 16 * 16; f.ksteps = (h->S + 1 + 31) / 32;
    f.trace = nullptr;
    p.close = 0;  // (the planner closed the batch; the pair kernel reads the plan, the list-mode launch the closed rows)
    p.plan = plan; p.tables = reinterpret_cast<float *>(wb + wl.tables);
    if (fused) p.fbslot = fslot;
    // The list-mode launch needs only the planner's list (normally empty: it exits at once) and the two score launches
    // write disjoint scores, so their order is free.  With the caller's promise the pair kernel goes LAST: the next call's
    // planner is then its direct dependent and plans while it drains (74.2 us per call instead of 79.1 at S=50, 1024 graphs);
    // without it the planner waits at its start anyway and the pair kernel goes first (79.1 instead of 81.4).
    if (p.late_wait) {
      rc = launch_lut(h, f, nsm, st);
      if (rc) return rc;
      rc = launch_kernel(k_score_tc<EPI_LUT, kPairCS, kPairCS, true>, p, grid, cta_threads(kPairCS), st, Lay<true>::BYTES);
    } else {
      rc = launch_kernel(k_score_tc<EPI_LUT, kPairCS, kPairCS, true>, p, grid, cta_threads(kPairCS), st, Lay<true>::BYTES);
      if (rc) return rc;
      rc = launch_lut(h, f, nsm, st);
    }
    fbslot = fslot;
  } else if (lut) {
    rc = launch_lut(h, p, grid, st, impl == NEMB200_SCORE_TC_LUT_MIX);
  } else if (lutm) {
    rc = h->lutm_levels == 2 ? launch_kernel(k_score_tc<EPI_LUTM, 0, 2>, p, grid, cta_threads(0), st, Lay<false, true>::BYTES)
                             : launch_kernel(k_score_tc<EPI_LUTM, 0, 3>, p, grid, cta_threads(0), st, Lay<false, true>::BYTES);
  } else if (impl == NEMB200_SCORE_TC_EX2) {
    rc = launch_kernel(k_score_tc<EPI_EX2, 0, 0>, p, grid, cta_threads(0), st);
  } else if (impl == NEMB200_SCORE_TC_EX2_2PASS) {
    rc = launch_kernel(k_score_tc<EPI_EX2_2PASS, 0, 0>, p, grid, cta_threads(0), st);
  } else {
    nemb200_set_error("unknown score impl %d", impl);
    return -1;
  }
  if (rc) return rc;
  if (p.done_counter != nullptr) return 0;  // finalized by the last CTA of the score kernel
  // 64-thread CTAs: small enough (registers) to run beside a resident score CTA of another stream (two-group MCMC loop)
  k_finalize<<<(n + 63) / 64, 64, 0, st>>>(p.partials, p.n_slots, G, h->n_mtiles, nsm, h->sumN, n, scores, fbslot,
                                               reinterpret_cast<const int *>(wb + wl.fcount),
                                               reinterpret_cast<const double *>(wb + wl.fb_partials), h->G_lut);
  NEMB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// Diagnostics (tests): how many graphs of the last NEMB200_SCORE_TC_PAIR launch on this workspace the planner gave up on
// (they went through the one-site kernel).  Synchronises the device.
extern "C" int nemb200_pair_fallback_count(const nemb200_dataset *h, const void *d_workspace, int n_graphs) {
  NEMB_REQUIRE(h != nullptr && d_workspace != nullptr && n_graphs >= 0 && h->pair_ok, "bad handle / workspace, or no pair-merged kernel for this dataset");
  const WsLay wl = ws_layout(h, n_graphs, true);
  std::vector<int> slot((size_t)n_graphs);
  NEMB_CHECK_CUDA(cudaDeviceSynchronize());
  NEMB_CHECK_CUDA(cudaMemcpy(slot.data(), reinterpret_cast<const uint8_t *>(d_workspace) + wl.fbslot, sizeof(int) * (size_t)n_graphs, cudaMemcpyDeviceToHost));
  int c = 0;
  for (int v : slot) c += v >= 0 ? 1 : 0;
  return c;
}

// Host-only self-test of the tile partition (no device work): every flat tile is owned by exactly one CTA, the items a
// CTA walks are contiguous and inside one N tile, slots are dense per N tile and never exceed kMaxSlots.  Returns 0 if
// all invariants hold, else a positive code.  Used by the CPU test-suite.
extern "C" int nemb200_selftest_partition(int n_ntiles, int n_mtiles, int max_ctas) {
  if (n_ntiles < 1 || n_mtiles < 1 || max_ctas < 1) return -1;
  if ((double)n_ntiles * n_mtiles * (max_ctas + 1.0) >= 4.0e9) return -1;
  const unsigned T = (unsigned)n_ntiles * (unsigned)n_mtiles;
  const int ctas = effective_ctas(T, n_mtiles, max_ctas);
  if (ctas < 1 || ctas > max_ctas) return 1;
  unsigned expect = 0;
  for (int c = 0; c < ctas; ++c) {
    const unsigned lo = cta_first_tile(c, T, ctas), hi = cta_first_tile(c + 1, T, ctas);
    if (lo != expect || hi < lo) return 2;
    expect = hi;
    for (unsigned x = lo; x < hi; ++x)
      if (cta_of_tile(x, T, ctas) != c) return 3;
  }
  if (expect != T) return 4;
  for (int nt = 0; nt < n_ntiles; ++nt) {
    const int cnt = slots_of_ntile(nt, n_mtiles, T, ctas);
    if (cnt < 1 || cnt > kMaxSlots) return 5;
    const unsigned lo = (unsigned)nt * n_mtiles;
    const int first = cta_of_tile(lo, T, ctas);
    for (int m = 0; m < n_mtiles; ++m) {
      const int slot = cta_of_tile(lo + m, T, ctas) - first;
      if (slot < 0 || slot >= cnt) return 6;
    }
  }
  return 0;
}

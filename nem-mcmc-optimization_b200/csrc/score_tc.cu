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
// Roles (384 threads, no CTA-wide barrier in the main loop): warp 0 = TMA producer (A tiles, 4-stage
// mbarrier ring), warp 1 = MMA issuer + TMEM owner (2 x 256-column accumulators), warps 2-3 = B-operand
// expanders (double-buffered, run one work item ahead), warps 4-11 = two epilogue warpgroups that
// alternate tiles.  Work items (N tile of graphs x chunk of E tiles) are statically strided over CTAs.
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int kThreads = 384;
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
constexpr int OFF_RED = OFF_ROWS + 1024 * 4;                // 8 warps x 16 doubles
constexpr int OFF_BAR = OFF_RED + 8 * kMaxG * 8;            // barriers + scalars
constexpr int kSmemBytes = OFF_BAR + 256;

struct Params {
  const uint32_t *phi;   // [n_graphs][S][W]
  const uint8_t *a_img;  // [n_mtiles][16 KB]
  const float *lut;      // [64][64]
  double *partials;      // [n_ntiles*G][n_chunks][2]
  int n_graphs, S, W, E, G, n_ntiles, n_mtiles, tpc, n_chunks, n_items, nmma, ksteps;
  float k1;              // EX2: x = k1 * acc + bias_j
  float kc;              // EX2: bias_j = kc * c_j - s
  float scale;           // s
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
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!done);
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
                                                      int ng, double (&part)[kMaxG]) {
  const int S = p.S;
#pragma unroll
  for (int g = 0; g < kMaxG; ++g) {
    if (g < ng) {
      float l = S >= 16 ? graph_lse2<EPI, 16>(taddr + g * S, S, sBias + g * kBiasStride * (S > 64 ? 2 : 1), p)
                        : graph_lse2<EPI, 8>(taddr + g * S, S, sBias + g * kBiasStride, p);
      if (valid) part[g] += (double)l;
    }
  }
}

// ---- specialised epilogue (compile-time S): straight-line code over the whole tile ----------------------
// Sites j < NLUT of every graph go through the shared-memory exp table (LSU pipe), the others through
// I2F + FFMA + ex2.approx (ALU/FMA/MUFU pipes).  The next 32 accumulator columns are prefetched from TMEM
// while the current 32 are consumed.
template <int S_T, int NLUT>
__device__ __forceinline__ void tile_epilogue_static(uint32_t taddr, const float *sBias, const Params &p, bool valid,
                                                     int ng, double (&part)[kMaxG]) {
  constexpr int G = (256 / S_T) < kMaxG ? (256 / S_T) : kMaxG;
  constexpr int NCOL = G * S_T;
  constexpr int NCH = (NCOL + 31) / 32;
  static_assert(NLUT % 4 == 0 || NLUT == S_T, "EX2 sites must start on a float4 boundary of the bias array");
  uint32_t r[2][32];
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
  const float null_term = ex2f(-p.scale);
  TMEM_LD32(taddr, r[0]);
  tmem_wait_ld();
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch) {
    if (ch + 1 < NCH) TMEM_LD32(taddr + (ch + 1) * 32, r[(ch + 1) & 1]);
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const int col = ch * 32 + i;
      if (col < NCOL) {
        const int g = col / S_T, j = col % S_T;
        if (j == 0) { acc[0] = null_term; acc[1] = 0.f; acc[2] = 0.f; acc[3] = 0.f; }
        float v;
        if (j < NLUT) {
          v = lds_f32(r[ch & 1][i]);
        } else {
          if (((j - NLUT) & 3) == 0) b4 = *reinterpret_cast<const float4 *>(sBias + g * kBiasStride + j);
          const float b = ((j - NLUT) & 3) == 0 ? b4.x : (((j - NLUT) & 3) == 1 ? b4.y : (((j - NLUT) & 3) == 2 ? b4.z : b4.w));
          v = ex2f(fmaf((float)(int)r[ch & 1][i], p.k1, b));
        }
        acc[j & 3] += v;
        if (j == S_T - 1) {
          const float l = log2_split((acc[0] + acc[1]) + (acc[2] + acc[3])) + p.scale;
          if (valid && g < ng) part[g] += (double)l;
        }
      }
    }
    if (ch + 1 < NCH) tmem_wait_ld();
  }
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
      mbar_init(bar_tfull + 8 * i, 1); mbar_init(bar_tempty + 8 * i, 4);
      mbar_init(bar_bfull + 8 * i, 2); mbar_init(bar_edone + 8 * i, 8);
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
          mbar_wait(bar_afull + 8 * st, ph);
          mbar_wait(bar_tempty + 8 * buf, bph ^ 1u);
          tc_fence_after();
          const uint32_t a0 = sbase + OFF_A + st * kTileBytes;
          for (int ks = 0; ks < p.ksteps; ++ks)
            umma_i8(tmem_base + buf * 256u, umma_desc(a0 + ks * 32), umma_desc(b0 + ks * 32), idesc, ks > 0 ? 1u : 0u);
          umma_commit(bar_aempty + 8 * st);  // A slot reusable once these MMAs retire
          umma_commit(bar_tfull + 8 * buf);  // accumulator ready for the epilogue
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
      asm volatile("bar.sync 3, 64;" ::: "memory");
      uint8_t *brow = smem + OFF_B + bb * 2 * kTileBytes;
      float *bias = reinterpret_cast<float *>(smem + OFF_BIAS) + bb * 1024;
      const int bstride = kBiasStride * (S > 64 ? 2 : 1);
      for (int col = te; col < 256; col += 64) {
        const int gi = col / S, j = col - gi * S;
        uint32_t anc[4] = {0u, 0u, 0u, 0u};
        int cnt = 0;
        const bool used = gi < ng;
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
        const bool lut_site = EPI == EPI_LUT && gi < G && (S_T == 0 || j < NLUT);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          uint32_t half = (anc[c >> 1] >> ((c & 1) * 16)) & 0xFFFFu;
          uint4 v;
          v.x = (((half & 0xFu) * 0x00204081u) & 0x01010101u) << 1;
          v.y = ((((half >> 4) & 0xFu) * 0x00204081u) & 0x01010101u) << 1;
          v.z = ((((half >> 8) & 0xFu) * 0x00204081u) & 0x01010101u) << 1;
          v.w = ((((half >> 12) & 0xFu) * 0x00204081u) & 0x01010101u) << 1;
          if (lut_site && c == (S >> 4)) {
            uint32_t ext = ((sbase + OFF_LUT) >> 7) << ((S & 3) * 8);
            int wsel = (S >> 2) & 3;
            if (wsel == 0) v.x |= ext; else if (wsel == 1) v.y |= ext; else if (wsel == 2) v.z |= ext; else v.w |= ext;
          }
          *reinterpret_cast<uint4 *>(brow + sw128((uint32_t)col, (uint32_t)c * 16u)) = v;
        }
        if (gi < G) bias[gi * bstride + j] = p.kc * (float)cnt - p.scale;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_bfull + 8 * bb);
      asm volatile("bar.sync 3, 64;" ::: "memory");  // sRows is reused by the next item
    }
  } else {
    // =============================== epilogue: warpgroup wg handles tiles with (tile index & 1) == wg ======
    const int wg = (warp - 4) >> 2, quarter = warp & 3, tw = threadIdx.x - 128 - wg * 128;
    uint32_t it = 0;
    for (int k = 0; k < n_my_items; ++k) {
      const int item = (int)blockIdx.x + k * (int)gridDim.x;
      const int chunk = item / p.n_ntiles, ntile = item % p.n_ntiles;
      const int m0 = chunk * p.tpc, nt = min(p.tpc, p.n_mtiles - m0);
      const int g0 = ntile * G, ng = min(G, p.n_graphs - g0);
      const uint32_t bb = k & 1u, bbph = (k >> 1) & 1u;
      const float *bias = reinterpret_cast<const float *>(smem + OFF_BIAS) + bb * 1024;
      mbar_wait(bar_bfull + 8 * bb, bbph);  // bias values of this item are written
      double part[kMaxG];
#pragma unroll
      for (int g = 0; g < kMaxG; ++g) part[g] = 0.0;
      for (int t = 0; t < nt; ++t, ++it) {
        const uint32_t buf = it & 1u, bph = (it >> 1) & 1u;
        if ((int)buf != wg) continue;
        mbar_wait(bar_tfull + 8 * buf, bph);
        tc_fence_after();
        const uint32_t taddr = tmem_base + buf * 256u + ((uint32_t)(quarter * 32) << 16);
        const bool valid = (m0 + t) * 128 + quarter * 32 + lane < p.E;
        if constexpr (S_T > 0) tile_epilogue_static<S_T, NLUT>(taddr, bias, p, valid, ng, part);
        else tile_epilogue_generic<EPI>(taddr, bias, p, valid, ng, part);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_tempty + 8 * buf);
      }
      // ---- per-item reduction inside the warpgroup -> one double per (graph, chunk, warpgroup)
#pragma unroll
      for (int g = 0; g < kMaxG; ++g) {
        if (g < ng) {
          double v = part[g];
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
          if (lane == 0) sRed[(warp - 4) * kMaxG + g] = v;
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_edone + 8 * bb);  // bias / B buffer of this item may be recycled
      if (wg == 0) asm volatile("bar.sync 1, 128;" ::: "memory"); else asm volatile("bar.sync 2, 128;" ::: "memory");
      if (tw < ng) {
        double v = 0.0;
        for (int w4 = 0; w4 < 4; ++w4) v += sRed[(wg * 4 + w4) * kMaxG + tw];
        p.partials[((size_t)(g0 + tw) * p.n_chunks + chunk) * 2 + wg] = v;
      }
      if (wg == 0) asm volatile("bar.sync 1, 128;" ::: "memory"); else asm volatile("bar.sync 2, 128;" ::: "memory");
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
  return 256 + sizeof(double) * 2 * (size_t)pl.n_ntiles * h->G * pl.n_chunks;
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
  p.nmma = (h->G * h->S + 15) / 16 * 16;
  p.ksteps = (h->S + (impl == NEMB200_SCORE_TC_LUT ? 1 : 0) + 31) / 32;
  // accumulator a = 4p + 256q = 256c - 252p  ->  x = alpha2*p + beta2*q = k1*a + kc*c
  const double kappa2 = h->alpha2 - h->beta2;
  p.k1 = (float)(-kappa2 / 252.0);
  p.kc = (float)(256.0 * kappa2 / 252.0 + h->beta2);
  p.scale = impl == NEMB200_SCORE_TC_EX2_2PASS ? 0.f : (float)h->scale;
  const int grid = pl.n_items < nsm ? pl.n_items : nsm;
  int rc;
  static const char *mix_env = getenv("NEMB200_NO_MIX");  // diagnostics: pure table epilogue
  const bool mix = mix_env == nullptr;
  if (impl == NEMB200_SCORE_TC_LUT) {
    switch (h->S) {
      case 10: rc = launch_kernel(k_score_tc<EPI_LUT, 10, 10>, p, grid, st); break;
      case 20: rc = mix ? launch_kernel(k_score_tc<EPI_LUT, 20, 12>, p, grid, st)
                        : launch_kernel(k_score_tc<EPI_LUT, 20, 20>, p, grid, st); break;
      case 30: rc = mix ? launch_kernel(k_score_tc<EPI_LUT, 30, 20>, p, grid, st)
                        : launch_kernel(k_score_tc<EPI_LUT, 30, 30>, p, grid, st); break;
      case 50: rc = mix ? launch_kernel(k_score_tc<EPI_LUT, 50, 32>, p, grid, st)
                        : launch_kernel(k_score_tc<EPI_LUT, 50, 50>, p, grid, st); break;
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
  k_finalize<<<(n + 255) / 256, 256, 0, st>>>(p.partials, p.n_chunks * 2, h->sumN, n, scores);
  NEMB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

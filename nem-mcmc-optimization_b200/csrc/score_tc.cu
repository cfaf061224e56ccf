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
// "EX2_2PASS" subtracts the running maximum first (any dynamic range).
// No B x S x E intermediate ever reaches HBM: per (graph, E-chunk) one double is written.
//
// Roles (320 threads): warp 0 = TMA producer (A tiles, 4-stage mbarrier ring), warp 1 = MMA issuer +
// TMEM owner (2 x 256-column accumulators), warps 2-9 = two epilogue warpgroups that alternate tiles.
#include "common.cuh"

namespace {

constexpr int kThreads = 320;
constexpr int kStages = 4;
constexpr int kTileBytes = 128 * 128;  // one operand atom: 128 rows x 128 B (SWIZZLE_128B)
constexpr int kLutBytes = 64 * 256;
constexpr int kMaxG = 16;

// shared memory carve-up (offsets from the 1024-aligned base)
constexpr int OFF_LUT = 0;                               // first: its base must be < 32 KB and 128-aligned
constexpr int OFF_B = OFF_LUT + kLutBytes;               // 256 x 128 B
constexpr int OFF_A = OFF_B + 2 * kTileBytes;            // kStages x 16 KB
constexpr int OFF_BIAS = OFF_A + kStages * kTileBytes;   // 256 floats
constexpr int OFF_ROWS = OFF_BIAS + 256 * 4;             // 1024 words of bit rows
constexpr int OFF_RED = OFF_ROWS + 1024 * 4;             // 8 warps x 16 doubles
constexpr int OFF_BAR = OFF_RED + 8 * kMaxG * 8;         // barriers + scalars
constexpr int kSmemBytes = OFF_BAR + 256;

struct Params {
  const uint32_t *phi;   // [n_graphs][S][W]
  const uint8_t *a_img;  // [n_mtiles][16 KB]
  const float *lut;      // [64][64]
  double *partials;      // [n_ntiles*G][n_chunks]
  unsigned int *counter; // work counter (zeroed before launch)
  int n_graphs, S, W, E, G, n_ntiles, n_mtiles, tpc, n_chunks, n_items, nmma;
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
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

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

// One graph's columns [col0, col0+S) of one accumulator tile for this thread's E-gene.
// Returns log2( 2^-s + sum_j term_j ) + s  (i.e. log2 of 1 + sum_j 2^x_j).
template <int EPI, int CH>
__device__ __forceinline__ float graph_lse2(uint32_t taddr, int S, const float *sBias, const Params &p) {
  static_assert(CH == 16 || CH == 8, "chunk");
  float acc0 = 0.f, acc1 = 0.f, m = 0.f;
  if (EPI == EPI_EX2_2PASS) {
    // pass 1: running max of x_j (and 0 for the null site)
    for (int c = 0; c < S; c += CH) {
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
      else v = ex2f(fmaf((float)(int)r[i], p.k1, sBias[cs + i]) - m);
      if (i >= skip) { if (i & 1) acc1 += v; else acc0 += v; }
    }
  }
  if (EPI == EPI_EX2_2PASS) return m + log2_split(ex2f(-m) + acc0 + acc1);
  return log2_split(ex2f(-p.scale) + acc0 + acc1) + p.scale;
}

template <int EPI>
__global__ void __launch_bounds__(kThreads, 1) k_score_tc(const Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t sbase = smem_u32(smem);
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + OFF_BAR);
  const uint32_t bar_full = sbase + OFF_BAR, bar_empty = bar_full + 8 * kStages, bar_tfull = bar_empty + 8 * kStages,
                 bar_tempty = bar_tfull + 16;
  volatile uint32_t *s_tmem = reinterpret_cast<volatile uint32_t *>(bars + 2 * kStages + 4);
  volatile int *s_item = reinterpret_cast<volatile int *>(bars + 2 * kStages + 5);
  float *sBias = reinterpret_cast<float *>(smem + OFF_BIAS);
  uint32_t *sRows = reinterpret_cast<uint32_t *>(smem + OFF_ROWS);
  double *sRed = reinterpret_cast<double *>(smem + OFF_RED);

  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) { mbar_init(bar_full + 8 * i, 1); mbar_init(bar_empty + 8 * i, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(bar_tfull + 8 * i, 1); mbar_init(bar_tempty + 8 * i, 4); }
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

  uint32_t it = 0;  // running tile counter: identical sequence in every role
  for (;;) {
    if (threadIdx.x == 0) *s_item = (int)atomicAdd(p.counter, 1u);
    __syncthreads();  // item id visible; previous item's B tile and accumulators are drained
    const int item = *s_item;
    if (item >= p.n_items) break;
    const int chunk = item / p.n_ntiles, ntile = item % p.n_ntiles;
    const int m0 = chunk * p.tpc, nt = min(p.tpc, p.n_mtiles - m0);
    const int g0 = ntile * G, ng = min(G, p.n_graphs - g0);

    if (warp >= 2) {
      // ---- expand the bit rows of this tile's graphs into the B operand (one thread per column)
      const int te = threadIdx.x - 64;
      const int nwords = ng * S * W;
      for (int i = te; i < nwords; i += 256) sRows[i] = p.phi[(size_t)g0 * S * W + i];
      epi_bar_sync();
      const int gi = te / S, j = te - gi * S;
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
      uint8_t *brow = smem + OFF_B;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        uint32_t half = (anc[c >> 1] >> ((c & 1) * 16)) & 0xFFFFu;
        uint4 v;
        v.x = (((half & 0xFu) * 0x00204081u) & 0x01010101u) << 1;
        v.y = ((((half >> 4) & 0xFu) * 0x00204081u) & 0x01010101u) << 1;
        v.z = ((((half >> 8) & 0xFu) * 0x00204081u) & 0x01010101u) << 1;
        v.w = ((((half >> 12) & 0xFu) * 0x00204081u) & 0x01010101u) << 1;
        if (EPI == EPI_LUT && used && c == (S >> 4)) {
          // extra K row S: B = lut_base/128, A = 128  ->  accumulator = shared address of lut[q][p]
          uint32_t ext = ((sbase + OFF_LUT) >> 7) << ((S & 3) * 8);
          int wsel = (S >> 2) & 3;
          if (wsel == 0) v.x |= ext; else if (wsel == 1) v.y |= ext; else if (wsel == 2) v.z |= ext; else v.w |= ext;
        }
        *reinterpret_cast<uint4 *>(brow + sw128((uint32_t)te, (uint32_t)c * 16u)) = v;
      }
      sBias[te] = p.kc * (float)cnt - p.scale;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();  // B tile + bias ready

    if (warp == 0) {
      if (lane == 0) {
        for (int t = 0; t < nt; ++t) {
          const uint32_t k = it + t, st = k % kStages, ph = (k / kStages) & 1u;
          mbar_wait(bar_empty + 8 * st, ph ^ 1u);
          mbar_expect_tx(bar_full + 8 * st, kTileBytes);
          bulk_g2s(sbase + OFF_A + st * kTileBytes, p.a_img + (size_t)(m0 + t) * kTileBytes, kTileBytes, bar_full + 8 * st);
        }
      }
    } else if (warp == 1) {
      if (lane == 0) {
        for (int t = 0; t < nt; ++t) {
          const uint32_t k = it + t, st = k % kStages, ph = (k / kStages) & 1u, buf = k & 1u, bph = (k >> 1) & 1u;
          mbar_wait(bar_full + 8 * st, ph);
          mbar_wait(bar_tempty + 8 * buf, bph ^ 1u);
          tc_fence_after();
          const uint32_t a0 = sbase + OFF_A + st * kTileBytes, b0 = sbase + OFF_B;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            umma_i8(tmem_base + buf * 256u, umma_desc(a0 + ks * 32), umma_desc(b0 + ks * 32), idesc, ks > 0 ? 1u : 0u);
          umma_commit(bar_empty + 8 * st);   // A slot reusable once these MMAs retire
          umma_commit(bar_tfull + 8 * buf);  // accumulator ready for the epilogue
        }
      }
    } else {
      // ---- epilogue: warpgroup wg handles tiles with (global tile index & 1) == wg
      const int wg = (warp - 2) >> 2, quarter = warp & 3;
      double part[kMaxG];
#pragma unroll
      for (int g = 0; g < kMaxG; ++g) part[g] = 0.0;
      for (int t = 0; t < nt; ++t) {
        const uint32_t k = it + t, buf = k & 1u, bph = (k >> 1) & 1u;
        if ((int)buf != wg) continue;
        mbar_wait(bar_tfull + 8 * buf, bph);
        tc_fence_after();
        const uint32_t taddr = tmem_base + buf * 256u + ((uint32_t)(quarter * 32) << 16);
        const bool valid = (m0 + t) * 128 + quarter * 32 + lane < p.E;
#pragma unroll
        for (int g = 0; g < kMaxG; ++g) {
          if (g < ng) {
            float l = S >= 16 ? graph_lse2<EPI, 16>(taddr + g * S, S, sBias + g * S, p)
                              : graph_lse2<EPI, 8>(taddr + g * S, S, sBias + g * S, p);
            if (valid) part[g] += (double)l;
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_tempty + 8 * buf);
      }
      // ---- per-item reduction over the 256 epilogue threads -> one double per (graph, chunk)
#pragma unroll
      for (int g = 0; g < kMaxG; ++g) {
        if (g < ng) {
          double v = part[g];
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
          if (lane == 0) sRed[(warp - 2) * kMaxG + g] = v;
        }
      }
      epi_bar_sync();
      const int te = threadIdx.x - 64;
      if (te < ng) {
        double v = 0.0;
        for (int w8 = 0; w8 < 8; ++w8) v += sRed[w8 * kMaxG + te];
        p.partials[(size_t)(g0 + te) * p.n_chunks + chunk] = v;
      }
    }
    it += (uint32_t)nt;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base));
}

__global__ void k_finalize(const double *partials, int n_chunks, const double *sumN, int n_graphs, double *scores) {
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n_graphs) return;
  double v = 0.0;
  for (int c = 0; c < n_chunks; ++c) v += partials[(size_t)g * n_chunks + c];
  scores[g] = *sumN + NEMB200_LN2 * v;
}

struct Plan { int n_ntiles, tpc, n_chunks, n_items; };
Plan make_plan(const nemb200_dataset *h, int n) {
  Plan pl;
  pl.n_ntiles = (n + h->G - 1) / h->G;
  // ~8 work items per SM, but at least 4 E tiles per item so the B expansion is amortised
  long total = (long)pl.n_ntiles * h->n_mtiles;
  int tpc = (int)(total / (148 * 8));
  if (tpc < 4) tpc = 4;
  if (tpc > h->n_mtiles) tpc = h->n_mtiles;
  pl.n_chunks = (h->n_mtiles + tpc - 1) / tpc;
  pl.tpc = (h->n_mtiles + pl.n_chunks - 1) / pl.n_chunks;
  pl.n_chunks = (h->n_mtiles + pl.tpc - 1) / pl.tpc;
  pl.n_items = pl.n_ntiles * pl.n_chunks;
  return pl;
}

}  // namespace

size_t score_tc_workspace_bytes(const nemb200_dataset *h, int n) {
  Plan pl = make_plan(h, n);
  return 256 + sizeof(double) * (size_t)pl.n_ntiles * h->G * pl.n_chunks;
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
  Plan pl = make_plan(h, n);
  Params p;
  p.phi = phi; p.a_img = h->a_lut; p.lut = h->lut;
  p.counter = reinterpret_cast<unsigned int *>(ws);
  p.partials = reinterpret_cast<double *>(reinterpret_cast<uint8_t *>(ws) + 256);
  p.n_graphs = n; p.S = h->S; p.W = h->W; p.E = h->E; p.G = h->G;
  p.n_ntiles = pl.n_ntiles; p.n_mtiles = h->n_mtiles; p.tpc = pl.tpc; p.n_chunks = pl.n_chunks; p.n_items = pl.n_items;
  p.nmma = (h->G * h->S + 15) / 16 * 16;
  // accumulator a = 4p + 256q = 256c - 252p  ->  x = alpha2*p + beta2*q = k1*a + kc*c
  const double kappa2 = h->alpha2 - h->beta2;
  p.k1 = (float)(-kappa2 / 252.0);
  p.kc = (float)(256.0 * kappa2 / 252.0 + h->beta2);
  p.scale = impl == NEMB200_SCORE_TC_EX2_2PASS ? 0.f : (float)h->scale;
  NEMB_CHECK_CUDA(cudaMemsetAsync(ws, 0, 256, st));
  int dev = 0, nsm = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
  int grid = pl.n_items < nsm ? pl.n_items : nsm;
  auto launch = [&](auto kern) -> int {
    NEMB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    kern<<<grid, kThreads, kSmemBytes, st>>>(p);
    NEMB_CHECK_CUDA(cudaGetLastError());
    return 0;
  };
  int rc;
  if (impl == NEMB200_SCORE_TC_LUT) rc = launch(k_score_tc<EPI_LUT>);
  else if (impl == NEMB200_SCORE_TC_EX2) rc = launch(k_score_tc<EPI_EX2>);
  else if (impl == NEMB200_SCORE_TC_EX2_2PASS) rc = launch(k_score_tc<EPI_EX2_2PASS>);
  else { nemb200_set_error("unknown score impl %d", impl); return -1; }
  if (rc) return rc;
  k_finalize<<<(n + 255) / 256, 256, 0, st>>>(p.partials, p.n_chunks, h->sumN, n, scores);
  NEMB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

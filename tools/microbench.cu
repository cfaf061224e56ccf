// Microbenchmarks that decide the scoring-kernel epilogue design on B200 (sm_100a):
//   1. tcgen05.ld (TMEM -> registers) bandwidth per SM for 4/8/16 warps
//   2. LDS table-lookup (+FADD) throughput with in-row (conflict-free) indices
//   3. MUFU.EX2 (+FFMA,FADD) throughput, and LDS/MUFU mixed
//   4. kind::i8 / kind::f16 tcgen05.mma smoke (does it run, are integer counts exact)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench microbench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int NCOL>
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r);
template <>
__device__ __forceinline__ void tmem_ld32<32>(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}

__global__ void __launch_bounds__(512, 1) k_tmem_bw(int iters, int* out, long long* cycles) {
  __shared__ uint32_t tmem_base;
  int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  uint32_t base = tmem_base + ((uint32_t)(warp & 3) * 32u << 16);
  uint32_t acc = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    uint32_t r[32], q[32];
    tmem_ld32<32>(base + ((it * 64) & 511), r);
    tmem_ld32<32>(base + ((it * 64 + 32) & 511), q);
    asm volatile("tcgen05.wait::ld.sync.aligned;");
#pragma unroll
    for (int i = 0; i < 32; ++i) acc += r[i] ^ q[i];
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base));
}

// mode 0: LDS lookup + FADD ; mode 1: FFMA + MUFU.EX2 + FADD ; mode 2: 2 LDS : 1 MUFU mix ; mode 3: LOP3+LDS+FADD
template <int MODE>
__global__ void __launch_bounds__(512, 1) k_epi(int iters, float* out, long long* cycles, const uint32_t* idx_in) {
  __shared__ float lut[64 * 52];
  for (int i = threadIdx.x; i < 64 * 52; i += blockDim.x) lut[i] = 1.0f / (1 + i);
  __syncthreads();
  // in-row indices: every lane of a warp shares q (row), p varies in 0..31
  uint32_t idx[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) idx[i] = idx_in[(threadIdx.x * 32 + i) & 16383];
  float acc0 = 0, acc1 = 0, acc2 = 0, acc3 = 0;
  uint32_t lb = smem_u32(lut);
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      float v0, v1, v2, v3;
      if (MODE == 0) {
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v0) : "r"(idx[i] + lb));
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v1) : "r"(idx[i + 1] + lb));
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v2) : "r"(idx[i + 2] + lb));
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v3) : "r"(idx[i + 3] + lb));
      } else if (MODE == 1) {
        v0 = exp2f(fmaf(__uint_as_float(idx[i]), 1.7f, -3.0f));
        v1 = exp2f(fmaf(__uint_as_float(idx[i + 1]), 1.7f, -3.0f));
        v2 = exp2f(fmaf(__uint_as_float(idx[i + 2]), 1.7f, -3.0f));
        v3 = exp2f(fmaf(__uint_as_float(idx[i + 3]), 1.7f, -3.0f));
      } else if (MODE == 2) {
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v0) : "r"(idx[i] + lb));
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v1) : "r"(idx[i + 1] + lb));
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v2) : "r"(idx[i + 2] + lb));
        v3 = exp2f(fmaf(__uint_as_float(idx[i + 3]), 1.7f, -3.0f));
      } else {
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v0) : "r"((idx[i] & 0x3ffc) + lb));
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v1) : "r"((idx[i + 1] & 0x3ffc) + lb));
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v2) : "r"((idx[i + 2] & 0x3ffc) + lb));
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v3) : "r"((idx[i + 3] & 0x3ffc) + lb));
      }
      acc0 += v0; acc1 += v1; acc2 += v2; acc3 += v3;
    }
    // perturb indices a little so nothing is hoisted (stay inside the row)
    idx[it & 31] ^= 4;
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc0 + acc1 + acc2 + acc3;
}

// ---- MMA smoke: D[128 x 64] = A[128 x K] * B[64 x K]^T with SWIZZLE_128B K-major operands ----
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}
__device__ __forceinline__ uint32_t sw128(uint32_t row, uint32_t byte_in_row) {
  return (row >> 3) * 1024 + (row & 7) * 128 + ((((byte_in_row >> 4) ^ (row & 7)) << 4) | (byte_in_row & 15));
}
// KIND 0: bf16 (K=64 elems/row), KIND 1: u8 (K=128 elems/row)
template <int KIND>
__global__ void __launch_bounds__(128, 1) k_mma(const uint8_t* a_vals, const uint8_t* b_vals, uint32_t* d_out) {
  extern __shared__ __align__(1024) uint8_t sm[];
  uint8_t* sa = sm;              // 128 rows x 128 B
  uint8_t* sb = sm + 16384;      // 64 rows x 128 B
  __shared__ uint32_t tmem_base;
  __shared__ __align__(8) uint64_t bar;
  const int KE = KIND == 0 ? 64 : 128;
  for (int i = threadIdx.x; i < 128 * KE; i += 128) {
    int r = i / KE, k = i % KE;
    if (KIND == 0) { uint16_t v = a_vals[i] ? (a_vals[i] == 1 ? 0x3F80 : (a_vals[i] == 4 ? 0x4080 : (a_vals[i] == 255 ? 0x4B00 : 0x4380))) : 0;
      *(uint16_t*)(sa + sw128(r, k * 2)) = v; }
    else sa[sw128(r, k)] = a_vals[i];
  }
  for (int i = threadIdx.x; i < 64 * KE; i += 128) {
    int r = i / KE, k = i % KE;
    if (KIND == 0) *(uint16_t*)(sb + sw128(r, k * 2)) = b_vals[i] ? 0x3F80 : 0;
    else sb[sw128(r, k)] = b_vals[i];
  }
  int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(&tmem_base)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  asm volatile("fence.proxy.async.shared::cta;");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  uint32_t tb = tmem_base;
  if (threadIdx.x == 0) {
    // idesc: c_format (F32=1 / S32=2) <<4, a_format<<7, b_format<<10, N>>3 <<17, M>>4 <<24
    uint32_t idesc = KIND == 0 ? ((1u << 4) | (1u << 7) | (1u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24))
                               : ((2u << 4) | (0u << 7) | (0u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24));
    for (int ks = 0; ks < 4; ++ks) {
      uint64_t da = make_desc(smem_u32(sa) + ks * 32), db = make_desc(smem_u32(sb) + ks * 32);
      uint32_t accum = ks > 0;
      if (KIND == 0)
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(tb), "l"(da), "l"(db), "r"(idesc), "r"(accum));
      else
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}" ::"r"(tb), "l"(da), "l"(db), "r"(idesc), "r"(accum));
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)));
  }
  // everyone waits for the MMA
  {
    uint32_t done = 0;
    while (!done) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0));
  }
  asm volatile("tcgen05.fence::after_thread_sync;");
  uint32_t r[32], q[32];
  uint32_t base = tb + ((uint32_t)(warp & 3) * 32u << 16);
  tmem_ld32<32>(base, r);
  tmem_ld32<32>(base + 32, q);
  asm volatile("tcgen05.wait::ld.sync.aligned;");
  for (int i = 0; i < 32; ++i) { d_out[threadIdx.x * 64 + i] = r[i]; d_out[threadIdx.x * 64 + 32 + i] = q[i]; }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tb));
}

int main() {
  int dev = 0; cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, dev));
  printf("device %s SMs %d clock %d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
  int nsm = p.multiProcessorCount;
  long long* d_cyc; CK(cudaMalloc(&d_cyc, 8)); long long cyc;
  int* d_out; CK(cudaMalloc(&d_out, nsm * 512 * 4));
  float* f_out = (float*)d_out;
  // ---- 1. TMEM bandwidth
  for (int warps : {4, 8, 16}) {
    int iters = 2000;
    k_tmem_bw<<<nsm, warps * 32>>>(10, d_out, d_cyc); CK(cudaDeviceSynchronize());
    k_tmem_bw<<<nsm, warps * 32>>>(iters, d_out, d_cyc); CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost));
    double bytes = (double)iters * warps * 32 * 64 * 4;
    printf("tmem_ld 32x32b.x32 x2/iter warps=%2d: %lld cyc, %.1f B/clk/SM\n", warps, cyc, bytes / cyc);
  }
  // ---- 2/3. epilogue pipes
  uint32_t* h_idx = new uint32_t[16384];
  for (int t = 0; t < 512; ++t) for (int i = 0; i < 32; ++i) {
    int lane = t & 31, warp = t >> 5; int q = (warp * 7 + i * 3) % 40; int pp = (lane * 5 + i) % 32;
    h_idx[t * 32 + i] = 4u * (pp + 64u * q);
  }
  uint32_t* d_idx; CK(cudaMalloc(&d_idx, 16384 * 4)); CK(cudaMemcpy(d_idx, h_idx, 16384 * 4, cudaMemcpyHostToDevice));
  for (int warps : {8, 16}) {
    int iters = 2000; double el = (double)iters * warps * 32 * 32;
#define RUN(M, name) k_epi<M><<<nsm, warps * 32>>>(10, f_out, d_cyc, d_idx); CK(cudaDeviceSynchronize()); \
    k_epi<M><<<nsm, warps * 32>>>(iters, f_out, d_cyc, d_idx); CK(cudaDeviceSynchronize()); \
    CK(cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost)); printf("%s warps=%2d: %.2f elem/clk/SM\n", name, warps, el / cyc);
    RUN(0, "LDS+FADD            ") RUN(3, "LOP3+LDS+FADD       ") RUN(1, "FFMA+EX2+FADD       ") RUN(2, "mix 3 LDS : 1 EX2   ")
  }
  // ---- 4. MMA smoke (exact integer counts)
  for (int kind = 0; kind < 2; ++kind) {
    int KE = kind == 0 ? 64 : 128;
    uint8_t* ha = new uint8_t[128 * KE]; uint8_t* hb = new uint8_t[64 * KE];
    for (int i = 0; i < 128 * KE; ++i) { int k = i % KE; ha[i] = k < 50 ? ((i * 2654435761u >> 13) & 1 ? 4 : 0) : 0; if (kind == 1 && ha[i]) ha[i] = 2; if (kind == 1 && !ha[i] && k < 50) ha[i] = 128; if (kind == 0 && !ha[i] && k < 50) ha[i] = 9; /* 9 -> 256.0 */ }
    for (int i = 0; i < 64 * KE; ++i) { int k = i % KE; hb[i] = k < 50 ? ((i * 40503u >> 7) & 1) : 0; if (kind == 1 && hb[i]) hb[i] = 2; }
    if (kind == 0) for (int r = 0; r < 128; ++r) ha[r * KE + 50] = 255;  // bias 2^23
    if (kind == 0) for (int r = 0; r < 64; ++r) hb[r * KE + 50] = 1;
    uint8_t *da, *db; uint32_t* dd; CK(cudaMalloc(&da, 128 * KE)); CK(cudaMalloc(&db, 64 * KE)); CK(cudaMalloc(&dd, 128 * 64 * 4));
    CK(cudaMemcpy(da, ha, 128 * KE, cudaMemcpyHostToDevice)); CK(cudaMemcpy(db, hb, 64 * KE, cudaMemcpyHostToDevice));
    if (kind == 0) { CK(cudaFuncSetAttribute(k_mma<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768)); k_mma<0><<<1, 128, 32768>>>(da, db, dd); }
    else { CK(cudaFuncSetAttribute(k_mma<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768)); k_mma<1><<<1, 128, 32768>>>(da, db, dd); }
    CK(cudaDeviceSynchronize());
    uint32_t* hd = new uint32_t[128 * 64]; CK(cudaMemcpy(hd, dd, 128 * 64 * 4, cudaMemcpyDeviceToHost));
    int bad = 0;
    for (int m = 0; m < 128; ++m) for (int n = 0; n < 64; ++n) {
      long ref = 0;
      for (int k = 0; k < 50; ++k) { int a = ha[m * KE + k], b = hb[n * KE + k]; if (kind == 0) { a = a == 4 ? 4 : (a == 9 ? 256 : 0); } ref += (long)a * b; }
      uint32_t got = hd[m * 64 + n];
      uint32_t want = kind == 0 ? 0x4B000000u + (uint32_t)ref : (uint32_t)ref;
      if (got != want) { if (bad < 5) printf("  kind %d mismatch m=%d n=%d got %08x want %08x\n", kind, m, n, got, want); ++bad; }
    }
    printf("mma kind %s: %d mismatches of %d\n", kind == 0 ? "f16(bf16)+bias2^23" : "i8(u8)", bad, 128 * 64);
  }
  return 0;
}

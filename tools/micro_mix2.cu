// Round-2 epilogue pipe study (B200): can FMA-pipe work run beside the shared-memory look-ups?
// Per element the epilogue turns an integer accumulator `a` into one term 2^x.  Recipes:
//   L  table look-up            : LDS + FADD (packed FADD2 per two look-ups)
//   P  polynomial exp2 (FMA pipe): int->float by magic number (IADD + FADD2), x = k1*a + bias (FFMA2), Cody-Waite
//      range reduction (FADD2 x3), degree-DEG polynomial (FFMA2 x DEG), exponent insert (ALU), FADD2
//   M  ex2.approx (MUFU)         : the same int->float + FFMA2, then MUFU.EX2, FADD2 (no I2F: I2F shares the XU pipe)
// One CTA per SM, `warps` warps, registers only; prints elements per clock per SM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float lds(uint32_t a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; }
__device__ __forceinline__ float ex2(float x) { float v; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(v) : "f"(x)); return v; }
__device__ __forceinline__ uint64_t pack(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void unpack(uint64_t v, float& a, float& b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) { uint64_t d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ uint64_t sub2(uint64_t a, uint64_t b) { uint64_t d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

template <int DEG>
__device__ __forceinline__ uint64_t poly2(uint64_t f) {
  // 2^f on [-0.5, 0.5] (Taylor-ish coefficients: throughput study only)
  const uint64_t c5 = pack(1.3333558e-3f, 1.3333558e-3f), c4 = pack(9.6181291e-3f, 9.6181291e-3f),
                 c3 = pack(5.5504109e-2f, 5.5504109e-2f), c2 = pack(2.4022651e-1f, 2.4022651e-1f),
                 c1 = pack(6.9314718e-1f, 6.9314718e-1f), c0 = pack(1.f, 1.f);
  uint64_t p;
  if (DEG >= 5) { p = fma2(c5, f, c4); p = fma2(p, f, c3); }
  else if (DEG == 4) { p = fma2(c4, f, c3); }
  else p = c3;
  p = fma2(p, f, c2); p = fma2(p, f, c1); p = fma2(p, f, c0);
  return p;
}
// two elements through the FMA pipe
template <int DEGS>
__device__ __forceinline__ uint64_t poly_pair(uint32_t a0, uint32_t a1, uint64_t k1, uint64_t bias) {
  constexpr int DEG = DEGS < 0 ? -DEGS : DEGS;  // DEGS < 0: the accumulator already is a float (kind::f16 MMA): no int->float step
  const uint64_t m23 = pack(8388608.f, 8388608.f), magic = pack(12582912.f, 12582912.f);
  uint64_t f = DEGS < 0 ? pack(__uint_as_float(a0), __uint_as_float(a1))
                        : sub2(pack(__uint_as_float(a0 + 0x4B000000u), __uint_as_float(a1 + 0x4B000000u)), m23);  // (float)a exactly
  uint64_t x = fma2(f, k1, bias);
  uint64_t t = add2(x, magic);
  uint64_t r = sub2(t, magic);
  uint64_t fr = sub2(x, r);
  uint64_t p = poly2<DEG>(fr);
  float t0, t1, p0, p1;
  unpack(t, t0, t1); unpack(p, p0, p1);
  p0 = __int_as_float(__float_as_int(p0) + (__float_as_int(t0) << 23));
  p1 = __int_as_float(__float_as_int(p1) + (__float_as_int(t1) << 23));
  return pack(p0, p1);
}
__device__ __forceinline__ uint64_t mufu_pair(uint32_t a0, uint32_t a1, uint64_t k1, uint64_t bias) {
  const uint64_t m23 = pack(8388608.f, 8388608.f);
  uint64_t f = sub2(pack(__uint_as_float(a0 + 0x4B000000u), __uint_as_float(a1 + 0x4B000000u)), m23);
  uint64_t x = fma2(f, k1, bias);
  float x0, x1; unpack(x, x0, x1);
  return pack(ex2(x0), ex2(x1));
}
// per group of 16 elements: NL look-ups, NM mufu, NP polynomial (all even)
template <int NL, int NM, int NP, int DEG>
__global__ void __launch_bounds__(768, 1) k(int iters, float* out, long long* cycles, const uint32_t* idx_in) {
  __shared__ float lut[64 * 52];
  for (int i = threadIdx.x; i < 64 * 52; i += blockDim.x) lut[i] = 1.0f / (1 + i);
  __syncthreads();
  const uint32_t lb = smem_u32(lut);
  uint32_t a[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) a[i] = idx_in[(threadIdx.x * 32 + i) & 16383] + lb;
  uint64_t acc0 = pack(0.f, 0.f), acc1 = acc0;
  const uint64_t k1 = pack(-1e-4f, -1e-4f), bias = pack(3.0f + threadIdx.x * 1e-3f, 2.0f);
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    // a warp-uniform offset that changes every iteration: nothing is loop-invariant (ptxas hoists plain ld.shared),
    // and it costs no instruction (LDS takes a uniform-register offset, the int->float IADD3 has a free third input)
    const uint32_t off = ((uint32_t)it * 4u) & 0xFCu;  // period 64: not unrollable into invariants
#pragma unroll
    for (int i = 0; i < 32; i += 16) {
#pragma unroll
      for (int e = 0; e < 16; e += 2) {
        uint64_t v;
        if (e < NL) v = pack(lds(a[i + e] + off), lds(a[i + e + 1] + off));
        else if (e < NL + NM) v = mufu_pair(a[i + e] + off, a[i + e + 1] + off, k1, bias);
        else v = poly_pair<DEG>(a[i + e] + off, a[i + e + 1] + off, k1, bias);
        if (e & 2) acc1 = add2(acc1, v); else acc0 = add2(acc0, v);
      }
    }
  }
  __syncthreads();
  lut[threadIdx.x] = __longlong_as_double(0) + threadIdx.x;  // (a store + barrier + load: the clock below is read after the barrier really released)
  __syncthreads();
  long long t1 = clock64() + (lut[(threadIdx.x + 1) % blockDim.x] < 0 ? 1 : 0);
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
  float s0, s1, s2, s3; unpack(acc0, s0, s1); unpack(acc1, s2, s3);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s0 + s1 + s2 + s3;
}
template <int NL, int NM, int NP, int DEG>
void run(const char* name, float* out, long long* d_cyc, const uint32_t* d_idx) {
  printf("%-34s", name);
  for (int warps : {8, 16, 20, 24}) {
    int iters = 4000; long long cyc;
    k<NL, NM, NP, DEG><<<148, warps * 32>>>(10, out, d_cyc, d_idx); cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<NL, NM, NP, DEG><<<148, warps * 32>>>(iters, out, d_cyc, d_idx);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) { printf(" CUDA error %s\n", cudaGetErrorString(err)); return; }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost);
    // clock64 is read by one warp right after a barrier that does not block at issue: cross-check with the event time
    printf("  w%2d: %6.2f (ev %6.2f)", warps, (double)iters * warps * 32 * 32 / cyc, (double)iters * warps * 32 * 32 / (ms * 1e-3 * 1.965e9));
  }
  printf("   elem/clk/SM\n");
}
int main() {
  long long* d_cyc; cudaMalloc(&d_cyc, 8);
  float* out; cudaMalloc(&out, 148 * 768 * 4);
  uint32_t* h_idx = new uint32_t[16384];
  for (int t = 0; t < 512; ++t) for (int i = 0; i < 32; ++i) {
    int lane = t & 31, warp = t >> 5; int q = (warp * 7 + i * 3) % 40; int pp = (lane * 5 + i) % 32;
    h_idx[t * 32 + i] = 4u * (pp + 64u * q);
  }
  uint32_t* d_idx; cudaMalloc(&d_idx, 16384 * 4); cudaMemcpy(d_idx, h_idx, 16384 * 4, cudaMemcpyHostToDevice);
  run<16, 0, 0, 5>("8 L", out, d_cyc, d_idx);
  run<0, 16, 0, 5>("8 M (no I2F)", out, d_cyc, d_idx);
  run<0, 0, 16, 5>("8 P deg5", out, d_cyc, d_idx);
  run<0, 0, 16, 4>("8 P deg4", out, d_cyc, d_idx);
  run<0, 0, 16, 3>("8 P deg3", out, d_cyc, d_idx);
  run<0, 0, 16, -5>("8 P deg5 float acc", out, d_cyc, d_idx);
  run<0, 0, 16, -4>("8 P deg4 float acc", out, d_cyc, d_idx);
  run<12, 0, 4, -5>("6 L : 2 P deg5 float acc", out, d_cyc, d_idx);
  run<8, 0, 8, -5>("4 L : 4 P deg5 float acc", out, d_cyc, d_idx);
  run<8, 0, 8, -4>("4 L : 4 P deg4 float acc", out, d_cyc, d_idx);
  run<12, 0, 4, 5>("6 L : 2 P deg5", out, d_cyc, d_idx);
  run<12, 0, 4, 4>("6 L : 2 P deg4", out, d_cyc, d_idx);
  run<8, 0, 8, 5>("4 L : 4 P deg5", out, d_cyc, d_idx);
  run<8, 0, 8, 4>("4 L : 4 P deg4", out, d_cyc, d_idx);
  run<14, 2, 0, 5>("7 L : 1 M", out, d_cyc, d_idx);
  run<12, 4, 0, 5>("6 L : 2 M", out, d_cyc, d_idx);
  run<10, 6, 0, 5>("5 L : 3 M", out, d_cyc, d_idx);
  run<10, 4, 2, 4>("5 L : 2 M : 1 P deg4", out, d_cyc, d_idx);
  run<10, 4, 2, -4>("5 L : 2 M : 1 P deg4 float", out, d_cyc, d_idx);
  run<8, 6, 2, 4>("4 L : 3 M : 1 P deg4", out, d_cyc, d_idx);
  run<14, 0, 2, 4>("7 L : 1 P deg4", out, d_cyc, d_idx);
  run<8, 8, 0, 5>("4 L : 4 M", out, d_cyc, d_idx);
  run<8, 4, 4, 5>("4 L : 2 M : 2 P deg5", out, d_cyc, d_idx);
  run<8, 4, 4, 4>("4 L : 2 M : 2 P deg4", out, d_cyc, d_idx);
  run<4, 4, 8, 4>("2 L : 2 M : 4 P deg4", out, d_cyc, d_idx);
  return 0;
}

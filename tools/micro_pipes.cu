// Throughput of the epilogue building blocks per SM (B200): table look-ups (LDS), ex2.approx (MUFU),
// FMA-pipe polynomial exp2, and mixes of them.  One CTA of 512 threads per SM, registers only.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float lds(uint32_t a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; }
__device__ __forceinline__ float ex2(float x) { float v; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(v) : "f"(x)); return v; }
// Cody-Waite + degree-3 polynomial exp2 on the FMA/ALU pipes (no MUFU): x in [-126, 126]
__device__ __forceinline__ float poly_ex2(float x) {
  float t = x + 12582912.f;                 // 1.5*2^23: round to nearest integer in the low mantissa bits
  float f = x - (t - 12582912.f);           // fractional part in [-0.5, 0.5]
  float p = fmaf(f, 0.0555041087f, 0.240226507f);
  p = fmaf(p, f, 0.693147181f);
  p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}
// MODE: per group of 8 elements, how many go through LDS / MUFU / POLY
template <int NL, int NM, int NP>
__global__ void __launch_bounds__(512, 1) k(int iters, float* out, long long* cycles, const uint32_t* idx_in) {
  __shared__ float lut[64 * 52];
  for (int i = threadIdx.x; i < 64 * 52; i += blockDim.x) lut[i] = 1.0f / (1 + i);
  __syncthreads();
  uint32_t lb = smem_u32(lut);
  uint32_t a[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) a[i] = idx_in[(threadIdx.x * 32 + i) & 16383] + lb;
  float acc0 = 0, acc1 = 0, acc2 = 0, acc3 = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 32; i += 8) {
      float v[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        if (e < NL) v[e] = lds(a[i + e]);
        else if (e < NL + NM) v[e] = ex2(fmaf((float)(int)a[i + e], 1e-4f, -3.0f));
        else v[e] = poly_ex2(fmaf((float)(int)a[i + e], 1e-4f, -3.0f));
      }
      acc0 += v[0] + v[4]; acc1 += v[1] + v[5]; acc2 += v[2] + v[6]; acc3 += v[3] + v[7];
    }
    a[it & 31] ^= 4;
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc0 + acc1 + acc2 + acc3;
}
template <int NL, int NM, int NP>
void run(const char* name, int nsm, float* out, long long* d_cyc, const uint32_t* d_idx) {
  for (int warps : {8, 16}) {
    int iters = 2000; long long cyc;
    k<NL, NM, NP><<<nsm, warps * 32>>>(10, out, d_cyc, d_idx); cudaDeviceSynchronize();
    k<NL, NM, NP><<<nsm, warps * 32>>>(iters, out, d_cyc, d_idx); cudaDeviceSynchronize();
    cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-28s warps=%2d: %6.2f elem/clk/SM\n", name, warps, (double)iters * warps * 32 * 32 / cyc);
  }
}
int main() {
  int nsm = 148;
  long long* d_cyc; cudaMalloc(&d_cyc, 8);
  float* out; cudaMalloc(&out, nsm * 512 * 4);
  uint32_t* h_idx = new uint32_t[16384];
  for (int t = 0; t < 512; ++t) for (int i = 0; i < 32; ++i) {
    int lane = t & 31, warp = t >> 5; int q = (warp * 7 + i * 3) % 40; int pp = (lane * 5 + i) % 32;
    h_idx[t * 32 + i] = 4u * (pp + 64u * q);
  }
  uint32_t* d_idx; cudaMalloc(&d_idx, 16384 * 4); cudaMemcpy(d_idx, h_idx, 16384 * 4, cudaMemcpyHostToDevice);
  run<8, 0, 0>("LDS only", nsm, out, d_cyc, d_idx);
  run<0, 8, 0>("MUFU ex2 only", nsm, out, d_cyc, d_idx);
  run<0, 0, 8>("FMA-poly exp2 only", nsm, out, d_cyc, d_idx);
  run<6, 2, 0>("6 LDS : 2 MUFU", nsm, out, d_cyc, d_idx);
  run<5, 3, 0>("5 LDS : 3 MUFU", nsm, out, d_cyc, d_idx);
  run<6, 0, 2>("6 LDS : 2 POLY", nsm, out, d_cyc, d_idx);
  run<5, 0, 3>("5 LDS : 3 POLY", nsm, out, d_cyc, d_idx);
  run<4, 0, 4>("4 LDS : 4 POLY", nsm, out, d_cyc, d_idx);
  run<4, 2, 2>("4 LDS : 2 MUFU : 2 POLY", nsm, out, d_cyc, d_idx);
  run<5, 1, 2>("5 LDS : 1 MUFU : 2 POLY", nsm, out, d_cyc, d_idx);
  run<0, 4, 4>("4 MUFU : 4 POLY", nsm, out, d_cyc, d_idx);
  return 0;
}

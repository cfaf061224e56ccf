// Do tcgen05.ld (TMEM -> registers) and LDS (shared-memory look-ups) share a pipe?  8 warps run TMEM loads,
// 8 warps run look-ups, alone and together (B200 diagnostics).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
#define LD32(taddr, r) asm volatile( \
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15," \
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];" \
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), \
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), \
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), \
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]) \
      : "r"(taddr))
// mode bit0: warps 0-7 do TMEM loads; bit1: warps 8-15 do look-ups
__global__ void __launch_bounds__(512, 1) k(int mode, int iters, float* out, long long* cyc_tmem, long long* cyc_lds) {
  __shared__ uint32_t tmem_base;
  __shared__ float lut[4096];
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 4096; i += 512) lut[i] = 1.0f / (1 + i);
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  float acc = 0;
  long long t0 = clock64(), t1 = t0;
  if (warp < 8) {
    if (mode & 1) {
      uint32_t base = tmem_base + ((uint32_t)(warp & 3) * 32u << 16) + (warp >> 2) * 256u;
      uint32_t a = 0;
      for (int it = 0; it < iters; ++it) {
        uint32_t r[32];
        LD32(base + ((it * 32) & 255), r);
        asm volatile("tcgen05.wait::ld.sync.aligned;");
#pragma unroll
        for (int i = 0; i < 32; i += 8) a += r[i];
      }
      acc = (float)a;
      t1 = clock64();
      if (threadIdx.x == 0 && blockIdx.x == 0) *cyc_tmem = t1 - t0;
    }
  } else {
    if (mode & 2) {
      uint32_t lb = smem_u32(lut);
      uint32_t idx[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) idx[i] = lb + 4u * (((lane * 5 + i) & 31) + 64u * ((warp * 7 + i * 3) % 40));
      float a0 = 0, a1 = 0, a2 = 0, a3 = 0;
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          float v0, v1, v2, v3;
          asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v0) : "r"(idx[i]));
          asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v1) : "r"(idx[i + 1]));
          asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v2) : "r"(idx[i + 2]));
          asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v3) : "r"(idx[i + 3]));
          a0 += v0; a1 += v1; a2 += v2; a3 += v3;
        }
        idx[it & 31] ^= 4;
      }
      acc = a0 + a1 + a2 + a3;
      t1 = clock64();
      if (threadIdx.x == 256 && blockIdx.x == 0) *cyc_lds = t1 - t0;
    }
  }
  out[blockIdx.x * 512 + threadIdx.x] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base));
}
int main() {
  float* out; cudaMalloc(&out, 148 * 512 * 4);
  long long *d; cudaMalloc(&d, 16); long long h[2];
  const int iters = 4000;
  for (int mode = 1; mode <= 3; ++mode) {
    cudaMemset(d, 0, 16);
    k<<<148, 512>>>(mode, 10, out, d, d + 1); cudaDeviceSynchronize();
    k<<<148, 512>>>(mode, iters, out, d, d + 1); cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("mode %d (%s):", mode, cudaGetErrorString(e));
    if (mode & 1) printf("  TMEM loads: %7lld clk -> %6.1f B/clk/SM (8 warps x 4 KB/iter)", h[0], 8.0 * 4096 * iters / h[0]);
    if (mode & 2) printf("  look-ups: %7lld clk -> %5.2f LDS-warp-instr/clk/SM", h[1], 8.0 * 32 * iters / h[1]);
    printf("\n");
  }
  return 0;
}

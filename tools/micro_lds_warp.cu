// How fast can a warp issue independent shared-memory look-ups, and does separating the loads from their
// consumers matter?  (B200 diagnostics; variants: 0 = compiler-scheduled LDS/FADD mix, 1 = all 32 LDS of a chunk
// first, then the 32 FADDs)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int VARIANT>
__global__ void __launch_bounds__(512, 1) k(int iters, float* out, long long* cycles) {
  __shared__ float lut[4096];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) lut[i] = 1.0f / (1 + i);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t lb = smem_u32(lut);
  uint32_t a[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) a[i] = lb + 4u * (((lane * 5 + i) & 31) + 64u * ((warp * 7 + i * 3) % 40));
  float acc0 = 0, acc1 = 0, acc2 = 0, acc3 = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v[i]) : "r"(a[i]));
    if (VARIANT == 1) {
      // consumers may not be scheduled between the loads: every result passes through one barrier statement
      asm volatile("" : "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3]), "+f"(v[4]), "+f"(v[5]), "+f"(v[6]), "+f"(v[7]),
                        "+f"(v[8]), "+f"(v[9]), "+f"(v[10]), "+f"(v[11]), "+f"(v[12]), "+f"(v[13]), "+f"(v[14]), "+f"(v[15]));
      asm volatile("" : "+f"(v[16]), "+f"(v[17]), "+f"(v[18]), "+f"(v[19]), "+f"(v[20]), "+f"(v[21]), "+f"(v[22]), "+f"(v[23]),
                        "+f"(v[24]), "+f"(v[25]), "+f"(v[26]), "+f"(v[27]), "+f"(v[28]), "+f"(v[29]), "+f"(v[30]), "+f"(v[31]));
    }
#pragma unroll
    for (int i = 0; i < 32; i += 4) { acc0 += v[i]; acc1 += v[i + 1]; acc2 += v[i + 2]; acc3 += v[i + 3]; }
    a[it & 31] ^= 4;
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc0 + acc1 + acc2 + acc3;
}
int main() {
  long long* d; cudaMalloc(&d, 8); long long cyc;
  float* out; cudaMalloc(&out, 148 * 512 * 4);
  const int iters = 2000;
  for (int variant = 0; variant < 2; ++variant)
    for (int warps : {4, 8, 12, 16}) {
      if (variant == 0) { k<0><<<148, warps * 32>>>(10, out, d); cudaDeviceSynchronize(); k<0><<<148, warps * 32>>>(iters, out, d); }
      else { k<1><<<148, warps * 32>>>(10, out, d); cudaDeviceSynchronize(); k<1><<<148, warps * 32>>>(iters, out, d); }
      cudaDeviceSynchronize();
      cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
      double per_sm = (double)iters * warps * 32 / cyc;
      printf("variant %d warps %2d: %.3f LDS/clk/SM, one warp issues an LDS every %.1f clk\n", variant, warps, per_sm, warps / per_sm);
    }
  return 0;
}

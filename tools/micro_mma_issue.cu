// How long does the issuing thread spend in tcgen05.mma / tcgen05.commit / fences? (B200 diagnostics)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do { asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(bar), "r"(parity) : "memory"); } while (!done);
}
__global__ void __launch_bounds__(128, 1) k(long long* out, int nmma, int N) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ uint32_t tmem_base;
  __shared__ __align__(8) uint64_t bar[2];
  for (int i = threadIdx.x; i < 49152 / 4; i += 128) ((uint32_t*)sm)[i] = 0x02020202u;
  int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[0])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[1])));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  asm volatile("fence.proxy.async.shared::cta;");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  if (threadIdx.x == 0) {
    uint32_t idesc = (2u << 4) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    uint32_t tb = tmem_base;
    for (int rep = 0; rep < 6; ++rep) {
      long long t0 = clock64();
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      long long t1 = clock64();
      for (int ks = 0; ks < nmma; ++ks) {
        uint64_t da = make_desc(smem_u32(sm) + ks * 32), db = make_desc(smem_u32(sm) + 16384 + ks * 32);
        uint32_t accum = ks > 0;
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}" ::"r"(tb + (rep & 1) * 256), "l"(da), "l"(db), "r"(idesc), "r"(accum) : "memory");
      }
      long long t2 = clock64();
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[0])) : "memory");
      long long t3 = clock64();
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[1])) : "memory");
      long long t4 = clock64();
      mbar_wait(smem_u32(&bar[0]), rep & 1);
      long long t5 = clock64();
      mbar_wait(smem_u32(&bar[1]), rep & 1);
      long long t6 = clock64();
      out[rep * 8 + 0] = t1 - t0; out[rep * 8 + 1] = t2 - t1; out[rep * 8 + 2] = t3 - t2; out[rep * 8 + 3] = t4 - t3;
      out[rep * 8 + 4] = t5 - t4; out[rep * 8 + 5] = t6 - t5; out[rep * 8 + 6] = t6 - t0;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base));
}
int main() {
  long long* d; cudaMalloc(&d, 8 * 64); long long h[64];
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 49152);
  for (int nmma : {0, 2, 4}) for (int N : {112, 208}) {
    k<<<1, 128, 49152>>>(d, nmma, N); cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h, d, 8 * 64, cudaMemcpyDeviceToHost);
    printf("nmma=%d N=%d (%s): per rep [fence | mma issue | commit1 | commit2 | wait1 | wait2 | total]\n", nmma, N, cudaGetErrorString(e));
    for (int r = 2; r < 6; ++r) printf("   %5lld %5lld %5lld %5lld %5lld %5lld %6lld\n", h[r*8], h[r*8+1], h[r*8+2], h[r*8+3], h[r*8+4], h[r*8+5], h[r*8+6]);
  }
  return 0;
}

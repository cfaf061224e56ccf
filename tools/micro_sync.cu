// Latency of the synchronisation primitives the score kernel's tile loop is built from (B200).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  } while (!done);
}
// mode 0: commit -> wait (one thread). mode 1: plain arrive -> wait (one thread). mode 2: two commits to two barriers, wait both.
// mode 3: thread A commits, 16 warps wait + arrive on second barrier, thread A waits second barrier (ping-pong)
__global__ void k(int mode, int iters, long long* out) {
  __shared__ __align__(8) uint64_t bars[4];
  __shared__ uint32_t tmem_base;
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[0])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[1])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 16;" ::"r"(smem_u32(&bars[2])));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(smem_u32(&tmem_base)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  __syncthreads();
  long long t0 = clock64();
  if (mode <= 2) {
    if (threadIdx.x == 0) {
      for (int i = 0; i < iters; ++i) {
        if (mode == 0) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bars[0])) : "memory");
        else if (mode == 1) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bars[0])) : "memory");
        else {
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bars[0])) : "memory");
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bars[1])) : "memory");
          mbar_wait(smem_u32(&bars[1]), i & 1);
        }
        mbar_wait(smem_u32(&bars[0]), i & 1);
      }
    }
  } else {
    if (threadIdx.x == 0) {
      for (int i = 0; i < iters; ++i) {
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bars[0])) : "memory");
        mbar_wait(smem_u32(&bars[2]), i & 1);
      }
    } else if (warp >= 4 && warp < 20) {
      for (int i = 0; i < iters; ++i) {
        mbar_wait(smem_u32(&bars[0]), i & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bars[2])) : "memory");
      }
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tmem_base));
}
int main() {
  long long* d; cudaMalloc(&d, 8 * 148); long long h;
  const char* names[] = {"commit->wait (1 thread)", "arrive->wait (1 thread)", "2 commits -> 2 waits", "commit -> 16 warps wait+arrive -> wait"};
  for (int grid : {1, 148}) for (int mode = 0; mode < 4; ++mode) {
    int iters = 2000;
    k<<<grid, 640>>>(mode, 10, d); cudaDeviceSynchronize();
    k<<<grid, 640>>>(mode, iters, d); cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("grid %3d %-42s %8.1f clk/iter (%s)\n", grid, names[mode], (double)h / iters, cudaGetErrorString(e));
  }
  return 0;
}

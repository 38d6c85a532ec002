// Micro-benchmark: cycles per tcgen05.mma.cta_group::2 (a CTA pair, M = 256 = 128 rows per CTA, bf16) as a function of N.
// Timing only: operands are zeroed shared memory.
#include <cstdio>
#include <cstdlib>
#include "../../contrast_gan_3d_b200/csrc/tc_common.cuh"

__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}

template <int UNROLL>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
bench2(int N, int nacc, int iters, int sw, int a_stride16, int commit_every, int mask, long long *out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar, dummy;
  __shared__ uint32_t tmem_ptr;
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0;
  if (threadIdx.x == 0) { tc::mbar_init(&bar, 1); tc::mbar_init(&dummy, 1); tc::fence_barrier_init(); }
  tc::fence_proxy_async();
  __syncthreads();
  cluster_sync_all();
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(&tmem_ptr)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc::tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc::tc_fence_after();
  const uint32_t tb = tmem_ptr;
  const uint32_t rank = cluster_rank();
  if (threadIdx.x < 32 && rank == 0) {
    const bool leader = tc::elect_one();
    const uint32_t idesc = tc::make_idesc_bf16(256, N, 0, 0);
    const uint32_t a_u32 = tc::smem_u32(smem), b_u32 = a_u32 + 128 * 1024;
    uint64_t a_d, b_d;
    if (sw == 0) { a_d = tc::make_desc(a_u32, 2048, 128); b_d = tc::make_desc(b_u32, (uint32_t)(N / 2) * 16, 128); }
    else { a_d = tc::make_desc_sw(a_u32, 8 * sw, sw); b_d = tc::make_desc_sw(b_u32, 8 * sw, sw); }
    long long t0 = 0, t1 = 0;
    for (int rep = 0; rep < 2; ++rep) {
      __syncwarp();
      t0 = clock64();
      for (int it = 0; it < iters; ++it) {
        if (leader) {
#pragma unroll
          for (int u = 0; u < UNROLL; ++u) {
            const int acc = (it * UNROLL + u) % nacc;
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "setp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                ::"r"(tb + acc * N), "l"(a_d + (uint64_t)(u * a_stride16)), "l"(b_d), "r"(idesc), "r"(1u)
                : "memory");
            if (commit_every && ((it * UNROLL + u + 1) & (commit_every - 1)) == 0)
              asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                               tc::smem_u32(&dummy)),
                           "h"((uint16_t)mask)
                           : "memory");
          }
        }
        __syncwarp();
      }
      if (leader)
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                         tc::smem_u32(&bar)),
                     "h"((uint16_t)1)
                     : "memory");
      __syncwarp();
      tc::mbar_wait(&bar, rep & 1);
      t1 = clock64();
    }
    if (threadIdx.x == 0) *out = t1 - t0;
  }
  tc::tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512) : "memory");
}

int main() {
  long long *d;
  cudaMalloc(&d, 8);
  cudaFuncSetAttribute(bench2<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int iters = 64, U = 8;
  printf("cta_group::2 M=256: N nacc sw a_stride cycles_per_mma\n");
  for (int sw : {0, 128})
    for (int N : {32, 64, 96, 128, 192, 256})
      for (int nacc : {1, 2})
        for (int astr : {0, 128}) {
          if (nacc * N > 512) continue;
          bench2<U><<<2, 128, 200 * 1024>>>(N, nacc, iters, sw, astr, 0, 3, d);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
          long long c;
          cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
          printf("%d %d %d %d %.1f\n", N, nacc, sw, astr, (double)c / (iters * U));
        }
  printf("commit overhead (N=64, sw=128): commit_every mask cycles_per_mma\n");
  for (int mask : {3, 1})
    for (int ce : {0, 1, 2, 4, 8, 16, 64}) {
      bench2<U><<<2, 128, 200 * 1024>>>(64, 2, iters, 128, 128, ce, mask, d);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      long long c;
      cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
      printf("%d %d %.1f\n", ce, mask, (double)c / (iters * U));
    }
  return 0;
}

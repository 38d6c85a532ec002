// Micro-benchmark: cycles per tcgen05.mma (cta_group::1, kind::f16, bf16) as a function of N, accumulator rotation and
// operand layout.  One CTA, one issuing thread, operands are whatever is in shared memory (timing only).
#include <cstdio>
#include <cstdlib>
#include "../../contrast_gan_3d_b200/csrc/tc_common.cuh"

template <int UNROLL>
__global__ void __launch_bounds__(128, 1) bench(int M, int N, int nacc, int iters, int sw, int a_stride16, long long *out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0;
  if (threadIdx.x == 0) { tc::mbar_init(&bar, 1); tc::fence_barrier_init(); }
  if (threadIdx.x < 32) { tc::tmem_alloc(&tmem_ptr, 512); tc::tmem_relinquish(); }
  tc::fence_proxy_async();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tb = tmem_ptr;
  if (threadIdx.x < 32) {
    const bool leader = tc::elect_one();
    const uint32_t idesc = tc::make_idesc_bf16(M, N, 0, 0);
    const uint32_t a_u32 = tc::smem_u32(smem), b_u32 = a_u32 + 128 * 1024;
    uint64_t a_d, b_d;
    if (sw == 0) { a_d = tc::make_desc(a_u32, 2048, 128); b_d = tc::make_desc(b_u32, (uint32_t)N * 16, 128); }
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
            tc::umma_bf16(tb + acc * N, a_d + (uint64_t)(u * a_stride16), b_d, idesc, 1u);
          }
        }
        __syncwarp();
      }
      if (leader) tc::umma_commit(&bar);
      __syncwarp();
      tc::mbar_wait(&bar, rep & 1);
      t1 = clock64();
    }
    if (threadIdx.x == 0) *out = t1 - t0;
  }
  tc::tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tc::tmem_dealloc(tb, 512);
}

int main() {
  long long *d;
  cudaMalloc(&d, 8);
  cudaFuncSetAttribute(bench<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int iters = 64, U = 8;
  printf("M N nacc sw a_stride cycles_per_mma\n");
  for (int M : {128, 64})
    for (int sw : {0, 32, 128})
      for (int N : {16, 32, 64, 96, 128, 192, 256})
        for (int nacc : {1, 2, 8})
          for (int astr : {0, 128}) {
            if (nacc * N > 512) continue;
            if (M == 64 && sw != 0) continue;
            bench<U><<<1, 128, 200 * 1024>>>(M, N, nacc, iters, sw, astr, d);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            long long c;
            cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
            printf("%d %d %d %d %d %.1f\n", M, N, nacc, sw, astr, (double)c / (iters * U));
          }
  return 0;
}

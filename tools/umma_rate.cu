// umma_rate.cu -- issue-rate microbenchmark for small-N tcgen05.mma (bf16, M=128, K=16) with a lean, unrolled issue loop.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(2);} } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mma_bf16(uint32_t tmem, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(acc));
}
template <int N, int NACC, int TF32>
__global__ void __launch_bounds__(128) rate_kernel(int iters, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  for (int i = tid; i < 48 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  asm volatile("fence.proxy.async.shared::cta;");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tmem_base_s;
  if (tid == 0) {
    const uint32_t fmt = TF32 ? 2u : 1u;
    const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t hi_a = ((uint64_t)(128 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)(16 >> 4) << 16);       // LBO 16 (aliased rows)
    const uint64_t hi_b = ((uint64_t)(128 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)((N * 16) >> 4) << 16);  // LBO N*16
    const uint32_t a0 = smem_u32(smem) >> 4, b0 = (smem_u32(smem) + 16384) >> 4;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int u = 0; u < 28; ++u) {
        const uint64_t ad = hi_a | (uint64_t)(a0 + u * 3 + (u % NACC) * 64);
        const uint64_t bd = hi_b | (uint64_t)(b0 + (u % 7) * ((2 * N * 16) >> 4));
        if (TF32) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem + (u % NACC) * N), "l"(ad), "l"(bd), "r"(idesc), "r"(1u));
        else mma_bf16(tmem + (u % NACC) * N, ad, bd, idesc, 1u);
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)));
    uint32_t done = 0;
    while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0));
    cycles[0] = clock64() - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}
template <int N, int NACC, int TF32>
void go(const char* what) {
  long long* cyc_d; long long cyc;
  CK(cudaMalloc(&cyc_d, 8));
  CK(cudaFuncSetAttribute(rate_kernel<N, NACC, TF32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  const int iters = 200;
  rate_kernel<N, NACC, TF32><<<1, 128, 64 * 1024>>>(iters, cyc_d);
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(&cyc, cyc_d, 8, cudaMemcpyDeviceToHost));
  printf("%s M128 N%-3d acc=%d: %.1f cycles/MMA (floor %d)\n", what, N, NACC, (double)cyc / (iters * 28.0), N / 2);
  cudaFree(cyc_d);
}
int main() {
  go<16, 1, 0>("bf16"); go<16, 4, 0>("bf16"); go<32, 1, 0>("bf16"); go<32, 4, 0>("bf16"); go<64, 1, 0>("bf16"); go<64, 4, 0>("bf16");
  go<128, 1, 0>("bf16"); go<128, 2, 0>("bf16"); go<256, 1, 0>("bf16"); go<256, 2, 0>("bf16");
  go<16, 4, 1>("tf32"); go<32, 4, 1>("tf32"); go<64, 4, 1>("tf32"); go<80, 1, 1>("tf32"); go<80, 4, 1>("tf32"); go<128, 2, 1>("tf32"); go<256, 1, 1>("tf32");
  return 0;
}

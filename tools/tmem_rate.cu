// tmem_rate.cu -- how fast can the epilogue warps drain TMEM?  tcgen05.ld.32x32b.xN throughput per SM as a function of
// the number of warps (each warp reads its own 32-lane quadrant) and of the load width, with and without a stream of
// tcgen05.mma (M = 128, N = 112, K = 16) running beside it.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/tmem_rate tools/tmem_rate.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(2);} } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int W>
__device__ __forceinline__ uint32_t ld_cols(uint32_t addr);
template <>
__device__ __forceinline__ uint32_t ld_cols<16>(uint32_t addr) {
  uint32_t r[16];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                 "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(addr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s ^= r[i];
  return s;
}
template <>
__device__ __forceinline__ uint32_t ld_cols<32>(uint32_t addr) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
        "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]),
        "=r"(r[30]), "=r"(r[31])
      : "r"(addr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 32; ++i) s ^= r[i];
  return s;
}

__device__ __forceinline__ void mma_f16(uint32_t tmem, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(acc));
}

// warps 0 .. n_ld-1 load; warp n_ld (if with_mma) issues MMAs into columns [256, 368)
template <int W>
__global__ void __launch_bounds__(1024) rate_kernel(int n_ld, int with_mma, int iters, long long* cycles, uint32_t* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  for (int i = tid; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  asm volatile("fence.proxy.async.shared::cta;");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tmem_base_s;
  long long t0 = clock64();
  if (warp < n_ld) {
    const uint32_t base = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t s = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int c = 0; c < 128; c += W) s ^= ld_cols<W>(base + c + ((warp >> 2) & 1) * 128);
    }
    if (s == 0x12345678u) sink[tid] = s;
    cycles[blockIdx.x * 64 + warp] = clock64() - t0;
  } else if (warp == n_ld && with_mma && (tid & 31) == 0) {
    const uint32_t idesc = (1u << 4) | ((uint32_t)(112 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t v1 = (uint64_t)1 << 46;
    const uint64_t hi_a = v1 | ((uint64_t)(128 >> 4) << 32) | ((uint64_t)(8192 >> 4) << 16);
    const uint64_t hi_b = v1 | ((uint64_t)(128 >> 4) << 32) | ((uint64_t)(1792 >> 4) << 16);
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem) + 32768;
    const int n_mma = iters * 128 * 4 / 64;  // about as long as the loads at 64 cycles per MMA
    for (int i = 0; i < n_mma; ++i) {
      const uint64_t ad = hi_a | (uint64_t)(((a0 + (i & 15) * 16) >> 4) & 0x3FFF);
      const uint64_t bd = hi_b | (uint64_t)((b0 >> 4) & 0x3FFF);
      mma_f16(tmem + 256 + (i & 1) * 128, ad, bd, idesc, 1u);
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)));
    uint32_t done = 0;
    while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0));
    cycles[blockIdx.x * 64 + 63] = clock64() - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

template <int W>
static void go(int n_ld, int with_mma) {
  const int n_cta = 148, iters = 200;
  long long* cyc_d;
  uint32_t* sink;
  CK(cudaMalloc(&cyc_d, 8 * 64 * n_cta));
  CK(cudaMemset(cyc_d, 0, 8 * 64 * n_cta));
  CK(cudaMalloc(&sink, 4 * 1024));
  CK(cudaFuncSetAttribute(rate_kernel<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  rate_kernel<W><<<n_cta, 32 * (n_ld + 1), 64 * 1024>>>(n_ld, with_mma, iters, cyc_d, sink);
  CK(cudaDeviceSynchronize());
  long long* cyc = new long long[64 * n_cta];
  CK(cudaMemcpy(cyc, cyc_d, 8 * 64 * n_cta, cudaMemcpyDeviceToHost));
  double mx = 0, mma = 0;
  for (int i = 0; i < n_cta; ++i) {
    for (int w = 0; w < n_ld; ++w) mx = cyc[i * 64 + w] > mx ? cyc[i * 64 + w] : mx;
    mma = cyc[i * 64 + 63] > mma ? cyc[i * 64 + 63] : mma;
  }
  const double bytes = (double)n_ld * iters * 128 * 32 * 4;  // per SM
  printf("ld.x%-2d warps=%-2d mma=%d: %8.0f cycles -> %6.1f B/cycle/SM (%5.1f B/cycle/warp)", W, n_ld, with_mma, mx, bytes / mx, bytes / mx / n_ld);
  if (with_mma) printf("   MMA stream: %6.1f cycles/MMA", mma / (iters * 128 * 4 / 64));
  printf("\n");
  delete[] cyc;
  cudaFree(cyc_d);
  cudaFree(sink);
}

int main() {
  for (int mma = 0; mma < 2; ++mma)
    for (int n : {1, 4, 8, 16}) {
      go<16>(n, mma);
      go<32>(n, mma);
    }
  return 0;
}

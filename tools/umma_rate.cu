// umma_rate.cu -- issue-rate microbenchmark for tcgen05.mma (kind::f16, fp16 inputs, M = 128 or 64, K = 16):
// how many cycles one MMA costs as a function of N, of the operand layout (SWIZZLE_NONE chunk planes with different
// LBO, or the SWIZZLE_128B K-major tile TMA writes) and of the number of accumulators in flight.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/umma_rate tools/umma_rate.cu
// The numbers steer the kernel designs in audio_key_estimation_b200/csrc (see DESIGN.md section "MMA cost model").
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(2);} } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mma_f16(uint32_t tmem, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(acc));
}

struct Cfg {
  int N, M, nacc;
  int a_swz;          // 0: SWIZZLE_NONE, 1: SWIZZLE_128B
  uint32_t a_lbo;     // bytes (SWIZZLE_NONE)
  uint32_t a_step;    // bytes the A start address advances per MMA (walks over the operand like a real kernel)
  int b_swz;
  uint32_t b_lbo;
};

__global__ void __launch_bounds__(128) rate_kernel(Cfg c, int iters, long long* cycles, int smem_bytes, uint32_t b_base) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(128));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  for (int i = tid; i < smem_bytes / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
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
    const uint32_t idesc = (1u << 4) | ((uint32_t)(c.N >> 3) << 17) | ((uint32_t)(c.M >> 4) << 24);
    const uint64_t v1 = (uint64_t)1 << 46;
    // SWIZZLE_NONE: LBO = chunk stride, SBO = 128 B (8 rows x 16 B).  SWIZZLE_128B (layout type 2): SBO = 1024 B.
    const uint64_t hi_a = c.a_swz ? (v1 | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 16) | ((uint64_t)2 << 61))
                                  : (v1 | ((uint64_t)(128 >> 4) << 32) | ((uint64_t)(c.a_lbo >> 4) << 16));
    const uint64_t hi_b = c.b_swz ? (v1 | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 16) | ((uint64_t)2 << 61))
                                  : (v1 | ((uint64_t)(128 >> 4) << 32) | ((uint64_t)(c.b_lbo >> 4) << 16));
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem) + b_base;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        const uint32_t aoff = c.a_swz ? (uint32_t)(u & 3) * 32 + (uint32_t)(u >> 2) * 16384 : (uint32_t)u * c.a_step;
        const uint32_t boff = c.b_swz ? (uint32_t)(u & 3) * 32 : (uint32_t)(u & 7) * 2 * c.b_lbo;
        const uint64_t ad = hi_a | (uint64_t)(((a0 + aoff) >> 4) & 0x3FFF);
        const uint64_t bd = hi_b | (uint64_t)(((b0 + boff) >> 4) & 0x3FFF);
        mma_f16(tmem + (u % c.nacc) * c.N, ad, bd, idesc, 1u);
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)));
    uint32_t done = 0;
    while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0));
    cycles[blockIdx.x] = clock64() - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128));
}

// Blocks of `per` MMAs into rotating accumulators (the first MMA of a block overwrites), one tcgen05.commit per block on a
// barrier nobody waits for: does switching accumulators / committing cost tensor-pipe cycles?
__global__ void __launch_bounds__(128) block_kernel(int N, int per, int nacc, int commit_each, int iters, long long* cycles, uint32_t blk_step,
                                                     uint32_t tap_step, uint32_t first_acc, int distinct, int n_issue) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) uint64_t bar, bars[8];
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  for (int i = tid; i < 96 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    for (int i = 0; i < 8; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[i])));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  asm volatile("fence.proxy.async.shared::cta;");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tmem_base_s;
  // converged warp + elect.sync, blocks unrolled: the issue path of the library's kernels (umma.cuh: single-lane issue)
  const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
  if (warp_u < n_issue) {
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t v1 = (uint64_t)1 << 46;
    const uint64_t hi_a = v1 | ((uint64_t)(128 >> 4) << 32) | ((uint64_t)(37344 >> 4) << 16);
    const uint64_t hi_b = v1 | ((uint64_t)(128 >> 4) << 32) | ((uint64_t)((uint32_t)N * 16 >> 4) << 16);
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem) + 80 * 1024;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const uint32_t d = tmem + (uint32_t)(it & (nacc - 1)) * 128 + (uint32_t)warp_u * 256;  // nacc is a power of two: no division on the issue path
      const uint32_t ab = a0 + (uint32_t)(it & 7) * blk_step;
      uint32_t el;
      asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(el));
      if (el) {
        if (per == 7 && distinct) {
          // every MMA gets its own copies of the accumulator address and the instruction descriptor (volatile moves the
          // compiler cannot merge): do back-to-back UTCHMMAs stall on shared uniform source registers?
          uint32_t dd[7], id[7];
#pragma unroll
          for (int u = 0; u < 7; ++u) {
            asm volatile("mov.b32 %0, %1;" : "=r"(dd[u]) : "r"(d));
            asm volatile("mov.b32 %0, %1;" : "=r"(id[u]) : "r"(idesc));
          }
#pragma unroll
          for (int u = 0; u < 7; ++u)
            mma_f16(dd[u], hi_a | (uint64_t)(((ab + (uint32_t)u * tap_step) >> 4) & 0x3FFF), hi_b | (uint64_t)(((b0 + (uint32_t)(u % 4) * 3584) >> 4) & 0x3FFF), id[u], u ? 1u : first_acc);
        } else if (per == 7) {
#pragma unroll
          for (int u = 0; u < 7; ++u)
            mma_f16(d, hi_a | (uint64_t)(((ab + (uint32_t)u * tap_step) >> 4) & 0x3FFF), hi_b | (uint64_t)(((b0 + (uint32_t)(u % 4) * 3584) >> 4) & 0x3FFF), idesc, u ? 1u : first_acc);
        } else {
#pragma unroll 4
          for (int u = 0; u < per; ++u)
            mma_f16(d, hi_a | (uint64_t)(((ab + (uint32_t)u * tap_step) >> 4) & 0x3FFF), hi_b | (uint64_t)(((b0 + (uint32_t)(u % 4) * 3584) >> 4) & 0x3FFF), idesc, u ? 1u : 0u);
        }
        if (commit_each) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bars[it & 7])));
      }
      __syncwarp();
    }
    if ((tid & 31) == 0) {
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(warp_u ? &bars[7] : &bar)));
      uint32_t done = 0;
      while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(smem_u32(warp_u ? &bars[7] : &bar)), "r"(0));
      if (warp_u == 0) cycles[blockIdx.x] = clock64() - t0;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

static void go_blocks(int N, int per, int nacc, int commit_each, uint32_t blk_step = 1952, uint32_t tap_step = 2512, uint32_t first_acc = 0, int distinct = 0, int n_issue = 1) {
  const int n_cta = 148, iters = 800;
  long long* cyc_d;
  CK(cudaMalloc(&cyc_d, 8 * n_cta));
  CK(cudaFuncSetAttribute(block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  block_kernel<<<n_cta, 128, 160 * 1024>>>(N, per, nacc, commit_each, iters, cyc_d, blk_step, tap_step, first_acc, distinct, n_issue);
  CK(cudaDeviceSynchronize());
  long long* cyc = new long long[n_cta];
  CK(cudaMemcpy(cyc, cyc_d, 8 * n_cta, cudaMemcpyDeviceToHost));
  double mx = 0;
  for (int i = 0; i < n_cta; ++i) mx = cyc[i] > mx ? cyc[i] : mx;
  printf("blocks of %2d MMAs N%-3d, %d accumulators, commit per block %d, A steps %u/%u B: %6.1f cycles/MMA (%6.1f per block)\n", per, N, nacc,
         commit_each, blk_step, tap_step,
         mx / ((double)iters * per), mx / iters);
  delete[] cyc;
  cudaFree(cyc_d);
}

static void go(const char* what, Cfg c, int ctas_per_sm = 1) {
  const int n_cta = 148 * ctas_per_sm;
  long long* cyc_d;
  CK(cudaMalloc(&cyc_d, 8 * n_cta));
  const int smem = ctas_per_sm == 1 ? 160 * 1024 : 48 * 1024;
  const uint32_t b_base = ctas_per_sm == 1 ? 96 * 1024 : 24 * 1024;
  CK(cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  const int iters = 400;
  rate_kernel<<<n_cta, 128, smem>>>(c, iters, cyc_d, smem, b_base);
  CK(cudaDeviceSynchronize());
  long long* cyc = new long long[n_cta];
  CK(cudaMemcpy(cyc, cyc_d, 8 * n_cta, cudaMemcpyDeviceToHost));
  double mx = 0;
  for (int i = 0; i < n_cta; ++i) mx = cyc[i] > mx ? cyc[i] : mx;
  printf("%-40s M%-3d N%-3d acc=%d CTAs/SM=%d: %7.1f cycles/MMA per CTA -> %6.1f cycles/MMA per SM\n", what, c.M, c.N, c.nacc, ctas_per_sm,
         mx / (iters * 16.0), mx / (iters * 16.0) / ctas_per_sm);
  delete[] cyc;
  cudaFree(cyc_d);
}

int main(int argc, char** argv) {
  if (argc > 1 && argv[1][0] == 'b') {
    for (int nacc : {1, 2, 4})
      for (int ce : {0, 1}) go_blocks(112, 7, nacc, ce);
    go_blocks(112, 14, 4, 1);
    go_blocks(112, 7, 4, 1, 1920, 2560);  // every A start 128-byte aligned
    go_blocks(112, 7, 4, 1, 2048, 2048);
    go_blocks(112, 7, 4, 1, 1952, 2560);
    go_blocks(112, 7, 4, 1, 0, 2512);  // loop-invariant descriptors: the bare UTCHMMA issue rate
    go_blocks(112, 7, 1, 0, 0, 2512);
    go_blocks(64, 8, 1, 0, 0, 2512);
    printf("-- first MMA of a block accumulates too (no overwrite):\n");
    go_blocks(112, 7, 4, 1, 1952, 2512, 1);
    go_blocks(112, 7, 1, 0, 0, 2512, 1);
    printf("-- distinct uniform registers per MMA:\n");
    go_blocks(112, 7, 4, 1, 1952, 2512, 0, 1);
    go_blocks(64, 7, 4, 1, 1952, 2512, 0, 1);
    go_blocks(64, 7, 4, 1, 1952, 2512, 0, 0);
    printf("-- two issuing warps (cycles per MMA of EACH warp: halve for the SM rate):\n");
    go_blocks(64, 7, 2, 1, 1952, 2512, 0, 0, 2);
    go_blocks(112, 7, 2, 1, 1952, 2512, 0, 0, 2);
    go_blocks(32, 7, 2, 1, 1952, 2512, 0, 0, 2);
    go_blocks(32, 7, 2, 1, 1952, 2512, 0, 0, 1);
    go_blocks(224, 12, 2, 1);
    go_blocks(64, 8, 4, 1);
    return 0;
  }
  if (argc > 1) {
    // p2p-like operand walk: which of {N, LBO between the two K chunks, start-address step, accumulators} costs cycles?
    for (int N : {96, 112, 128})
      for (uint32_t lbo : {2064u, 37344u})
        for (uint32_t step : {16u, 2512u})
          for (int nacc : {1, 4}) {
            if (N * nacc > 128) continue;
            char what[64];
            snprintf(what, sizeof what, "A LBO=%u step=%u", lbo, step);
            go(what, Cfg{N, 128, nacc, 0, lbo, step, 0, (uint32_t)N * 16}, 1);
          }
    return 0;
  }
  for (int N : {16, 32, 64, 96, 128})
    for (int k : {1, 2, 3, 4}) go("none A LBO=2064, B none", Cfg{N, 128, 1, 0, 2064, 16, 0, (uint32_t)N * 16}, k);
  return 0;
}

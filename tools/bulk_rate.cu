// bulk_rate.cu -- how much HBM bandwidth do cp.async.bulk copies deliver per SM, as a function of the bytes kept in
// flight (stages x stage size), of the size of one copy and of who issues them?  One persistent CTA per SM streams a
// 2 GiB buffer (>> L2) through a ring of shared-memory stages; a consumer warp only waits for each stage and frees it.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/bulk_rate tools/bulk_rate.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(2);} } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
}

// warp 0: producer (lane l issues pieces l, l + 32, ...), warp 1: consumer
__global__ void __launch_bounds__(64) bulk_kernel(const uint8_t* src, size_t total, int stages, uint32_t stage_bytes, uint32_t piece, int ctas) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t full[16], empty[16];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full[s])));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&empty[s])));
    }
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  __syncthreads();
  const size_t n_chunks = total / stage_bytes;
  const int pieces = stage_bytes / piece;
  int k = 0;
  for (size_t c = blockIdx.x; c < n_chunks; c += ctas, ++k) {
    const int s = k % stages;
    const uint32_t ph = (k / stages) & 1;
    if (warp == 0) {
      mbar_wait(&empty[s], ph ^ 1);
      if (lane == 0)
        asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}\n" ::"r"(smem_u32(&full[s])), "r"(stage_bytes) : "memory");
      __syncwarp();
      for (int p = lane; p < pieces; p += 32)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         smem_u32(smem + (size_t)s * stage_bytes + (size_t)p * piece)),
                     "l"(src + c * stage_bytes + (size_t)p * piece), "r"(piece), "r"(smem_u32(&full[s]))
                     : "memory");
    } else {
      mbar_wait(&full[s], ph);
      if (lane == 0) asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}\n" ::"r"(smem_u32(&empty[s])) : "memory");
      __syncwarp();
    }
  }
}

// plain vector loads for comparison: `warps` warps per CTA, `unroll` independent 16-byte loads in flight per thread
template <int U>
__global__ void ldg_kernel(const uint4* src, size_t n16, uint32_t* sink) {
  uint32_t acc = 0;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + (U - 1) * stride < n16; i += U * stride) {
    uint4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = __ldg(src + i + u * stride);
#pragma unroll
    for (int u = 0; u < U; ++u) acc ^= v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
  }
  if (acc == 0x12345678u) sink[threadIdx.x] = acc;
}

int main() {
  const size_t total = (size_t)2 << 30;
  uint8_t* buf;
  CK(cudaMalloc(&buf, total));
  CK(cudaMemset(buf, 1, total));
  uint32_t* sink;
  CK(cudaMalloc(&sink, 4096));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  CK(cudaFuncSetAttribute(bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  struct C { int stages; uint32_t stage_bytes, piece; int per_sm; };
  const C cfgs[] = {{2, 65536, 2048, 1},  {2, 65536, 16384, 1}, {2, 65536, 65536, 1}, {3, 65536, 2048, 1}, {4, 32768, 2048, 1},
                    {4, 32768, 32768, 1}, {8, 16384, 16384, 1}, {8, 16384, 2048, 1},  {12, 16384, 16384, 1}, {2, 32768, 2048, 2},
                    {2, 32768, 32768, 2}, {4, 16384, 16384, 2}, {3, 16384, 16384, 4}, {2, 16384, 2048, 4}, {6, 32768, 32768, 1}};
  for (const C& c : cfgs) {
    const int ctas = 148 * c.per_sm;
    const size_t smem = (size_t)c.stages * c.stage_bytes;
    for (int rep = 0; rep < 2; ++rep) {
      CK(cudaEventRecord(e0));
      bulk_kernel<<<ctas, 64, smem>>>(buf, total, c.stages, c.stage_bytes, c.piece, ctas);
      CK(cudaEventRecord(e1));
      CK(cudaDeviceSynchronize());
    }
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("bulk: %2d stages x %6u B (pieces of %6u B), %d CTA/SM, %4zu KB in flight/SM: %7.1f GB/s\n", c.stages, c.stage_bytes, c.piece,
           c.per_sm, smem * c.per_sm / 1024, total / ms / 1e6);
  }
  for (int warps : {8, 16, 32}) {
    for (int per_sm : {1, 2}) {
      float ms;
      for (int rep = 0; rep < 2; ++rep) {
        CK(cudaEventRecord(e0));
        ldg_kernel<4><<<148 * per_sm, 32 * warps>>>(reinterpret_cast<const uint4*>(buf), total / 16, sink);
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
      }
      CK(cudaEventElapsedTime(&ms, e0, e1));
      printf("ldg.128 x4: %2d warps x %d CTA/SM: %7.1f GB/s\n", warps, per_sm, total / ms / 1e6);
      for (int rep = 0; rep < 2; ++rep) {
        CK(cudaEventRecord(e0));
        ldg_kernel<8><<<148 * per_sm, 32 * warps>>>(reinterpret_cast<const uint4*>(buf), total / 16, sink);
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
      }
      CK(cudaEventElapsedTime(&ms, e0, e1));
      printf("ldg.128 x8: %2d warps x %d CTA/SM: %7.1f GB/s\n", warps, per_sm, total / ms / 1e6);
    }
  }
  return 0;
}

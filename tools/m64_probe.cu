// m64_probe.cu -- where does tcgen05.mma (cta_group::1, kind::f16) put the 64 rows of an M = 64 accumulator in TMEM?
// D[r, n] = r for every column: A[r, 0] = r, B[n, 0] = 1, everything else 0.  Prints TMEM lane -> row for all 128 lanes.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I audio_key_estimation_b200/csrc -o tools/bin/m64_probe tools/m64_probe.cu
#include <cstdio>
#include <cuda_fp16.h>
#include "umma.cuh"
using namespace ake::umma;

__global__ void probe(float* out, int M) {
  __shared__ __align__(1024) uint8_t a_img[2 * 128 * 16];  // [chunk 2][row 128][8 halves]
  __shared__ __align__(1024) uint8_t b_img[2 * 16 * 16];   // [chunk 2][n 16][8 halves]
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 2 * 128 * 8; i += blockDim.x) reinterpret_cast<__half*>(a_img)[i] = __float2half(0.f);
  for (int i = tid; i < 2 * 16 * 8; i += blockDim.x) reinterpret_cast<__half*>(b_img)[i] = __float2half(0.f);
  __syncthreads();
  if (tid < 128) reinterpret_cast<__half*>(a_img)[tid * 8] = __float2half((float)tid);
  if (tid < 16) reinterpret_cast<__half*>(b_img)[tid * 8] = __float2half(1.f);
  if (warp == 0) tmem_alloc(&slot, 32);
  if (tid == 0) mbar_init(&bar, 1), mbar_init_fence();
  fence_proxy_async();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = slot;
  // pre-fill the accumulator lanes with -1 so untouched lanes are visible: an M = 128 MMA with zero A, then the probe MMA on top
  if (warp == 0) {
    if (elect_one()) {
      mma_f16(tmem, make_desc(desc_hi(128 * 16), smem_u32(a_img)), make_desc(desc_hi(16 * 16), smem_u32(b_img)), idesc_f16(16, M), 0u);
      commit(&bar);
    }
    __syncwarp();
  }
  mbar_wait(&bar, 0);
  fence_after_sync();
  float v[8];
  tmem_ld8(tmem + ((uint32_t)(warp * 32) << 16), v);
  out[tid] = v[0];
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 32);
}

int main() {
  float* d;
  cudaMalloc(&d, 128 * sizeof(float));
  for (int M : {128, 64}) {
    cudaMemset(d, 0xFF, 128 * sizeof(float));
    probe<<<1, 128>>>(d, M);
    float h[128];
    cudaError_t e = cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
    printf("M=%d (%s): lane -> value\n", M, cudaGetErrorString(e));
    for (int i = 0; i < 128; ++i) printf("%s%3d:%-6.0f", i % 16 == 0 ? "\n" : " ", i, h[i]);
    printf("\n");
  }
  return 0;
}

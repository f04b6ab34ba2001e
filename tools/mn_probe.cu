// mn_probe.cu -- do MN-major SWIZZLE_NONE operands with OVERLAPPING core matrices work?  (the weight-gradient GEMM of the 7x7 conv:
// K = 16 consecutive positions, M = 8 time taps x 8 input channels read from ONE activation row at SBO = 16 B, N = 7 row taps x 8 output
// channels read from seven gradient rows at SBO = row pitch; both operands are position-major chunk planes, i.e. MN-major)
//   D[(g, ci), (j, co)] = sum_{k < 16} X[k + g][ci] * G[j * Wg + k][co]
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I audio_key_estimation_b200/csrc -o tools/bin/mn_probe tools/mn_probe.cu
#include <cstdio>
#include <cuda_fp16.h>
#include "umma.cuh"
using namespace ake::umma;

constexpr int kWg = 24, kXPos = 64, kGPos = 7 * kWg + 16;

__host__ __device__ inline float xval(int pos, int ci) { return (float)((pos * 3 + ci * 5) % 7 - 3); }
__host__ __device__ inline float gval(int pos, int co) { return (float)((pos * 2 + co * 3) % 5 - 2); }

__global__ void probe(float* out, uint32_t a_sbo, uint32_t a_lbo, uint32_t b_sbo, uint32_t b_lbo) {
  __shared__ __align__(1024) uint8_t x_img[kXPos * 16];
  __shared__ __align__(1024) uint8_t g_img[kGPos * 16];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < kXPos * 8; i += blockDim.x) reinterpret_cast<__half*>(x_img)[i] = __float2half(xval(i / 8, i % 8));
  for (int i = tid; i < kGPos * 8; i += blockDim.x) reinterpret_cast<__half*>(g_img)[i] = __float2half(gval(i / 8, i % 8));
  if (warp == 0) tmem_alloc(&slot, 64);
  if (tid == 0) mbar_init(&bar, 1), mbar_init_fence();
  fence_proxy_async();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = slot;
  if (warp == 0) {
    if (elect_one()) {
      const uint32_t idesc = idesc_f16(56, 64) | (1u << 15) | (1u << 16);  // A and B MN-major
      mma_f16(tmem, make_desc(desc_hi(a_lbo, a_sbo), smem_u32(x_img)), make_desc(desc_hi(b_lbo, b_sbo), smem_u32(g_img)), idesc, 0u);
      commit(&bar);
    }
    __syncwarp();
  }
  mbar_wait(&bar, 0);
  fence_after_sync();
  // M = 64: accumulator rows 16 q .. 16 q + 15 sit in lanes 0..15 of TMEM quadrant q
  const int lane = tid & 31;
  for (int c0 = 0; c0 < 56; c0 += 8) {
    float v[8];
    tmem_ld8(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
    if (lane < 16)
      for (int e = 0; e < 8; ++e) out[(16 * warp + lane) * 56 + c0 + e] = v[e];
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 64);
}

int main() {
  float* d;
  cudaMalloc(&d, 64 * 56 * sizeof(float));
  struct { uint32_t a_sbo, a_lbo, b_sbo, b_lbo; const char* what; } cfgs[] = {
      {16, 128, kWg * 16, 128, "SBO = group stride, LBO = 8-position stride"},
      {128, 16, 128, kWg * 16, "swapped roles"},
  };
  for (auto& c : cfgs) {
    cudaMemset(d, 0, 64 * 56 * sizeof(float));
    probe<<<1, 128>>>(d, c.a_sbo, c.a_lbo, c.b_sbo, c.b_lbo);
    static float h[64 * 56];
    cudaError_t e = cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
    int bad = 0;
    double worst = 0;
    for (int m = 0; m < 64; ++m)
      for (int n = 0; n < 56; ++n) {
        const int g = m / 8, ci = m % 8, j = n / 8, co = n % 8;
        float want = 0.f;
        for (int k = 0; k < 16; ++k) want += xval(k + g, ci) * gval(j * kWg + k, co);
        const double err = fabs((double)h[m * 56 + n] - want);
        if (err > 1e-3) {
          if (bad < 6) printf("  D[%d (g %d ci %d)][%d (j %d co %d)] = %g, want %g\n", m, g, ci, n, j, co, h[m * 56 + n], want);
          ++bad;
        }
        worst = err > worst ? err : worst;
      }
    printf("%s (%s): %d of %d wrong, worst %.3g\n", c.what, cudaGetErrorString(e), bad, 64 * 56, worst);
  }
  return 0;
}

// umma_probe.cu -- stand-alone probe of the tcgen05 building blocks the conv / filter-bank kernels rely on.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/bin/umma_probe tools/umma_probe.cu
// It checks, against a CPU result, that
//   * SWIZZLE_NONE K-major shared-memory descriptors address   chunk(r, c) = start + r*16 B + c*LBO   (SBO = 128 B),
//   * the start address may be offset by any multiple of 16 B (row shifts: the "shift-GEMM" convolution),
//   * LBO = 16 B (chunk c of row r aliases chunk 0 of row r + c) is accepted (two time taps per bf16 MMA),
//   * kind::tf32 reads raw fp32 bits (reports whether it truncates or rounds the low 13 bits),
//   * accumulate / commit / tcgen05.ld behave as the production kernels assume,
// and it measures the issue-to-completion cost of back-to-back MMAs of the shapes we use.
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x)                                                                        \
  do {                                                                               \
    cudaError_t e = (x);                                                             \
    if (e != cudaSuccess) {                                                          \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
      exit(2);                                                                       \
    }                                                                                \
  } while (0)

struct Case {
  int kind;        // 0 tf32, 1 bf16
  int N;           // 16..256
  int n_mma;       // MMAs accumulated into the same D
  uint32_t a_bytes, b_bytes;                 // smem image sizes
  uint32_t a_off0, a_step, b_off0, b_step;   // start offset of MMA j = off0 + j*step
  uint32_t a_lbo, a_sbo, b_lbo, b_sbo;
  int repeat;      // timing: issue the whole MMA list this many times (accumulating garbage) when > 1
  int n_acc;       // timing: round-robin over this many independent accumulators (TMEM column blocks of N)
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version 1 (Blackwell)
  return d;                // base_offset 0, lbo_mode 0, layout SWIZZLE_NONE
}

__device__ __forceinline__ uint32_t make_idesc(int kind, int M, int N) {
  uint32_t d = 0;
  d |= 1u << 4;                          // D format f32
  const uint32_t f = kind == 0 ? 2u : 1u;  // tf32 : bf16
  d |= f << 7;
  d |= f << 10;
  d |= (uint32_t)(N >> 3) << 17;
  d |= (uint32_t)(M >> 4) << 24;
  return d;  // K-major A and B, no negate, dense
}

__global__ void __launch_bounds__(128) probe_kernel(Case c, const uint8_t* __restrict__ a_img, const uint8_t* __restrict__ b_img,
                                                    float* __restrict__ d_out, long long* __restrict__ cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) uint64_t bar;
  uint8_t* a_s = smem;
  uint8_t* b_s = smem + ((c.a_bytes + 1023) / 1024) * 1024;
  const int tid = threadIdx.x, warp = tid >> 5;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  for (uint32_t i = tid * 16; i < c.a_bytes; i += 128 * 16) *reinterpret_cast<uint4*>(a_s + i) = *reinterpret_cast<const uint4*>(a_img + i);
  for (uint32_t i = tid * 16; i < c.b_bytes; i += 128 * 16) *reinterpret_cast<uint4*>(b_s + i) = *reinterpret_cast<const uint4*>(b_img + i);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  asm volatile("fence.proxy.async.shared::cta;");  // generic-proxy smem writes -> visible to the tensor core (async proxy)
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tmem_base_s;

  long long t0 = 0, t1 = 0;
  const uint32_t tmem0 = tmem;
  if (tid == 0) {
    const uint32_t idesc = make_idesc(c.kind, 128, c.N);
    t0 = clock64();
    for (int rep = 0; rep < c.repeat; ++rep) {
      for (int j = 0; j < c.n_mma; ++j) {
        const uint64_t ad = make_desc(smem_u32(a_s) + c.a_off0 + j * c.a_step, c.a_lbo, c.a_sbo);
        const uint64_t bd = make_desc(smem_u32(b_s) + c.b_off0 + j * c.b_step, c.b_lbo, c.b_sbo);
        const uint32_t acc = (j > 0 || rep > 0) ? 1u : 0u;
        const uint32_t tmem = tmem0 + (uint32_t)((rep * c.n_mma + j) % c.n_acc) * c.N;
        if (c.kind == 0) {
          asm volatile(
              "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
              "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem),
              "l"(ad), "l"(bd), "r"(idesc), "r"(acc));
        } else {
          asm volatile(
              "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
              "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem),
              "l"(ad), "l"(bd), "r"(idesc), "r"(acc));
        }
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)));
  }
  // everyone waits for the MMAs (phase 0)
  {
    uint32_t done = 0;
    while (!done) {
      asm volatile(
          "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
          : "=r"(done)
          : "r"(smem_u32(&bar)), "r"(0));
    }
  }
  if (tid == 0) {
    t1 = clock64();
    cycles[0] = t1 - t0;
  }
  asm volatile("tcgen05.fence::after_thread_sync;");
  // D row = TMEM lane, column = TMEM column; warp w reads lanes 32w .. 32w+31
  for (int col0 = 0; col0 < c.N; col0 += 8) {
    uint32_t v[8];
    const uint32_t addr = tmem + ((uint32_t)(warp * 32) << 16) + col0;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(addr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 8; ++j) d_out[tid * c.N + col0 + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256));
}

// ------------------------------------------------------------------------------------------------ host
static float tf32_trunc(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  u &= 0xFFFFE000u;
  memcpy(&x, &u, 4);
  return x;
}
static float tf32_rn(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  u += 0x1000u;
  u &= 0xFFFFE000u;
  memcpy(&x, &u, 4);
  return x;
}

struct Problem {
  Case c;
  std::vector<uint8_t> a_img, b_img;
  std::vector<double> expect, expect_alt;  // expect_alt: tf32 round-to-nearest interpretation
  const char* name;
};

static uint32_t rng_state = 12345;
static float rnd_small() {  // multiples of 1/8 in [-2, 2): exact in bf16 and tf32
  rng_state = rng_state * 1664525u + 1013904223u;
  return (float)((int)((rng_state >> 16) % 33) - 16) / 8.0f;
}
static float rnd_full() {  // full-mantissa fp32 values
  rng_state = rng_state * 1664525u + 1013904223u;
  return ((float)(rng_state >> 8) / 16777216.0f) * 2.0f - 1.0f;
}

// Generic builder.  Logical problem: D[r][n] = sum_j sum_k A_j[r][k] * B_j[n][k], j < n_mma, k < K (8 tf32 / 16 bf16),
// where A_j[r][k] = pool value at position (a_pos0 + j*a_pos_step + r + (k / CH) * a_chunk_pos_step), channel (k % CH) of
// channel-group (k / CH) * a_chunk_grp_step ... kept simple: we build images from an explicit address function instead.
template <class AddrA, class AddrB>
static Problem build(const char* name, int kind, int N, int n_mma, uint32_t a_bytes, uint32_t b_bytes, uint32_t a_lbo, uint32_t a_sbo,
                     uint32_t b_lbo, uint32_t b_sbo, uint32_t a_off0, uint32_t a_step, uint32_t b_off0, uint32_t b_step, bool full_mantissa,
                     AddrA addr_a, AddrB addr_b) {
  Problem p;
  p.name = name;
  p.c = Case{kind, N, n_mma, a_bytes, b_bytes, a_off0, a_step, b_off0, b_step, a_lbo, a_sbo, b_lbo, b_sbo, 1, 1};
  p.a_img.assign(a_bytes, 0), p.b_img.assign(b_bytes, 0);
  const int esz = kind == 0 ? 4 : 2;
  // fill the images with random values (every element slot), then read the logical matrices back through the address functions
  auto fill = [&](std::vector<uint8_t>& img, bool full) {
    for (size_t i = 0; i + esz <= img.size(); i += esz) {
      float v = full ? rnd_full() : rnd_small();
      if (kind == 0) memcpy(&img[i], &v, 4);
      else {
        __nv_bfloat16 h = __float2bfloat16(v);
        memcpy(&img[i], &h, 2);
      }
    }
  };
  fill(p.a_img, full_mantissa);
  fill(p.b_img, false);
  auto rd = [&](const std::vector<uint8_t>& img, uint32_t off) -> float {
    if (off + esz > img.size()) {
      printf("[%s] address function out of range (%u)\n", name, off);
      exit(3);
    }
    if (kind == 0) {
      float v;
      memcpy(&v, &img[off], 4);
      return v;
    }
    __nv_bfloat16 h;
    memcpy(&h, &img[off], 2);
    return __bfloat162float(h);
  };
  const int K = kind == 0 ? 8 : 16;
  p.expect.assign(128 * N, 0.0), p.expect_alt.assign(128 * N, 0.0);
  for (int j = 0; j < n_mma; ++j)
    for (int r = 0; r < 128; ++r)
      for (int n = 0; n < N; ++n) {
        double s = 0, s2 = 0;
        for (int k = 0; k < K; ++k) {
          const float a = rd(p.a_img, addr_a(j, r, k)), b = rd(p.b_img, addr_b(j, n, k));
          if (kind == 0) s += (double)tf32_trunc(a) * tf32_trunc(b), s2 += (double)tf32_rn(a) * tf32_rn(b);
          else s += (double)a * b, s2 += (double)a * b;
        }
        p.expect[r * N + n] += s, p.expect_alt[r * N + n] += s2;
      }
  return p;
}

static bool run(Problem& p, int repeat = 1, double* cyc_out = nullptr, int n_acc = 1) {
  uint8_t *a_d, *b_d;
  float* d_d;
  long long* cyc_d;
  CK(cudaMalloc(&a_d, p.a_img.size() + 16));
  CK(cudaMalloc(&b_d, p.b_img.size() + 16));
  CK(cudaMalloc(&d_d, sizeof(float) * 128 * p.c.N));
  CK(cudaMalloc(&cyc_d, 8));
  CK(cudaMemcpy(a_d, p.a_img.data(), p.a_img.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(b_d, p.b_img.data(), p.b_img.size(), cudaMemcpyHostToDevice));
  Case c = p.c;
  c.repeat = repeat;
  c.n_acc = n_acc;
  const size_t smem = ((c.a_bytes + 1023) / 1024) * 1024 + ((c.b_bytes + 1023) / 1024) * 1024 + 1024;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  probe_kernel<<<1, 128, smem>>>(c, a_d, b_d, d_d, cyc_d);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  std::vector<float> d(128 * c.N);
  long long cyc = 0;
  CK(cudaMemcpy(d.data(), d_d, sizeof(float) * d.size(), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&cyc, cyc_d, 8, cudaMemcpyDeviceToHost));
  cudaFree(a_d), cudaFree(b_d), cudaFree(d_d), cudaFree(cyc_d);
  if (cyc_out) *cyc_out = (double)cyc;
  if (repeat > 1) return true;
  double e1 = 0, e2 = 0, mx = 0;
  for (size_t i = 0; i < d.size(); ++i) {
    e1 = fmax(e1, fabs(d[i] - p.expect[i])), e2 = fmax(e2, fabs(d[i] - p.expect_alt[i]));
    mx = fmax(mx, fabs(p.expect[i]));
  }
  const bool ok = e1 <= 1e-4 * fmax(1.0, mx) || e2 <= 1e-4 * fmax(1.0, mx);
  printf("%-58s %s  max|D|=%8.3f  err(trunc)=%.3e  err(rn)=%.3e  cycles=%lld\n", p.name, ok ? "PASS" : "FAIL", mx, e1, e2, cyc);
  return ok;
}

int main() {
  bool all = true;
  const uint32_t R = 192;  // rows available in the A image (so that shifted starts stay inside)
  // 1. tf32 K=8 N=16: chunk(r, c) = c*LBO + r*16, element k%4 inside the chunk
  {
    auto aa = [&](int j, int r, int k) { return (uint32_t)((k / 4) * (R * 16) + r * 16 + (k % 4) * 4); };
    auto ab = [&](int j, int n, int k) { return (uint32_t)((k / 4) * (16 * 16) + n * 16 + (k % 4) * 4); };
    Problem p = build("tf32 K8 N16 plain (LBO=R*16, SBO=128)", 0, 16, 1, 2 * R * 16, 2 * 16 * 16, R * 16, 128, 16 * 16, 128, 0, 0, 0, 0, false, aa, ab);
    all &= run(p);
  }
  // 2. tf32, start address shifted by 5 rows (80 B)
  {
    auto aa = [&](int j, int r, int k) { return (uint32_t)((k / 4) * (R * 16) + (r + 5) * 16 + (k % 4) * 4); };
    auto ab = [&](int j, int n, int k) { return (uint32_t)((k / 4) * (16 * 16) + n * 16 + (k % 4) * 4); };
    Problem p = build("tf32 K8 N16 A start +5 rows", 0, 16, 1, 2 * R * 16, 2 * 16 * 16, R * 16, 128, 16 * 16, 128, 5 * 16, 0, 0, 0, false, aa, ab);
    all &= run(p);
  }
  // 3. tf32, 7 accumulated MMAs: A start advances by 3 rows, B advances to the next weight block
  {
    const uint32_t bblk = 2 * 16 * 16;
    auto aa = [&](int j, int r, int k) { return (uint32_t)((k / 4) * (R * 16) + (r + 3 * j) * 16 + (k % 4) * 4); };
    auto ab = [&](int j, int n, int k) { return (uint32_t)(j * bblk + (k / 4) * (16 * 16) + n * 16 + (k % 4) * 4); };
    Problem p = build("tf32 K8 N16 x7 accumulate, A shift 3 rows/step", 0, 16, 7, 2 * R * 16, 7 * bblk, R * 16, 128, 16 * 16, 128, 0, 3 * 16, 0, bblk, false, aa, ab);
    all &= run(p);
  }
  // 4. tf32 full-mantissa A: does the tensor core truncate or round the low 13 bits?
  {
    auto aa = [&](int j, int r, int k) { return (uint32_t)((k / 4) * (R * 16) + r * 16 + (k % 4) * 4); };
    auto ab = [&](int j, int n, int k) { return (uint32_t)((k / 4) * (16 * 16) + n * 16 + (k % 4) * 4); };
    Problem p = build("tf32 K8 N16 full-mantissa A (trunc vs rn)", 0, 16, 1, 2 * R * 16, 2 * 16 * 16, R * 16, 128, 16 * 16, 128, 0, 0, 0, 0, true, aa, ab);
    all &= run(p);
  }
  // 5. tf32 N=80 (filter-bank shape), 4 k-steps
  {
    const uint32_t NB = 80, KC = 8;  // 8 chunks of 4 = K 32 in the image, MMA j uses chunks 2j, 2j+1
    auto aa = [&](int j, int r, int k) { return (uint32_t)((2 * j + k / 4) * (128 * 16) + r * 16 + (k % 4) * 4); };
    auto ab = [&](int j, int n, int k) { return (uint32_t)((2 * j + k / 4) * (NB * 16) + n * 16 + (k % 4) * 4); };
    Problem p = build("tf32 K8 N80 x4 k-steps", 0, NB, 4, KC * 128 * 16, KC * NB * 16, 128 * 16, 128, NB * 16, 128, 0, 2 * 128 * 16, 0, 2 * NB * 16, false, aa, ab);
    all &= run(p);
  }
  // 6. bf16 K=16 N=32, LBO=16 on A: chunk 1 of row r is chunk 0 of row r+1 (two time taps per MMA)
  {
    auto aa = [&](int j, int r, int k) { return (uint32_t)((r + k / 8) * 16 + (k % 8) * 2); };
    auto ab = [&](int j, int n, int k) { return (uint32_t)((k / 8) * (32 * 16) + n * 16 + (k % 8) * 2); };
    Problem p = build("bf16 K16 N32, A LBO=16 (row-aliased chunks)", 1, 32, 1, R * 16, 2 * 32 * 16, 16, 128, 32 * 16, 128, 0, 0, 0, 0, false, aa, ab);
    all &= run(p);
  }
  // 7. bf16 K=16 N=32: 21 accumulated MMAs, A start moves by arbitrary rows, B by blocks
  {
    const uint32_t bblk = 2 * 32 * 16;
    auto aa = [&](int j, int r, int k) { return (uint32_t)((r + 2 * j + 1 + k / 8) * 16 + (k % 8) * 2); };
    auto ab = [&](int j, int n, int k) { return (uint32_t)(j * bblk + (k / 8) * (32 * 16) + n * 16 + (k % 8) * 2); };
    Problem p = build("bf16 K16 N32 x21 accumulate, row-shifted A", 1, 32, 21, R * 16, 21 * bblk, 16, 128, 32 * 16, 128, 16, 2 * 16, 0, bblk, false, aa, ab);
    all &= run(p);
    // timing: cycles per MMA when 21*40 MMAs are issued back to back by one thread
    for (int nacc : {1, 2, 4, 8}) {
      double cyc = 0;
      run(p, 40, &cyc, nacc);
      printf("   timing bf16 M128 N32 K16, %d accumulators: %d MMAs in %.0f cycles -> %.1f cycles/MMA (floor 16)\n", nacc, 21 * 40, cyc, cyc / (21.0 * 40));
    }
  }
  // 8. bf16 K=16 N=16 (equivariant conv shape) timing + correctness with a plain LBO
  {
    auto aa = [&](int j, int r, int k) { return (uint32_t)((k / 8) * (R * 16) + (r + j) * 16 + (k % 8) * 2); };
    auto ab = [&](int j, int n, int k) { return (uint32_t)(j * 512 + (k / 8) * (16 * 16) + n * 16 + (k % 8) * 2); };
    Problem p = build("bf16 K16 N16 x12 accumulate, planar chunks", 1, 16, 12, 2 * R * 16, 12 * 512, R * 16, 128, 16 * 16, 128, 0, 16, 0, 512, false, aa, ab);
    all &= run(p);
    for (int nacc : {1, 4, 8, 16}) {
      double cyc = 0;
      run(p, 80, &cyc, nacc);
      printf("   timing bf16 M128 N16 K16, %d accumulators: %d MMAs in %.0f cycles -> %.1f cycles/MMA (floor 8)\n", nacc, 12 * 80, cyc, cyc / (12.0 * 80));
    }
  }
  // 9. tf32 N=80 timing
  {
    const uint32_t NB = 80, KC = 8;
    auto aa = [&](int j, int r, int k) { return (uint32_t)((2 * j + k / 4) * (128 * 16) + r * 16 + (k % 4) * 4); };
    auto ab = [&](int j, int n, int k) { return (uint32_t)((2 * j + k / 4) * (NB * 16) + n * 16 + (k % 4) * 4); };
    Problem p = build("tf32 K8 N80 x4 (timing)", 0, NB, 4, KC * 128 * 16, KC * NB * 16, 128 * 16, 128, NB * 16, 128, 0, 2 * 128 * 16, 0, 2 * NB * 16, false, aa, ab);
    for (int nacc : {1, 2, 3}) {
      double cyc = 0;
      run(p, 100, &cyc, nacc);
      printf("   timing tf32 M128 N80 K8, %d accumulators: %d MMAs in %.0f cycles -> %.1f cycles/MMA (floor 40)\n", nacc, 400, cyc, cyc / 400.0);
    }
  }
  // 10. calibration: bf16 N=256 (the shape cuBLAS-class GEMMs use; floor 128 cycles)
  {
    const uint32_t NB = 256;
    auto aa = [&](int j, int r, int k) { return (uint32_t)((2 * j + k / 8) * (128 * 16) + r * 16 + (k % 8) * 2); };
    auto ab = [&](int j, int n, int k) { return (uint32_t)((2 * j + k / 8) * (NB * 16) + n * 16 + (k % 8) * 2); };
    Problem p = build("bf16 K16 N256 x4", 1, NB, 4, 8 * 128 * 16, 8 * NB * 16, 128 * 16, 128, NB * 16, 128, 0, 2 * 128 * 16, 0, 2 * NB * 16, false, aa, ab);
    all &= run(p);
    for (int nacc : {1, 2}) {
      double cyc = 0;
      run(p, 50, &cyc, nacc);
      printf("   timing bf16 M128 N256 K16, %d accumulators: %d MMAs in %.0f cycles -> %.1f cycles/MMA (floor 128)\n", nacc, 200, cyc, cyc / 200.0);
    }
  }
  printf(all ? "ALL PASS\n" : "SOME FAILED\n");
  return all ? 0 : 1;
}

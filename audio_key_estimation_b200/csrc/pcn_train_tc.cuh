// pcn_train_tc.cuh -- the TRAINING step's convolutions on the tensor cores (BASELINE config 5):
//   * the 7x7 circular convolutions (models.py:228-234; 64.6 % of the network's MACs): forward, data gradient (this comment) and
//     weight gradient (p2p_wgrad_umma_kernel below: one MN-major GEMM over positions);
//   * the equivariant 12x7 convolutions of both PitchClass2PitchClass stacks (models.py:22-51, 191-197): forward and data gradient through
//     raw-output variants of the eval-mode kernels (eq_pack_planes_kernel below feeds them);
//   * the first conv of the tonic and key heads, forward (heads_raw_ss_kernel below builds its epilogue table);
//   * every operand image of the step in one launch (tc_pack_all_weights_kernel).
// Host side: Fwd::tc_conv / tc_wgrad / eq_conv / heads_tc in pcn.cu, wired into the kept forward / backward in pcn_train.cuh.
//
// The eval-mode kernel p2p_umma_kernel<false, RAW = true> (pcn_umma.cuh) already is "7x7 circular conv, fp16 hi/lo three-product
// operands, raw fp32 accumulators out"; the training step keeps planar fp32 activations (B, C, R, T) for its BatchNorm / weight-gradient
// kernels, so the conv is wrapped by two streaming kernels:
//   tc_pack_planes_kernel : planar fp32 [+ a second, row-tiled tensor: cat[mel, tile(up)]] -> hi / lo chunk planes [B][P+6][T+6][8] with
//                           circular halos, times an exact power of two that brings the tensor's max |x| to [2^13, 2^14) (gradients are
//                           ~1e-5: the lo halves would fall into fp16's subnormals otherwise)
//   p2p_umma_kernel<0, 1> : raw accumulators (B, P, T, 8)
//   tc_unpack_kernel      : -> planar z (+ bias, scales divided out) and, for the forward, the BatchNorm batch statistics of z in the
//                           same pass (replaces bn_stats_kernel's extra read)
// The data gradient is the same conv with the weights transposed and both taps flipped (tc_pack_all_weights_kernel, flip = 1):
//   dX[ci, p, t] = sum_{co, dp, dt} W[co, ci, 6 - dp, 6 - dt] dZ[co, p + dp - 3, t + dt - 3]   (indices circular).
#pragma once
#include "pcn_umma.cuh"

namespace ake {

struct TcPackArgs {
  const float* in0;   // (B, c0, P, T)
  const float* in1;   // (B, c1, rows1, T), row p of the conv input = row p % rows1 (PitchClass2Pitch tiling), or unused (c1 = 0)
  long long bs0, bs1;
  int c0, c1, rows1;
  int B, P, T, Wd;
  int col_shift;      // plane column of frame 0: 3 (circular halo columns, the conv operand) or 0
  int wrap_cols;      // 1: columns outside [0, T) hold the circular copies; 0: they hold zeros (the weight-gradient operand)
  const unsigned* maxbits;  // NULL: scale 1
  __half* hi;
  __half* lo;
};

__global__ void __launch_bounds__(256) tc_pack_planes_kernel(const TcPackArgs a) {
  const float mul = a.maxbits ? tc_scale_of(__ldg(a.maxbits)) : 1.f;
  const long long n = (long long)a.B * (a.P + 6) * a.Wd;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int col = (int)(i % a.Wd);
    const long long q = i / a.Wd;
    const int row = (int)(q % (a.P + 6)), b = (int)(q / (a.P + 6));
    int p = (row - 3) % a.P, t = col - a.col_shift;
    const bool live = a.wrap_cols || (t >= 0 && t < a.T);
    t %= a.T;
    p += p < 0 ? a.P : 0, t += t < 0 ? a.T : 0;
    float v[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      float x = 0.f;
      if (!live) x = 0.f;
      else if (c < a.c0) x = __ldg(a.in0 + b * a.bs0 + ((long long)c * a.P + p) * a.T + t);
      else if (c < a.c0 + a.c1) x = __ldg(a.in1 + b * a.bs1 + ((long long)(c - a.c0) * a.rows1 + p % a.rows1) * a.T + t);
      v[c] = x * mul;
    }
    store_split8(a.hi + i * 8, a.lo + i * 8, v);
  }
}

struct TcUnpackArgs {
  const float* raw;   // (B, P, T, 8) accumulators (weights carried kWScale, the input the scale of `maxbits`)
  float* out;         // (B, C, P, T)
  const float* bias;  // C or NULL
  const unsigned* maxbits;
  int B, C, P, T;
  double* stats;      // NULL, or stats[2 c] += sum, stats[2 c + 1] += sum of squares over (B, P, T) (bn_stats_kernel's contract)
};

__global__ void __launch_bounds__(256) tc_unpack_kernel(const TcUnpackArgs a) {
  const float mul = (1.f / kWScale) / (a.maxbits ? tc_scale_of(__ldg(a.maxbits)) : 1.f);  // exact: powers of two
  const long long PT = (long long)a.P * a.T, n = (long long)a.B * PT;
  float bias[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) bias[c] = (a.bias && c < a.C) ? __ldg(a.bias + c) : 0.f;
  double s[8], ss[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) s[c] = 0.0, ss[c] = 0.0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / PT, e = i - b * PT;
    const float4 r0 = __ldg(reinterpret_cast<const float4*>(a.raw + i * 8)), r1 = __ldg(reinterpret_cast<const float4*>(a.raw + i * 8 + 4));
    const float r[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      if (c < a.C) {
        const float v = fmaf(r[c], mul, bias[c]);
        a.out[(b * a.C + c) * PT + e] = v;
        s[c] += (double)v, ss[c] += (double)v * (double)v;
      }
    }
  }
  if (!a.stats) return;
  __shared__ double sh[8][16];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    for (int o = 16; o; o >>= 1) {
      s[c] += __shfl_xor_sync(0xffffffffu, s[c], o);
      ss[c] += __shfl_xor_sync(0xffffffffu, ss[c], o);
    }
    if (lane == 0) sh[warp][2 * c] = s[c], sh[warp][2 * c + 1] = ss[c];
  }
  __syncthreads();
  if (threadIdx.x < 2 * a.C) {
    double v = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) v += sh[w][threadIdx.x];
    atomicAdd(a.stats + threadIdx.x, v);
  }
}

// ---- weight gradient of the 7x7 circular conv as ONE tensor-core GEMM over positions ---------------------------------------------------
//   dW[co, ci, dp, dt] = sum_{b, r, t} dZ[b, co, r, t] * x~[b, ci, r + dp, t + dt]        (x~ = the halo'd input, circular)
// Both tensors are position-major chunk planes ([position][8 channels] fp16), i.e. MN-MAJOR tcgen05 operands whose K index is the
// position -- no transposition pass.  With s = r + dp - 3 (the un-halo'd input row, taken modulo P) the sum over (r, dp) becomes a sum
// over s in [0, P) and the seven gradient rows s - 3 ... s + 3 (halo rows above / below are the circular copies):
//   A (M = 64): rows (g = dt, ci), K = 16 consecutive frames of input row s: x~[s + 3][t0 + k + g]  -- eight core matrices 16 B apart
//               (OVERLAPPING: SBO = 16 B; verified by tools/mn_probe.cu), the eighth (g = 7) is discarded
//   B (N = 56): columns (j = 6 - dp, co): dZ plane row s + j, frames t0 + k                           -- SBO = the plane's row pitch
//   D[(dt, ci), (6 - dp, co)] += A . B, three hi/lo products per 16 frames, accumulated in TMEM over every tile of the CTA.
// The gradient planes carry zeros beyond frame T - 1 (tc_pack_planes_kernel, wrap_cols = 0), so partial 16-frame blocks add nothing.
// Persistent CTA per SM: warps 0-3 drain the accumulator once at the end (M = 64: rows 16 q ... 16 q + 15 in lanes 0-15 of TMEM quadrant
// q) into a per-CTA partial; warp 4 loads the tiles (R input rows + R + 6 gradient rows, two buffers), warp 5 issues the MMAs.
// wgrad_tc_reduce_kernel sums the partials in CTA order (deterministic, no atomics) and writes dW.
constexpr int kWgR = 6;                 // input rows per tile
constexpr int kWgThreads = 32 * 6;
struct WgradTcArgs {
  const __half* x_hi;
  const __half* x_lo;   // [B][P+6][Wd][8]: the conv's input operand planes (circular halos, scale 1)
  const __half* g_hi;
  const __half* g_lo;   // [B][P+6][Wg][8]: output gradient, halo ROWS circular, column = frame, zeros beyond T, scaled by tc_scale_of(maxbits)
  float* partial;       // [gridDim.x][64][56]
  int B, P, T, Wd, Wg;
  int n_rtiles, n_tiles;
};
__host__ __device__ inline uint32_t wgrad_tc_x_bytes(int Wd) { return (uint32_t)(kWgR * Wd + 32) * 16; }      // + slack: the last block reads past the row
__host__ __device__ inline uint32_t wgrad_tc_g_bytes(int Wg) { return (uint32_t)((kWgR + 6) * Wg + 16) * 16; }
__host__ __device__ inline size_t wgrad_tc_smem_bytes(int Wd, int Wg) { return (size_t)2 * 2 * (wgrad_tc_x_bytes(Wd) + wgrad_tc_g_bytes(Wg)); }

__global__ void __launch_bounds__(kWgThreads, 1) p2p_wgrad_umma_kernel(const WgradTcArgs a) {
  using namespace umma;
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t full_bar[2], empty_bar[2], done_bar;
  __shared__ uint32_t tmem_slot;
  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const uint32_t XB = wgrad_tc_x_bytes(a.Wd), GB = wgrad_tc_g_bytes(a.Wg), BUF = 2 * (XB + GB);  // buffer: [x_hi][x_lo][g_hi][g_lo]
  if (warp == 5) tmem_alloc(&tmem_slot, 64);
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) mbar_init(&full_bar[i], 1), mbar_init(&empty_bar[i], 1);
    mbar_init(&done_bar, 1);
    mbar_init_fence();
  }
  // the slack behind the rows is read (by blocks whose gradient frames are zero) but never written by the copies: keep it finite
  for (uint32_t i = threadIdx.x; i < 2 * BUF / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_slot;
  const int n_blk = (a.T + 15) / 16;

  if (warp == 4) {
    // ------------------------------------------------------------ loader: rows are contiguous in the planes, one copy per plane
    int k = 0;
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++k) {
      const int s = k & 1;
      const int b = tile / a.n_rtiles, s0 = (tile - b * a.n_rtiles) * kWgR, R = min(kWgR, a.P - s0);
      mbar_wait_relaxed(&empty_bar[s], ((k >> 1) & 1) ^ 1);
      if (lane == 0) {
        const uint32_t xb = (uint32_t)(R * a.Wd) * 16, gb = (uint32_t)((R + 6) * a.Wg) * 16;
        mbar_arrive_expect_tx(&full_bar[s], 2 * (xb + gb));
        uint8_t* dst = smem + (size_t)s * BUF;
        const long long xo = (((long long)b * (a.P + 6) + s0 + 3) * a.Wd) * 8, go = (((long long)b * (a.P + 6) + s0) * a.Wg) * 8;
        bulk_g2s(dst, a.x_hi + xo, xb, &full_bar[s]);
        bulk_g2s(dst + XB, a.x_lo + xo, xb, &full_bar[s]);
        bulk_g2s(dst + 2 * XB, a.g_hi + go, gb, &full_bar[s]);
        bulk_g2s(dst + 2 * XB + GB, a.g_lo + go, gb, &full_bar[s]);
      }
      __syncwarp();
    }
  } else if (warp == 5) {
    // ------------------------------------------------------------ MMA issuer (converged warp, one elected lane issues)
    constexpr uint64_t A_DESC = desc_hi(128, 16);                      // MN-major: K groups of 8 frames 128 B apart, M groups (time taps) 16 B apart
    const uint64_t B_DESC = desc_hi(128, (uint32_t)a.Wg * 16);          // N groups (row taps) one plane row apart
    constexpr uint32_t IDESC = idesc_f16(56, 64) | (1u << 15) | (1u << 16);  // A and B MN-major
    int k = 0;
    uint32_t first = 0;  // the CTA's very first MMA overwrites the accumulator
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++k) {
      const int s = k & 1;
      const int b = tile / a.n_rtiles, s0 = (tile - b * a.n_rtiles) * kWgR, R = min(kWgR, a.P - s0);
      const uint32_t x0 = smem_u32(smem + (size_t)s * BUF), g0 = x0 + 2 * XB;
      mbar_wait(&full_bar[s], (k >> 1) & 1);
      fence_after_sync();
      if (elect_one()) {
        for (int i = 0; i < R; ++i) {
          for (int c = 0; c < n_blk; ++c) {
            const uint32_t xa = x0 + (uint32_t)(i * a.Wd + 16 * c) * 16, ga = g0 + (uint32_t)(i * a.Wg + 16 * c) * 16;
            mma_f16(tmem, make_desc(A_DESC, xa), make_desc(B_DESC, ga), IDESC, first);            // x_hi . g_hi
            first = 1u;
            mma_f16(tmem, make_desc(A_DESC, xa + XB), make_desc(B_DESC, ga), IDESC, 1u);          // x_lo . g_hi
            mma_f16(tmem, make_desc(A_DESC, xa), make_desc(B_DESC, ga + GB), IDESC, 1u);          // x_hi . g_lo
          }
        }
        commit(&empty_bar[s]);
      }
      __syncwarp();
    }
    if (elect_one()) commit(&done_bar);
    __syncwarp();
  } else {
    // ------------------------------------------------------------ final drain: partial[cta][m][n], m = 8 dt + ci, n = 8 (6 - dp) + co
    mbar_wait_relaxed(&done_bar, 0);
    fence_after_sync();
    float* dst = a.partial + (size_t)blockIdx.x * 64 * 56;
    for (int c0 = 0; c0 < 56; c0 += 8) {
      float v[8];
      tmem_ld8(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
      if (lane < 16) {
        float4* d4 = reinterpret_cast<float4*>(dst + (16 * warp + lane) * 56 + c0);
        d4[0] = make_float4(v[0], v[1], v[2], v[3]), d4[1] = make_float4(v[4], v[5], v[6], v[7]);
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 5) tmem_dealloc(tmem, 64);
}

// dW[co][ci][dp][dt] = (sum over the CTAs' partials, in CTA order) / the gradient planes' scale.  One thread per weight.
__global__ void wgrad_tc_reduce_kernel(const float* __restrict__ partial, int n_cta, const unsigned* __restrict__ maxbits, int Cout, int Cin,
                                       float* __restrict__ dw) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Cout * Cin * 49) return;
  const int dt = i % 7, dp = (i / 7) % 7, ci = (i / 49) % Cin, co = i / (49 * Cin);
  const int m = 8 * dt + ci, n = 8 * (6 - dp) + co;
  float s = 0.f;
  for (int c = 0; c < n_cta; ++c) s += __ldg(partial + ((size_t)c * 64 + m) * 56 + n);
  dw[i] = s / (maxbits ? tc_scale_of(__ldg(maxbits)) : 1.f);
}

// ---- equivariant 12 x 7 convolutions of the PitchClass2PitchClass stacks in train mode ------------------------------------------------
// pc2pc_umma_kernel<3> (<= 16 -> <= 16 channels) and pc8_umma_kernel<2> with raw = 1 (<= 8 -> <= 8) write raw planar fp32; their operand
// planes [B][G][23][T + 6][8] (zero halo columns = the "same" padding in time, rows 12..22 = rows 0..10) come from the planar activations:
struct EqPackArgs {
  const float* in;  // (B, C, 12, T); the planes hold its channels [c0, c0 + 8 G)
  int B, C, G, T, Wd;
  int col_shift;            // plane column of frame 0: 3 ("same" convs), 0 (valid convs: Wd = T) or 6 (the data gradient of a valid conv)
  int c0;
  int row_shift;            // plane row i holds pitch class (i + row_shift) % 12: 0 (operand of the conv), 1 (gradient operand of eq_wgrad_umma_kernel)
  const unsigned* maxbits;  // NULL: scale 1
  __half* hi;
  __half* lo;
};
__global__ void __launch_bounds__(256) eq_pack_planes_kernel(const EqPackArgs a) {
  const float mul = a.maxbits ? tc_scale_of(__ldg(a.maxbits)) : 1.f;
  const long long n = (long long)a.B * a.G * 23 * a.Wd;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int col = (int)(i % a.Wd);
    long long q = i / a.Wd;
    const int row = (int)(q % 23);
    q /= 23;
    const int g = (int)(q % a.G), b = (int)(q / a.G);
    const int c = (row + a.row_shift) % 12, t = col - a.col_shift;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int ch = a.c0 + g * 8 + e;
      v[e] = (ch < a.C && t >= 0 && t < a.T) ? __ldg(a.in + (((long long)b * a.C + ch) * 12 + c) * a.T + t) * mul : 0.f;
    }
    store_split8(a.hi + i * 8, a.lo + i * 8, v);
  }
}
// ---- weight gradient of an equivariant 12 x 7 conv as a tensor-core GEMM over positions -------------------------------------------------------
//   dW[co, ci, dp, dt] = sum_{b, c, t} dZ[b, co, c, t] * x~[b, ci, (c + dp) mod 12, t + dt]
// (x~: the conv's operand planes -- three zero halo columns for the "same" convs of the stacks, none for the heads' valid convs)
// The scheme of p2p_wgrad_umma_kernel with twelve row taps: for input row rho the gradient rows (rho - dp) mod 12, dp = 11 ... 0, are the
// CONSECUTIVE rows rho ... rho + 11 of a 23-row gradient plane whose row i holds pitch class (i + 1) mod 12 (eq_pack_planes_kernel, row_shift 1):
//   A (M = 64): rows (dt, ci) of one channel group, K = 16 frames of x~ row rho (time taps 16 B apart: overlapping core matrices)
//   B (N = 96): columns (j = 11 - dp, co) of one gradient channel group, rows rho + j (SBO = the plane's row pitch)
// A work item is (clip, input channel group, output channel group); a CTA keeps its (group, group) pair over all its items (the grid is a
// multiple of the number of pairs), so its accumulator stays in TMEM; eq_wgrad_reduce_kernel sums the partials of each pair in CTA order.
constexpr int kEqWgThreads = 32 * 6;
struct EqWgradArgs {
  const __half* x_hi;
  const __half* x_lo;   // [B][2][23][Wx][8]
  const __half* g_hi;
  const __half* g_lo;   // [B][2][23][Wg][8]: zeros beyond T, scaled by tc_scale_of(maxbits)
  float* partial;       // [gridDim.x][64][96]
  int B, T, Wx, Wg;
  int n_gi, n_go;       // channel groups of the input / of the gradient (planes per clip); gridDim.x is a multiple of n_gi * n_go
};
__host__ __device__ inline uint32_t eq_wgrad_x_bytes(int Wx) { return (uint32_t)(12 * Wx + 32) * 16; }
__host__ __device__ inline uint32_t eq_wgrad_g_bytes(int Wg) { return (uint32_t)(23 * Wg + 16) * 16; }
__host__ __device__ inline size_t eq_wgrad_smem_bytes(int Wx, int Wg) { return (size_t)2 * (eq_wgrad_x_bytes(Wx) + eq_wgrad_g_bytes(Wg)); }

__global__ void __launch_bounds__(kEqWgThreads, 1) eq_wgrad_umma_kernel(const EqWgradArgs a) {
  using namespace umma;
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t full_bar, empty_bar, done_bar;
  __shared__ uint32_t tmem_slot;
  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const uint32_t XB = eq_wgrad_x_bytes(a.Wx), GB = eq_wgrad_g_bytes(a.Wg);  // [x_hi][x_lo][g_hi][g_lo]
  if (warp == 5) tmem_alloc(&tmem_slot, 128);
  if (threadIdx.x == 0) {
    mbar_init(&full_bar, 1), mbar_init(&empty_bar, 1), mbar_init(&done_bar, 1);
    mbar_init_fence();
  }
  for (uint32_t i = threadIdx.x; i < 2 * (XB + GB) / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_slot;
  const int n_pairs = a.n_gi * a.n_go, n_items = n_pairs * a.B, n_blk = (a.T + 15) / 16;
  const int pair = (int)blockIdx.x % n_pairs, gi = pair / a.n_go, go = pair - gi * a.n_go;  // fixed per CTA: gridDim.x is a multiple of n_pairs

  if (warp == 4) {
    // ------------------------------------------------------------ loader: 12 input rows and 23 gradient rows, contiguous in their planes
    int k = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++k) {
      const int b = item / n_pairs;
      mbar_wait_relaxed(&empty_bar, (k & 1) ^ 1);
      if (lane == 0) {
        const uint32_t xb = (uint32_t)(12 * a.Wx) * 16, gb = (uint32_t)(23 * a.Wg) * 16;
        mbar_arrive_expect_tx(&full_bar, 2 * (xb + gb));
        const long long xo = (((long long)b * a.n_gi + gi) * 23) * a.Wx * 8, gof = (((long long)b * a.n_go + go) * 23) * a.Wg * 8;
        bulk_g2s(smem, a.x_hi + xo, xb, &full_bar);
        bulk_g2s(smem + XB, a.x_lo + xo, xb, &full_bar);
        bulk_g2s(smem + 2 * XB, a.g_hi + gof, gb, &full_bar);
        bulk_g2s(smem + 2 * XB + GB, a.g_lo + gof, gb, &full_bar);
      }
      __syncwarp();
    }
  } else if (warp == 5) {
    // ------------------------------------------------------------ MMA issuer (converged warp, one elected lane issues)
    constexpr uint64_t A_DESC = desc_hi(128, 16);
    const uint64_t B_DESC = desc_hi(128, (uint32_t)a.Wg * 16);
    constexpr uint32_t IDESC = idesc_f16(96, 64) | (1u << 15) | (1u << 16);  // A and B MN-major
    const uint32_t x0 = smem_u32(smem), g0 = x0 + 2 * XB;
    int k = 0;
    uint32_t first = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++k) {
      mbar_wait(&full_bar, k & 1);
      fence_after_sync();
      if (elect_one()) {
        for (int rho = 0; rho < 12; ++rho) {
          for (int c = 0; c < n_blk; ++c) {
            const uint32_t xa = x0 + (uint32_t)(rho * a.Wx + 16 * c) * 16, ga = g0 + (uint32_t)(rho * a.Wg + 16 * c) * 16;
            mma_f16(tmem, make_desc(A_DESC, xa), make_desc(B_DESC, ga), IDESC, first);
            first = 1u;
            mma_f16(tmem, make_desc(A_DESC, xa + XB), make_desc(B_DESC, ga), IDESC, 1u);
            mma_f16(tmem, make_desc(A_DESC, xa), make_desc(B_DESC, ga + GB), IDESC, 1u);
          }
        }
        commit(&empty_bar);
      }
      __syncwarp();
    }
    if (elect_one()) commit(&done_bar);
    __syncwarp();
  } else {
    // ------------------------------------------------------------ final drain (M = 64: rows 16 q ... 16 q + 15 in lanes 0-15 of quadrant q)
    mbar_wait_relaxed(&done_bar, 0);
    fence_after_sync();
    float* dst = a.partial + (size_t)blockIdx.x * 64 * 96;
    for (int c0 = 0; c0 < 96; c0 += 8) {
      float v[8];
      tmem_ld8(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
      if (lane < 16) {
        float4* d4 = reinterpret_cast<float4*>(dst + (16 * warp + lane) * 96 + c0);
        d4[0] = make_float4(v[0], v[1], v[2], v[3]), d4[1] = make_float4(v[4], v[5], v[6], v[7]);
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 5) tmem_dealloc(tmem, 128);
}

// dW[co][ci][dp][dt] = (sum over the partials of the CTAs that own (ci / 8, co / 8), in CTA order) / the gradient planes' scale
__global__ void eq_wgrad_reduce_kernel(const float* __restrict__ partial, int n_cta, const unsigned* __restrict__ maxbits, int Cout, int Cin,
                                       int n_gi, int n_go, float* __restrict__ dw) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Cout * Cin * 84) return;
  const int dt = i % 7, dp = (i / 7) % 12, ci = (i / 84) % Cin, co = i / (84 * Cin);
  const int pair = (ci >> 3) * n_go + (co >> 3), m = 8 * dt + (ci & 7), n = 8 * (11 - dp) + (co & 7);
  float s = 0.f;
  for (int c = pair; c < n_cta; c += n_gi * n_go) s += __ldg(partial + ((size_t)c * 64 + m) * 96 + n);
  dw[i] = s / (maxbits ? tc_scale_of(__ldg(maxbits)) : 1.f);
}

// epilogue table of the heads' fused first conv in train mode: scale 1, shift = [tonic bias | key bias]
__global__ void heads_raw_ss_kernel(const float* __restrict__ bias_t, const float* __restrict__ bias_k, float* __restrict__ ss) {
  const int i = threadIdx.x;
  if (i < 64) ss[i] = 1.f, ss[64 + i] = i < 32 ? (bias_t ? bias_t[i] : 0.f) : (bias_k ? bias_k[i - 32] : 0.f);
}

// ---- every operand image of a training step's tensor-core convolutions (forward + tap-flipped data-gradient images of the 7x7, the
// <= 8-channel and the 16-channel equivariant convs) in ONE launch: the table travels as a kernel argument, blockIdx.y = entry.
struct TcWeightEntry {
  long long w_off;   // floats into the flat parameters
  long long img_off; // halves into the step's image block
  int Cout, Cin;     // of the convolution as stored (Cout, Cin, KH, 7)
  int kind;          // 0: 7x7 (p2p_umma_kernel), 1: <= 8 channels equivariant (pc8_umma_kernel), 2: 16 channels equivariant (pc2pc_umma_kernel)
  int flip;          // 1: the data-gradient image
  int ci0;           // kind 2, flip: first gradient channel of this image (a data gradient over > 16 channels is summed over several images)
};
constexpr int kTcWeightMax = 24;
struct TcWeightTable {
  int n;
  TcWeightEntry e[kTcWeightMax];
};
__global__ void __launch_bounds__(256) tc_pack_all_weights_kernel(const TcWeightTable t, const float* __restrict__ params, __half* __restrict__ images) {
  const TcWeightEntry& en = t.e[blockIdx.y];
  const float* w = params + en.w_off;
  __half* img = images + en.img_off;
  const int Cout = en.Cout, Cin = en.Cin;
  if (en.kind == 0) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 7 * 56 * 8; i += gridDim.x * blockDim.x) {
      const int ci = i % 8, fc = (i / 8) % 56, dp = i / 448, f = fc / 8, co = fc % 8;
      float v = 0.f;
      if (!en.flip) {
        if (ci < Cin && co < Cout) v = w[(((long long)co * Cin + ci) * 7 + dp) * 7 + f] * kWScale;
      } else if (ci < Cout && co < Cin) {
        v = w[(((long long)ci * Cin + co) * 7 + (6 - dp)) * 7 + (6 - f)] * kWScale;
      }
      p2p_img_store(img, dp, f, co, ci, v);
    }
  } else if (en.kind == 1) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 12 * 56 * 8; i += gridDim.x * blockDim.x) {
      const int ci = i % 8, fc = (i / 8) % 56, dp = i / 448, f = fc / 8, co = fc % 8;
      float v = 0.f;
      if (!en.flip) {
        if (ci < Cin && co < Cout) v = w[(((long long)co * Cin + ci) * 12 + dp) * 7 + f] * kWScale;
      } else if (ci < Cout && co < Cin) {
        v = w[(((long long)ci * Cin + co) * 12 + (12 - dp) % 12) * 7 + (6 - f)] * kWScale;
      }
      const __half hi = __float2half_rn(v);
      const __half lo = __float2half_rn(v - __half2float(hi));
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        img[((dp * 2 + c) * 112 + 16 * f + co) * 8 + ci] = hi;
        img[((dp * 2 + c) * 112 + 16 * f + 8 + co) * 8 + ci] = lo;
      }
    }
  } else {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 12 * 2 * 112 * 8; i += gridDim.x * blockDim.x) {
      const int e = i % 8, n = (i / 8) % 112, g = (i / (8 * 112)) % 2, dp = i / (16 * 112);
      const int f = n / 16, co = n % 16, ci = g * 8 + e;
      float v = 0.f;
      if (!en.flip) {
        if (ci < Cin && co < Cout) v = w[(((long long)co * Cin + ci) * 12 + dp) * 7 + f] * kWScale;
      } else if (en.ci0 + ci < Cout && co < Cin) {
        v = w[(((long long)(en.ci0 + ci) * Cin + co) * 12 + (12 - dp) % 12) * 7 + (6 - f)] * kWScale;
      }
      const __half hi = __float2half_rn(v);
      const __half lo = __float2half_rn(v - __half2float(hi));
      img[((dp * 2 + g) * 224 + n) * 8 + e] = hi;
      img[((dp * 2 + g) * 224 + 112 + n) * 8 + e] = lo;
    }
  }
}

}  // namespace ake

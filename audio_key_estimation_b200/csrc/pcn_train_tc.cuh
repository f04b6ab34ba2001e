// pcn_train_tc.cuh -- the 7x7 circular convolutions of the TRAINING step on the tensor cores (BASELINE config 5; models.py:228-234 in
// train mode and its data gradient).  They are 64.6 % of the network's MACs, forward and backward.
//
// The eval-mode kernel p2p_umma_kernel<false, RAW = true> (pcn_umma.cuh) already is "7x7 circular conv, fp16 hi/lo three-product
// operands, raw fp32 accumulators out"; the training step keeps planar fp32 activations (B, C, R, T) for its BatchNorm / weight-gradient
// kernels, so the conv is wrapped by two streaming kernels:
//   tc_pack_planes_kernel : planar fp32 [+ a second, row-tiled tensor: cat[mel, tile(up)]] -> hi / lo chunk planes [B][P+6][T+6][8] with
//                           circular halos, times an exact power of two that brings the tensor's max |x| to [8, 16) (gradients are
//                           ~1e-5: the lo halves would fall into fp16's subnormals otherwise)
//   p2p_umma_kernel<0, 1> : raw accumulators (B, P, T, 8)
//   tc_unpack_kernel      : -> planar z (+ bias, scales divided out) and, for the forward, the BatchNorm batch statistics of z in the
//                           same pass (replaces bn_stats_kernel's extra read)
// The data gradient is the same conv with the weights transposed and both taps flipped (p2p_pack_weights_flip_kernel):
//   dX[ci, p, t] = sum_{co, dp, dt} W[co, ci, 6 - dp, 6 - dt] dZ[co, p + dp - 3, t + dt - 3]   (indices circular).
#pragma once
#include "pcn_umma.cuh"

namespace ake {

// power of two that brings a tensor whose largest |x| has the bit pattern `maxbits` into [8, 16)
__device__ __forceinline__ float tc_scale_of(unsigned maxbits) {
  const int e = (int)((maxbits >> 23) & 0xffu);
  if (e < 3 || e == 255) return 1.f;  // zero / denormal / non-finite: leave the tensor alone
  return __uint_as_float((unsigned)(257 - e) << 23);  // 2^(3 - (e - 127))
}

struct TcPackArgs {
  const float* in0;   // (B, c0, P, T)
  const float* in1;   // (B, c1, rows1, T), row p of the conv input = row p % rows1 (PitchClass2Pitch tiling), or unused (c1 = 0)
  long long bs0, bs1;
  int c0, c1, rows1;
  int B, P, T, Wd;
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
    int p = (row - 3) % a.P, t = (col - 3) % a.T;
    p += p < 0 ? a.P : 0, t += t < 0 ? a.T : 0;
    float v[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      float x = 0.f;
      if (c < a.c0) x = __ldg(a.in0 + b * a.bs0 + ((long long)c * a.P + p) * a.T + t);
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

// weight image of the data-gradient conv: the conv (Cout' = Cin, Cin' = Cout) with w'[co'][ci'][dp][f] = w[ci'][co'][6 - dp][6 - f]
__global__ void p2p_pack_weights_flip_kernel(const float* __restrict__ w, int Cout, int Cin, __half* __restrict__ img) {
  const int n_items = 7 * 56 * 8;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_items; i += gridDim.x * blockDim.x) {
    const int ci = i % 8, fc = (i / 8) % 56, dp = i / 448;  // ci: channel of dZ (< Cout), co: channel of dX (< Cin)
    const int f = fc / 8, co = fc % 8;
    float v = 0.f;
    if (ci < Cout && co < Cin) v = w[(((long long)ci * Cin + co) * 7 + (6 - dp)) * 7 + (6 - f)] * kWScale;
    p2p_img_store(img, dp, f, co, ci, v);
  }
}

}  // namespace ake

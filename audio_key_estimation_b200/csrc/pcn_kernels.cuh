// pcn_kernels.cuh -- fp32 CUDA-core kernels of the PitchClassNet forward (sm_100a).
//
// Every convolution of the network is an instance of
// one "row convolution":
//   out[co, r, t] = sum_{ci, dr<KH, dt<KW} W[co,ci,dr,dt] * x[ci, rowmap(r*SR + off + dr), tmap(t - pad + dt)]
// with circular or zero row/time maps.  conv_rows_kernel stages an input tile (8 input channels at a
// time) plus the matching weight slice in shared memory and lets every thread own a CO_T x RT register
// tile of outputs (CO_T output channels of RT consecutive frames of one row), so the inner loop is
// pure FFMA fed by 128-bit shared-memory loads (weights are warp-uniform -> broadcast).
//
// Bit-exact transposition equivariance (SURVEY.md hard part 6): every output element accumulates its
// K = Cin*KH*KW products in the same order (ci, dr, dt ascending) with the same instructions no matter
// which row it is; rows only enter through load addresses (mod rows); no atomics are used.
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"

namespace ake {

struct ConvArgs {
  const float* in0;  // channels [0, c0)
  const float* in1;  // channels [c0, c0 + c1) -- rows tiled modulo rows1 (PitchClass2Pitch, models.py:135-143)
  int c0, c1;
  int rows0, rows1;    // physical rows per channel
  long long bs0, bs1;  // batch strides (floats)
  int T_in;
  int rows_v;     // logical input rows
  int row_circ;   // 1: rows wrap, 0: rows outside read as zero
  int row_off;    // input row of (out row 0, tap 0)
  int pad_t;      // input frame of (out frame 0, tap 0) is -pad_t
  int time_circ;  // 1: frames wrap (padding_mode="circular"), 0: zero padding / valid
  int rows_out, T_out;
  int Cin, Cout;
  const float* w;  // packed [Cin][KH][KW][cout_pad]
  int cout_pad;
  const float* scale;  // epilogue y = act(acc * scale[c] + shift[c])
  const float* shift;
  int act;
  float* out;
  long long obs, ocs;  // output batch / channel strides (floats)
  int out_coff;        // channel offset in the output tensor (free torch.cat, models.py:383,392)
  int T_store;         // stored frames per row (T_out, or T_out/2 with pool_t)
  int pool_t;          // fuse MaxPool2d((1,2)) (models.py:349-350)
  int n_row_tiles;
  int tgroups;  // time groups per block: block covers tgroups*RT output frames
  int xp;       // shared-memory row pitch (floats, multiple of 4, >= tgroups*RT + KW - 1)
  int accum;    // add to the stored values instead of overwriting them (data gradients with several consumers)
  double* stats;  // train mode (raw outputs, no pool): stats[2 ch] += sum, stats[2 ch + 1] += sum of squares of the stored values
                  // (bn_stats_kernel's contract, out of the conv's own epilogue); NULL otherwise
};

constexpr int kConvCI = 8;  // input channels staged per shared-memory pass

__device__ __forceinline__ float leaky(float v) { return v > 0.f ? v : kLeakySlope * v; }
// activation codes of the conv / affine epilogues: 0 none, 1 nn.LeakyReLU(), 2 nn.ReLU() (dense layers, models.py:461-466)
__device__ __forceinline__ float apply_act(float v, int act) { return act == 1 ? leaky(v) : (act == 2 ? fmaxf(v, 0.f) : v); }

template <int KH, int KW, int SR, int RB, int CO_T, int RT>
__global__ void __launch_bounds__(256) conv_rows_kernel(const ConvArgs a) {
  constexpr int RIN = (RB - 1) * SR + KH;  // input rows a row tile touches
  constexpr int NX = RT + KW - 1;          // input frames a thread touches per (ci, dr)
  static_assert(NX % 2 == 0 && RT % 2 == 0, "vector loads need even windows");
  extern __shared__ float4 smem4[];
  float* xs = reinterpret_cast<float*>(smem4);  // [kConvCI][RIN][XP]
  const int XP = a.xp;
  float* wsm = xs + kConvCI * RIN * XP;         // [kConvCI][KH][KW][CO_T]

  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const int TB = a.tgroups * RT;
  const int TBW = TB + KW - 1;
  const int row_tile = blockIdx.y % a.n_row_tiles;
  const int cob = blockIdx.y / a.n_row_tiles;
  const int b = blockIdx.z;
  const int t0 = blockIdx.x * TB;
  const int out_row0 = row_tile * RB;
  const bool active = tid < RB * a.tgroups;
  const int r = tid % RB, tg = tid / RB;

  // frame index of each shared-memory column this lane fills (-1 = zero)
  int tm[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    int j = lane + 32 * k;
    int t = t0 - a.pad_t + j;
    if (a.time_circ) {
      t %= a.T_in;
      if (t < 0) t += a.T_in;
    } else if (t < 0 || t >= a.T_in) {
      t = -1;
    }
    tm[k] = (j < TBW) ? t : -2;  // -2: column not present
  }

  float acc[CO_T][RT];
#pragma unroll
  for (int c = 0; c < CO_T; ++c)
#pragma unroll
    for (int j = 0; j < RT; ++j) acc[c][j] = 0.f;

  for (int cbase = 0; cbase < a.Cin; cbase += kConvCI) {
    const int nci = min(kConvCI, a.Cin - cbase);
    __syncthreads();
    // ---- stage the input tile: one warp per (channel, row) line
    // (four lines per round: their loads are all in flight before the first store -- with few warps per block the staging
    // is latency-bound otherwise)
    const int n_lines = nci * RIN;
    for (int line0 = warp; line0 < n_lines; line0 += 4 * nwarps) {
      float xv[4][3];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int line = line0 + u * nwarps;
        const int ci = line / RIN, row = line - ci * RIN;
        int v = out_row0 * SR + a.row_off + row;
        const float* src = nullptr;
        if (a.row_circ) {
          v %= a.rows_v;
          if (v < 0) v += a.rows_v;
        }
        if (line < n_lines && v >= 0 && v < a.rows_v) {
          const int ch = cbase + ci;
          src = (ch < a.c0) ? a.in0 + b * a.bs0 + (long long)(ch * a.rows0 + v) * a.T_in
                            : a.in1 + b * a.bs1 + (long long)((ch - a.c0) * a.rows1 + v % a.rows1) * a.T_in;
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) xv[u][k] = (src != nullptr && tm[k] >= 0) ? __ldg(src + tm[k]) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int line = line0 + u * nwarps;
        if (line >= n_lines) break;
        float* dst = xs + line * XP;  // line = ci * RIN + row
#pragma unroll
        for (int k = 0; k < 3; ++k)
          if (tm[k] != -2) dst[lane + 32 * k] = xv[u][k];
      }
    }
    // ---- stage the weight slice [nci][KH][KW][CO_T]
    {
      const float* wsrc = a.w + (long long)cbase * KH * KW * a.cout_pad + cob * CO_T;
      for (int i = tid; i < nci * KH * KW * CO_T; i += blockDim.x) {
        const int k = i / CO_T, co = i - k * CO_T;
        wsm[i] = __ldg(wsrc + (long long)k * a.cout_pad + co);
      }
    }
    __syncthreads();
    if (active) {
      for (int ci = 0; ci < nci; ++ci) {
#pragma unroll 1
        for (int dr = 0; dr < KH; ++dr) {
          const float* xr = xs + (ci * RIN + r * SR + dr) * XP + tg * RT;
          float x[NX];
#pragma unroll
          for (int q = 0; q < NX / 4; ++q) {
            const float4 v4 = *reinterpret_cast<const float4*>(xr + 4 * q);
            x[4 * q] = v4.x, x[4 * q + 1] = v4.y, x[4 * q + 2] = v4.z, x[4 * q + 3] = v4.w;
          }
          if (NX % 4) {
            const float2 v2 = *reinterpret_cast<const float2*>(xr + (NX / 4) * 4);
            x[NX - 2] = v2.x, x[NX - 1] = v2.y;
          }
          const float* wr = wsm + (ci * KH + dr) * KW * CO_T;
#pragma unroll
          for (int dt = 0; dt < KW; ++dt) {
            float wv[CO_T];
            if constexpr (CO_T % 4 == 0) {
#pragma unroll
              for (int q = 0; q < CO_T / 4; ++q) {
                const float4 w4 = *reinterpret_cast<const float4*>(wr + dt * CO_T + 4 * q);
                wv[4 * q] = w4.x, wv[4 * q + 1] = w4.y, wv[4 * q + 2] = w4.z, wv[4 * q + 3] = w4.w;
              }
            } else {
#pragma unroll
              for (int c = 0; c < CO_T; ++c) wv[c] = wr[dt * CO_T + c];
            }
#pragma unroll
            for (int c = 0; c < CO_T; ++c)
#pragma unroll
              for (int j = 0; j < RT; ++j) acc[c][j] = fmaf(wv[c], x[j + dt], acc[c][j]);
          }
        }
      }
    }
  }

  const int out_row = out_row0 + r;
  if (a.stats != nullptr) {
    // raw outputs + their BatchNorm sums: every lane of every warp takes part in the shuffles, stores are predicated
    const bool store = active && out_row < a.rows_out;
#pragma unroll
    for (int c = 0; c < CO_T; ++c) {
      const int ch = cob * CO_T + c;  // (warp-uniform)
      if (ch >= a.Cout) break;
      const float s = a.scale[ch], h = a.shift[ch];
      float* op = a.out + b * a.obs + (a.out_coff + ch) * a.ocs + (long long)out_row * a.T_store;
      double s1 = 0.0, s2 = 0.0;
#pragma unroll
      for (int j = 0; j < RT; ++j) {
        const int t = t0 + tg * RT + j;
        if (store && t < a.T_out) {
          const float v = fmaf(acc[c][j], s, h);
          op[t] = v;
          s1 += (double)v, s2 += (double)v * (double)v;
        }
      }
      for (int o = 16; o; o >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
      }
      if (lane == 0 && tid < RB * a.tgroups) {
        atomicAdd(a.stats + 2 * ch, s1);
        atomicAdd(a.stats + 2 * ch + 1, s2);
      }
    }
    return;
  }
  if (!active || out_row >= a.rows_out) return;
#pragma unroll
  for (int c = 0; c < CO_T; ++c) {
    const int ch = cob * CO_T + c;
    if (ch >= a.Cout) break;
    const float s = a.scale[ch], h = a.shift[ch];
    float* op = a.out + b * a.obs + (a.out_coff + ch) * a.ocs + (long long)out_row * a.T_store;
    if (!a.pool_t) {
#pragma unroll
      for (int j = 0; j < RT; ++j) {
        const int t = t0 + tg * RT + j;
        if (t < a.T_out) {
          float v = apply_act(fmaf(acc[c][j], s, h), a.act);
          op[t] = a.accum ? op[t] + v : v;
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < RT / 2; ++j) {
        const int u = (t0 + tg * RT) / 2 + j;
        if (2 * u + 1 < a.T_out) {
          float v0 = fmaf(acc[c][2 * j], s, h), v1 = fmaf(acc[c][2 * j + 1], s, h);
          v0 = apply_act(v0, a.act), v1 = apply_act(v1, a.act);
          op[u] = fmaxf(v0, v1);
        }
      }
    }
  }
}

// ---- last convolution of the classifier heads (models.py:716-733): 32 -> 1 channel, valid in time ----------------
// tonic / key: EquivariantPitchClassConvolutionSimple(32 -> 1, 12 x 7, pitch classes wrap); genre: Conv2d(32, 1, (2, 7))
// (11 rows, no wrap).  A one-output-channel conv starves the tiled kernel above (no weight reuse across channels), so
// each thread here owns 8 consecutive frames of one output row and walks (ci, dp) with the frame window in registers.
// grid (B, heads, time tiles of kHeadTile frames), 128 threads.  No BatchNorm / activation follows (the per-frame
// logits go to the masked mean).
struct HeadTailArgs {
  const float* x[3];     // (B, 32, 12, T1) per head
  const float* w[3];     // (1, 32, KH, 7)
  const float* bias[3];  // 1
  float* out[3];         // (B, 1, rows_out, T1 - 6)
  int KH[3];
  int T1, Cin;
};

constexpr int kHeadCi = 4;     // input channels staged per pass
constexpr int kHeadTile = 80;  // output frames per block (10 strips of 8)
constexpr int kHeadTP = 96;    // staged row pitch: tile + 6 taps + strip overhang, multiple of 4

__global__ void __launch_bounds__(128) head_tail_kernel(const HeadTailArgs a) {
  extern __shared__ __align__(16) float hsm[];
  const int h = blockIdx.y, b = blockIdx.x;
  const int KH = a.KH[h], rows_out = 12 - (KH == 12 ? 0 : KH - 1), rows_in = KH == 12 ? 23 : 12;
  const int T1 = a.T1, Tf = T1 - 6;
  const int tb = blockIdx.z * kHeadTile;       // first output frame of this tile
  constexpr int TP = kHeadTP;
  float* ws = hsm;                             // [Cin][KH][8]
  float* xs = hsm + a.Cin * KH * 8;            // [kHeadCi][rows_in][TP]
  for (int i = threadIdx.x; i < a.Cin * KH * 8; i += blockDim.x) {
    const int dt = i % 8, k = i / 8;
    ws[i] = dt < 7 ? __ldg(a.w[h] + (long long)k * 7 + dt) : 0.f;
  }
  const int strips = (min(kHeadTile, Tf - tb) + 7) / 8;
  const bool active = threadIdx.x < rows_out * strips;
  const int c = threadIdx.x / strips, t0 = (threadIdx.x - c * strips) * 8;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  const float* xb = a.x[h] + (long long)b * a.Cin * 12 * T1;
  for (int c0 = 0; c0 < a.Cin; c0 += kHeadCi) {
    __syncthreads();
    for (int i = threadIdx.x; i < kHeadCi * rows_in * TP; i += blockDim.x) {
      const int t = i % TP, r = (i / TP) % rows_in, ci = i / (TP * rows_in);
      xs[i] = tb + t < T1 ? __ldg(xb + ((long long)(c0 + ci) * 12 + (r % 12)) * T1 + tb + t) : 0.f;
    }
    __syncthreads();
    if (active) {
      for (int ci = 0; ci < kHeadCi; ++ci) {
        for (int dp = 0; dp < KH; ++dp) {
          const float* xr = xs + (ci * rows_in + c + dp) * TP + t0;
          float x[16];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 v = *reinterpret_cast<const float4*>(xr + 4 * q);
            x[4 * q] = v.x, x[4 * q + 1] = v.y, x[4 * q + 2] = v.z, x[4 * q + 3] = v.w;
          }
          const float* wr = ws + ((c0 + ci) * KH + dp) * 8;
          const float4 w0 = *reinterpret_cast<const float4*>(wr), w1 = *reinterpret_cast<const float4*>(wr + 4);
          const float wv[7] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z};
#pragma unroll
          for (int dt = 0; dt < 7; ++dt)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = fmaf(wv[dt], x[j + dt], acc[j]);
        }
      }
    }
  }
  if (!active) return;
  const float bias = __ldg(a.bias[h]);
  float* dst = a.out[h] + ((long long)b * rows_out + c) * Tf;
#pragma unroll
  for (int j = 0; j < 8; ++j)
    if (tb + t0 + j < Tf) dst[tb + t0 + j] = acc[j] + bias;
}

// ---- ConvTranspose2d(C,C,(3,1),stride=(3,1)) (+affine+act): models.py:325-327 ------------------
// out[b,co,3c+r,t] = act(scale[co] * sum_ci W[ci,co,r] * pc[b,ci,c,t] + shift[co]);  W is (Cin,Cout,3,1).
__global__ void upsixth_kernel(const float* __restrict__ pc, const float* __restrict__ w,
                               const float* __restrict__ scale, const float* __restrict__ shift, int act,
                               float* __restrict__ out, int B, int C, int T) {
  const long long n = (long long)B * C * 36 * T;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int t = i % T;
    long long q = i / T;
    const int row = q % 36;
    q /= 36;
    const int co = q % C;
    const int b = q / C;
    const int c = row / 3, r = row - 3 * c;
    float acc = 0.f;
    for (int ci = 0; ci < C; ++ci)
      acc = fmaf(__ldg(w + (ci * C + co) * 3 + r), __ldg(pc + (((long long)b * C + ci) * 12 + c) * T + t), acc);
    float v = fmaf(acc, scale[co], shift[co]);
    out[i] = act ? leaky(v) : v;
  }
}

// ---- Pitch2PitchClassPool (models.py:82-106): pc[c] = max_o x[c + 12 o], optional affine+act first ----
__global__ void octmax_kernel(const float* __restrict__ in, int B, int C, int R, int T, const float* __restrict__ scale,
                              const float* __restrict__ shift, int affine_act, float* __restrict__ out, int C_total,
                              int coff) {
  const long long n = (long long)B * C * 12 * T;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int t = i % T;
    long long q = i / T;
    const int pc = q % 12;
    q /= 12;
    const int c = q % C;
    const int b = q / C;
    const float* src = in + (((long long)b * C + c) * R) * T + t;
    float s = 1.f, h = 0.f;
    if (affine_act) s = scale[c], h = shift[c];
    float m = -INFINITY;
    for (int row = pc; row < R; row += 12) {
      float v = __ldg(src + (long long)row * T);
      if (affine_act) v = leaky(fmaf(v, s, h));
      m = fmaxf(m, v);
    }
    out[(((long long)b * C_total + coff + c) * 12 + pc) * T + t] = m;
  }
}

// ---- layer 0 in one pass (eval mode): pool_semi Conv2d(1,1,3,stride (3,1),time-circular) + BN + LeakyReLU (models.py:313-315,
// 361-363) and Pitch2PitchClassPool (models.py:368) straight from the log-CQT.  One thread per (pitch class, frame): it walks
// the octaves, so every semitone is accumulated by the same instruction sequence (transposition equivariance stays bit exact).
// grid (12 * frames / 128, 1, B): the (pitch class, frame) pairs of a clip flattened over the blocks (151 frames in blocks of 128 would
// leave 41 % of the second block's threads idle).  `semi` (B,1,S,T) keeps the pre-pool map for the parity taps.
__global__ void __launch_bounds__(128) l0_semitone_pool_kernel(const float* __restrict__ mel, const float* __restrict__ w,
                                                                const float* __restrict__ scale, const float* __restrict__ shift,
                                                                float* __restrict__ semi, float* __restrict__ pc, int P, int T,
                                                                int C_total, int coff, __half* __restrict__ pl_hi = nullptr,
                                                                __half* __restrict__ pl_lo = nullptr) {
  const int idx = blockIdx.x * 128 + threadIdx.x, b = blockIdx.z;
  if (idx >= 12 * T) return;
  const int c = idx / T, t = idx - c * T;
  const int tm = t == 0 ? T - 1 : t - 1, tp = t == T - 1 ? 0 : t + 1;
  float k[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) k[i] = __ldg(w + i);
  const float sc = __ldg(scale), sh = __ldg(shift);
  const int S = P / 3;
  float best = -INFINITY;
  for (int j = c; j < S; j += 12) {
    const float* r0 = mel + ((long long)b * P + 3 * j) * T;
    float acc = 0.f;
#pragma unroll
    for (int dp = 0; dp < 3; ++dp) {
      const float* r = r0 + (long long)dp * T;
      acc = fmaf(k[dp * 3 + 0], __ldg(r + tm), acc);
      acc = fmaf(k[dp * 3 + 1], __ldg(r + t), acc);
      acc = fmaf(k[dp * 3 + 2], __ldg(r + tp), acc);
    }
    const float v = leaky(fmaf(acc, sc, sh));
    semi[((long long)b * S + j) * T + t] = v;
    best = fmaxf(best, v);
  }
  pc[(((long long)b * C_total + coff) * 12 + c) * T + t] = best;
  if (pl_hi) {
    // the same map as 8-channel chunk planes [B][23][T + 6][8] (channel 0 only) for the tensor-core equivariant convs:
    // column t + 3, wrap rows 12..22 = rows 0..10, zero halo columns (the "same" padding in time)
    const __half h = __float2half_rn(best), l = __float2half_rn(best - __half2float(h));
    const uint4 hv = make_uint4((uint32_t)__half_as_ushort(h), 0, 0, 0), lv = make_uint4((uint32_t)__half_as_ushort(l), 0, 0, 0);
    const uint4 z = make_uint4(0, 0, 0, 0);
    const int Wd = T + 6;
    for (int row = c; row < 23; row += 12) {
      const long long q = ((long long)b * 23 + row) * Wd;
      reinterpret_cast<uint4*>(pl_hi)[q + t + 3] = hv, reinterpret_cast<uint4*>(pl_lo)[q + t + 3] = lv;
      if (t < 3) {
        reinterpret_cast<uint4*>(pl_hi)[q + t] = z, reinterpret_cast<uint4*>(pl_lo)[q + t] = z;
        reinterpret_cast<uint4*>(pl_hi)[q + T + 3 + t] = z, reinterpret_cast<uint4*>(pl_lo)[q + T + 3 + t] = z;
      }
    }
  }
}

// ---- train-mode BatchNorm pieces (batch statistics over (B, rows, T) per channel) ---------------
// stats[2c] += sum, stats[2c+1] += sum of squares (double); in is (B, C_total, R, T) viewed at channel coff + c.
__global__ void bn_stats_kernel(const float* __restrict__ in, int B, int C_total, int coff, int RT_elems,
                                double* __restrict__ stats) {
  const int c = blockIdx.y;
  const long long n = (long long)B * RT_elems;
  double s = 0.0, ss = 0.0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / RT_elems, e = i - b * RT_elems;
    const double v = (double)__ldg(in + ((b * C_total + coff + c) * (long long)RT_elems) + e);
    s += v;
    ss += v * v;
  }
  __shared__ double sh[2][32];
  for (int o = 16; o; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    ss += __shfl_xor_sync(0xffffffffu, ss, o);
  }
  if ((threadIdx.x & 31) == 0) sh[0][threadIdx.x >> 5] = s, sh[1][threadIdx.x >> 5] = ss;
  __syncthreads();
  if (threadIdx.x < 32) {
    const int nw = blockDim.x >> 5;
    s = threadIdx.x < nw ? sh[0][threadIdx.x] : 0.0;
    ss = threadIdx.x < nw ? sh[1][threadIdx.x] : 0.0;
    for (int o = 16; o; o >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, o);
      ss += __shfl_xor_sync(0xffffffffu, ss, o);
    }
    if (threadIdx.x == 0) {
      atomicAdd(stats + 2 * c, s);
      atomicAdd(stats + 2 * c + 1, ss);
    }
  }
}

// scale = gamma / sqrt(var + eps), shift = beta - mean * scale (the conv bias is already inside the raw values)
__global__ void bn_finalize_kernel(const double* __restrict__ stats, double count, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, int C, float* __restrict__ scale,
                                   float* __restrict__ shift, float* __restrict__ stats_out, float* __restrict__ mean_out = nullptr,
                                   float* __restrict__ invstd_out = nullptr) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double mean = stats[2 * c] / count;
  double var = stats[2 * c + 1] / count - mean * mean;
  if (var < 0.0) var = 0.0;
  const double is = 1.0 / sqrt(var + (double)kBnEps);
  const double s = (double)gamma[c] * is;
  scale[c] = (float)s;
  shift[c] = (float)((double)beta[c] - mean * s);
  if (stats_out) stats_out[c] = (float)mean, stats_out[C + c] = (float)var;
  if (mean_out) mean_out[c] = (float)mean, invstd_out[c] = (float)is;  // kept for the BatchNorm backward (pcn_train.cuh)
}

// in-place y = act(x*scale[c] + shift[c]) over channels [coff, coff+C) of a (B, C_total, R*T) tensor
__global__ void affine_act_kernel(float* __restrict__ x, int B, int C_total, int coff, int C, int RT_elems,
                                  const float* __restrict__ scale, const float* __restrict__ shift, int act) {
  const long long n = (long long)B * C * RT_elems;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long e = i % RT_elems;
    long long q = i / RT_elems;
    const int c = q % C;
    const long long b = q / C;
    float* p = x + ((b * C_total + coff + c) * (long long)RT_elems) + e;
    *p = apply_act(fmaf(*p, scale[c], shift[c]), act);
  }
}

// MaxPool2d((1,2)) with optional affine+act first: (B,C,R,T) -> (B,C,R,T/2)
__global__ void timepool_kernel(const float* __restrict__ in, int B, int C, int R, int T, const float* __restrict__ scale,
                                const float* __restrict__ shift, int affine_act, float* __restrict__ out) {
  const int T2 = T / 2;
  const long long n = (long long)B * C * R * T2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int u = i % T2;
    const long long line = i / T2;
    const int c = (line / R) % C;
    float v0 = __ldg(in + line * T + 2 * u), v1 = __ldg(in + line * T + 2 * u + 1);
    if (affine_act) {
      v0 = leaky(fmaf(v0, scale[c], shift[c]));
      v1 = leaky(fmaf(v1, scale[c], shift[c]));
    }
    out[i] = fmaxf(v0, v1);
  }
}

// ---- non-default architecture pieces -----------------------------------------------------------
// Pitch2PitchClassConv (opt.p2pc_conv, models.py:108-133): Conv2d(C, C, (KS, 1), dilation (12, 1)) over the semitone rows --
// out[co, c, t] = sum_{ci, o < KS} W[co, ci, o] * x[ci, c + 12 o, t] -- the learned replacement of the octave max pool.  The
// epilogue y = act(acc * scale + shift) carries the conv bias and eval-mode BatchNorm (train mode: scale 1, shift = bias, no
// act; BatchNorm follows from the batch statistics).  Writes channels [coff, coff + C) of a (B, C_total, 12, T) tensor.
__global__ void p2pc_conv_kernel(const float* __restrict__ in, const float* __restrict__ w, int B, int C, int R, int KS, int T,
                                 const float* __restrict__ scale, const float* __restrict__ shift, int act, float* __restrict__ out,
                                 int C_total, int coff) {
  const long long n = (long long)B * C * 12 * T;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int t = i % T;
    long long q = i / T;
    const int pc = q % 12;
    q /= 12;
    const int co = q % C;
    const int b = q / C;
    float acc = 0.f;
    for (int ci = 0; ci < C; ++ci) {
      const float* src = in + (((long long)b * C + ci) * R + pc) * T + t;
      const float* wr = w + ((long long)co * C + ci) * KS;
      for (int o = 0; o < KS; ++o) acc = fmaf(__ldg(wr + o), __ldg(src + (long long)(12 * o) * T), acc);
    }
    const float v = fmaf(acc, scale[co], shift[co]);
    out[(((long long)b * C_total + coff + co) * 12 + pc) * T + t] = act ? leaky(v) : v;
  }
}

// PitchClass2Pitch_MemoryVariant (opt.pc2p_mem, models.py:145-166): the up-sampled pitch-class features (B, Cpc, 36, T) are summed
// over groups of Cpc / Cp channels and ADDED to the pitch-wise features instead of being concatenated:
//   out[b, c, p, t] = x[b, c, p, t] + sum_{g < Cpc/Cp} six[b, c * (Cpc/Cp) + g, p / (P/36), t]
// (the reference reshapes the pitch axis as (36, P/36): row p pairs with table row p / (P/36), NOT p % 36).
__global__ void pc2p_mem_add_kernel(const float* __restrict__ x, const float* __restrict__ six, int B, int Cp, int Cpc, int P, int T,
                                    float* __restrict__ out) {
  const long long n = (long long)B * Cp * P * T;
  const int G = Cpc / Cp, per = P / 36;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int t = i % T;
    long long q = i / T;
    const int pr = q % P;
    q /= P;
    const int c = q % Cp;
    const int b = q / Cp;
    float acc = 0.f;
    for (int g = 0; g < G; ++g) acc += __ldg(six + (((long long)b * Cpc + c * G + g) * 36 + pr / per) * T + t);
    out[i] = __ldg(x + i) + acc;
  }
}

// Tail of ResBlock / ResBlockEquivariant (models.py:402-454): out = LeakyReLU(x + bn2(conv2(..))).  `z` holds conv2's output:
// already normalised (eval: BatchNorm folded into the conv epilogue, scale == NULL) or raw (train: scale / shift from the batch).
__global__ void residual_act_kernel(const float* __restrict__ z, const float* __restrict__ x, int B, int C, int RT_elems,
                                    const float* __restrict__ scale, const float* __restrict__ shift, float* __restrict__ out) {
  const long long n = (long long)B * C * RT_elems;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int c = (i / RT_elems) % C;
    float v = __ldg(z + i);
    if (scale) v = fmaf(v, scale[c], shift[c]);
    out[i] = leaky(__ldg(x + i) + v);
  }
}

// opt.local heads (models.py:720-722): MaxPool2d((1, W), stride 1) along time behind the last head conv, (optionally) the
// sigmoid of the key head.  in (B, R, Tf) -> out (B, R, Tf - W + 1).  (W = 1: plain copy, the genre head has no pool.)
__global__ void slide_max_kernel(const float* __restrict__ in, int B, int R, int Tf, int W, int sigmoid, float* __restrict__ out) {
  const int To = Tf - W + 1;
  const long long n = (long long)B * R * To;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int t = i % To;
    const long long line = i / To;
    const float* src = in + line * Tf + t;
    float m = -INFINITY;
    for (int j = 0; j < W; ++j) m = fmaxf(m, __ldg(src + j));
    out[i] = sigmoid ? 1.f / (1.f + expf(-m)) : m;
  }
}

// Dense layers (opt.denseblock, models.py:456-648) normalise the CONCATENATION of everything before them: the first C channels
// of a (B, C_total_in, RT) tensor -> act(x * scale + shift) as a dense (B, C, RT) tensor (the raw features stay for later layers).
__global__ void bn_act_copy_kernel(const float* __restrict__ x, int B, int C_total_in, int C, int RT_elems, const float* __restrict__ scale,
                                   const float* __restrict__ shift, int act, float* __restrict__ out) {
  const long long n = (long long)B * C * RT_elems;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long e = i % RT_elems;
    long long q = i / RT_elems;
    const int c = q % C;
    const long long b = q / C;
    out[i] = apply_act(fmaf(__ldg(x + ((b * C_total_in + c) * (long long)RT_elems) + e), scale[c], shift[c]), act);
  }
}

// PitchClass2Pitch (models.py:135-143) materialised: rows r of the destination take row r % R_src of the source; written into
// channels [coff, coff + C) of a (B, C_total, R_dst, T) tensor (R_src == R_dst: a plain channel-slice copy).
__global__ void tile_rows_kernel(const float* __restrict__ src, int B, int C, int R_src, int R_dst, int T, float* __restrict__ dst,
                                 int C_total, int coff) {
  const long long n = (long long)B * C * R_dst * T;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int t = i % T;
    long long q = i / T;
    const int r = q % R_dst;
    q /= R_dst;
    const int c = q % C;
    const long long b = q / C;
    dst[(((b * C_total + coff + c) * R_dst) + r) * (long long)T + t] = __ldg(src + (((b * C + c) * R_src) + r % R_src) * (long long)T + t);
  }
}

// Frames the masked temporal mean runs over: actual_seq_length = floor(seq / pool^(L-1)) - (k-1) * head_layers, then the
// Python slice x[..., :actual_seq_length] (models.py:757-770): a positive length clamps to the Th frames that exist, a
// NEGATIVE one (very short clips) counts from the end -- Th + length frames -- and an empty slice gives torch.mean = nan.
__device__ __forceinline__ int masked_frames(int seq, int pool_div, int head_shrink, int Th) {
  const int m = seq / pool_div - head_shrink;
  return m >= 0 ? min(Th, m) : max(0, Th + m);
}

// ---- masked temporal mean / max + sigmoid (models.py:754-804) -----------------------------------
// One warp per output element (clip, head row).  frames: key (B,12,Th), tonic (B,12,Th), genre (B,11,Th).
// Fixed summation order (lane-strided partial sums, then a shuffle tree) independent of the row index.
__global__ void head_reduce_kernel(const float* __restrict__ key_f, const float* __restrict__ tonic_f,
                                   const float* __restrict__ genre_f, int B, int Th, const int* __restrict__ seq_len,
                                   int pool_div, int head_shrink, int max_pool, float* __restrict__ key_out,
                                   float* __restrict__ tonic_out, float* __restrict__ genre_out, int os_key = 12,
                                   int os_tonic = 12, int os_genre = 11) {  // os_*: floats between consecutive clips' outputs
  const int rows_per_clip = genre_f ? 35 : 24;
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (gw >= B * rows_per_clip) return;
  const int b = gw / rows_per_clip, rr = gw - b * rows_per_clip;
  const float* src;
  float* dst;
  int is_key = 0;
  if (rr < 12) {
    src = key_f + ((long long)b * 12 + rr) * Th, dst = key_out + (long long)b * os_key + rr, is_key = 1;
  } else if (rr < 24) {
    src = tonic_f + ((long long)b * 12 + rr - 12) * Th, dst = tonic_out + (long long)b * os_tonic + rr - 12;
  } else {
    src = genre_f + ((long long)b * 11 + rr - 24) * Th, dst = genre_out + (long long)b * os_genre + rr - 24;
  }
  int n = Th;
  bool use_max = max_pool != 0;
  if (seq_len) {
    // actual_seq_length = floor(seq / pool^(L-1)) - (k-1)*head_layers  (models.py:757-760); slicing clamps to Th
    n = masked_frames(seq_len[b], pool_div, head_shrink, Th);
    use_max = max_pool && b == 0;  // reference quirk: max_pool honoured for sample 0 only (models.py:765-785)
  }
  float v;
  if (n <= 0) {
    v = use_max ? -INFINITY : __int_as_float(0x7fc00000);  // empty slice: torch.mean -> nan
  } else if (use_max) {
    float m = -INFINITY;
    for (int t = lane; t < n; t += 32) m = fmaxf(m, __ldg(src + t));
    for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    v = m;
  } else {
    float s = 0.f;
    for (int t = lane; t < n; t += 32) s += __ldg(src + t);
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    v = s / (float)n;
  }
  if (is_key) v = 1.f / (1.f + expf(-v));
  if (lane == 0) *dst = v;
}

// ---- last head convolution folded into the temporal mean (eval fast path, max_pool off) ------------------------------------
// models.py:716-742, 754-804: out[c] = mean_{t < n} conv2(y)[c, t] with conv2 linear (32 -> 1 channel, KH x 7, valid in time, no
// activation behind it), hence  out[c] = b + (1/n) sum_{ci, dp, dt} W[ci, dp, dt] * S[ci, row(c, dp), dt],
// S[ci, r, dt] = sum_{t < n} y[ci, r, t + dt]  -- 7 windowed time sums per (channel, row) instead of a convolution over every frame.
// One block per (clip, head).  y = first-conv output as fp16 hi/lo chunk planes [B][G_total][R][T1][8] (equiv_umma_kernel EPI 3).
// Every row / pitch class runs the same instruction sequence (bit-exact transposition equivariance of key and tonic).
struct HeadFoldArgs {
  const __half* in_hi[3];
  const __half* in_lo[3];
  const float* w[3];      // (1, 32, KH, 7) fp32, reference layout
  const float* bias[3];
  float* out[3];          // key_out (B,12) [sigmoid], tonic_out (B,12), genre_out (B,11)
  int G_total[3], g0[3], R[3], KH[3], rows_out[3], wrap[3], sigmoid[3];
  const int* seq_len;     // or NULL
  int T1, Tf, pool_div, head_shrink;
  int out_stride[3];      // floats between consecutive clips' outputs (12 / 12 / 11, or 35 when the heads write (B, 35) result rows)
};

__global__ void __launch_bounds__(512) head_fold_kernel(const HeadFoldArgs a) {
  __shared__ float S[32 * 12 * 7];     // [ci][row][dt]
  __shared__ float edge[4 * 12 * 12 * 8];  // [group][row][j 12 = head 0..5, tail 0..5][8 ch]: y[j] and y[n + j]
  const int h = blockIdx.y, b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int n = a.Tf;
  if (a.seq_len) n = masked_frames(a.seq_len[b], a.pool_div, a.head_shrink, a.Tf);
  float* out = a.out[h] + (long long)b * a.out_stride[h];
  if (n <= 0) {  // empty slice: torch.mean -> nan
    if (threadIdx.x < a.rows_out[h]) out[threadIdx.x] = __int_as_float(0x7fc00000);
    return;
  }
  const int R = a.R[h];
  // ---- windowed sums: unit (group g, row r) per warp; lanes stride over frames
  constexpr int kWarps = 16;
  for (int u = warp; u < 48; u += kWarps) {
    const int g = u / 12, r = u - 12 * g;
    const long long base = ((((long long)b * a.G_total[h] + a.g0[h] + g) * R + r) * a.T1) * 8;
    float s0[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) s0[e] = 0.f;
#pragma unroll 3
    for (int t = lane; t < n + 6; t += 32) {
      const uint4 hv = __ldg(reinterpret_cast<const uint4*>(a.in_hi[h] + base + (long long)t * 8));
      const uint4 lv = __ldg(reinterpret_cast<const uint4*>(a.in_lo[h] + base + (long long)t * 8));
      const uint32_t hw[4] = {hv.x, hv.y, hv.z, hv.w}, lw[4] = {lv.x, lv.y, lv.z, lv.w};
      float y[8];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hw[e]));
        const float2 lf = __half22float2(*reinterpret_cast<const __half2*>(&lw[e]));
        y[2 * e] = hf.x + lf.x, y[2 * e + 1] = hf.y + lf.y;
      }
      if (t < n) {
#pragma unroll
        for (int e = 0; e < 8; ++e) s0[e] += y[e];
      }
      if (t < 6 || t >= n) {
        const int j = t < 6 ? t : 6 + (t - n);  // a frame can be both (n < 6): the tail copy is written by the second branch below
        float* d = edge + ((g * 12 + r) * 12 + j) * 8;
#pragma unroll
        for (int e = 0; e < 8; ++e) d[e] = y[e];
        if (t < 6 && t >= n) {
          float* d2 = edge + ((g * 12 + r) * 12 + 6 + (t - n)) * 8;
#pragma unroll
          for (int e = 0; e < 8; ++e) d2[e] = y[e];
        }
      }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e)
      for (int o = 16; o; o >>= 1) s0[e] += __shfl_xor_sync(0xffffffffu, s0[e], o);
    __syncwarp();
    // S[dt] = S[dt - 1] - y[dt - 1] + y[n + dt - 1]: lanes 0..7 = the 8 channels of the group
    if (lane < 8) {
      float sv = s0[lane];
      const float* ed = edge + (g * 12 + r) * 12 * 8 + lane;
      float* dst = S + ((g * 8 + lane) * 12 + r) * 7;
      dst[0] = sv;
#pragma unroll
      for (int dt = 1; dt < 7; ++dt) {
        sv = sv - ed[(dt - 1) * 8] + ed[(6 + dt - 1) * 8];
        dst[dt] = sv;
      }
    }
  }
  __syncthreads();
  // ---- out[c] = b + (1/n) sum W[ci][dp][dt] S[ci][row][dt]: one warp per output row, lanes over the (ci, dp) pairs, the seven time
  // taps of a pair in sequence (one index decode per seven MACs: decoding every (ci, dp, dt) triple cost ~50 integer instructions per MAC
  // and was most of the kernel's instruction count)
  const int KH = a.KH[h], n_pairs = 32 * KH;
  const float* w = a.w[h];
  for (int c = warp; c < a.rows_out[h]; c += kWarps) {
    float acc = 0.f;
    for (int q = lane; q < n_pairs; q += 32) {
      const int ci = KH == 12 ? q / 12 : q / KH, dp = q - ci * KH;
      int row = c + dp;
      if (a.wrap[h]) row -= row >= 12 ? 12 : 0;
      const float* wq = w + q * 7;
      const float* sq = S + (ci * 12 + row) * 7;
#pragma unroll
      for (int dt = 0; dt < 7; ++dt) acc = fmaf(__ldg(wq + dt), sq[dt], acc);
    }
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) {
      float v = acc / (float)n + __ldg(a.bias[h]);
      if (a.sigmoid[h]) v = 1.f / (1.f + expf(-v));
      out[c] = v;
    }
  }
}

// ---- key / tonic / genre decode (models.py:1083-1085, 1096, 923) --------------------------------
__constant__ unsigned short kKeySignatureBits[21] = {
    // bit (11 - pc) set when pitch class pc belongs to the signature; rows of utils/key_signatures.py:19-42
    0x5AB, 0x56B, 0xD6A, 0xD5A, 0xB5A, 0xB56, 0xAD6, 0xAD5, 0xAB5, 0x6B5, 0x6AD,
    0x5AD, 0x5AB, 0x56B, 0xD6A, 0x6B5, 0x5AD, 0x6AD, 0xB5A, 0xD5A, 0xB56};

// first maximum of (v, i) over the lanes of a warp (torch.argmax's tie rule); lanes that do not take part pass v = -inf
__device__ __forceinline__ int warp_argmax_first(float v, int i) {
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, v, o);
    const int oi = __shfl_xor_sync(0xffffffffu, i, o);
    if (ov > v || (ov == v && oi < i)) v = ov, i = oi;
  }
  return i;
}

// One WARP per clip: lane r < 21 scores key signature r (the same twelve conditional adds in the same order as a serial loop would do),
// lanes < 12 / < 11 hold the tonic / genre logits; first-maximum reductions by shuffles.  (One thread per clip walked 21 x 12 table bits
// serially: 8 us for 256 clips.)
__global__ void __launch_bounds__(128) decode_kernel(const float* __restrict__ key_out, const float* __restrict__ tonic_out,
                                                     const float* __restrict__ genre_out, int B, int* __restrict__ key_id,
                                                     int* __restrict__ tonic_id, int* __restrict__ genre_id, int os_key = 12, int os_tonic = 12,
                                                     int os_genre = 11) {
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (b >= B) return;
  if (key_id) {
    float k[12], nk = 0.f;
    for (int i = 0; i < 12; ++i) k[i] = key_out[(long long)b * os_key + i], nk = fmaf(k[i], k[i], nk);
    const float inv = 1.f / (fmaxf(sqrtf(nk), 1e-8f) * sqrtf(7.f));
    float d = -INFINITY;
    if (lane < 21) {
      d = 0.f;
      const unsigned bits = kKeySignatureBits[lane];
      for (int i = 0; i < 12; ++i)
        if ((bits >> (11 - i)) & 1) d += k[i];
      d *= inv;
    }
    const int best = warp_argmax_first(d, lane);
    if (lane == 0) key_id[b] = best;
  }
  if (tonic_id) {
    const float v = lane < 12 ? tonic_out[(long long)b * os_tonic + lane] : -INFINITY;
    const int best = warp_argmax_first(v, lane);
    if (lane == 0) tonic_id[b] = best;
  }
  if (genre_id) {
    int best = -1;
    if (genre_out) {
      const float v = lane < 11 ? genre_out[(long long)b * os_genre + lane] : -INFINITY;
      best = warp_argmax_first(v, lane);
    }
    if (lane == 0) genre_id[b] = best;
  }
}

// ---- MIREX-weighted key score (models.py:1065-1116): per-clip category + batch counters ----------
// counters[9] (accumulated): samples, correct, fifths, relative, parallel, other, all 12 key bits right ("accuracy"),
// correct tonics, key bits right.  One thread per clip; the categories follow the reference's if-chain (first match wins).
__global__ void __launch_bounds__(128) mirex_kernel(const float* __restrict__ key_out, const float* __restrict__ tonic_out,
                                                    const float* __restrict__ key_labels, const float* __restrict__ tonic_labels,
                                                    const float* __restrict__ key_sig_id, int sig_w, int B, unsigned long long* __restrict__ counters,
                                                    float* __restrict__ sim_out, int* __restrict__ cat_out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned int c[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  if (b < B) {
    float k[12], lab[12], nk = 0.f, nl = 0.f, dot = 0.f;
    for (int i = 0; i < 12; ++i) {
      k[i] = key_out[b * 12 + i], lab[i] = key_labels[b * 12 + i];
      nk = fmaf(k[i], k[i], nk), nl = fmaf(lab[i], lab[i], nl), dot = fmaf(k[i], lab[i], dot);
    }
    // argmax_r cos(key_out, MAP[r]) exactly as decode_kernel
    const float inv = 1.f / (fmaxf(sqrtf(nk), 1e-8f) * sqrtf(7.f));
    int pred = 0;
    float bv = -INFINITY;
    for (int r = 0; r < 21; ++r) {
      float d = 0.f;
      for (int i = 0; i < 12; ++i)
        if ((kKeySignatureBits[r] >> (11 - i)) & 1) d += k[i];
      d *= inv;
      if (d > bv) bv = d, pred = r;
    }
    // torch.argmax(key_signature_id[i]) over whatever width the data layer delivers (24-wide one-hot: KeyDataset.py:366, 447)
    int label = 0;
    for (int r = 1; r < sig_w; ++r)
      if (key_sig_id[(long long)b * sig_w + r] > key_sig_id[(long long)b * sig_w + label]) label = r;
    int bits = 0;
    for (int i = 0; i < 12; ++i) bits += (float)((kKeySignatureBits[pred] >> (11 - i)) & 1) == lab[i];
    int tl = 0, tp = 0;
    for (int i = 1; i < 12; ++i) {
      if (tonic_labels[b * 12 + i] > tonic_labels[b * 12 + tl]) tl = i;
      if (tonic_out[b * 12 + i] > tonic_out[b * 12 + tp]) tp = i;
    }
    const bool tonic_ok = tl == tp, keys_ok = bits == 12;
    const int diff = abs(pred - label);
    int cat;
    if (diff == 1 && !(tonic_ok && keys_ok)) cat = 1;   // fifth
    else if (tonic_ok && keys_ok) cat = 0;               // correct
    else if (keys_ok) cat = 2;                           // relative
    else if (tonic_ok) cat = 3;                          // parallel
    else cat = 4;                                        // other
    c[0] = 1, c[1 + cat] = 1, c[6] = keys_ok, c[7] = tonic_ok, c[8] = (unsigned)bits;
    if (cat_out) cat_out[b] = cat;
    if (sim_out) sim_out[b] = dot / (fmaxf(sqrtf(nk), 1e-8f) * fmaxf(sqrtf(nl), 1e-8f));  // nn.CosineSimilarity(dim=0), :1094
  }
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    unsigned int v = c[i];
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(counters + i, (unsigned long long)v);
  }
}

// Conv weights (Cout, Cin, KH, KW) -> [Cin][KH*KW][cout_pad] (zero padded output channels), the eval-mode epilogue (BatchNorm folded on
// the running statistics) and the raw epilogue (bias only):
// ---- every convolution's repacked weights and folded epilogues in ONE launch (a training loop uploads new parameters each step:
// ~35 separate 3 us launches otherwise).  The table travels as a kernel argument; blockIdx.y = table entry.
struct PackEntry {
  long long w_off, packed_off, b_off, gamma, beta, mean, var;  // float offsets into the flat parameter / packed buffers
  int Cout, Cin, KHW, cout_pad, ss_off;
  int pack;             // 1: repack the weights (false for transposed convs and norm-only sites)
  int has_bias, has_bn;
};
constexpr int kPackTableMax = 32;
struct PackTable {
  int n, n_ss;
  PackEntry e[kPackTableMax];
};
__global__ void __launch_bounds__(256) pack_all_kernel(const PackTable t, const float* __restrict__ params, float* __restrict__ packed,
                                                       float* __restrict__ ss_eval, float* __restrict__ ss_raw) {
  const PackEntry& e = t.e[blockIdx.y];
  if (e.pack) {
    const float* w = params + e.w_off;
    float* dst = packed + e.packed_off;
    const int n = e.Cin * e.KHW * e.cout_pad;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
      const int co = i % e.cout_pad, k = i / e.cout_pad;  // k = ci*KHW + tap
      dst[i] = co < e.Cout ? w[(long long)co * e.Cin * e.KHW + k] : 0.f;
    }
  }
  if (blockIdx.x == 0) {
    for (int c = threadIdx.x; c < e.Cout; c += blockDim.x) {
      const double bb = e.has_bias ? (double)params[e.b_off + c] : 0.0;
      ss_raw[e.ss_off + c] = 1.f;
      ss_raw[t.n_ss + e.ss_off + c] = (float)bb;
      if (e.has_bn) {
        const double sc = (double)params[e.gamma + c] / sqrt((double)params[e.var + c] + (double)kBnEps);
        ss_eval[e.ss_off + c] = (float)sc;
        ss_eval[t.n_ss + e.ss_off + c] = (float)((bb - (double)params[e.mean + c]) * sc + (double)params[e.beta + c]);
      } else {
        ss_eval[e.ss_off + c] = 1.f;
        ss_eval[t.n_ss + e.ss_off + c] = (float)bb;
      }
    }
  }
}

}  // namespace ake

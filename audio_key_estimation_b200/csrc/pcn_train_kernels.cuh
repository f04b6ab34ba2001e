// pcn_train_kernels.cuh -- backward kernels of the PitchClassNet training step (BASELINE config 5; SURVEY.md section 8
// a-15): gradients of models.py:352-399, 713-817 with train-mode BatchNorm, fp32 CUDA cores.
//
//   data gradient of a stride-1 row convolution  = the forward conv_rows_kernel run on dY with flipped, transposed weights
//   weight gradient                               = conv_wgrad_kernel (below), fp32 atomics into the flat gradient buffer
//   BatchNorm + LeakyReLU backward                = bn_bwd_reduce_kernel + bn_bwd_apply_kernel (two passes per site)
//   pools / up-sampling / masked mean             = one small kernel each
// The gradient buffer has the layout of the flat parameter buffer (ake_pcn_set_params_f32), so a data-parallel job
// all-reduces it as ONE bucket.
#pragma once
#include "pcn_kernels.cuh"

namespace ake {

// a = leaky(z * scale[c] + shift[c]), out of place (z is kept for the backward pass).  z, a: (B, C, RT).
__global__ void bn_act_out_kernel(const float* __restrict__ z, float* __restrict__ a, int B, int C, int RT,
                                  const float* __restrict__ scale, const float* __restrict__ shift) {
  const long long n = (long long)B * C * RT;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int c = (i / RT) % C;
    a[i] = leaky(fmaf(z[i], scale[c], shift[c]));
  }
}

// Weights of the data-gradient convolutions: dst[co][KH-1-dr][KW-1-dt][ci (padded)] = w[co][ci][dr][dt]
// (the forward packing [Cin'][KH][KW][cout_pad'] with Cin' = Cout, Cout' = Cin, taps flipped), every site of a backward pass in one
// launch: the table travels as a kernel argument, blockIdx.y = table entry.
struct DgradPackEntry {
  long long w_off, dst_off;  // float offsets into the flat parameters / the pass's scratch block
  int Cout, Cin, KH, KW, cin_pad;
};
constexpr int kDgradPackMax = 32;
struct DgradPackTable {
  int n;
  DgradPackEntry e[kDgradPackMax];
};
__global__ void __launch_bounds__(256) pack_dgrad_all_kernel(const DgradPackTable t, const float* __restrict__ params, float* __restrict__ scratch) {
  const DgradPackEntry& e = t.e[blockIdx.y];
  const float* w = params + e.w_off;
  float* dst = scratch + e.dst_off;
  const int n = e.Cout * e.KH * e.KW * e.cin_pad;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int ci = i % e.cin_pad;
    int q = i / e.cin_pad;
    const int dt = q % e.KW;
    q /= e.KW;
    const int dr = q % e.KH, co = q / e.KH;
    dst[i] = ci < e.Cin ? w[(((long long)co * e.Cin + ci) * e.KH + (e.KH - 1 - dr)) * e.KW + (e.KW - 1 - dt)] : 0.f;
  }
}

// ---- weight gradient of a row convolution ---------------------------------------------------------------------------
//   dW[co, ci, dr, dt] = sum_{b, r, t} dZ[b, co, r, t] * x[b, ci, rowmap(r*SR + off + dr), tmap(t - pad + dt)]
// A block owns (clip, row tile, 8 input channels, 8 output channels, a range of 16-frame tiles).  A thread owns one (ci, dr) and
// every RS-th output row of the tile; it keeps the 8 x KW accumulators of its 8 output channels in registers over all its frames
// (per row and tile: 22 input frames and 16 x 8 gradients from shared memory feed 16 x 8 x KW FMAs), the RS row shares are summed
// with warp shuffles and added to the gradient buffer once per block.
struct WgradArgs {
  const float* in0;
  const float* in1;
  int c0, c1, rows0, rows1;
  long long bs0, bs1;
  int T_in, rows_v, row_circ, row_off, pad_t, time_circ;
  int rows_out, T_out, Cin, Cout;
  const float* dz;  // (B, Cout, rows_out, T_out)
  float* dw;        // (Cout, Cin, KH, KW), accumulated
  int t_splits;     // blocks along the frames of a (clip, row tile)
  int n_cob;        // blocks of 8 output channels
};

constexpr int kWgTB = 16;             // frames per tile
constexpr int kWgCI = 8, kWgCO = 8;   // input / output channels per block
constexpr int kWgGP = kWgTB * kWgCO + 4;  // gradient row pitch (floats): rows land in different banks
__host__ __device__ constexpr int wg_xp(int kw) {  // input row pitch: a multiple of 4 floats with an odd quotient (conflict-free float4 rows)
  int xp = (kWgTB + kw - 1 + 3) / 4 * 4;
  return (xp / 4) % 2 ? xp : xp + 4;
}

template <int KH, int KW, int SR, int RB, int RS>
__global__ void __launch_bounds__(kWgCI* KH* RS) conv_wgrad_kernel(const WgradArgs a) {
  constexpr int RIN = (RB - 1) * SR + KH;
  constexpr int XW = kWgTB + KW - 1;
  constexpr int XP = wg_xp(KW);
  constexpr int NT = kWgCI * KH * RS;
  static_assert(RB % RS == 0 && (RS & (RS - 1)) == 0 && RS <= 32 && NT % 32 == 0 && XW <= 32, "thread layout");
  extern __shared__ float4 smem4[];
  float* xs = reinterpret_cast<float*>(smem4);  // [kWgCI][RIN][XP]
  float* gs = xs + kWgCI * RIN * XP;            // [RB][kWgTB][kWgCO] (+ 4 floats per row)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = NT / 32;
  const int rs = tid % RS, dr = (tid / RS) % KH, cil = tid / (RS * KH);
  const int row_tile = blockIdx.x / a.t_splits, tsp = blockIdx.x - row_tile * a.t_splits;
  const int cib = blockIdx.y / a.n_cob, cob = blockIdx.y - cib * a.n_cob;
  const int b = blockIdx.z;
  const int out_row0 = row_tile * RB;
  const int n_tiles = (a.T_out + kWgTB - 1) / kWgTB, per = (n_tiles + a.t_splits - 1) / a.t_splits;
  const int tile_lo = tsp * per, tile_hi = min(n_tiles, tile_lo + per);

  float acc[kWgCO][KW];
#pragma unroll
  for (int c = 0; c < kWgCO; ++c)
#pragma unroll
    for (int d = 0; d < KW; ++d) acc[c][d] = 0.f;

  for (int tile = tile_lo; tile < tile_hi; ++tile) {
    const int t0 = tile * kWgTB;
    __syncthreads();
    // ---- inputs: one warp per (channel, row) line, four lines of loads in flight
    {
      int t = t0 - a.pad_t + lane;
      if (a.time_circ) {
        t %= a.T_in;
        if (t < 0) t += a.T_in;
      } else if (t < 0 || t >= a.T_in) {
        t = -1;
      }
      if (lane >= XW) t = -1;
      for (int line0 = warp; line0 < kWgCI * RIN; line0 += 4 * NW) {
        float xv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int line = line0 + u * NW;
          const int ci = line / RIN, row = line - ci * RIN, ch = cib * kWgCI + ci;
          int v = out_row0 * SR + a.row_off + row;
          if (a.row_circ) {
            v %= a.rows_v;
            if (v < 0) v += a.rows_v;
          }
          xv[u] = 0.f;
          if (line < kWgCI * RIN && ch < a.Cin && v >= 0 && v < a.rows_v && t >= 0) {
            const float* src = (ch < a.c0) ? a.in0 + b * a.bs0 + (long long)(ch * a.rows0 + v) * a.T_in
                                           : a.in1 + b * a.bs1 + (long long)((ch - a.c0) * a.rows1 + v % a.rows1) * a.T_in;
            xv[u] = __ldg(src + t);
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int line = line0 + u * NW;
          if (line < kWgCI * RIN && lane < XP) xs[line * XP + lane] = xv[u];
        }
      }
    }
    // ---- output gradients, transposed to [row][frame][co]
    for (int i = tid; i < kWgCO * RB * kWgTB; i += NT) {
      const int j = i % kWgTB, r = (i / kWgTB) % RB, co = i / (kWgTB * RB);
      const int orow = out_row0 + r, t = t0 + j, cg = cob * kWgCO + co;
      gs[r * kWgGP + j * kWgCO + co] =
          (cg < a.Cout && orow < a.rows_out && t < a.T_out) ? __ldg(a.dz + (((long long)b * a.Cout + cg) * a.rows_out + orow) * a.T_out + t) : 0.f;
    }
    __syncthreads();
#pragma unroll 1
    for (int r = rs; r < RB; r += RS) {
      const float* xr = xs + (cil * RIN + r * SR + dr) * XP;
      const float* gr = gs + r * kWgGP;
      float x[XP];
#pragma unroll
      for (int q = 0; q < XP / 4; ++q) {
        const float4 v4 = *reinterpret_cast<const float4*>(xr + 4 * q);
        x[4 * q] = v4.x, x[4 * q + 1] = v4.y, x[4 * q + 2] = v4.z, x[4 * q + 3] = v4.w;
      }
#pragma unroll
      for (int j = 0; j < kWgTB; ++j) {
        const float4 g0 = *reinterpret_cast<const float4*>(gr + j * kWgCO), g1 = *reinterpret_cast<const float4*>(gr + j * kWgCO + 4);
        const float g[kWgCO] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
        for (int c = 0; c < kWgCO; ++c)
#pragma unroll
          for (int d = 0; d < KW; ++d) acc[c][d] = fmaf(g[c], x[j + d], acc[c][d]);
      }
    }
  }
  // ---- sum the RS row shares (consecutive lanes), one atomic per weight and block
#pragma unroll
  for (int c = 0; c < kWgCO; ++c)
#pragma unroll
    for (int d = 0; d < KW; ++d) {
      float v = acc[c][d];
#pragma unroll
      for (int o = RS / 2; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      acc[c][d] = v;
    }
  const int ch = cib * kWgCI + cil;
  if (rs == 0 && ch < a.Cin && tile_lo < tile_hi) {
#pragma unroll
    for (int c = 0; c < kWgCO; ++c) {
      const int cg = cob * kWgCO + c;
      if (cg >= a.Cout) break;
      float* dst = a.dw + (((long long)cg * a.Cin + ch) * KH + dr) * KW;
#pragma unroll
      for (int d = 0; d < KW; ++d) atomicAdd(dst + d, acc[c][d]);
    }
  }
}

// out[c] += sum over (B, RT) of x[b, c, :]   (bias gradients).  grid (chunks, C)
__global__ void channel_sum_kernel(const float* __restrict__ x, int B, int C, int RT, float* __restrict__ out) {
  const int c = blockIdx.y;
  const long long n = (long long)B * RT;
  float s = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / RT, e = i - b * RT;
    s += x[(b * C + c) * (long long)RT + e];
  }
  __shared__ float sh[32];
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.f;
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0) atomicAdd(out + c, s);
  }
}

// ---- BatchNorm (batch statistics) + LeakyReLU backward ----------------------------------------------------------------
//   y = z*scale + shift, xhat = (z - mean)*invstd, a = leaky(y);  dyh = da * (y > 0 ? 1 : slope)
//   dbeta = sum dyh, dgamma = sum dyh*xhat, dz = scale * (dyh - dbeta/N - xhat*dgamma/N)
// pass 1: sums[2c] += sum dyh, sums[2c+1] += sum dyh*xhat (double).  grid (chunks, C)
__global__ void bn_bwd_reduce_kernel(const float* __restrict__ z, const float* __restrict__ da, int B, int C, int RT,
                                     const float* __restrict__ scale, const float* __restrict__ shift,
                                     const float* __restrict__ mean, const float* __restrict__ invstd,
                                     double* __restrict__ sums) {
  const int c = blockIdx.y;
  const long long n = (long long)B * RT;
  const float sc = scale[c], sh_ = shift[c], mu = mean[c], is = invstd[c];
  double s1 = 0.0, s2 = 0.0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / RT, e = i - b * RT;
    const long long idx = (b * C + c) * (long long)RT + e;
    const float zz = z[idx];
    const float y = fmaf(zz, sc, sh_);
    const float g = da[idx] * (y > 0.f ? 1.f : kLeakySlope);
    s1 += g, s2 += g * ((zz - mu) * is);
  }
  __shared__ double sh[2][32];
  for (int o = 16; o; o >>= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  if ((threadIdx.x & 31) == 0) sh[0][threadIdx.x >> 5] = s1, sh[1][threadIdx.x >> 5] = s2;
  __syncthreads();
  if (threadIdx.x < 32) {
    const int nw = blockDim.x >> 5;
    s1 = threadIdx.x < nw ? sh[0][threadIdx.x] : 0.0;
    s2 = threadIdx.x < nw ? sh[1][threadIdx.x] : 0.0;
    for (int o = 16; o; o >>= 1) {
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    if (threadIdx.x == 0) {
      atomicAdd(sums + 2 * c, s1);
      atomicAdd(sums + 2 * c + 1, s2);
    }
  }
}

// pass 2: dz; block 0 also writes dgamma / dbeta.
__global__ void bn_bwd_apply_kernel(const float* __restrict__ z, const float* __restrict__ da, float* __restrict__ dz, int B,
                                    int C, int RT, const float* __restrict__ scale, const float* __restrict__ shift,
                                    const float* __restrict__ mean, const float* __restrict__ invstd,
                                    const double* __restrict__ sums, double count, float* __restrict__ dgamma,
                                    float* __restrict__ dbeta, unsigned* __restrict__ maxbits = nullptr) {
  unsigned mx = 0;  // largest |dz| as a float bit pattern (max is order-free): the scale of the tensor-core data gradient (pcn_train_tc.cuh)
  if (blockIdx.x == 0 && threadIdx.x < C) {
    dbeta[threadIdx.x] = (float)sums[2 * threadIdx.x];
    dgamma[threadIdx.x] = (float)sums[2 * threadIdx.x + 1];
  }
  const long long n = (long long)B * C * RT;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int c = (i / RT) % C;
    const float sc = scale[c], mu = mean[c], is = invstd[c];
    const float zz = z[i];
    const float y = fmaf(zz, sc, shift[c]);
    const float g = da[i] * (y > 0.f ? 1.f : kLeakySlope);
    const float xh = (zz - mu) * is;
    const float m1 = (float)(sums[2 * c] / count), m2 = (float)(sums[2 * c + 1] / count);
    const float d = sc * (g - m1 - xh * m2);
    dz[i] = d;
    mx = max(mx, __float_as_uint(fabsf(d)));
  }
  if (maxbits) {
    mx = __reduce_max_sync(0xffffffffu, mx);
    if ((threadIdx.x & 31) == 0 && mx) atomicMax(maxbits, mx);
  }
}

// ---- pools ------------------------------------------------------------------------------------------------------------
// Pitch2PitchClassPool backward (models.py:82-106): the gradient of pc[c] goes to the first maximal octave of
// a = leaky(z*scale + shift) (recomputed).  z, da: (B, C, R, T); dcat: (B, C_total, 12, T) read at channel coff + c.
__global__ void octmax_bwd_kernel(const float* __restrict__ z, int B, int C, int R, int T, const float* __restrict__ scale,
                                  const float* __restrict__ shift, const float* __restrict__ dcat, int C_total, int coff,
                                  float* __restrict__ da) {
  const long long n = (long long)B * C * 12 * T;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int t = i % T;
    long long q = i / T;
    const int pc = q % 12;
    q /= 12;
    const int c = q % C;
    const int b = q / C;
    const long long base = (((long long)b * C + c) * R) * T + t;
    const float s = scale[c], h = shift[c];
    float m = -INFINITY;
    int best = pc;
    for (int row = pc; row < R; row += 12) {
      const float v = leaky(fmaf(z[base + (long long)row * T], s, h));
      if (v > m) m = v, best = row;
    }
    const float g = dcat[(((long long)b * C_total + coff + c) * 12 + pc) * T + t];
    for (int row = pc; row < R; row += 12) da[base + (long long)row * T] = row == best ? g : 0.f;
  }
}

// MaxPool2d((1,2)) backward with the affine + activation recomputed: z, da (B*C*R lines of T), dpooled lines of T/2.
__global__ void timepool_bwd_kernel(const float* __restrict__ z, int B, int C, int R, int T, const float* __restrict__ scale,
                                    const float* __restrict__ shift, const float* __restrict__ dpooled,
                                    float* __restrict__ da) {
  const int T2 = T / 2;
  const long long n = (long long)B * C * R * T2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int u = i % T2;
    const long long line = i / T2;
    const int c = (line / R) % C;
    const float v0 = leaky(fmaf(z[line * T + 2 * u], scale[c], shift[c]));
    const float v1 = leaky(fmaf(z[line * T + 2 * u + 1], scale[c], shift[c]));
    const float g = dpooled[i];
    da[line * T + 2 * u] = v1 > v0 ? 0.f : g;
    da[line * T + 2 * u + 1] = v1 > v0 ? g : 0.f;
    if ((T & 1) && u == T2 - 1) da[line * T + T - 1] = 0.f;
  }
}

// ---- ConvTranspose2d(C,C,(3,1),stride (3,1)) backward (models.py:325): out[co,3c+r] = b + sum_ci W[ci,co,r] pc[ci,c] ----
//   dpc[b,ci,c,t] += sum_{co,r} W[ci,co,r] dz[b,co,3c+r,t];  dW[ci,co,r] += sum_{b,c,t} pc[b,ci,c,t] dz[b,co,3c+r,t]
// One thread per (b, c, t); C <= 8.  The weight gradient is reduced over the warp before the atomics.
__global__ void upsixth_bwd_kernel(const float* __restrict__ pc, const float* __restrict__ w, const float* __restrict__ dz,
                                   int B, int C, int T, float* __restrict__ dpc, float* __restrict__ dw) {
  const long long n = (long long)B * 12 * T;
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const bool ok = i < n;
  const int t = ok ? (int)(i % T) : 0;
  const int c = ok ? (int)((i / T) % 12) : 0;
  const int b = ok ? (int)(i / ((long long)T * 12)) : 0;
  float x[8], g[8][3];
  for (int ci = 0; ci < C; ++ci) x[ci] = ok ? pc[(((long long)b * C + ci) * 12 + c) * T + t] : 0.f;
  for (int co = 0; co < C; ++co)
    for (int r = 0; r < 3; ++r) g[co][r] = ok ? dz[(((long long)b * C + co) * 36 + 3 * c + r) * T + t] : 0.f;
  for (int ci = 0; ci < C; ++ci) {
    float d = 0.f;
    for (int co = 0; co < C; ++co)
      for (int r = 0; r < 3; ++r) {
        d = fmaf(__ldg(w + (ci * C + co) * 3 + r), g[co][r], d);
        float p = x[ci] * g[co][r];
        for (int o = 16; o; o >>= 1) p += __shfl_xor_sync(0xffffffffu, p, o);
        if ((threadIdx.x & 31) == 0) atomicAdd(dw + (ci * C + co) * 3 + r, p);
      }
    if (ok) dpc[(((long long)b * C + ci) * 12 + c) * T + t] += d;
  }
}

// ---- semitone conv (3x3, stride (3,1), time-circular) data gradient (models.py:337) ------------------------------------
//   dx[b,ci,3s+dp,t] = sum_{co,dt} W[co,ci,dp,dt] dz[b,co,s,(t - dt + 1) mod T]
__global__ void semitone_dgrad_kernel(const float* __restrict__ dz, const float* __restrict__ w, int B, int Cin, int Cout,
                                      int P, int T, float* __restrict__ dx) {
  const long long n = (long long)B * Cin * P * T;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int t = i % T;
    long long q = i / T;
    const int p = q % P;
    q /= P;
    const int ci = q % Cin;
    const int b = q / Cin;
    const int s = p / 3, dp = p - 3 * s, S = P / 3;
    float acc = 0.f;
    for (int co = 0; co < Cout; ++co) {
      const float* g = dz + (((long long)b * Cout + co) * S + s) * T;
      const float* wk = w + ((co * Cin + ci) * 3 + dp) * 3;
#pragma unroll
      for (int dt = 0; dt < 3; ++dt) {
        int tt = t - dt + 1;
        tt += tt < 0 ? T : 0, tt -= tt >= T ? T : 0;
        acc = fmaf(__ldg(wk + dt), g[tt], acc);
      }
    }
    dx[i] = acc;
  }
}

// PitchClass2Pitch backward (models.py:135-143): du[b,c,r36,t] = sum_o dx[b, coff + c, r36 + 36 o, t]
__global__ void tile_sum_kernel(const float* __restrict__ dx, int B, int C_total, int coff, int C, int P, int T,
                                float* __restrict__ du) {
  const long long n = (long long)B * C * 36 * T;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int t = i % T;
    long long q = i / T;
    const int r = q % 36;
    q /= 36;
    const int c = q % C;
    const int b = q / C;
    float s = 0.f;
    for (int row = r; row < P; row += 36) s += dx[(((long long)b * C_total + coff + c) * P + row) * T + t];
    du[i] = s;
  }
}

// dst[b, c, :] (+)= src[b, soff + c, :]   for c < C
__global__ void copy_channels_kernel(const float* __restrict__ src, int Cs, int soff, float* __restrict__ dst, int Cd, int doff,
                                     int B, int C, int RT, int accumulate) {
  const long long n = (long long)B * C * RT;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long e = i % RT;
    long long q = i / RT;
    const int c = q % C;
    const long long b = q / C;
    const float v = src[(b * Cs + soff + c) * RT + e];
    float* d = dst + (b * Cd + doff + c) * RT + e;
    *d = accumulate ? *d + v : v;
  }
}

// ---- masked temporal mean (+ sigmoid) backward (models.py:754-804) ------------------------------------------------------
// d_frames[b, row, t] = d_out[b, row] * (sigmoid' for the key head) / n_b for t < n_b, else 0.
__global__ void head_reduce_bwd_kernel(const float* __restrict__ d_key_out, const float* __restrict__ d_tonic_out,
                                       const float* __restrict__ d_genre_out, const float* __restrict__ key_out, int B, int Th,
                                       const int* __restrict__ seq_len, int pool_div, int head_shrink, float* __restrict__ d_key_f,
                                       float* __restrict__ d_tonic_f, float* __restrict__ d_genre_f) {
  const int rows_per_clip = d_genre_f ? 35 : 24;
  const long long n = (long long)B * rows_per_clip * Th;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int t = i % Th;
    const int rr = (i / Th) % rows_per_clip, b = i / ((long long)Th * rows_per_clip);
    int len = Th;
    if (seq_len) len = min(Th, seq_len[b] / pool_div - head_shrink);
    float g;
    float* dst;
    if (rr < 12) {
      const float p = key_out[b * 12 + rr];
      g = (d_key_out ? d_key_out[b * 12 + rr] : 0.f) * p * (1.f - p);
      dst = d_key_f + ((long long)b * 12 + rr) * Th + t;
    } else if (rr < 24) {
      g = d_tonic_out ? d_tonic_out[b * 12 + rr - 12] : 0.f;
      dst = d_tonic_f + ((long long)b * 12 + rr - 12) * Th + t;
    } else {
      g = d_genre_out ? d_genre_out[b * 11 + rr - 24] : 0.f;
      dst = d_genre_f + ((long long)b * 11 + rr - 24) * Th + t;
    }
    *dst = (len > 0 && t < len) ? g / (float)len : 0.f;
  }
}

// ---- training objective (models.py:855-896, global key estimation) -----------------------------------------------------
//   loss = key_weight * BCELoss(key_out, key_labels) + tonic_weight * CE(tonic_out, tonic_idx)
//        + genre_weight * CE(genre_out[mask], genre_idx[mask])          (mask: genre_idx >= 0; skipped when empty)
// One block.  loss_out[4] = {total, bce, tonic, genre}; the gradients w.r.t. the three outputs are written too.
__global__ void loss_kernel(const float* __restrict__ key_out, const float* __restrict__ tonic_out, const float* __restrict__ genre_out,
                            const float* __restrict__ key_labels, const int* __restrict__ tonic_idx, const int* __restrict__ genre_idx,
                            int B, float key_weight, float tonic_weight, float genre_weight, float* __restrict__ loss_out,
                            float* __restrict__ d_key, float* __restrict__ d_tonic, float* __restrict__ d_genre) {
  __shared__ float red[3][32];
  __shared__ int s_nm;
  const int tid = threadIdx.x;
  if (tid == 0) {
    int nm = 0;
    if (genre_out && genre_idx)
      for (int b = 0; b < B; ++b) nm += genre_idx[b] >= 0;
    s_nm = nm;
  }
  __syncthreads();
  const int nm = s_nm;
  float bce = 0.f, ce_t = 0.f, ce_g = 0.f;
  for (int i = tid; i < B * 12; i += blockDim.x) {
    const float p = key_out[i], y = key_labels[i];
    bce -= y * fmaxf(logf(p), -100.f) + (1.f - y) * fmaxf(log1pf(-p), -100.f);  // nn.BCELoss clamps the logs at -100
    d_key[i] = key_weight * (p - y) / fmaxf((1.f - p) * p, 1e-12f) / (float)(B * 12);
  }
  for (int b = tid; b < B; b += blockDim.x) {
    {
      const float* x = tonic_out + b * 12;
      float m = x[0];
      for (int i = 1; i < 12; ++i) m = fmaxf(m, x[i]);
      float s = 0.f;
      for (int i = 0; i < 12; ++i) s += expf(x[i] - m);
      const float lse = m + logf(s);
      ce_t += lse - x[tonic_idx[b]];
      for (int i = 0; i < 12; ++i) d_tonic[b * 12 + i] = tonic_weight * (expf(x[i] - lse) - (i == tonic_idx[b] ? 1.f : 0.f)) / (float)B;
    }
    if (genre_out) {
      const float* x = genre_out + b * 11;
      const bool on = genre_idx && genre_idx[b] >= 0 && nm > 0;
      float m = x[0];
      for (int i = 1; i < 11; ++i) m = fmaxf(m, x[i]);
      float s = 0.f;
      for (int i = 0; i < 11; ++i) s += expf(x[i] - m);
      const float lse = m + logf(s);
      if (on) ce_g += lse - x[genre_idx[b]];
      for (int i = 0; i < 11; ++i)
        d_genre[b * 11 + i] = on ? genre_weight * (expf(x[i] - lse) - (i == genre_idx[b] ? 1.f : 0.f)) / (float)nm : 0.f;
    }
  }
  for (int o = 16; o; o >>= 1) {
    bce += __shfl_xor_sync(0xffffffffu, bce, o);
    ce_t += __shfl_xor_sync(0xffffffffu, ce_t, o);
    ce_g += __shfl_xor_sync(0xffffffffu, ce_g, o);
  }
  if ((tid & 31) == 0) red[0][tid >> 5] = bce, red[1][tid >> 5] = ce_t, red[2][tid >> 5] = ce_g;
  __syncthreads();
  if (tid == 0) {
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) a0 += red[0][w], a1 += red[1][w], a2 += red[2][w];
    a0 /= (float)(B * 12), a1 /= (float)B, a2 = nm > 0 ? a2 / (float)nm : 0.f;
    loss_out[1] = a0, loss_out[2] = a1, loss_out[3] = a2;
    loss_out[0] = key_weight * a0 + tonic_weight * a1 + (nm > 0 ? genre_weight * a2 : 0.f);
  }
}

// ---- fused optimizer step (models.py:1017-1027: torch.optim.Adam(betas=(0.9, 0.999), lr, weight_decay=reg)) ----------
// One launch over the flat gradient bucket: element i belongs to the tensor t with off[t] <= i < off[t + 1] (binary search
// over the <= 64 tensor offsets); tensors without a parameter (running statistics) have a null pointer and are skipped.
// torch.optim.Adam, single-tensor form: g += wd p; m = b1 m + (1 - b1) g; v = b2 v + (1 - b2) g^2;
// p -= (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps).
struct AdamArgs {
  const float* grads;          // flat, layout of the parameter buffer
  float* m;
  float* v;                    // flat moments, same layout
  float* const* params;        // device table: parameter storage of every tensor (null: not trainable)
  const long long* offsets;    // device table: n_tensors + 1 offsets into the flat buffers
  int n_tensors;
  long long total;
  float lr, beta1, beta2, eps, weight_decay, grad_scale, bias1, bias2_sqrt;  // bias1 = 1 - b1^t, bias2_sqrt = sqrt(1 - b2^t)
};

__global__ void __launch_bounds__(256) adam_step_kernel(const AdamArgs a) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < a.total; i += (long long)gridDim.x * blockDim.x) {
    int lo = 0, hi = a.n_tensors;  // offsets[lo] <= i < offsets[hi]
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (__ldg(a.offsets + mid) <= i) lo = mid;
      else hi = mid;
    }
    float* p = a.params[lo];
    if (!p) continue;
    p += i - __ldg(a.offsets + lo);
    const float w = *p;
    const float g = fmaf(a.weight_decay, w, a.grads[i] * a.grad_scale);
    const float m = a.beta1 * a.m[i] + (1.f - a.beta1) * g;   // m.lerp_(g, 1 - beta1) rounds this way
    const float v = a.beta2 * a.v[i] + (1.f - a.beta2) * g * g;
    a.m[i] = m, a.v[i] = v;
    const float denom = sqrtf(v) / a.bias2_sqrt + a.eps;
    *p = w - (a.lr / a.bias1) * (m / denom);
  }
}

// ---- last convolution of a classifier head (Cin -> 1 channel, KH x 7, valid in time; models.py:716-742) in train mode --------------------
// 16 MMAC per step at batch 8: the generic row-tiled conv spends ~50 us of pure latency on it (one output channel per block pass, four
// staged input-channel passes); here one thread owns one output frame of one row and a quarter of the input channels.  (A matching
// dedicated data-gradient kernel was measured too: 8.7 us against the generic kernel's 6.5 us for the single input channel -- dropped.)
//   out[b, 0, r, t] = bias + sum_{ci, dp, dt} W[0, ci, dp, dt] x[b, ci, row(r + dp), t + dt]      row() wraps modulo R_in when `wrap`
// grid (R_out, B), 256 threads = 4 channel quarters x 64 frames (T_out <= 64 per pass, looped otherwise)
template <int KH>
__global__ void __launch_bounds__(256) tail_conv_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                                                            float* __restrict__ out, int Cin, int R_in, int R_out, int T_in, int T_out,
                                                            int wrap) {
  extern __shared__ float tail_smem[];
  float* ws = tail_smem;                 // [Cin][KH][7]
  float* red = ws + Cin * KH * 7;        // [4][64]
  const int r = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, q = tid >> 6, tl = tid & 63;
  for (int i = tid; i < Cin * KH * 7; i += 256) ws[i] = __ldg(w + i);
  __syncthreads();
  const int cq = (Cin + 3) / 4;
  for (int t0 = 0; t0 < T_out; t0 += 64) {
    const int t = t0 + tl;
    float acc = 0.f;
    if (t < T_out) {
      for (int ci = q * cq; ci < min(Cin, (q + 1) * cq); ++ci) {
        // all KH x 7 loads of a channel are in flight before the first FMA: every (channel, row) line is a first touch for this
        // block, i.e. an L2 round trip -- one per channel instead of one per row tap
        float xv[KH][7];
        const float* xc = x + ((long long)b * Cin + ci) * R_in * T_in + t;
#pragma unroll
        for (int dp = 0; dp < KH; ++dp) {
          int row = r + dp;
          if (wrap) row -= row >= R_in ? R_in : 0;
#pragma unroll
          for (int dt = 0; dt < 7; ++dt) xv[dp][dt] = __ldg(xc + row * T_in + dt);
        }
        const float* wc = ws + ci * KH * 7;
#pragma unroll
        for (int dp = 0; dp < KH; ++dp)
#pragma unroll
          for (int dt = 0; dt < 7; ++dt) acc = fmaf(wc[dp * 7 + dt], xv[dp][dt], acc);
      }
    }
    red[q * 64 + tl] = acc;
    __syncthreads();
    if (q == 0 && t < T_out)
      out[((long long)b * R_out + r) * T_out + t] = ((red[tl] + red[64 + tl]) + (red[128 + tl] + red[192 + tl])) + (bias ? __ldg(bias) : 0.f);
    __syncthreads();
  }
}

}  // namespace ake

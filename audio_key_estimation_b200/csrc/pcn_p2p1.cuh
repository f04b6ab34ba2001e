// pcn_p2p1.cuh -- the FIRST convolution of the Pitch2Pitch stack (models.py:228-234 applied to cat[p, tile(up_sixth(pc))],
// models.py:372-383), split by input channel:
//
//   out[co,p,t] = act(bn( sum_{dp,dt} W[co,0,dp,dt] mel[(p+dp-3) mod P, (t+dt-3) mod T]                       (1 channel, P rows)
//                       + sum_{ci=1..4} sum_{dp,dt} W[co,ci,dp,dt] up[ci-1, (p+dp-3) mod 36, (t+dt-3) mod T] ))   (4 channels, 36 rows)
//
// PitchClass2Pitch (models.py:135-143) tiles the 36 up-sampled rows over all pitches and 36 divides P, so the second sum has
// period 36 in p: it is computed once per clip on a 36-row image (`p2p_umma_kernel<false, true>` on the planes written by
// upsixth_planes_kernel; raw fp32 accumulators -> `tab`) instead of P / 36 = 8 times.  The first sum has ONE input channel,
// so the K = 8 slots of an operand chunk hold the seven TIME taps of a position instead of channels:
//   chunk(row, col) = [mel(row, col-3) ... mel(row, col+3), 0]   (fp16 hi plane, fp16 lo plane)
//   MMA per row tap dp and 128 anchors: K = 16 = [chunk_hi | chunk_lo],  N = 16 = 8 co x {W_hi, W_lo}
// -- no time-tap phases in N, hence no phase realignment in the epilogue (one 16-column TMEM load per anchor instead of seven
// plus 48 shuffles), and the tile carries no garbage anchors.  MACs: (P + 4 * 36) * 392 T instead of 5 P * 392 T (3.3x fewer).
// Every output is still accumulated in the same order whatever its pitch: transposition equivariance stays bit exact.
#pragma once
#include "pcn_umma.cuh"

namespace ake {

constexpr int kF1Groups = 4;       // epilogue groups of 4 warps = accumulator buffers of 16 TMEM columns; a multiple of kF1Issuers, so that
                                   // every accumulator belongs to ONE issuer (mbarrier parity waits must not skip a phase)
constexpr int kF1GenWarps = 9;     // tile-generator warps: 288 threads >= (8 + 6) rows x 20 column groups, one item per thread and tile
constexpr int kF1Issuers = 2;      // MMA-issuer warps (alternate blocks; one warp issues at most one MMA per ~61 cycles)
constexpr int kF1Threads = 32 * (4 * kF1Groups + kF1GenWarps + kF1Issuers);
constexpr int kF1Bufs = 2;
constexpr uint32_t kF1WBytes = 7 * 2 * 16 * 16;  // [dp 7][chunk 2][n 16][k 8] fp16

__host__ __device__ inline uint32_t f1_plane_positions(int TB) { return (uint32_t)((kP2PRows + 6) * TB + 136); }
__host__ __device__ inline size_t f1_smem_bytes(int TB) { return (size_t)2 * kF1Bufs * f1_plane_positions(TB) * 16 + kF1WBytes; }

// weight image of the mel channel: n < 8: W_hi of output channel n, n >= 8: W_lo of output channel n - 8; k = time tap (k = 7: 0).
// Both K chunks (x_hi, x_lo) hold the same weights.
__global__ void p2p1_pack_mel_kernel(const float* __restrict__ w, int Cout, int Cin, __half* __restrict__ img) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 7 * 8 * 8) return;
  const int k = i % 8, co = (i / 8) % 8, dp = i / 64;
  float v = 0.f;
  if (k < 7 && co < Cout) v = w[(((long long)co * Cin + 0) * 7 + dp) * 7 + k] * kWScale;
  const __half hi = __float2half_rn(v);
  const __half lo = __float2half_rn(v - __half2float(hi));
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    img[((dp * 2 + c) * 16 + co) * 8 + k] = hi;
    img[((dp * 2 + c) * 16 + 8 + co) * 8 + k] = lo;
  }
}

// p2p_pack_weights_kernel for input channels [ci0, ci0 + n_ci) of a conv with Cin channels (the periodic part)
__global__ void p2p_pack_weights_sub_kernel(const float* __restrict__ w, int Cout, int Cin, int ci0, int n_ci, __half* __restrict__ img) {
  const int n_items = 7 * 56 * 8;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_items; i += gridDim.x * blockDim.x) {
    const int ci = i % 8, fc = (i / 8) % 56, dp = i / 448;
    const int f = fc / 8, co = fc % 8;
    float v = 0.f;
    if (ci < n_ci && co < Cout) v = w[(((long long)co * Cin + ci0 + ci) * 7 + dp) * 7 + f] * kWScale;
    p2p_img_store(img, dp, f, co, ci, v);
  }
}

// up_sixth (models.py:372-374) as chunk planes of a 36-row image with circular halos: [B][36 + 6][Wd][8] (channels 4..7 zero).
// grid (ceil(36 T / 128), 1, B)
__global__ void __launch_bounds__(128) upsixth_planes_kernel(const float* __restrict__ pc, const float* __restrict__ w_up,
                                                             const float* __restrict__ scale, const float* __restrict__ shift,
                                                             __half* __restrict__ out_hi, __half* __restrict__ out_lo, int T, int Wd) {
  const int idx = blockIdx.x * 128 + threadIdx.x, b = blockIdx.z;  // (row, frame) pairs flattened over the blocks: no idle tail threads
  if (idx >= 36 * T) return;
  const int p36 = idx / T, t = idx - p36 * T;
  const int c = p36 / 3, r = p36 - 3 * c;
  float x[4], y[8];
#pragma unroll
  for (int ci = 0; ci < 4; ++ci) x[ci] = __ldg(pc + (((long long)b * 4 + ci) * 12 + c) * T + t);
#pragma unroll
  for (int co = 0; co < 4; ++co) {
    float acc = 0.f;
#pragma unroll
    for (int ci = 0; ci < 4; ++ci) acc = fmaf(__ldg(w_up + (ci * 4 + co) * 3 + r), x[ci], acc);
    y[co] = leaky_f(fmaf(acc, __ldg(scale + co), __ldg(shift + co)));
    y[4 + co] = 0.f;
  }
  uint32_t h[4], l[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) umma::split_f16x2(y[2 * e], y[2 * e + 1], h[e], l[e]);
  const uint4 hv = make_uint4(h[0], h[1], h[2], h[3]), lv = make_uint4(l[0], l[1], l[2], l[3]);
  const long long base = (long long)b * 42;
  const int row2 = (p36 < 3) ? p36 + 3 + 36 : ((p36 >= 33) ? p36 + 3 - 36 : -1);
  const int col2 = (t < 3) ? t + 3 + T : ((t >= T - 3) ? t + 3 - T : -1);
  const long long q00 = ((base + p36 + 3) * Wd + t + 3) * 8;
  *reinterpret_cast<uint4*>(out_hi + q00) = hv, *reinterpret_cast<uint4*>(out_lo + q00) = lv;
  if (col2 >= 0) {
    const long long q = ((base + p36 + 3) * Wd + col2) * 8;
    *reinterpret_cast<uint4*>(out_hi + q) = hv, *reinterpret_cast<uint4*>(out_lo + q) = lv;
  }
  if (row2 >= 0) {
    const long long q = ((base + row2) * Wd + t + 3) * 8;
    *reinterpret_cast<uint4*>(out_hi + q) = hv, *reinterpret_cast<uint4*>(out_lo + q) = lv;
    if (col2 >= 0) {
      const long long q2 = ((base + row2) * Wd + col2) * 8;
      *reinterpret_cast<uint4*>(out_hi + q2) = hv, *reinterpret_cast<uint4*>(out_lo + q2) = lv;
    }
  }
}

struct P2P1Args {
  const float* mel;      // (B, 1, P, T)
  const float* tab;      // (B, 36, T, 8): periodic part, raw accumulators (weights scaled by kWScale)
  __half* out_hi;
  __half* out_lo;        // [B][P+6][Wd][8]
  const __half* wimg;    // p2p1_pack_mel_kernel image
  const float* scale;    // 8: eval-mode BN scale (the 1/kWScale factor is applied in the kernel)
  const float* shift;    // 8
  int P, T, Wd;
  int TB, n_ttiles;      // frames per tile, tiles along time (T >= 64)
  int n_rtiles, n_tiles;
};

// Persistent CTA (one per SM), warp-specialised:
//   warps [0, 4G)        : G epilogue groups of 4 warps; group g drains blocks g, g + G, ... (thread = TMEM lane = anchor)
//   warps [4G, 4G + NGEN): tile generators -- (rows + 6) x TB chunks of seven time taps, fp16 hi / lo planes, two tile buffers
//   last kF1Issuers warps: MMA issuers (converged warps, one elected lane issues), alternate blocks
__global__ void __launch_bounds__(kF1Threads, 1) p2p1_umma_kernel(const P2P1Args a) {
  using namespace umma;
  constexpr int G = kF1Groups, NGEN = kF1GenWarps, ISSUER0 = 4 * G + NGEN, NI = kF1Issuers;
  static_assert(G % NI == 0, "every accumulator must belong to one issuer");
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t w_bar, full_bar[kF1Bufs], empty_bar[kF1Bufs], acc_full[G], acc_empty[G];
  __shared__ uint32_t tmem_slot;
  __shared__ float s_scale[8], s_shift[8];

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const int TB = a.TB;
  const uint32_t plane = f1_plane_positions(TB) * 16;
  uint8_t* s_w = smem + 2 * kF1Bufs * plane;
  const int tiles_per_clip = a.n_rtiles * a.n_ttiles;

  struct Geom {
    int b, p0, t0, PB, TBv, n_anchor, n_mb;
  };
  const uint32_t tpc_magic = 0xFFFFFFFFu / (uint32_t)tiles_per_clip + 1, ntt_magic = 0xFFFFFFFFu / (uint32_t)a.n_ttiles + 1;
  auto geom = [&](int tile) {
    Geom g;
    g.b = tiles_per_clip == 1 ? tile : (int)__umulhi((uint32_t)tile, tpc_magic);
    const int r = tile - g.b * tiles_per_clip;
    const int rtile = a.n_ttiles == 1 ? r : (int)__umulhi((uint32_t)r, ntt_magic), ttile = r - rtile * a.n_ttiles;
    g.p0 = rtile * kP2PRows, g.t0 = ttile * TB;
    g.PB = min(kP2PRows, a.P - g.p0);
    g.TBv = min(TB, a.T - g.t0);
    g.n_anchor = g.PB * TB;
    g.n_mb = (g.n_anchor + 127) >> 7;
    return g;
  };

  if (warp == ISSUER0) tmem_alloc(&tmem_slot, 64);
  if (threadIdx.x == 0) {
    mbar_init(&w_bar, 1);
    for (int i = 0; i < kF1Bufs; ++i) mbar_init(&full_bar[i], NGEN), mbar_init(&empty_bar[i], NI);
    for (int i = 0; i < G; ++i) mbar_init(&acc_full[i], 1), mbar_init(&acc_empty[i], 128);
    mbar_init_fence();
  }
  if (threadIdx.x < 8) s_scale[threadIdx.x] = a.scale[threadIdx.x] * (1.f / kWScale), s_shift[threadIdx.x] = a.shift[threadIdx.x];
  // positions the generators never write (tail padding, columns of a narrow last time tile) only feed discarded anchors, but
  // they must hold finite values
  for (uint32_t i = threadIdx.x; i < 2 * kF1Bufs * plane / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_slot;

  if (warp >= 4 * G && warp < ISSUER0) {
    // ------------------------------------------------------------ tile generators
    const int gt = threadIdx.x - 128 * G;
    if (gt == 0) {
      mbar_arrive_expect_tx(&w_bar, kF1WBytes);
      bulk_g2s(s_w, a.wimg, kF1WBytes, &w_bar);
    }
    // Warp item = (tile row, segment of 26 columns): lane L loads ONE sample (input column 26 seg - 3 + L), splits it into
    // fp16 hi / lo once, and the seven time taps of a position come from the neighbouring lanes by shuffle; lanes 3..28 store
    // the two 16-byte chunks of their position (consecutive lanes = consecutive positions: conflict-free).  The samples of
    // the NEXT tile are loaded before this tile is converted (software pipeline: the generators would otherwise sit out one
    // L2 round trip per tile).
    const int gw = warp - 4 * G;
    const int n_seg = (TB + 25) / 26;
    const uint32_t seg_magic = 0xFFFFFFFFu / (uint32_t)n_seg + 1;
    constexpr int MAXI = 11;  // items per warp and tile: (8 + 6) rows x 7 segments (TB <= 182) / 9 warps
    auto fetch_tile = [&](const Geom& g, float (&x)[MAXI]) {
      const int n_it = (g.PB + 6) * n_seg;
      const float* mel_b = a.mel + (long long)g.b * a.P * a.T;
#pragma unroll
      for (int i = 0; i < MAXI; ++i) {
        const int item = gw + i * NGEN;
        x[i] = 0.f;
        if (item < n_it) {
          const int row = (int)__umulhi((uint32_t)item, seg_magic), sgm = item - row * n_seg;
          int p = g.p0 + row - 3;
          p += (p < 0) ? a.P : 0, p -= (p >= a.P) ? a.P : 0;
          int t = g.t0 + 26 * sgm - 3 + lane;  // circular in time (T >= 64: one wrap each way suffices)
          t += (t < 0) ? a.T : 0, t -= (t >= a.T) ? a.T : 0;
          x[i] = __ldg(mel_b + (long long)p * a.T + t);
        }
      }
    };
    auto emit_tile = [&](const Geom& g, int s, const float (&x)[MAXI]) {
      const int n_it = (g.PB + 6) * n_seg;
      uint4* d_hi = reinterpret_cast<uint4*>(smem + (size_t)s * 2 * plane);
      uint4* d_lo = reinterpret_cast<uint4*>(smem + (size_t)s * 2 * plane + plane);
#pragma unroll
      for (int i = 0; i < MAXI; ++i) {
        const int item = gw + i * NGEN;
        if (item < n_it) {  // warp-uniform
          const int row = (int)__umulhi((uint32_t)item, seg_magic), sgm = item - row * n_seg;
          const __half h = __float2half_rn(x[i]);
          const __half l = __float2half_rn(x[i] - __half2float(h));
          const uint32_t u = pack_h2(h, l);  // hi in the low half, lo in the high half
          uint32_t n[7];
#pragma unroll
          for (int d = 0; d < 7; ++d) n[d] = __shfl_sync(0xffffffffu, u, lane + d - 3);
          const int col = 26 * sgm + lane - 3;
          if (lane >= 3 && lane < 29 && col < TB) {
            const int q = row * TB + col;
            d_hi[q] = make_uint4(__byte_perm(n[0], n[1], 0x5410), __byte_perm(n[2], n[3], 0x5410), __byte_perm(n[4], n[5], 0x5410), n[6] & 0xFFFFu);
            d_lo[q] = make_uint4(__byte_perm(n[0], n[1], 0x7632), __byte_perm(n[2], n[3], 0x7632), __byte_perm(n[4], n[5], 0x7632), n[6] >> 16);
          }
        }
      }
    };
    float x[MAXI], xn[MAXI];
    Geom g = geom(min((int)blockIdx.x, a.n_tiles - 1));
    fetch_tile(g, x);
    int k = 0;
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++k) {
      const int s = k % kF1Bufs;
      const int tile_n = tile + (int)gridDim.x;
      Geom gn = g;
      if (tile_n < a.n_tiles) {
        gn = geom(tile_n);
        fetch_tile(gn, xn);
      }
      mbar_wait_relaxed(&empty_bar[s], ((k / kF1Bufs) & 1) ^ 1);
      emit_tile(g, s, x);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&full_bar[s]);
#pragma unroll
      for (int i = 0; i < MAXI; ++i) x[i] = xn[i];
      g = gn;
    }
  } else if (warp >= ISSUER0) {
    // ------------------------------------------------------------ MMA issuers: warp iw issues blocks j with j % NI == iw
    const int iw = warp - ISSUER0;
    const uint64_t A_DESC = desc_hi(plane);        // chunk 1 = the lo plane at the same position
    constexpr uint64_t B_DESC = desc_hi(16 * 16);  // chunk stride: 16 rows x 16 B
    constexpr uint32_t IDESC = idesc_f16(16);
    const uint32_t w0 = smem_u32(s_w);
    mbar_wait(&w_bar, 0);
    int k = 0;
    uint32_t j = 0;  // block counter of this CTA (all roles walk the same sequence)
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++k) {
      const int s = k % kF1Bufs;
      const Geom g = geom(tile);
      const uint32_t hi0 = smem_u32(smem + (size_t)s * 2 * plane);
      mbar_wait(&full_bar[s], (k / kF1Bufs) & 1);
      fence_after_sync();
      for (int m = 0; m < g.n_mb; ++m, ++j) {
        if ((int)(j % NI) != iw) continue;
        const uint32_t buf = j % G;
        mbar_wait(&acc_empty[buf], ((j / G) & 1) ^ 1);
        fence_after_sync();
        const uint32_t d = tmem + buf * 16;
        const uint32_t a_off = hi0 + (uint32_t)(m * 128) * 16;
        if (elect_one()) {
#pragma unroll
          for (int dp = 0; dp < 7; ++dp)
            mma_f16(d, make_desc(A_DESC, a_off + (uint32_t)(dp * TB) * 16), make_desc(B_DESC, w0 + dp * (2 * 16 * 16)), IDESC, dp ? 1u : 0u);
          commit(&acc_full[buf]);
        }
        __syncwarp();
      }
      // the tile buffer is free once the MMAs of every issuer have read it (each issuer commits after its last block of the tile)
      if (elect_one()) commit(&empty_bar[s]);
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------ epilogue: thread = TMEM lane = anchor
    const int grp = warp >> 2, wq = warp & 3, tid = threadIdx.x & 127;
    const uint32_t acc = tmem + ((uint32_t)(wq * 32) << 16) + grp * 16;
    const uint32_t tb_magic = 0xFFFFFFFFu / (uint32_t)TB + 1;
    uint64_t sc2[4], sh2[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) sc2[e] = f2_pack(s_scale[2 * e], s_scale[2 * e + 1]), sh2[e] = f2_pack(s_shift[2 * e], s_shift[2 * e + 1]);
    uint32_t n_done = 0;
    uint32_t j = grp, j0 = 0;
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
      const Geom g = geom(tile);
      const long long base = (long long)g.b * (a.P + 6);
      const int p36_0 = g.p0 % 36;
      for (; j < j0 + g.n_mb; j += G, ++n_done) {
        const int m = (int)(j - j0);
        const int anchor = m * 128 + tid;
        const int pl = (int)__umulhi((uint32_t)anchor, tb_magic), tl = anchor - pl * TB;
        const bool valid = anchor < g.n_anchor && tl < g.TBv;
        const int p = g.p0 + pl, t = g.t0 + tl;
        // the periodic part of this output: issued before the accumulator is waited for
        float4 q0 = make_float4(0.f, 0.f, 0.f, 0.f), q1 = q0;
        if (valid) {
          int p36 = p36_0 + pl;
          p36 -= (p36 >= 36) ? 36 : 0;
          const float4* src = reinterpret_cast<const float4*>(a.tab + (((long long)g.b * 36 + p36) * a.T + t) * 8);
          q0 = __ldg(src), q1 = __ldg(src + 1);
        }
        mbar_wait_sleep<160>(&acc_full[grp], n_done & 1);
        fence_after_sync();
        uint32_t v[16];
        tmem_ld16_issue(acc, v);
        tmem_ld_wait16(v);
        fence_before_sync();
        mbar_arrive(&acc_empty[grp]);  // accumulator drained: the issuer may start block j + G
        if (valid) {
          uint64_t o[4];
#pragma unroll
          for (int e = 0; e < 4; ++e)
            o[e] = f2_add(f2_pack(__uint_as_float(v[2 * e]), __uint_as_float(v[2 * e + 1])),
                          f2_pack(__uint_as_float(v[8 + 2 * e]), __uint_as_float(v[9 + 2 * e])));
          o[0] = f2_add(o[0], f2_pack(q0.x, q0.y)), o[1] = f2_add(o[1], f2_pack(q0.z, q0.w));
          o[2] = f2_add(o[2], f2_pack(q1.x, q1.y)), o[3] = f2_add(o[3], f2_pack(q1.z, q1.w));
          uint32_t h[4], l[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float y0, y1;
            f2_unpack(f2_fma(o[e], sc2[e], sh2[e]), y0, y1);
            y0 = fmaxf(y0, kLeakySlope * y0), y1 = fmaxf(y1, kLeakySlope * y1);  // LeakyReLU (slope < 1)
            split_f16x2(y0, y1, h[e], l[e]);
          }
          const uint4 hv = make_uint4(h[0], h[1], h[2], h[3]), lv = make_uint4(l[0], l[1], l[2], l[3]);
          // home position + circular halo copies (rows p +- P, columns t +- T)
          const int row2 = (p < 3) ? p + 3 + a.P : ((p >= a.P - 3) ? p + 3 - a.P : -1);
          const int col2 = (t < 3) ? t + 3 + a.T : ((t >= a.T - 3) ? t + 3 - a.T : -1);
          const long long q00 = ((base + p + 3) * a.Wd + t + 3) * 8;
          *reinterpret_cast<uint4*>(a.out_hi + q00) = hv, *reinterpret_cast<uint4*>(a.out_lo + q00) = lv;
          if (col2 >= 0) {
            const long long q = ((base + p + 3) * a.Wd + col2) * 8;
            *reinterpret_cast<uint4*>(a.out_hi + q) = hv, *reinterpret_cast<uint4*>(a.out_lo + q) = lv;
          }
          if (row2 >= 0) {
            const long long q = ((base + row2) * a.Wd + t + 3) * 8;
            *reinterpret_cast<uint4*>(a.out_hi + q) = hv, *reinterpret_cast<uint4*>(a.out_lo + q) = lv;
            if (col2 >= 0) {
              const long long q2 = ((base + row2) * a.Wd + col2) * 8;
              *reinterpret_cast<uint4*>(a.out_hi + q2) = hv, *reinterpret_cast<uint4*>(a.out_lo + q2) = lv;
            }
          }
        }
      }
      j0 += g.n_mb;
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == ISSUER0) tmem_dealloc(tmem, 64);
}

}  // namespace ake

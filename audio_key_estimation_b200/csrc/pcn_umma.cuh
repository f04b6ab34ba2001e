// pcn_umma.cuh -- tensor-core (tcgen05) kernels of the PitchClassNet forward, eval mode, default channel plan.
//
// Activation layout ("chunk planes"): a tensor of 8*G channels over a (rows x cols) grid is stored as
//   plane[b][g][row][col][8 channels] in fp16, twice: a `hi` plane and a `lo` plane with  x ~= hi + lo  (22 bits).
// One position of one channel group is 16 bytes = one K chunk of the MMA operand layout of umma.cuh, so a tile of the
// tensor copied verbatim into shared memory IS a valid A operand, and a convolution tap (dp, dt) is the same operand
// started (dp * pitch + dt) * 16 bytes further on ("shift-GEMM": no im2col, no register traffic).  Circular padding
// (padding_mode="circular", models.py:230-232) is materialised as halo rows / columns by the producing kernel.
//
// Pitch2Pitch 7x7 circular convolution (models.py:228-234; 64.6 % of the forward's MACs):
//   out[a] = sum_{dp<7} sum_{dt<7} X[a + dp*Wt + dt] . W[dp][dt]        (a = flattened anchor, 8 -> 8 channels)
// is issued as ONE MMA per row tap and 128 anchors:
//   K = 16 : chunk 0 = X_hi[a' + dp*Wt], chunk 1 = X_lo[a' + dp*Wt]            (LBO = distance between the hi and lo planes)
//   N = 112: 7 "phases" f (time tap dt = f) x 8 output channels x {W_hi, W_lo}   (both K chunks meet the same weights)
// and the epilogue adds the phases back together one row apart:  out[a] = sum_f D_f[a + f]  (fixed order, so every
// output is accumulated identically whatever its pitch -> transposition equivariance stays bit exact).
// Cost model (tools/umma_rate.cu): 7 x max(56, 32 + 28) = 420 cycles per 122 anchors, against 7 x (48 + 40) = 616 for the
// earlier two-MMA form (time-tap pairs in K, x_hi and x_lo in separate MMAs).
#pragma once
#include "common.cuh"
#include "umma.cuh"

namespace ake {

constexpr float kWScale = 64.f;     // conv weights are scaled into fp16's normal range; folded back in the epilogue
constexpr int kP2PStride = 122;     // anchors produced per 128-row MMA block (6 rows feed the phase shifts)
constexpr int kP2PRows = 8;         // pitch rows per CTA tile
constexpr int kP2PBufs = 2;         // tile buffers of the load ring
constexpr int kP2PMaxTB = 160;      // frames per CTA tile (upper bound)
#ifndef AKE_P2P_GROUPS
#define AKE_P2P_GROUPS 5
#endif
#ifndef AKE_P2P_ISSUERS
#define AKE_P2P_ISSUERS 2
#endif
constexpr int kP2PGroups = AKE_P2P_GROUPS;  // epilogue groups of 4 warps = accumulator buffers of 64 TMEM columns (56 used)
constexpr int kP2PGroupsGen = AKE_P2P_GROUPS < 4 ? AKE_P2P_GROUPS : 4;    // ... of the tile-generating variant (eight more warps: the register file allows four groups)
constexpr int kP2PGenWarps = 8;      // tile-generator warps of the first conv (instead of the one loader warp)
constexpr int kP2PThreads = 32 * (4 * kP2PGroups + 1 + AKE_P2P_ISSUERS);        // + the loader warp and the MMA-issuer warps
constexpr int kP2PThreadsGen = 32 * (4 * kP2PGroupsGen + kP2PGenWarps + 1);  // + the generator warps and one MMA-issuer warp
constexpr int kP2PMmas = 11;        // MMAs per block: 7 row taps (x_hi + x_lo) . W_hi, then x_hi . W_lo of the row taps in pairs

__device__ __forceinline__ float leaky_f(float v) { return v > 0.f ? v : kLeakySlope * v; }

// Training step (pcn_train_tc.cuh): operand planes of gradients carry an exact power of two that brings the tensor's largest |x|
// (given as its float bit pattern) into [2^13, 2^14) -- the top of fp16's range, so that elements down to 2^-16 of the largest keep
// both halves of the hi/lo split in fp16's normal range; the kernels divide it out again.
__device__ __forceinline__ float tc_scale_of(unsigned maxbits) {
  const int e = (int)((maxbits >> 23) & 0xffu);
  if (e < 14 || e == 255) return 1.f;  // zero / denormal / non-finite: leave the tensor alone
  return __uint_as_float((unsigned)(267 - e) << 23);  // 2^(13 - (e - 127))
}

__device__ __forceinline__ void store_split8(__half* hi_dst, __half* lo_dst, const float (&v)[8]) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) umma::split_f16x2(v[2 * e], v[2 * e + 1], h[e], l[e]);
  *reinterpret_cast<uint4*>(hi_dst) = make_uint4(h[0], h[1], h[2], h[3]);
  *reinterpret_cast<uint4*>(lo_dst) = make_uint4(l[0], l[1], l[2], l[3]);
}

// ---- layer >= 1 input: cat[p (1 ch), tile(up_sixth(pc)) (4 ch)] (models.py:372-383) --------------------------------------
// Never materialised: the first 7x7 convolution generates its input tiles in shared memory (p2p_umma_kernel<true>) from the
// log-CQT and the table below.
// up_sixth of models.py:372-374 as a table: ConvTranspose2d(4,4,(3,1),stride (3,1)) + BN + LeakyReLU maps the 12 pitch classes to
// 36 rows; PitchClass2Pitch (:135-143) then tiles those 36 rows over all pitches, so row p of the conv input is table row p % 36.
// grid (ceil(T / 128), 36, B) -> up[b][p36][t] = 4 channels.
__global__ void __launch_bounds__(128) upsixth_table_kernel(const float* __restrict__ pc, const float* __restrict__ w_up,
                                                            const float* __restrict__ scale, const float* __restrict__ shift,
                                                            float4* __restrict__ up, int T) {
  const int t = blockIdx.x * 128 + threadIdx.x, p36 = blockIdx.y, b = blockIdx.z;
  if (t >= T) return;
  const int c = p36 / 3, r = p36 - 3 * c;
  float x[4], y[4];
#pragma unroll
  for (int ci = 0; ci < 4; ++ci) x[ci] = __ldg(pc + (((long long)b * 4 + ci) * 12 + c) * T + t);
#pragma unroll
  for (int co = 0; co < 4; ++co) {
    float acc = 0.f;
#pragma unroll
    for (int ci = 0; ci < 4; ++ci) acc = fmaf(__ldg(w_up + (ci * 4 + co) * 3 + r), x[ci], acc);
    y[co] = leaky_f(fmaf(acc, __ldg(scale + co), __ldg(shift + co)));
  }
  up[((long long)b * 36 + p36) * T + t] = make_float4(y[0], y[1], y[2], y[3]);
}

// ---- weight image for the 7x7 convolution: [mma 11][chunk 2][n 56][ci 8] fp16 -------------------------------------------
// n = 8 f + co (time tap f, output channel co).  The three products x_hi W_hi + x_lo W_hi + x_hi W_lo all accumulate into ONE
// 56-column accumulator (no hi / lo halves to add in the epilogue, half the TMEM columns per block: more blocks in flight):
//   MMA dp < 7   : A chunks = (x_hi, x_lo) of one position (LBO = plane distance), B chunks = (W_hi[dp], W_hi[dp])
//   MMA 7 + j    : A chunks = x_hi of row taps 2 j and 2 j + 1 (LBO = row pitch),  B chunks = (W_lo[2 j], W_lo[2 j + 1] or 0)
__device__ __forceinline__ void p2p_img_store(__half* __restrict__ img, int dp, int f, int co, int ci, float v) {
  const __half hi = __float2half_rn(v);
  const __half lo = __float2half_rn(v - __half2float(hi));
  const int n = 8 * f + co;
  img[((dp * 2 + 0) * 56 + n) * 8 + ci] = hi;
  img[((dp * 2 + 1) * 56 + n) * 8 + ci] = hi;
  img[(((7 + dp / 2) * 2 + (dp & 1)) * 56 + n) * 8 + ci] = lo;
  if (dp == 6) img[((10 * 2 + 1) * 56 + n) * 8 + ci] = __float2half_rn(0.f);  // row tap 7 does not exist
}
__global__ void p2p_pack_weights_kernel(const float* __restrict__ w, int Cout, int Cin, __half* __restrict__ img) {
  const int n_items = 7 * 56 * 8;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_items; i += gridDim.x * blockDim.x) {
    const int ci = i % 8, fc = (i / 8) % 56, dp = i / 448;
    const int f = fc / 8, co = fc % 8;
    float v = 0.f;
    if (ci < Cin && co < Cout) v = w[(((long long)co * Cin + ci) * 7 + dp) * 7 + f] * kWScale;
    p2p_img_store(img, dp, f, co, ci, v);
  }
}

struct P2PArgs {
  const __half* in_hi;
  const __half* in_lo;   // [B][P+6][Wd][8]
  __half* out_hi;
  __half* out_lo;        // same geometry
  const __half* wimg;    // p2p_pack_weights_kernel image (kP2PWBytes)
  const float* scale;    // 8: eval-mode BN scale (the 1/kWScale factor is applied in the kernel)
  const float* shift;    // 8
  int P, T, Wd;          // Wd = T + 6
  int TB, n_ttiles;      // frames per tile, tiles along time
  int n_rtiles, n_tiles; // pitch-row tiles per clip; tiles in the launch (B * n_rtiles * n_ttiles)
  // GEN = true (first conv of the stack): the input tile cat[p, tile(up_sixth(pc))] (models.py:372-383) is generated in shared
  // memory from the log-CQT and the 36-row up-sampled table instead of being read from operand planes
  const float* mel;      // (B, 1, P, T)
  const float4* up;      // (B, 36, T) x 4 channels: act(bn(ConvTranspose2d(pc))) per pitch class third (upsixth_table_kernel)
  // RAW = true (periodic part of the first conv, pcn_p2p1.cuh): the accumulators, before BN, as fp32 (B, P, T, 8)
  float* raw_out;
};

constexpr uint32_t kP2PWBytes = kP2PMmas * 2 * 56 * 16;
constexpr uint32_t kP2PPubBytes = 2 * 3 * 21 * 32;  // per epilogue group: [parity 2][warp 1..3][phase f = 1..6: f lanes][co 8] floats

__host__ __device__ inline uint32_t p2p_plane_positions(int Wt) { return (uint32_t)((kP2PRows + 6) * Wt + 136); }
__host__ __device__ inline size_t p2p_smem_bytes(int Wt) {
  return (size_t)2 * kP2PBufs * p2p_plane_positions(Wt) * 16 + kP2PWBytes + kP2PGroups * kP2PPubBytes;
}

// Persistent CTA (one per SM), warp-specialised:
//   warp 4G     : loader -- lands the next tile's (rows + 6) x Wt positions of the hi and lo planes in the free tile buffer
//                 with one bulk async copy per row and plane (two tile buffers: the load of tile k+1 overlaps the MMAs of k)
//   warp 4G + 1 : one lane issues 7 MMAs per 128-anchor block into accumulator buffer (block % G) of TMEM
//   warps 0..4G-1: G epilogue groups of 4 warps; group g drains blocks g, g + G, ... (thread = TMEM lane = anchor):
//                 phase realignment (shuffles + a 6-lane hand-over between neighbouring warps), BN + LeakyReLU, fp16
//                 hi/lo split, stores of the home position and of the circular halo copies.
template <bool GEN, bool RAW = false>
__global__ void __launch_bounds__(GEN ? kP2PThreadsGen : kP2PThreads, 1) p2p_umma_kernel(const P2PArgs a) {
  using namespace umma;
  constexpr int G = GEN ? kP2PGroupsGen : kP2PGroups;
  // MMA-issuer warps, taking alternate blocks: one warp issues an MMA of N <= 64 every ~61 cycles, the tensor pipe executes one
  // every ~46 (tools/umma_rate.cu).  The accumulator buffers are shared between the issuers (G is odd), so a buffer's "drained"
  // hand-over alternates between TWO mbarriers: a waiter two uses ahead of a single barrier would pass on a stale parity.
  constexpr int NI = GEN ? 1 : AKE_P2P_ISSUERS;
  constexpr uint32_t TMEM_COLS = 64 * G <= 256 ? 256 : 512;
  constexpr int NLOAD = GEN ? kP2PGenWarps : 1, ISSUER = 4 * G + NLOAD;  // warp roles: [0, 4G) epilogue, [4G, 4G + NLOAD) tile producers, ISSUER
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t w_bar, full_bar[kP2PBufs], empty_bar[kP2PBufs], acc_full[G], acc_empty[G][2];
  static_assert(G <= kP2PGroups, "shared-memory sizing assumes at most kP2PGroups hand-over buffers");
  __shared__ uint32_t tmem_slot;
  __shared__ float s_scale[8], s_shift[8];

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const int Wt = a.TB + 6;                          // tile pitch (positions per staged row)
  const uint32_t plane = p2p_plane_positions(Wt) * 16;
  uint8_t* s_w = smem + 2 * kP2PBufs * plane;
  uint8_t* s_pub = s_w + kP2PWBytes;
  const int tiles_per_clip = a.n_rtiles * a.n_ttiles;

  struct Geom {
    int b, p0, t0, PB, TBv, cols_in, n_anchor, n_mb;
  };
  // tile -> (clip, row tile, time tile) by multiply-high: the issuer warp decodes a tile between two MMA blocks, and a
  // hardware-less integer division is ~150 dependent instructions there (x / d == umulhi(x, 2^32 / d + 1) for x * d < 2^32)
  const uint32_t tpc_magic = 0xFFFFFFFFu / (uint32_t)tiles_per_clip + 1, ntt_magic = 0xFFFFFFFFu / (uint32_t)a.n_ttiles + 1;
  auto geom = [&](int tile) {
    Geom g;
    g.b = tiles_per_clip == 1 ? tile : (int)__umulhi((uint32_t)tile, tpc_magic);
    const int r = tile - g.b * tiles_per_clip;
    const int rtile = a.n_ttiles == 1 ? r : (int)__umulhi((uint32_t)r, ntt_magic), ttile = r - rtile * a.n_ttiles;
    g.p0 = rtile * kP2PRows, g.t0 = ttile * a.TB;
    g.PB = min(kP2PRows, a.P - g.p0);          // valid output rows of this tile
    g.TBv = min(a.TB, a.T - g.t0);             // valid output frames
    g.cols_in = min(Wt, a.Wd - g.t0);          // the last time tile may be narrower than Wt
    g.n_anchor = g.PB * Wt;
    g.n_mb = (g.n_anchor + kP2PStride - 1) / kP2PStride;
    return g;
  };

  if (warp == ISSUER) tmem_alloc(&tmem_slot, TMEM_COLS);
  if (threadIdx.x == 0) {
    mbar_init(&w_bar, 1);
    for (int i = 0; i < kP2PBufs; ++i) mbar_init(&full_bar[i], GEN ? 32 * NLOAD : 1), mbar_init(&empty_bar[i], NI);
    for (int i = 0; i < G; ++i) mbar_init(&acc_full[i], 1), mbar_init(&acc_empty[i][0], 128), mbar_init(&acc_empty[i][1], 128);
    mbar_init_fence();
  }
  if (threadIdx.x < 8) s_scale[threadIdx.x] = a.scale[threadIdx.x] * (1.f / kWScale), s_shift[threadIdx.x] = a.shift[threadIdx.x];
  // Positions the bulk copies never write (tail padding, the column gap of a narrow last time tile) only feed anchors that
  // are discarded, but they must hold finite values once: a NaN bit pattern would be harmless too, zero is tidier.
  for (uint32_t i = threadIdx.x; i < 2 * kP2PBufs * plane / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_slot;

  if (warp >= 4 * G && warp < ISSUER) {
    if constexpr (GEN) {
      // ------------------------------------------------------------ tile generators (first conv): position q of the tile =
      // [log-CQT (1 ch) | up-sampled pitch-class features (4 ch) | 0 0 0], circular in pitch and time, split into fp16 hi / lo
      const int gt = threadIdx.x - 128 * G;
      if (gt == 0) {
        mbar_arrive_expect_tx(&w_bar, kP2PWBytes);
        bulk_g2s(s_w, a.wimg, kP2PWBytes, &w_bar);
      }
      const int gw = warp - 4 * G;  // generator warp: rows gw, gw + NLOAD, ... of the tile, 32 consecutive columns per round
      int k = 0;
      for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++k) {
        const int s = k % kP2PBufs;
        const Geom g = geom(tile);
        mbar_wait_relaxed(&empty_bar[s], ((k / kP2PBufs) & 1) ^ 1);
        uint4* d_hi = reinterpret_cast<uint4*>(smem + (size_t)s * 2 * plane);
        uint4* d_lo = reinterpret_cast<uint4*>(smem + (size_t)s * 2 * plane + plane);
        const float* mel_b = a.mel + (long long)g.b * a.P * a.T;
        const float4* up_b = a.up + (long long)g.b * 36 * a.T;
        for (int row = gw; row < g.PB + 6; row += NLOAD) {
          // per row: the (circular) pitch and its table row are fixed; columns beyond the planes' width (narrow last time
          // tile) only feed discarded anchors and are skipped
          int p = g.p0 + row - 3;
          p += (p < 0) ? a.P : 0, p -= (p >= a.P) ? a.P : 0;
          const float* mel_r = mel_b + (long long)p * a.T;
          const float4* up_r = up_b + (p % 36) * a.T;
          uint4* r_hi = d_hi + row * Wt;
          uint4* r_lo = d_lo + row * Wt;
          constexpr int kBatch = 5;  // 5 x 32 columns >= kP2PMaxTB + 6: all loads of a row in flight before the first split
          float m0[kBatch];
          float4 u[kBatch];
#pragma unroll
          for (int i = 0; i < kBatch; ++i) {
            const int col = lane + 32 * i;
            m0[i] = 0.f, u[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (col < g.cols_in) {
              int t = g.t0 + col - 3;
              t += (t < 0) ? a.T : 0, t -= (t >= a.T) ? a.T : 0;
              m0[i] = __ldg(mel_r + t);
              u[i] = __ldg(up_r + t);
            }
          }
#pragma unroll
          for (int i = 0; i < kBatch; ++i) {
            const int col = lane + 32 * i;
            if (col < g.cols_in) {
              uint32_t h[3], l[3];
              split_f16x2(m0[i], u[i].x, h[0], l[0]);
              split_f16x2(u[i].y, u[i].z, h[1], l[1]);
              split_f16x2(u[i].w, 0.f, h[2], l[2]);
              r_hi[col] = make_uint4(h[0], h[1], h[2], 0), r_lo[col] = make_uint4(l[0], l[1], l[2], 0);
            }
          }
          for (int col = lane + 32 * kBatch; col < g.cols_in; col += 32) {  // (wider tiles than kP2PMaxTB allows: not reached)
            int t = g.t0 + col - 3;
            t += (t < 0) ? a.T : 0, t -= (t >= a.T) ? a.T : 0;
            const float mm = __ldg(mel_r + t);
            const float4 uu = __ldg(up_r + t);
            uint32_t h[3], l[3];
            split_f16x2(mm, uu.x, h[0], l[0]);
            split_f16x2(uu.y, uu.z, h[1], l[1]);
            split_f16x2(uu.w, 0.f, h[2], l[2]);
            r_hi[col] = make_uint4(h[0], h[1], h[2], 0), r_lo[col] = make_uint4(l[0], l[1], l[2], 0);
          }
        }
        fence_proxy_async();
        mbar_arrive(&full_bar[s]);
      }
    } else {
      // ------------------------------------------------------------ loader
      if (lane == 0) {
        mbar_arrive_expect_tx(&w_bar, kP2PWBytes);
        bulk_g2s(s_w, a.wimg, kP2PWBytes, &w_bar);
      }
      int k = 0;
      for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++k) {
        const int s = k % kP2PBufs;
        const Geom g = geom(tile);
        const int rows_in = g.PB + 6;
        const uint32_t row_bytes = (uint32_t)g.cols_in * 16;
        mbar_wait_relaxed(&empty_bar[s], ((k / kP2PBufs) & 1) ^ 1);
        if (lane == 0) mbar_arrive_expect_tx(&full_bar[s], 2u * rows_in * row_bytes);
        __syncwarp();
        uint8_t* d_hi = smem + (size_t)s * 2 * plane;
        for (int rr = lane; rr < rows_in; rr += 32) {
          const long long src = (((long long)g.b * (a.P + 6) + g.p0 + rr) * a.Wd + g.t0) * 8;
          bulk_g2s(d_hi + (size_t)rr * Wt * 16, a.in_hi + src, row_bytes, &full_bar[s]);
          bulk_g2s(d_hi + plane + (size_t)rr * Wt * 16, a.in_lo + src, row_bytes, &full_bar[s]);
        }
        // The tile after this one cannot be copied before a buffer frees up (one tile time from now): pull it into L2
        // meanwhile, so that copy pays the L2 latency instead of the HBM latency.
        const int tile2 = tile + (kP2PBufs - 1) * (int)gridDim.x;
        if (tile2 < a.n_tiles) {
          const Geom g2 = geom(tile2);
          const uint32_t rb2 = (uint32_t)g2.cols_in * 16;
          for (int rr = lane; rr < g2.PB + 6; rr += 32) {
            const long long src = (((long long)g2.b * (a.P + 6) + g2.p0 + rr) * a.Wd + g2.t0) * 8;
            bulk_prefetch_l2(a.in_hi + src, rb2);
            bulk_prefetch_l2(a.in_lo + src, rb2);
          }
        }
      }
    }
  } else if (warp >= ISSUER) {
    // ------------------------------------------------------------ MMA issuers (converged warps, one elected lane issues)
    const uint32_t my = (uint32_t)(warp - ISSUER);  // this warp issues the blocks j with j % NI == my
    const uint64_t A_DESC = desc_hi(plane);                     // chunk 1 = the x_lo plane at the same position
    const uint64_t A_DESC_ROW = desc_hi((uint32_t)Wt * 16);     // chunk 1 = the x_hi plane one pitch row further (the next row tap)
    constexpr uint64_t B_DESC = desc_hi(56 * 16);               // chunk stride: 56 rows x 16 B
    constexpr uint32_t IDESC = idesc_f16(56);
    constexpr uint32_t W_MMA = 2 * 56 * 16;                     // bytes of one MMA's weight block
    const uint32_t w0 = smem_u32(s_w);
    mbar_wait(&w_bar, 0);
    int k = 0;
    uint32_t j = 0;  // block counter of this CTA (all roles walk the same sequence)
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++k) {
      const int s = k % kP2PBufs;
      const Geom g = geom(tile);
      const uint32_t hi0 = smem_u32(smem + (size_t)s * 2 * plane);
      mbar_wait(&full_bar[s], (k / kP2PBufs) & 1);
      for (int m = 0; m < g.n_mb; ++m, ++j) {
        if (NI > 1 && j % NI != my) continue;
        const uint32_t buf = j % G, use = j / G;  // the buffer's previous use (use - 1) must have been drained
        if (use > 0) mbar_wait(&acc_empty[buf][(use - 1) & 1], ((use - 1) >> 1) & 1);
        fence_after_sync();
        const uint32_t d = tmem + buf * 64;
        const uint32_t a_off = hi0 + (uint32_t)(m * kP2PStride) * 16;
        if (elect_one()) {
#pragma unroll
          for (int dp = 0; dp < 7; ++dp)  // (x_hi + x_lo) . W_hi of row tap dp
            mma_f16(d, make_desc(A_DESC, a_off + (uint32_t)(dp * Wt) * 16), make_desc(B_DESC, w0 + dp * W_MMA), IDESC, dp ? 1u : 0u);
#pragma unroll
          for (int j = 0; j < 4; ++j)     // x_hi . W_lo of the row taps 2 j, 2 j + 1 (tap 7: zero weights on whatever follows the tile)
            mma_f16(d, make_desc(A_DESC_ROW, a_off + (uint32_t)(2 * j * Wt) * 16), make_desc(B_DESC, w0 + (7 + j) * W_MMA), IDESC, 1u);
          commit(&acc_full[buf]);
        }
        __syncwarp();
      }
      if (elect_one()) commit(&empty_bar[s]);  // the tile buffer is free once every issuer's MMAs of this tile have read it
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------ epilogue: thread = TMEM lane = anchor row
    const int grp = warp >> 2, wq = warp & 3, tid = threadIdx.x & 127;
    const uint32_t acc = tmem + ((uint32_t)(wq * 32) << 16) + grp * 64;
    float4* pub = reinterpret_cast<float4*>(s_pub + grp * kP2PPubBytes);
    const uint32_t wt_magic = 0xFFFFFFFFu / (uint32_t)Wt + 1;  // anchor / Wt == umulhi(anchor, magic) for anchor * Wt < 2^32
    uint64_t sc2[4], sh2[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) sc2[e] = f2_pack(s_scale[2 * e], s_scale[2 * e + 1]), sh2[e] = f2_pack(s_shift[2 * e], s_shift[2 * e + 1]);
    uint32_t n_done = 0;   // blocks this group has drained
    uint32_t j = grp, j0 = 0;  // next block of this group; first block of the current tile
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
      const Geom g = geom(tile);
      const long long base = (long long)g.b * (a.P + 6);
      for (; j < j0 + g.n_mb; j += G, ++n_done) {
        const int m = (int)(j - j0);
        mbar_wait_relaxed(&acc_full[grp], n_done & 1);
        fence_after_sync();
        // phase f = time tap f: out[a] = sum_f D_f[a + f], D_f of this thread's row = columns [8 f, 8 f + 8).  Rows a + f live f
        // lanes further on: warp shuffles for lane + f < 32, the first lanes of the next warp (through shared memory)
        // otherwise -- ascending f for every anchor either way.
        float4* pub_w = pub + (n_done & 1) * (3 * 21 * 2);
        uint64_t o[4];
        uint32_t r[2][8];
        tmem_ld8_issue(acc, r[0]);
#pragma unroll
        for (int f = 0; f < 7; ++f) {
          uint32_t(&v)[8] = r[f & 1];
          tmem_ld_wait8(v);
          if (f < 6) tmem_ld8_issue(acc + 8 * (f + 1), r[(f + 1) & 1]);
          uint64_t u[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) u[e] = f2_pack(__uint_as_float(v[2 * e]), __uint_as_float(v[2 * e + 1]));
          if (f == 0) {
#pragma unroll
            for (int e = 0; e < 4; ++e) o[e] = u[e];
          } else {
            float x[8];
#pragma unroll
            for (int e = 0; e < 4; ++e) f2_unpack(u[e], x[2 * e], x[2 * e + 1]);
            if (wq > 0 && lane < f) {
              float4* dst = pub_w + ((wq - 1) * 21 + f * (f - 1) / 2 + lane) * 2;
              dst[0] = make_float4(x[0], x[1], x[2], x[3]), dst[1] = make_float4(x[4], x[5], x[6], x[7]);
            }
#pragma unroll
            for (int c = 0; c < 8; ++c) x[c] = __shfl_down_sync(0xffffffffu, x[c], f);
            // o += x on the lanes whose source lane exists: x * 1 + o is the same single rounding as x + o, x * 0 + o = o
            // (an out-of-range shuffle returns the lane's own, finite, value); predicated FADD2s compile to FADD2 + 2 SEL
            const float mk = (lane + f < 32) ? 1.f : 0.f;
            const uint64_t mk2 = f2_pack(mk, mk);
#pragma unroll
            for (int e = 0; e < 4; ++e) o[e] = f2_fma(f2_pack(x[2 * e], x[2 * e + 1]), mk2, o[e]);
          }
        }
        fence_before_sync();
        mbar_arrive(&acc_empty[grp][n_done & 1]);  // accumulator drained: block j + G may be issued
        asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
        if (wq < 3 && lane >= 26) {
#pragma unroll
          for (int f = 1; f < 7; ++f) {
            const bool take = lane + f >= 32;
            const float4* src = pub_w + (wq * 21 + f * (f - 1) / 2 + (take ? lane + f - 32 : 0)) * 2;
            const float4 x0 = src[0], x1 = src[1];
            const float mk = take ? 1.f : 0.f;
            const uint64_t mk2 = f2_pack(mk, mk);
            o[0] = f2_fma(f2_pack(x0.x, x0.y), mk2, o[0]), o[1] = f2_fma(f2_pack(x0.z, x0.w), mk2, o[1]);
            o[2] = f2_fma(f2_pack(x1.x, x1.y), mk2, o[2]), o[3] = f2_fma(f2_pack(x1.z, x1.w), mk2, o[3]);
          }
        }
        const int anchor = m * kP2PStride + tid;
        if (tid < kP2PStride && anchor < g.n_anchor) {
          const int pl = (int)__umulhi((uint32_t)anchor, wt_magic), tl = anchor - pl * Wt;
          if (RAW && tl < g.TBv) {
            float y[8];
#pragma unroll
            for (int e = 0; e < 4; ++e) f2_unpack(o[e], y[2 * e], y[2 * e + 1]);
            float4* dst = reinterpret_cast<float4*>(a.raw_out + (((long long)g.b * a.P + g.p0 + pl) * a.T + g.t0 + tl) * 8);
            dst[0] = make_float4(y[0], y[1], y[2], y[3]), dst[1] = make_float4(y[4], y[5], y[6], y[7]);
          } else if (tl < g.TBv) {
            uint32_t h[4], l[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float y0, y1;
              f2_unpack(f2_fma(o[e], sc2[e], sh2[e]), y0, y1);
              y0 = fmaxf(y0, kLeakySlope * y0), y1 = fmaxf(y1, kLeakySlope * y1);  // LeakyReLU (slope < 1)
              split_f16x2(y0, y1, h[e], l[e]);
            }
            const uint4 hv = make_uint4(h[0], h[1], h[2], h[3]), lv = make_uint4(l[0], l[1], l[2], l[3]);
            const int p = g.p0 + pl, t = g.t0 + tl;
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
      }
      j0 += g.n_mb;
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == ISSUER) tmem_dealloc(tmem, TMEM_COLS);
}

// ---- pool_semi conv (3x3, stride (3,1), time-circular) + BN + LeakyReLU + octave max pool, from chunk planes ------
// models.py:337-339, 386-389: s[co, j, t] = act(bn(sum_{ci,dp<3,dt<3} W x[ci, 3j+dp, (t+dt-1) mod T])), pc[co,c,t] = max_o s[co, c+12o, t].
// One thread per (c, column).  The result is concatenated behind the previous pitch-class features (models.py:392) and
// written as 16-channel chunk planes [B][2][23][T+6][8] for the equivariant convolutions: channels [pc (4) | pooled (8) |
// zero (4)], wrap rows 12..22 = rows 0..10 (models.py:27-28), zero halo columns (the "same" padding of models.py:45-47).
struct SemiArgs {
  const __half* in_hi;
  const __half* in_lo;  // [B][P+6][Wd][8]
  const float* w;       // (8 co, 8 ci, 3, 3)
  const float* scale;
  const float* shift;
  const float* pc_prev; // (B, 4, 12, T) fp32
  __half* out_hi;
  __half* out_lo;       // [B][2][23][Wd][8]
  int B, P, T, Wd;
};

__global__ void __launch_bounds__(128) semitone_pool_chunks_kernel(const SemiArgs a) {
  __shared__ float w[8 * 8 * 9];  // [dp][dt][ci][co]
  __shared__ float sc[8], sh[8];
  for (int i = threadIdx.x; i < 576; i += blockDim.x) {
    const int co = i % 8, ci = (i / 8) % 8, tap = i / 64;
    w[i] = a.w[(co * 8 + ci) * 9 + tap];
  }
  if (threadIdx.x < 8) sc[threadIdx.x] = a.scale[threadIdx.x], sh[threadIdx.x] = a.shift[threadIdx.x];
  __syncthreads();
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  const int c = blockIdx.y, b = blockIdx.z;
  if (col >= a.Wd) return;
  const int t = col - 3;
  float g0[8], g1[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) g0[e] = 0.f, g1[e] = 0.f;
  if (t >= 0 && t < a.T) {
    const int n_oct = a.P / 36;
    float best[8];
#pragma unroll
    for (int co = 0; co < 8; ++co) best[co] = -INFINITY;
    for (int o = 0; o < n_oct; ++o) {
      float acc[8];
#pragma unroll
      for (int co = 0; co < 8; ++co) acc[co] = 0.f;
      const int row0 = 3 * (c + 12 * o) + 3;  // halo'd row of tap dp = 0
#pragma unroll
      for (int dp = 0; dp < 3; ++dp) {
#pragma unroll
        for (int dt = 0; dt < 3; ++dt) {
          const long long q = (((long long)b * (a.P + 6) + row0 + dp) * a.Wd + t + 2 + dt) * 8;  // column (t + dt - 1) + 3
          const uint4 hv = __ldg(reinterpret_cast<const uint4*>(a.in_hi + q));
          const uint4 lv = __ldg(reinterpret_cast<const uint4*>(a.in_lo + q));
          const uint32_t hw[4] = {hv.x, hv.y, hv.z, hv.w}, lw[4] = {lv.x, lv.y, lv.z, lv.w};
          float x[8];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hw[e]));
            const float2 lf = __half22float2(*reinterpret_cast<const __half2*>(&lw[e]));
            x[2 * e] = hf.x + lf.x, x[2 * e + 1] = hf.y + lf.y;
          }
          const float* wt = w + (dp * 3 + dt) * 64;
#pragma unroll
          for (int ci = 0; ci < 8; ++ci)
#pragma unroll
            for (int co = 0; co < 8; ++co) acc[co] = fmaf(wt[ci * 8 + co], x[ci], acc[co]);
        }
      }
#pragma unroll
      for (int co = 0; co < 8; ++co) best[co] = fmaxf(best[co], leaky_f(fmaf(acc[co], sc[co], sh[co])));
    }
#pragma unroll
    for (int ci = 0; ci < 4; ++ci) g0[ci] = __ldg(a.pc_prev + (((long long)b * 4 + ci) * 12 + c) * a.T + t);
#pragma unroll
    for (int e = 0; e < 4; ++e) g0[4 + e] = best[e], g1[e] = best[4 + e];
  }
  const long long q0 = ((((long long)b * 2 + 0) * 23 + c) * a.Wd + col) * 8, q1 = ((((long long)b * 2 + 1) * 23 + c) * a.Wd + col) * 8;
  store_split8(a.out_hi + q0, a.out_lo + q0, g0);
  store_split8(a.out_hi + q1, a.out_lo + q1, g1);
  if (c < 11) {
    const long long w0 = q0 + (long long)12 * a.Wd * 8, w1 = q1 + (long long)12 * a.Wd * 8;
    store_split8(a.out_hi + w0, a.out_lo + w0, g0);
    store_split8(a.out_hi + w1, a.out_lo + w1, g1);
  }
}

// ---- pool_semi + octave pool on tensor cores ---------------------------------------------------------------------------
// Same operator as semitone_pool_chunks_kernel (models.py:337-339, 386-392), as a shift-GEMM: for pitch class c and octave o
// the three input rows 3 (c + 12 o) + dp are three row taps, K = 16 = [x_hi | x_lo] chunks, N = 48 = 3 time-tap phases x
// {W_hi 8, W_lo 8}; the accumulators of ALL octaves of a work item (b, c, time tile) sit side by side in TMEM (48 columns
// each), and the epilogue re-aligns the phases, applies BN + LeakyReLU and takes the octave max in registers.  The kernel
// streams the last Pitch2Pitch output once (HBM-bound) instead of spending 4608 FFMAs per output on it.
// Persistent CTA per SM: 4 epilogue groups of 4 warps (thread = anchor = frame; group g drains the octaves g, g + 4, ... of
// every item, so the ~1400-instruction drain of an item is four independent chains instead of one; the groups' octave maxima
// meet in shared memory), then the loader warp (triple-buffered tiles) and the MMA-issuer warp.
constexpr int kSemiMaxTB = 126;     // frames per tile: TB + 2 <= 128 anchors
constexpr int kSemiMaxOct = 10;     // 48 TMEM columns per octave
constexpr int kSemiGroups = 4;      // epilogue groups
constexpr int kSemiOctPerGroup = (kSemiMaxOct + kSemiGroups - 1) / kSemiGroups;
constexpr int kSemiThreads = 32 * (4 * kSemiGroups + 2);
constexpr int kSemiBufs = 3;        // tile buffers: two loads in flight while one tile is multiplied (a single 64 KB load in flight per SM
                                    // runs at its ~3 us loaded latency, far below the HBM rate)
constexpr uint32_t kSemiWBytes = 3 * 2 * 48 * 16;

// [dp 3][chunk 2][n 48][ci 8]: n = 16 dt + j; j < 8: W_hi of output channel j, j >= 8: W_lo; both chunks hold the same weights
__global__ void semi_pack_weights_kernel(const float* __restrict__ w, __half* __restrict__ img) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 3 * 3 * 8 * 8) return;
  const int ci = i % 8, co = (i / 8) % 8, dt = (i / 64) % 3, dp = i / 192;
  const float v = w[((co * 8 + ci) * 3 + dp) * 3 + dt] * kWScale;
  const __half hi = __float2half_rn(v);
  const __half lo = __float2half_rn(v - __half2float(hi));
  for (int c = 0; c < 2; ++c) {
    img[((dp * 2 + c) * 48 + 16 * dt + co) * 8 + ci] = hi;
    img[((dp * 2 + c) * 48 + 16 * dt + 8 + co) * 8 + ci] = lo;
  }
}

struct SemiUmmaArgs {
  const __half* in_hi;
  const __half* in_lo;  // [B][P+6][Wd][8]
  const __half* wimg;   // semi_pack_weights_kernel image
  const float* scale;
  const float* shift;   // 8
  const float* pc_prev; // (B, 4, 12, T) fp32
  __half* out_hi;
  __half* out_lo;       // [B][2][23][Wd][8]: channels [pc_prev 4 | pooled 8 | 0 x 4]
  int B, P, T, Wd, n_oct, TB, n_ttiles, n_items;
};

__host__ __device__ inline uint32_t semi_plane_positions(int n_oct, int Wt) { return (uint32_t)(3 * n_oct * Wt + 136); }
__host__ __device__ inline size_t semi_smem_bytes(int n_oct, int Wt) {
  return (size_t)kSemiBufs * 2 * semi_plane_positions(n_oct, Wt) * 16 + kSemiWBytes + (size_t)kSemiMaxOct * 3 * 3 * 32 +
         (size_t)kSemiGroups * 128 * 32;  // + the octave maxima the groups exchange
}

__global__ void __launch_bounds__(kSemiThreads, 1) semi_umma_kernel(const SemiUmmaArgs a) {
  using namespace umma;
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t w_bar, full_bar[kSemiBufs], empty_bar[kSemiBufs], acc_full, acc_empty;
  __shared__ uint32_t tmem_slot;
  __shared__ float s_scale[8], s_shift[8];

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const int Wt = a.TB + 2, n_oct = a.n_oct, rows_in = 3 * n_oct;
  const uint32_t plane = semi_plane_positions(n_oct, Wt) * 16;  // a tile buffer holds [hi][lo]
  uint8_t* s_w = smem + 2 * kSemiBufs * plane;
  float4* pub = reinterpret_cast<float4*>(s_w + kSemiWBytes);    // [octave][warp 1..3][slot 3 = (f 1: lane 0), (f 2: lanes 0, 1)][8 floats]
  float4* s_best = pub + kSemiMaxOct * 3 * 3 * 2;                // [group][thread 128][8 floats]
  constexpr int G = kSemiGroups, LOADER = 4 * G, ISSUER = 4 * G + 1;
  const int per_clip = 12 * a.n_ttiles;
  const uint32_t pc_magic = 0xFFFFFFFFu / (uint32_t)per_clip + 1, tt_magic = 0xFFFFFFFFu / (uint32_t)a.n_ttiles + 1;
  auto decode = [&](int item, int& b, int& c, int& t0) {
    b = (int)__umulhi((uint32_t)item, pc_magic);
    const int r = item - b * per_clip;
    c = a.n_ttiles == 1 ? r : (int)__umulhi((uint32_t)r, tt_magic);
    t0 = (r - c * a.n_ttiles) * a.TB;
  };

  if (warp == ISSUER) tmem_alloc(&tmem_slot, 512);
  if (threadIdx.x == 0) {
    mbar_init(&w_bar, 1);
    for (int i = 0; i < kSemiBufs; ++i) mbar_init(&full_bar[i], 1), mbar_init(&empty_bar[i], 1);
    mbar_init(&acc_full, 1), mbar_init(&acc_empty, 128 * G);
    mbar_init_fence();
  }
  if (threadIdx.x < 8) s_scale[threadIdx.x] = a.scale[threadIdx.x] * (1.f / kWScale), s_shift[threadIdx.x] = a.shift[threadIdx.x];
  for (uint32_t i = threadIdx.x; i < 2 * kSemiBufs * plane / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_slot;

  if (warp == LOADER) {
    // ------------------------------------------------------------ loader: rows 3 (c + 12 o) + dp of the halo'd planes
    if (lane == 0) {
      mbar_arrive_expect_tx(&w_bar, kSemiWBytes);
      bulk_g2s(s_w, a.wimg, kSemiWBytes, &w_bar);
    }
    int k = 0;
    for (int item = blockIdx.x; item < a.n_items; item += gridDim.x, ++k) {
      const int s = k % kSemiBufs;
      int b, c, t0;
      decode(item, b, c, t0);
      const int cols_in = min(Wt, a.Wd - (t0 + 2));
      const uint32_t row_bytes = (uint32_t)cols_in * 16;
      mbar_wait_relaxed(&empty_bar[s], ((k / kSemiBufs) & 1) ^ 1);
      if (lane == 0) mbar_arrive_expect_tx(&full_bar[s], 2u * rows_in * row_bytes);
      __syncwarp();
      uint8_t* dst = smem + (size_t)s * 2 * plane;
      for (int r = lane; r < rows_in; r += 32) {
        const int o = r / 3, dp = r - 3 * o;
        const long long src = (((long long)b * (a.P + 6) + 3 * (c + 12 * o) + dp + 3) * a.Wd + t0 + 2) * 8;
        bulk_g2s(dst + (size_t)r * Wt * 16, a.in_hi + src, row_bytes, &full_bar[s]);
        bulk_g2s(dst + plane + (size_t)r * Wt * 16, a.in_lo + src, row_bytes, &full_bar[s]);
      }
    }
  } else if (warp == ISSUER) {
    // ------------------------------------------------------------ MMA issuer (converged warp, one elected lane issues)
    const uint64_t A_DESC = desc_hi(plane);        // chunk 1 = the x_lo plane at the same position
    constexpr uint64_t B_DESC = desc_hi(48 * 16);  // chunk stride: 48 rows x 16 B
    constexpr uint32_t IDESC = idesc_f16(48);
    const uint32_t w0 = smem_u32(s_w);
    mbar_wait(&w_bar, 0);
    int k = 0;
    for (int item = blockIdx.x; item < a.n_items; item += gridDim.x, ++k) {
      const int s = k % kSemiBufs;
      const uint32_t hi0 = smem_u32(smem + (size_t)s * 2 * plane);
      mbar_wait(&full_bar[s], (k / kSemiBufs) & 1);
      mbar_wait(&acc_empty, (k & 1) ^ 1);
      fence_after_sync();
      if (elect_one()) {
        for (int o = 0; o < n_oct; ++o) {
#pragma unroll
          for (int dp = 0; dp < 3; ++dp)
            mma_f16(tmem + o * 48, make_desc(A_DESC, hi0 + (uint32_t)((3 * o + dp) * Wt) * 16), make_desc(B_DESC, w0 + dp * (2 * 48 * 16)), IDESC,
                    dp ? 1u : 0u);
        }
        commit(&acc_full);
        commit(&empty_bar[s]);
      }
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------ epilogue: thread = TMEM lane = frame t0 + tid; group grp owns
    // the octaves grp, grp + G, ...
    const int grp = warp >> 2, wq = warp & 3, tid = threadIdx.x & 127;
    const uint32_t acc = tmem + ((uint32_t)(wq * 32) << 16);
    int k = 0;
    for (int item = blockIdx.x; item < a.n_items; item += gridDim.x, ++k) {
      int b, c, t0;
      decode(item, b, c, t0);
      // the previous layer's pitch-class features of this frame (channels 0..3 of the output): loaded before the wait so that
      // their latency hides behind the MMAs
      const int t = t0 + tid;
      float pcv[4] = {0.f, 0.f, 0.f, 0.f};
      if ((grp & 1) == 0 && tid < a.TB && t < a.T) {
#pragma unroll
        for (int ci = 0; ci < 4; ++ci) pcv[ci] = __ldg(a.pc_prev + (((long long)b * 4 + ci) * 12 + c) * a.T + t);
      }
      mbar_wait_relaxed(&acc_full, k & 1);
      fence_after_sync();
      // pass 1: per octave, D_dt = columns [16 dt, 16 dt + 8) + [16 dt + 8, 16 dt + 16); s = D_0[a] + D_1[a + 1] + D_2[a + 2] with
      // the in-warp part by shuffles; lanes 0, 1 of warps 1..3 publish what lanes 30, 31 of the previous warp still need
      uint64_t sacc[kSemiOctPerGroup][4];
#pragma unroll
      for (int oi = 0; oi < kSemiOctPerGroup; ++oi) {
        const int o = grp + G * oi;
        if (o < n_oct) {
          uint32_t v0[16], v1[16], v2[16];
          tmem_ld16_issue(acc + o * 48, v0);
          tmem_ld16_issue(acc + o * 48 + 16, v1);
          tmem_ld16_issue(acc + o * 48 + 32, v2);
          tmem_ld_wait16(v0), tmem_ld_wait16(v1), tmem_ld_wait16(v2);
#pragma unroll
          for (int e = 0; e < 4; ++e)
            sacc[oi][e] = f2_add(f2_pack(__uint_as_float(v0[2 * e]), __uint_as_float(v0[2 * e + 1])),
                                 f2_pack(__uint_as_float(v0[8 + 2 * e]), __uint_as_float(v0[9 + 2 * e])));
#pragma unroll
          for (int f = 1; f < 3; ++f) {
            const uint32_t(&v)[16] = f == 1 ? v1 : v2;
            float x[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) x[e] = __uint_as_float(v[e]) + __uint_as_float(v[8 + e]);
            if (wq > 0 && lane < f) {
              float4* dst = pub + ((o * 3 + (wq - 1)) * 3 + (f - 1) + lane) * 2;
              dst[0] = make_float4(x[0], x[1], x[2], x[3]), dst[1] = make_float4(x[4], x[5], x[6], x[7]);
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) x[e] = __shfl_down_sync(0xffffffffu, x[e], f);
            const float mk = (lane + f < 32) ? 1.f : 0.f;  // masked FMA = predicated add (see p2p_umma_kernel)
            const uint64_t mk2 = f2_pack(mk, mk);
#pragma unroll
            for (int e = 0; e < 4; ++e) sacc[oi][e] = f2_fma(f2_pack(x[2 * e], x[2 * e + 1]), mk2, sacc[oi][e]);
          }
        }
      }
      fence_before_sync();
      mbar_arrive(&acc_empty);  // this group's accumulators drained: the issuer may start the next item once every group has
      asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
      // pass 2: the cross-warp contributions (ascending tap order on every lane), BN + LeakyReLU, octave max
      float best[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) best[e] = -INFINITY;
#pragma unroll
      for (int oi = 0; oi < kSemiOctPerGroup; ++oi) {
        const int o = grp + G * oi;
        if (o < n_oct) {
          if (wq < 3 && lane >= 30) {
#pragma unroll
            for (int f = 1; f < 3; ++f) {
              const bool take = lane + f >= 32;
              const float4* src = pub + ((o * 3 + wq) * 3 + (f - 1) + (take ? lane + f - 32 : 0)) * 2;
              const float4 x0 = src[0], x1 = src[1];
              const float mk = take ? 1.f : 0.f;
              const uint64_t mk2 = f2_pack(mk, mk);
              sacc[oi][0] = f2_fma(f2_pack(x0.x, x0.y), mk2, sacc[oi][0]), sacc[oi][1] = f2_fma(f2_pack(x0.z, x0.w), mk2, sacc[oi][1]);
              sacc[oi][2] = f2_fma(f2_pack(x1.x, x1.y), mk2, sacc[oi][2]), sacc[oi][3] = f2_fma(f2_pack(x1.z, x1.w), mk2, sacc[oi][3]);
            }
          }
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float y0, y1;
            f2_unpack(f2_fma(sacc[oi][e], f2_pack(s_scale[2 * e], s_scale[2 * e + 1]), f2_pack(s_shift[2 * e], s_shift[2 * e + 1])), y0, y1);
            best[2 * e] = fmaxf(best[2 * e], fmaxf(y0, kLeakySlope * y0)), best[2 * e + 1] = fmaxf(best[2 * e + 1], fmaxf(y1, kLeakySlope * y1));
          }
        }
      }
      // every group publishes the maxima of its octaves and reads the others' (max is exact and order-free), then stores one
      // quarter of the item: group 0 / 1 the two channel groups of the home row, groups 2 / 3 their wrap-row copies
      {
        float4* d = s_best + (grp * 128 + tid) * 2;
        d[0] = make_float4(best[0], best[1], best[2], best[3]), d[1] = make_float4(best[4], best[5], best[6], best[7]);
      }
      asm volatile("bar.sync %0, %1;" ::"r"(1 + G), "r"(128 * G) : "memory");
#pragma unroll
      for (int g2 = 1; g2 < G; ++g2) {
        const int og = (grp + g2) & (G - 1);
        const float4 b0 = s_best[(og * 128 + tid) * 2], b1 = s_best[(og * 128 + tid) * 2 + 1];
        best[0] = fmaxf(best[0], b0.x), best[1] = fmaxf(best[1], b0.y), best[2] = fmaxf(best[2], b0.z), best[3] = fmaxf(best[3], b0.w);
        best[4] = fmaxf(best[4], b1.x), best[5] = fmaxf(best[5], b1.y), best[6] = fmaxf(best[6], b1.z), best[7] = fmaxf(best[7], b1.w);
      }
      asm volatile("bar.sync %0, %1;" ::"r"(2 + G), "r"(128 * G) : "memory");  // s_best read: the next item may overwrite it
      const long long row0 = (((long long)b * 2 + 0) * 23 + c) * a.Wd, row1 = (((long long)b * 2 + 1) * 23 + c) * a.Wd;
      const long long wr = (long long)12 * a.Wd * 8;
      if (tid < a.TB && t < a.T && (grp < 2 || c < 11)) {  // wrap rows 12..22 = rows 0..10 (models.py:27-28)
        float gv[8];
        if ((grp & 1) == 0) {
#pragma unroll
          for (int e = 0; e < 4; ++e) gv[e] = pcv[e], gv[4 + e] = best[e];
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e) gv[e] = best[4 + e], gv[4 + e] = 0.f;
        }
        const long long q = (((grp & 1) ? row1 : row0) + t + 3) * 8 + (grp >= 2 ? wr : 0);
        store_split8(a.out_hi + q, a.out_lo + q, gv);
      }
      if (grp > 0) continue;
      if (t0 == 0 && tid < 6) {
        // zero halo columns 0..2 and T + 3..T + 5 (the "same" padding of the equivariant convs, models.py:45-47)
        const int col = tid < 3 ? tid : a.T + tid;
        const uint4 z = make_uint4(0, 0, 0, 0);
        for (int g = 0; g < 2; ++g) {
          const long long q = ((g ? row1 : row0) + col) * 8;
          *reinterpret_cast<uint4*>(a.out_hi + q) = z, *reinterpret_cast<uint4*>(a.out_lo + q) = z;
          if (c < 11) *reinterpret_cast<uint4*>(a.out_hi + q + wr) = z, *reinterpret_cast<uint4*>(a.out_lo + q + wr) = z;
        }
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == ISSUER) tmem_dealloc(tmem, 512);
}

// ---- equivariant pitch-class convolution (models.py:22-51) on tensor cores -----------------------------------------
//   out[co, c, t] = sum_{ci<16, dp<12, dt<7} W[co,ci,dp,dt] x[ci, (c+dp) mod 12, t + dt - pad]
// Input: 16-channel chunk planes [B][2 groups][23 rows][Wd][8] (rows 12..22 wrap, zero halo columns for "same" layers).
// K = 16 = the two channel groups of ONE position (LBO = group plane), anchors a = c*Wt + tl as in the 7x7 kernel.
//   PH = 2 (NCO = 16, the PitchClass2PitchClass stack): two time taps per start as N-phases, out[a] = D_0[a] + D_1[a+1];
//   PH = 1 (NCO = 64, both classifier heads' first conv in one pass): one tap per start.
// The weights (84 taps x 16 ci x NCO x {hi,lo}) do not fit beside the tile: they stream through a 2 x 16 KB ring.
// EPI 0: BN + LeakyReLU -> chunk planes (next layer);  EPI 1: + MaxPool2d((1,2)) (models.py:349-350, 396) -> chunk planes
// without halos + fp32 (B,16,12,T/2);  EPI 2: BN + LeakyReLU -> fp32 (B,32,12,T_out) for each head.
struct EquivArgs {
  const __half* in_hi;
  const __half* in_lo;
  int Wd_in, T_out, TB, n_ttiles;
  const __half* wimg;
  const float* scale;
  const float* shift;
  __half* out_hi;
  __half* out_lo;
  int Wd_out, col_off;
  float* out_f32;
  float* out_f32_b;
  int out_rows;   // EPI 3: rows per group of the output planes (23: wrapped pitch classes, 12: plain rows)
  int raw;        // EPI 2, training step: no activation (y = acc * scale / kWScale + shift)
};

constexpr uint32_t kEqStageBytes = 16384;
constexpr int kEqWStages = 2;

__host__ __device__ inline uint32_t equiv_group_positions(int Wt) { return (uint32_t)(23 * Wt + 136); }
__host__ __device__ inline size_t equiv_smem_bytes(int Wt) {
  return (size_t)4 * equiv_group_positions(Wt) * 16 + kEqWStages * kEqStageBytes + 2 * 4 * 128 * 16;
}

// Weight image: [start j][chunk g 2][n N1][ci 8] fp16; N1 = 2 * NCO * PH: rows [0, N1/2) = W_hi of (phase f = n / NCO,
// co = n % NCO), rows [N1/2, N1) = W_lo.  PH = 2: j = dp*4 + s/2, time tap dt = s + f (zero for dt = 7); PH = 1: j = dp*7 + dt.
// Output channels [0, split) come from w0 (Cout0 = split), [split, NCO) from w1.
__global__ void equiv_pack_weights_kernel(const float* __restrict__ w0, const float* __restrict__ w1, int split, int NCO, int Cin, int PH,
                                          __half* __restrict__ img, int KH = 12) {
  // KH = 1 (the genre head's Conv2d(16, 32, (1, 7)), models.py:724): 7 taps + one all-zero start, PH = 1
  const int N2 = NCO * PH, NS = KH == 12 ? 12 * (PH == 2 ? 4 : 7) : 8;
  const int n_items = NS * 2 * N2 * 8;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_items; i += gridDim.x * blockDim.x) {
    const int e = i % 8, n = (i / 8) % N2, g = (i / (8 * N2)) % 2, j = i / (16 * N2);
    const int f = n / NCO, co = n % NCO, ci = g * 8 + e;
    const int dp = KH == 12 ? (PH == 2 ? j / 4 : j / 7) : 0, dt = KH == 12 ? (PH == 2 ? 2 * (j % 4) + f : j % 7) : j;
    float v = 0.f;
    if (dt < 7 && ci < Cin) {
      const float* w = co < split ? w0 : w1;
      const int cc = co < split ? co : co - split;
      v = w[(((long long)cc * Cin + ci) * KH + dp) * 7 + dt] * kWScale;
    }
    const __half hi = __float2half_rn(v);
    const __half lo = __float2half_rn(v - __half2float(hi));
    const long long base = ((long long)(j * 2 + g) * 2 * N2) * 8;
    img[base + (long long)n * 8 + e] = hi;
    img[base + (long long)(N2 + n) * 8 + e] = lo;
  }
}

template <int NCO, int PH, int EPI, int KHT = 12>
__global__ void __launch_bounds__(192) equiv_umma_kernel(const EquivArgs a) {
  using namespace umma;
  constexpr int N2 = NCO * PH, N1 = 2 * N2;
  constexpr int ACC = N1;                        // TMEM columns per accumulator (one 128-anchor block)
  constexpr int NACC = ACC <= 64 ? 2 : 1;        // blocks sharing one pass over the weight stream
  constexpr int NS = KHT == 12 ? 12 * (PH == 2 ? 4 : 7) : 8;  // MMA starts per block (KHT = 1: 7 taps + a zero-weight start)
  static_assert(KHT == 12 || (KHT == 1 && PH == 1), "row taps: 12 (equivariant) or 1");
  constexpr uint32_t START_BYTES = 32u * N1;
  constexpr int SPS = kEqStageBytes / START_BYTES;
  constexpr int NSTG = NS / SPS;
  static_assert(NS % SPS == 0, "stages must tile the tap list");
  constexpr int STRIDE = PH == 2 ? (EPI == 1 ? 126 : 127) : 128;

  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t tile_bar, w_full[kEqWStages], w_empty[kEqWStages], acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_slot;
  __shared__ float s_scale[NCO], s_shift[NCO];

  const int tid = threadIdx.x, warp = uniform_warp_idx(), lane = tid & 31;
  const int b = blockIdx.y, t0 = blockIdx.x * a.TB;
  const int TBv = min(a.TB, a.T_out - t0);
  const int Wt = a.TB + 6;
  const int cols_in = min(Wt, a.Wd_in - t0);
  const uint32_t GPpos = equiv_group_positions(Wt), GP = GPpos * 16;
  const int n_anchor = 12 * Wt;
  const int n_mb = (n_anchor + STRIDE - 1) / STRIDE, n_grp = (n_mb + NACC - 1) / NACC;
  uint8_t* s_hi = smem;             // [g 2][GP]
  uint8_t* s_lo = smem + 2 * GP;
  uint8_t* s_w = smem + 4 * GP;
  float4* s_ex = reinterpret_cast<float4*>(smem + 4 * GP + kEqWStages * kEqStageBytes);  // 2 x [4][128]

  if (warp == 4) tmem_alloc(&tmem_slot, 256);
  if (tid == 0) {
    mbar_init(&tile_bar, 1);
    for (int s = 0; s < kEqWStages; ++s) mbar_init(&w_full[s], 1), mbar_init(&w_empty[s], 1);
    mbar_init(&acc_full[0], 1), mbar_init(&acc_full[1], 1);
    mbar_init(&acc_empty[0], 128), mbar_init(&acc_empty[1], 128);
    mbar_init_fence();
  }
  for (int i = tid; i < NCO; i += blockDim.x) s_scale[i] = a.scale[i] * (1.f / kWScale), s_shift[i] = a.shift[i];
  {
    // zero what the bulk copies leave untouched (see p2p_umma_kernel)
    const uint4 z = make_uint4(0, 0, 0, 0);
    const int gap = Wt - cols_in;
    for (int pl = 0; pl < 4; ++pl) {
      uint4* base = reinterpret_cast<uint4*>(smem + (size_t)pl * GP);
      for (uint32_t i = 23u * Wt + tid; i < GPpos; i += blockDim.x) base[i] = z;
      if (gap > 0)
        for (int i = tid; i < 23 * gap; i += blockDim.x) base[(uint32_t)(i / gap) * Wt + cols_in + i % gap] = z;
    }
    fence_proxy_async();
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_slot;

  if (warp == 4) {
    // ------------------------------------------------------------ loader: activation tile, then the weight stream
    const uint32_t row_bytes = (uint32_t)cols_in * 16;
    if (lane == 0) mbar_arrive_expect_tx(&tile_bar, 4u * 23u * row_bytes);
    __syncwarp();
    for (int idx = lane; idx < 46; idx += 32) {
      const int g = idx / 23, r = idx - g * 23;
      const long long src = ((((long long)b * 2 + g) * 23 + r) * a.Wd_in + t0) * 8;
      bulk_g2s(s_hi + (size_t)g * GP + (size_t)r * Wt * 16, a.in_hi + src, row_bytes, &tile_bar);
      bulk_g2s(s_lo + (size_t)g * GP + (size_t)r * Wt * 16, a.in_lo + src, row_bytes, &tile_bar);
    }
    // the weight stream: converged warp, one elected lane issues (see umma.cuh: single-lane issue)
    const int n_it = n_grp * NSTG;
    for (int it = 0; it < n_it; ++it) {
      const int s = it % kEqWStages;
      mbar_wait(&w_empty[s], ((it / kEqWStages) & 1) ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&w_full[s], kEqStageBytes);
        bulk_g2s(s_w + (size_t)s * kEqStageBytes, reinterpret_cast<const uint8_t*>(a.wimg) + (size_t)(it % NSTG) * kEqStageBytes,
                 kEqStageBytes, &w_full[s]);
      }
      __syncwarp();
    }
  } else if (warp == 5) {
    // ------------------------------------------------------------ MMA issuer (converged warp, one elected lane issues)
    mbar_wait(&tile_bar, 0);
    const uint64_t A_DESC = desc_hi(GP);  // chunk 1 = the other channel group of the same position
    constexpr uint64_t B_DESC = desc_hi(N1 * 16);
    constexpr uint32_t IDESC1 = idesc_f16(N1), IDESC2 = idesc_f16(N2);
    const uint32_t hi0 = smem_u32(s_hi), lo0 = smem_u32(s_lo), w0 = smem_u32(s_w);
    int it = 0;
    for (int grp = 0; grp < n_grp; ++grp) {
      const int gb = grp & 1;
      mbar_wait(&acc_empty[gb], ((grp >> 1) & 1) ^ 1);
      fence_after_sync();
      for (int stg = 0; stg < NSTG; ++stg, ++it) {
        const int s = it % kEqWStages;
        mbar_wait(&w_full[s], (it / kEqWStages) & 1);
        fence_after_sync();
        if (elect_one()) {
#pragma unroll
          for (int u = 0; u < SPS; ++u) {
            const int j = stg * SPS + u;
            const int dp = KHT == 12 ? (PH == 2 ? j / 4 : j / 7) : 0, sft = KHT == 12 ? (PH == 2 ? 2 * (j % 4) : j % 7) : j;
            const uint64_t bd = make_desc(B_DESC, w0 + s * kEqStageBytes + u * START_BYTES);
#pragma unroll
            for (int acc = 0; acc < NACC; ++acc) {
              const int m = grp * NACC + acc;
              if (m < n_mb) {
                const uint32_t off = (uint32_t)(m * STRIDE + dp * Wt + sft) * 16;
                const uint32_t d = tmem + gb * (NACC * ACC) + acc * ACC;
                mma_f16(d, make_desc(A_DESC, hi0 + off), bd, IDESC1, j ? 1u : 0u);
                mma_f16(d, make_desc(A_DESC, lo0 + off), bd, IDESC2, 1u);
              }
            }
          }
          commit(&w_empty[s]);
          if (stg == NSTG - 1) commit(&acc_full[gb]);
        }
        __syncwarp();
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue: thread = TMEM lane = anchor row
    const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
    float4* exA = s_ex;
    float4* exB = s_ex + 4 * 128;
    for (int grp = 0; grp < n_grp; ++grp) {
      const int gb = grp & 1;
      mbar_wait(&acc_full[gb], (grp >> 1) & 1);
      fence_after_sync();
      for (int acc = 0; acc < NACC; ++acc) {
        const int m = grp * NACC + acc;
        if (m >= n_mb) break;
        const uint32_t d = lane_base + gb * (NACC * ACC) + acc * ACC;
        const int anchor = m * STRIDE + tid;
        const int c = anchor / Wt, tl = anchor - c * Wt;
        const int t = t0 + tl;
        if constexpr (PH == 2) {
          // NCO == 16: D_0 = cols [0,16) + [32,48), D_1 = cols [16,32) + [48,64)
          float o[16], d1[16];
          {
            float u[16], w[16];
            tmem_ld16(d, u), tmem_ld16(d + 32, w);
#pragma unroll
            for (int j = 0; j < 16; ++j) o[j] = u[j] + w[j];
            tmem_ld16(d + 16, u), tmem_ld16(d + 48, w);
#pragma unroll
            for (int j = 0; j < 16; ++j) d1[j] = u[j] + w[j];
          }
          float4* ex = (EPI == 0) ? (m & 1 ? exB : exA) : exA;
#pragma unroll
          for (int q = 0; q < 4; ++q) ex[q * 128 + tid] = make_float4(d1[4 * q], d1[4 * q + 1], d1[4 * q + 2], d1[4 * q + 3]);
          asm volatile("bar.sync 1, 128;" ::: "memory");
          if (tid < 127) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float4 x = ex[q * 128 + tid + 1];
              o[4 * q] += x.x, o[4 * q + 1] += x.y, o[4 * q + 2] += x.z, o[4 * q + 3] += x.w;
            }
          }
#pragma unroll
          for (int j = 0; j < 16; ++j) o[j] = leaky_f(fmaf(o[j], s_scale[j], s_shift[j]));
          if constexpr (EPI == 0) {
            if (tid < STRIDE && anchor < n_anchor && tl < TBv) {
              float g0[8], g1[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) g0[e] = o[e], g1[e] = o[8 + e];
              const long long q0 = ((((long long)b * 2 + 0) * 23 + c) * a.Wd_out + t + a.col_off) * 8;
              const long long q1 = ((((long long)b * 2 + 1) * 23 + c) * a.Wd_out + t + a.col_off) * 8;
              store_split8(a.out_hi + q0, a.out_lo + q0, g0);
              store_split8(a.out_hi + q1, a.out_lo + q1, g1);
              if (c < 11) {
                const long long wr = (long long)12 * a.Wd_out * 8;
                store_split8(a.out_hi + q0 + wr, a.out_lo + q0 + wr, g0);
                store_split8(a.out_hi + q1 + wr, a.out_lo + q1 + wr, g1);
              }
            }
          } else {
            // fused MaxPool2d((1,2)): frames (2u, 2u+1) sit in adjacent rows (t0 and TB are even)
#pragma unroll
            for (int q = 0; q < 4; ++q) exB[q * 128 + tid] = make_float4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
            asm volatile("bar.sync 1, 128;" ::: "memory");
            const int Th = a.T_out / 2;
            if (tid < STRIDE && anchor < n_anchor && tl < TBv && (t & 1) == 0 && (t >> 1) < Th) {
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const float4 x = exB[q * 128 + tid + 1];
                o[4 * q] = fmaxf(o[4 * q], x.x), o[4 * q + 1] = fmaxf(o[4 * q + 1], x.y);
                o[4 * q + 2] = fmaxf(o[4 * q + 2], x.z), o[4 * q + 3] = fmaxf(o[4 * q + 3], x.w);
              }
              const int uo = t >> 1;
              float g0[8], g1[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) g0[e] = o[e], g1[e] = o[8 + e];
              const long long q0 = ((((long long)b * 2 + 0) * 23 + c) * a.Wd_out + uo + a.col_off) * 8;
              const long long q1 = ((((long long)b * 2 + 1) * 23 + c) * a.Wd_out + uo + a.col_off) * 8;
              store_split8(a.out_hi + q0, a.out_lo + q0, g0);
              store_split8(a.out_hi + q1, a.out_lo + q1, g1);
              if (c < 11) {
                const long long wr = (long long)12 * a.Wd_out * 8;
                store_split8(a.out_hi + q0 + wr, a.out_lo + q0 + wr, g0);
                store_split8(a.out_hi + q1 + wr, a.out_lo + q1 + wr, g1);
              }
#pragma unroll
              for (int j = 0; j < 16; ++j) a.out_f32[(((long long)b * 16 + j) * 12 + c) * Th + uo] = o[j];
            }
          }
        } else {
          // PH == 1, NCO == 64: out = cols [0,64) + [64,128); heads: channels [0,32) tonic, [32,64) key (EPI 2)
          const bool valid = anchor < n_anchor && tl < TBv;
#pragma unroll
          for (int h = 0; h < NCO / 16; ++h) {
            float u[16], w[16];
            tmem_ld16(d + h * 16, u), tmem_ld16(d + N2 + h * 16, w);
            if constexpr (EPI == 3) {
              // BN + LeakyReLU -> chunk planes [B][NCO / 8 groups][out_rows][T_out][8] for the tensor-core head tails
              if (valid) {
                float g0[8], g1[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  g0[j] = leaky_f(fmaf(u[j] + w[j], s_scale[h * 16 + j], s_shift[h * 16 + j]));
                  g1[j] = leaky_f(fmaf(u[8 + j] + w[8 + j], s_scale[h * 16 + 8 + j], s_shift[h * 16 + 8 + j]));
                }
                const long long q0 = ((((long long)b * (NCO / 8) + 2 * h) * a.out_rows + c) * a.T_out + t) * 8;
                const long long q1 = q0 + (long long)a.out_rows * a.T_out * 8;
                store_split8(a.out_hi + q0, a.out_lo + q0, g0);
                store_split8(a.out_hi + q1, a.out_lo + q1, g1);
                if (a.out_rows == 23 && c < 11) {
                  const long long wr = (long long)12 * a.T_out * 8;
                  store_split8(a.out_hi + q0 + wr, a.out_lo + q0 + wr, g0);
                  store_split8(a.out_hi + q1 + wr, a.out_lo + q1 + wr, g1);
                }
              }
              continue;
            }
            if (valid) {
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const int co = h * 16 + j;
                const float yl = fmaf(u[j] + w[j], s_scale[co], s_shift[co]);
                const float y = a.raw ? yl : leaky_f(yl);
                float* dst = co < 32 ? a.out_f32 : a.out_f32_b;
                dst[(((long long)b * 32 + (co & 31)) * 12 + c) * a.T_out + t] = y;
              }
            }
          }
        }
      }
      fence_before_sync();
      mbar_arrive(&acc_empty[gb]);
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem, 256);
}

// ---- PitchClass2PitchClass stack (16 -> 16 channels, models.py:191-197) as a persistent 7-phase shift-GEMM ---------------
// Same scheme as p2p_umma_kernel, with the row taps running over the 23 wrapped pitch-class rows:
//   out[a] = sum_{dp<12} sum_{f<7} X[a + dp*Wt + f] . W[dp][f]       (a = c*Wt + tl, 16 -> 16 channels, "same" zero padding in time)
//   MMA 1 (per dp): A = x_hi (K = 16: the two channel groups, LBO = plane pitch), B = [W_hi | W_lo], N = 224 = 2 x (7 phases x 16 co)
//   MMA 2 (per dp): A = x_lo, B = W_hi, N = 112 (accumulates on the W_hi columns)
// i.e. 12 x (112 + 60) = 2064 MMA cycles per 122 anchors, against 48 x (48 + 40) = 4224 per 127 for the two-phase form, and
// the whole weight image (86 KB) stays resident in shared memory instead of streaming through a ring once per CTA.
// EPI 0: BN + LeakyReLU -> chunk planes with wrap rows (next conv);  EPI 1: + MaxPool2d((1,2)) (models.py:349-350, 396)
// -> chunk planes without halos + fp32 (B,16,12,T/2).
constexpr int kPcStride = 122;
constexpr int kPcGroups = 2;        // epilogue groups of 4 warps = accumulator buffers of 224 TMEM columns
constexpr int kPcThreads = 32 * (4 * kPcGroups + 2);
constexpr uint32_t kPcWBytes = 12 * 2 * 224 * 16;
constexpr uint32_t kPcPubBytes = 3 * 21 * 64;  // per epilogue group: [warp 1..3][phase f = 1..6: f lanes][co 16] floats

// [dp 12][chunk g 2][n 224][ci 8]: n < 112: W_hi of (phase f = n / 16, co = n % 16); n >= 112: W_lo
__global__ void pc2pc_pack_weights_kernel(const float* __restrict__ w, int Cout, int Cin, __half* __restrict__ img) {
  const int n_items = 12 * 2 * 112 * 8;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_items; i += gridDim.x * blockDim.x) {
    const int e = i % 8, n = (i / 8) % 112, g = (i / (8 * 112)) % 2, dp = i / (16 * 112);
    const int f = n / 16, co = n % 16, ci = g * 8 + e;
    float v = 0.f;
    if (ci < Cin && co < Cout) v = w[(((long long)co * Cin + ci) * 12 + dp) * 7 + f] * kWScale;
    const __half hi = __float2half_rn(v);
    const __half lo = __float2half_rn(v - __half2float(hi));
    img[((dp * 2 + g) * 224 + n) * 8 + e] = hi;
    img[((dp * 2 + g) * 224 + 112 + n) * 8 + e] = lo;
  }
}

struct Pc2PcArgs {
  const __half* in_hi;
  const __half* in_lo;   // [B][2][23][Wd_in][8]: zero halo columns (3 each side), wrap rows 12..22
  int Wd_in, T_out, TB, n_ttiles, n_tiles;
  const __half* wimg;    // pc2pc_pack_weights_kernel image (kPcWBytes)
  const float* scale;    // 16: eval-mode BN scale (the 1/kWScale factor is applied in the kernel)
  const float* shift;
  __half* out_hi;
  __half* out_lo;        // EPI 0: [B][2][23][Wd_out][8] written at column t + col_off; EPI 1: same with t / 2
  int Wd_out, col_off;
  float* out_f32;        // EPI 1: (B,16,12,T_out/2);  EPI 3: (B, Cout_store, 12, T_out)
  // EPI 3 (training step): y = acc * scale / kWScale / tc_scale_of(*maxbits) + shift, NO activation, planar fp32
  int Cout_store;
  const unsigned* maxbits;
  int accumulate;        // EPI 3: add to the stored values (a data gradient summed over several launches: > 16 input channels, several consumers)
};

__host__ __device__ inline uint32_t pc2pc_plane_positions(int Wt) { return (uint32_t)(23 * Wt + 136); }
__host__ __device__ inline size_t pc2pc_smem_bytes(int Wt) {
  return (size_t)2 * 4 * pc2pc_plane_positions(Wt) * 16 + kPcWBytes + kPcGroups * kPcPubBytes;
}

template <int EPI>
__global__ void __launch_bounds__(kPcThreads, 1) pc2pc_umma_kernel(const Pc2PcArgs a) {
  using namespace umma;
  constexpr int G = kPcGroups;
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t w_bar, full_bar[2], empty_bar[2], acc_full[G], acc_empty[G];
  __shared__ uint32_t tmem_slot;
  __shared__ float s_scale[16], s_shift[16];

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const int Wt = a.TB + 6;
  const uint32_t plane = pc2pc_plane_positions(Wt) * 16;  // one (hi | lo, channel group) plane; a tile buffer holds [hi g0][hi g1][lo g0][lo g1]
  uint8_t* s_w = smem + 8 * plane;
  uint8_t* s_pub = s_w + kPcWBytes;
  const int n_anchor = 12 * Wt;
  const int n_mb = (n_anchor + kPcStride - 1) / kPcStride;

  if (warp == 4 * G + 1) tmem_alloc(&tmem_slot, 512);
  if (threadIdx.x == 0) {
    mbar_init(&w_bar, 1);
    for (int i = 0; i < 2; ++i) mbar_init(&full_bar[i], 1), mbar_init(&empty_bar[i], 1);
    for (int i = 0; i < G; ++i) mbar_init(&acc_full[i], 1), mbar_init(&acc_empty[i], 128);
    mbar_init_fence();
  }
  if (threadIdx.x < 16) {
    const float inv = (EPI == 3 && a.maxbits) ? 1.f / tc_scale_of(__ldg(a.maxbits)) : 1.f;  // exact: a power of two
    s_scale[threadIdx.x] = a.scale[threadIdx.x] * (1.f / kWScale) * inv, s_shift[threadIdx.x] = a.shift[threadIdx.x];
  }
  // positions the bulk copies never write only feed discarded anchors; give them finite values once
  for (uint32_t i = threadIdx.x; i < 8 * plane / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_slot;

  if (warp == 4 * G) {
    // ------------------------------------------------------------ loader
    if (lane == 0) {
      mbar_arrive_expect_tx(&w_bar, kPcWBytes);
      bulk_g2s(s_w, a.wimg, kPcWBytes, &w_bar);
    }
    int k = 0;
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++k) {
      const int s = k & 1;
      const int b = tile / a.n_ttiles, t0 = (tile - b * a.n_ttiles) * a.TB;
      const int cols_in = min(Wt, a.Wd_in - t0);
      const uint32_t row_bytes = (uint32_t)cols_in * 16;
      mbar_wait_relaxed(&empty_bar[s], ((k >> 1) & 1) ^ 1);
      if (lane == 0) mbar_arrive_expect_tx(&full_bar[s], 92u * row_bytes);
      __syncwarp();
      uint8_t* dst = smem + (size_t)s * 4 * plane;
      for (int idx = lane; idx < 46; idx += 32) {
        const int g = idx / 23, r = idx - g * 23;
        const long long src = ((((long long)b * 2 + g) * 23 + r) * a.Wd_in + t0) * 8;
        bulk_g2s(dst + (size_t)g * plane + (size_t)r * Wt * 16, a.in_hi + src, row_bytes, &full_bar[s]);
        bulk_g2s(dst + (size_t)(2 + g) * plane + (size_t)r * Wt * 16, a.in_lo + src, row_bytes, &full_bar[s]);
      }
    }
  } else if (warp == 4 * G + 1) {
    // ------------------------------------------------------------ MMA issuer (converged warp, one elected lane issues)
    const uint64_t A_DESC = desc_hi(plane);         // chunk 1 = the other channel group of the same position
    constexpr uint64_t B_DESC = desc_hi(224 * 16);  // chunk stride: 224 rows x 16 B
    constexpr uint32_t IDESC_WIDE = idesc_f16(224), IDESC_NARROW = idesc_f16(112);
    const uint32_t w0 = smem_u32(s_w);
    mbar_wait(&w_bar, 0);
    int k = 0;
    uint32_t j = 0;
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++k) {
      const int s = k & 1;
      const uint32_t hi0 = smem_u32(smem + (size_t)s * 4 * plane), lo0 = hi0 + 2 * plane;
      mbar_wait(&full_bar[s], (k >> 1) & 1);
      for (int m = 0; m < n_mb; ++m, ++j) {
        const uint32_t buf = j % G;
        mbar_wait(&acc_empty[buf], ((j / G) & 1) ^ 1);
        fence_after_sync();
        const uint32_t d = tmem + buf * 256;
        const uint32_t a_off = (uint32_t)(m * kPcStride) * 16;
        if (elect_one()) {
#pragma unroll
          for (int dp = 0; dp < 12; ++dp) {
            const uint32_t off = a_off + (uint32_t)(dp * Wt) * 16;
            const uint64_t bd = make_desc(B_DESC, w0 + dp * (2 * 224 * 16));
            mma_f16(d, make_desc(A_DESC, hi0 + off), bd, IDESC_WIDE, dp ? 1u : 0u);
            mma_f16(d, make_desc(A_DESC, lo0 + off), bd, IDESC_NARROW, 1u);
          }
          commit(&acc_full[buf]);
          if (m == n_mb - 1) commit(&empty_bar[s]);
        }
        __syncwarp();
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue: thread = TMEM lane = anchor row
    const int grp = warp >> 2, wq = warp & 3, tid = threadIdx.x & 127;
    const uint32_t acc = tmem + ((uint32_t)(wq * 32) << 16) + grp * 256;
    float4* pub = reinterpret_cast<float4*>(s_pub + grp * kPcPubBytes);
    const uint32_t wt_magic = 0xFFFFFFFFu / (uint32_t)Wt + 1;
    uint32_t n_done = 0;
    uint32_t j = grp, j0 = 0;
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
      const int b = tile / a.n_ttiles, t0 = (tile - b * a.n_ttiles) * a.TB;
      const int TBv = min(a.TB, a.T_out - t0);
      for (; j < j0 + n_mb; j += G, ++n_done) {
        const int m = (int)(j - j0);
        mbar_wait_relaxed(&acc_full[grp], n_done & 1);
        fence_after_sync();
        // phase f = time tap f: out[a] = sum_f D_f[a + f]; D_f = columns [16 f, 16 f + 16) (x . W_hi) + [112 + 16 f, ..) (x_hi . W_lo)
        uint64_t o[8];
#pragma unroll
        for (int f = 0; f < 7; ++f) {
          uint32_t vh[16], vl[16];
          tmem_ld16_issue(acc + 16 * f, vh);
          tmem_ld16_issue(acc + 112 + 16 * f, vl);
          tmem_ld_wait16(vh);
          tmem_ld_wait16(vl);
          uint64_t u[8];
#pragma unroll
          for (int e = 0; e < 8; ++e)
            u[e] = f2_add(f2_pack(__uint_as_float(vh[2 * e]), __uint_as_float(vh[2 * e + 1])),
                          f2_pack(__uint_as_float(vl[2 * e]), __uint_as_float(vl[2 * e + 1])));
          if (f == 0) {
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = u[e];
          } else {
            float x[16];
#pragma unroll
            for (int e = 0; e < 8; ++e) f2_unpack(u[e], x[2 * e], x[2 * e + 1]);
            if (wq > 0 && lane < f) {
              float4* dst = pub + ((wq - 1) * 21 + f * (f - 1) / 2 + lane) * 4;
#pragma unroll
              for (int q = 0; q < 4; ++q) dst[q] = make_float4(x[4 * q], x[4 * q + 1], x[4 * q + 2], x[4 * q + 3]);
            }
#pragma unroll
            for (int c = 0; c < 16; ++c) x[c] = __shfl_down_sync(0xffffffffu, x[c], f);
            const float mk = (lane + f < 32) ? 1.f : 0.f;  // see p2p_umma_kernel: masked FMA = predicated add
            const uint64_t mk2 = f2_pack(mk, mk);
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = f2_fma(f2_pack(x[2 * e], x[2 * e + 1]), mk2, o[e]);
          }
        }
        fence_before_sync();
        mbar_arrive(&acc_empty[grp]);
        asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
        if (wq < 3 && lane >= 26) {
#pragma unroll
          for (int f = 1; f < 7; ++f) {
            const bool take = lane + f >= 32;
            const float4* src = pub + (wq * 21 + f * (f - 1) / 2 + (take ? lane + f - 32 : 0)) * 4;
            const float mk = take ? 1.f : 0.f;
            const uint64_t mk2 = f2_pack(mk, mk);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float4 x = src[q];
              o[2 * q] = f2_fma(f2_pack(x.x, x.y), mk2, o[2 * q]), o[2 * q + 1] = f2_fma(f2_pack(x.z, x.w), mk2, o[2 * q + 1]);
            }
          }
        }
        asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");  // the hand-over buffer is reused by the next block
        float y[16];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          f2_unpack(f2_fma(o[e], f2_pack(s_scale[2 * e], s_scale[2 * e + 1]), f2_pack(s_shift[2 * e], s_shift[2 * e + 1])), y[2 * e], y[2 * e + 1]);
          if constexpr (EPI != 3) y[2 * e] = fmaxf(y[2 * e], kLeakySlope * y[2 * e]), y[2 * e + 1] = fmaxf(y[2 * e + 1], kLeakySlope * y[2 * e + 1]);
        }
        const int anchor = m * kPcStride + tid;
        const int c = (int)__umulhi((uint32_t)anchor, wt_magic), tl = anchor - c * Wt;
        const int t = t0 + tl;
        bool valid = tid < kPcStride && anchor < n_anchor && tl < TBv;
        int col = t + a.col_off;
        if constexpr (EPI == 3) {  // raw planar fp32 (the training step's conv / data gradient)
          if (valid) {
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (i < a.Cout_store) {
                float* dst = a.out_f32 + (((long long)b * a.Cout_store + i) * 12 + c) * a.T_out + t;
                *dst = a.accumulate ? *dst + y[i] : y[i];
              }
          }
          valid = false;
        }
        if constexpr (EPI == 1) {
          // fused MaxPool2d((1,2)): frames (2u, 2u+1) are neighbouring anchors of one warp (t0, TB, Wt and the block stride are even)
#pragma unroll
          for (int i = 0; i < 16; ++i) y[i] = fmaxf(y[i], __shfl_down_sync(0xffffffffu, y[i], 1));
          const int Th = a.T_out / 2;
          valid = valid && (t & 1) == 0 && (t >> 1) < Th;
          col = (t >> 1) + a.col_off;
          if (valid) {
#pragma unroll
            for (int i = 0; i < 16; ++i) a.out_f32[(((long long)b * 16 + i) * 12 + c) * Th + (t >> 1)] = y[i];
          }
        }
        if (valid) {
          float g0[8], g1[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) g0[e] = y[e], g1[e] = y[8 + e];
          const long long q0 = ((((long long)b * 2 + 0) * 23 + c) * a.Wd_out + col) * 8;
          const long long q1 = ((((long long)b * 2 + 1) * 23 + c) * a.Wd_out + col) * 8;
          store_split8(a.out_hi + q0, a.out_lo + q0, g0);
          store_split8(a.out_hi + q1, a.out_lo + q1, g1);
          if (c < 11) {  // wrap rows 12..22 = rows 0..10 (models.py:27-28)
            const long long wr = (long long)12 * a.Wd_out * 8;
            store_split8(a.out_hi + q0 + wr, a.out_lo + q0 + wr, g0);
            store_split8(a.out_hi + q1 + wr, a.out_lo + q1 + wr, g1);
          }
        }
      }
      j0 += n_mb;
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 4 * G + 1) tmem_dealloc(tmem, 512);
}

// ---- narrow equivariant pitch-class convolutions (<= 8 -> <= 8 channels: the layer-0 PitchClass2PitchClass stack) ------------
// models.py:22-51, 191-197 with 1 -> 4 -> 4 -> 4 channels.  One channel group, so the K = 16 of an MMA holds [x_hi | x_lo]
// (as in p2p_umma_kernel) and ONE MMA per row tap does all three products: N = 112 = 7 phases x {W_hi 8, W_lo 8}, 12 row taps
// = 720 MMA cycles per 122 anchors.  Persistent CTA per SM, resident weights (43 KB), double-buffered tiles, 4 epilogue groups.
// EPI 0: BN + LeakyReLU -> chunk planes with wrap rows (next conv);  EPI 2: -> fp32 (B, Cout, 12, T).
constexpr int kPc8Groups = 4;
constexpr int kPc8Threads = 32 * (4 * kPc8Groups + 2);
constexpr uint32_t kPc8WBytes = 12 * 2 * 112 * 16;
constexpr int kPc8MaxTB = 76;

// [dp 12][chunk 2][n 112][ci 8]: n = 16 f + j; j < 8: W_hi of output channel j, j >= 8: W_lo; both chunks hold the same weights
__global__ void pc8_pack_weights_kernel(const float* __restrict__ w, int Cout, int Cin, __half* __restrict__ img) {
  const int n_items = 12 * 56 * 8;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_items; i += gridDim.x * blockDim.x) {
    const int ci = i % 8, fc = (i / 8) % 56, dp = i / 448;
    const int f = fc / 8, co = fc % 8;
    float v = 0.f;
    if (ci < Cin && co < Cout) v = w[(((long long)co * Cin + ci) * 12 + dp) * 7 + f] * kWScale;
    const __half hi = __float2half_rn(v);
    const __half lo = __float2half_rn(v - __half2float(hi));
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      img[((dp * 2 + c) * 112 + 16 * f + co) * 8 + ci] = hi;
      img[((dp * 2 + c) * 112 + 16 * f + 8 + co) * 8 + ci] = lo;
    }
  }
}

struct Pc8Args {
  const __half* in_hi;
  const __half* in_lo;   // [B][23][Wd_in][8]: zero halo columns (3 each side), wrap rows 12..22
  int Wd_in, T_out, TB, n_ttiles, n_tiles;
  const __half* wimg;    // pc8_pack_weights_kernel image (kPc8WBytes)
  const float* scale;    // Cout: eval-mode BN scale (the 1/kWScale factor is applied in the kernel)
  const float* shift;
  int Cout;
  __half* out_hi;
  __half* out_lo;        // EPI 0: [B][23][Wd_out][8] written at column t + 3
  int Wd_out;
  float* out_f32;        // EPI 2: (B, Cout, 12, T_out)
  int raw;               // EPI 2, training step: no activation, scale additionally divided by tc_scale_of(*maxbits)
  const unsigned* maxbits;
};

__host__ __device__ inline uint32_t pc8_plane_positions(int Wt) { return (uint32_t)(23 * Wt + 136); }
__host__ __device__ inline size_t pc8_smem_bytes(int Wt) {
  return (size_t)2 * 2 * pc8_plane_positions(Wt) * 16 + kPc8WBytes + kPc8Groups * kP2PPubBytes;
}

template <int EPI>
__global__ void __launch_bounds__(kPc8Threads, 1) pc8_umma_kernel(const Pc8Args a) {
  using namespace umma;
  constexpr int G = kPc8Groups;
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t w_bar, full_bar[2], empty_bar[2], acc_full[G], acc_empty[G];
  __shared__ uint32_t tmem_slot;
  __shared__ float s_scale[8], s_shift[8];

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const int Wt = a.TB + 6;
  const uint32_t plane = pc8_plane_positions(Wt) * 16;  // a tile buffer holds [hi][lo]
  uint8_t* s_w = smem + 4 * plane;
  uint8_t* s_pub = s_w + kPc8WBytes;
  const int n_anchor = 12 * Wt;
  const int n_mb = (n_anchor + kP2PStride - 1) / kP2PStride;
  const uint32_t ntt_magic = 0xFFFFFFFFu / (uint32_t)a.n_ttiles + 1;
  auto clip_of = [&](int tile) { return a.n_ttiles == 1 ? tile : (int)__umulhi((uint32_t)tile, ntt_magic); };

  if (warp == 4 * G + 1) tmem_alloc(&tmem_slot, 128 * G);
  if (threadIdx.x == 0) {
    mbar_init(&w_bar, 1);
    for (int i = 0; i < 2; ++i) mbar_init(&full_bar[i], 1), mbar_init(&empty_bar[i], 1);
    for (int i = 0; i < G; ++i) mbar_init(&acc_full[i], 1), mbar_init(&acc_empty[i], 128);
    mbar_init_fence();
  }
  if (threadIdx.x < 8) {
    const bool on = (int)threadIdx.x < a.Cout;
    const float inv = (a.raw && a.maxbits) ? 1.f / tc_scale_of(__ldg(a.maxbits)) : 1.f;  // exact: a power of two
    s_scale[threadIdx.x] = on ? a.scale[threadIdx.x] * (1.f / kWScale) * inv : 0.f, s_shift[threadIdx.x] = on ? a.shift[threadIdx.x] : 0.f;
  }
  // positions the bulk copies never write only feed discarded anchors; give them finite values once
  for (uint32_t i = threadIdx.x; i < 4 * plane / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_slot;

  if (warp == 4 * G) {
    // ------------------------------------------------------------ loader
    if (lane == 0) {
      mbar_arrive_expect_tx(&w_bar, kPc8WBytes);
      bulk_g2s(s_w, a.wimg, kPc8WBytes, &w_bar);
    }
    int k = 0;
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++k) {
      const int s = k & 1;
      const int b = clip_of(tile), t0 = (tile - b * a.n_ttiles) * a.TB;
      const int cols_in = min(Wt, a.Wd_in - t0);
      const uint32_t row_bytes = (uint32_t)cols_in * 16;
      mbar_wait_relaxed(&empty_bar[s], ((k >> 1) & 1) ^ 1);
      if (lane == 0) mbar_arrive_expect_tx(&full_bar[s], 46u * row_bytes);
      __syncwarp();
      uint8_t* dst = smem + (size_t)s * 2 * plane;
      if (lane < 23) {
        const long long src = (((long long)b * 23 + lane) * a.Wd_in + t0) * 8;
        bulk_g2s(dst + (size_t)lane * Wt * 16, a.in_hi + src, row_bytes, &full_bar[s]);
        bulk_g2s(dst + plane + (size_t)lane * Wt * 16, a.in_lo + src, row_bytes, &full_bar[s]);
      }
    }
  } else if (warp == 4 * G + 1) {
    // ------------------------------------------------------------ MMA issuer (converged warp, one elected lane issues)
    const uint64_t A_DESC = desc_hi(plane);         // chunk 1 = the x_lo plane at the same position
    constexpr uint64_t B_DESC = desc_hi(112 * 16);
    constexpr uint32_t IDESC = idesc_f16(112);
    const uint32_t w0 = smem_u32(s_w);
    mbar_wait(&w_bar, 0);
    int k = 0;
    uint32_t j = 0;
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++k) {
      const int s = k & 1;
      const uint32_t hi0 = smem_u32(smem + (size_t)s * 2 * plane);
      mbar_wait(&full_bar[s], (k >> 1) & 1);
      for (int m = 0; m < n_mb; ++m, ++j) {
        const uint32_t buf = j % G;
        mbar_wait(&acc_empty[buf], ((j / G) & 1) ^ 1);
        fence_after_sync();
        const uint32_t d = tmem + buf * 128;
        const uint32_t a_off = hi0 + (uint32_t)(m * kP2PStride) * 16;
        if (elect_one()) {
#pragma unroll
          for (int dp = 0; dp < 12; ++dp)
            mma_f16(d, make_desc(A_DESC, a_off + (uint32_t)(dp * Wt) * 16), make_desc(B_DESC, w0 + dp * (2 * 112 * 16)), IDESC, dp ? 1u : 0u);
          commit(&acc_full[buf]);
          if (m == n_mb - 1) commit(&empty_bar[s]);
        }
        __syncwarp();
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue (as p2p_umma_kernel): thread = TMEM lane = anchor row
    const int grp = warp >> 2, wq = warp & 3, tid = threadIdx.x & 127;
    const uint32_t acc = tmem + ((uint32_t)(wq * 32) << 16) + grp * 128;
    float4* pub = reinterpret_cast<float4*>(s_pub + grp * kP2PPubBytes);
    const uint32_t wt_magic = 0xFFFFFFFFu / (uint32_t)Wt + 1;
    uint64_t sc2[4], sh2[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) sc2[e] = f2_pack(s_scale[2 * e], s_scale[2 * e + 1]), sh2[e] = f2_pack(s_shift[2 * e], s_shift[2 * e + 1]);
    uint32_t n_done = 0;
    uint32_t j = grp, j0 = 0;
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
      const int b = clip_of(tile), t0 = (tile - b * a.n_ttiles) * a.TB;
      const int TBv = min(a.TB, a.T_out - t0);
      for (; j < j0 + n_mb; j += G, ++n_done) {
        const int m = (int)(j - j0);
        mbar_wait_relaxed(&acc_full[grp], n_done & 1);
        fence_after_sync();
        float4* pub_w = pub + (n_done & 1) * (3 * 21 * 2);
        uint64_t o[4];
        uint32_t r[2][16];
        tmem_ld16_issue(acc, r[0]);
#pragma unroll
        for (int f = 0; f < 7; ++f) {
          uint32_t(&v)[16] = r[f & 1];
          tmem_ld_wait16(v);
          if (f < 6) tmem_ld16_issue(acc + 16 * (f + 1), r[(f + 1) & 1]);
          uint64_t u[4];
#pragma unroll
          for (int e = 0; e < 4; ++e)
            u[e] = f2_add(f2_pack(__uint_as_float(v[2 * e]), __uint_as_float(v[2 * e + 1])),
                          f2_pack(__uint_as_float(v[8 + 2 * e]), __uint_as_float(v[9 + 2 * e])));
          if (f == 0) {
#pragma unroll
            for (int e = 0; e < 4; ++e) o[e] = u[e];
          } else {
            float x[8];
#pragma unroll
            for (int e = 0; e < 4; ++e) f2_unpack(u[e], x[2 * e], x[2 * e + 1]);
            if (wq > 0 && lane < f) {
              float4* dst = pub_w + ((wq - 1) * 21 + f * (f - 1) / 2 + lane) * 2;
              dst[0] = make_float4(x[0], x[1], x[2], x[3]), dst[1] = make_float4(x[4], x[5], x[6], x[7]);
            }
#pragma unroll
            for (int c = 0; c < 8; ++c) x[c] = __shfl_down_sync(0xffffffffu, x[c], f);
            const float mk = (lane + f < 32) ? 1.f : 0.f;
            const uint64_t mk2 = f2_pack(mk, mk);
#pragma unroll
            for (int e = 0; e < 4; ++e) o[e] = f2_fma(f2_pack(x[2 * e], x[2 * e + 1]), mk2, o[e]);
          }
        }
        fence_before_sync();
        mbar_arrive(&acc_empty[grp]);
        asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
        if (wq < 3 && lane >= 26) {
#pragma unroll
          for (int f = 1; f < 7; ++f) {
            const bool take = lane + f >= 32;
            const float4* src = pub_w + (wq * 21 + f * (f - 1) / 2 + (take ? lane + f - 32 : 0)) * 2;
            const float4 x0 = src[0], x1 = src[1];
            const float mk = take ? 1.f : 0.f;
            const uint64_t mk2 = f2_pack(mk, mk);
            o[0] = f2_fma(f2_pack(x0.x, x0.y), mk2, o[0]), o[1] = f2_fma(f2_pack(x0.z, x0.w), mk2, o[1]);
            o[2] = f2_fma(f2_pack(x1.x, x1.y), mk2, o[2]), o[3] = f2_fma(f2_pack(x1.z, x1.w), mk2, o[3]);
          }
        }
        const int anchor = m * kP2PStride + tid;
        if (tid < kP2PStride && anchor < n_anchor) {
          const int c = (int)__umulhi((uint32_t)anchor, wt_magic), tl = anchor - c * Wt;
          if (tl < TBv) {
            float y[8];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              f2_unpack(f2_fma(o[e], sc2[e], sh2[e]), y[2 * e], y[2 * e + 1]);
              if (!a.raw) y[2 * e] = fmaxf(y[2 * e], kLeakySlope * y[2 * e]), y[2 * e + 1] = fmaxf(y[2 * e + 1], kLeakySlope * y[2 * e + 1]);
            }
            const int t = t0 + tl;
            if constexpr (EPI == 0) {
              const long long q0 = (((long long)b * 23 + c) * a.Wd_out + t + 3) * 8;
              store_split8(a.out_hi + q0, a.out_lo + q0, y);
              if (c < 11) {  // wrap rows 12..22 = rows 0..10 (models.py:27-28)
                const long long wr = (long long)12 * a.Wd_out * 8;
                store_split8(a.out_hi + q0 + wr, a.out_lo + q0 + wr, y);
              }
            } else {
#pragma unroll
              for (int co = 0; co < 8; ++co)
                if (co < a.Cout) a.out_f32[(((long long)b * a.Cout + co) * 12 + c) * a.T_out + t] = y[co];
            }
          }
        }
      }
      j0 += n_mb;
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 4 * G + 1) tmem_dealloc(tmem, 128 * G);
}

// ---- last convolution of the classifier heads on tensor cores (models.py:716-733: 32 -> 1 channel, valid in time) -----
//   out[c, t] = bias + sum_{ci<32, dp<KH, dt<7} W[ci, dp, dt] x[ci, c + dp, t + dt]     (rows wrap for tonic / key: 23-row planes)
// One output channel would leave the MMA's N idle, so the 7 time taps are the N-phases:
//   D_f[a] = sum_{ci, dp} W[ci, dp, f] x[ci, a + dp * Wt],   out[a] = sum_f D_f[a + f]       (a = c * Wt + tl, as in the 7x7 kernel)
// K = 16 = two channel groups of one position (LBO = group plane), two K-halves per row tap; N = 16 = 8 phases x {W_hi, W_lo};
// both the x_hi and the x_lo MMA use the full [W_hi | W_lo] operand (N = 16 is the minimum: the extra x_lo * W_lo term is free).
// Input: chunk planes [B][G_total][R][T1][8] written by equiv_umma_kernel EPI 3.  grid (time tiles, heads, B).
struct HeadUmmaArgs {
  const __half* in_hi[3];
  const __half* in_lo[3];
  int G_total[3], g0[3], R[3], KH[3], rows_out[3];
  const __half* wimg[3];  // [KH * 2 starts][chunk 2][n 16][ci 8]
  const float* bias[3];
  float* out[3];          // (B, 1, rows_out, Tf)
  int T1, Tf, TB, n_ttiles;
};

constexpr int kHtStride = 122;  // anchors produced per 128-row block (6 rows feed the phase shifts)

__host__ __device__ inline uint32_t head_tail_group_positions(int Wt) { return (uint32_t)(23 * Wt + 136); }
__host__ __device__ inline size_t head_tail_smem_bytes(int Wt) {
  return (size_t)8 * head_tail_group_positions(Wt) * 16 + 24 * 512 + 6 * 128 * 4;
}

// W (1, 32, KH, 7) -> operand image: start j = dp * 2 + kh covers channels [16 kh, 16 kh + 16); n < 8: W_hi of time tap n
// (zero for n = 7), n >= 8: W_lo.
__global__ void head_tail_pack_kernel(const float* __restrict__ w, int KH, __half* __restrict__ img) {
  const int n_items = KH * 2 * 2 * 8 * 8;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_items; i += gridDim.x * blockDim.x) {
    const int e = i % 8, f = (i / 8) % 8, g = (i / 64) % 2, j = i / 128;
    const int dp = j / 2, kh = j % 2, ci = kh * 16 + g * 8 + e;
    const float v = f < 7 ? w[((long long)ci * KH + dp) * 7 + f] * kWScale : 0.f;
    const __half hi = __float2half_rn(v);
    const __half lo = __float2half_rn(v - __half2float(hi));
    img[((j * 2 + g) * 16 + f) * 8 + e] = hi;
    img[((j * 2 + g) * 16 + 8 + f) * 8 + e] = lo;
  }
}

__global__ void __launch_bounds__(160) head_tail_umma_kernel(const HeadUmmaArgs a) {
  using namespace umma;
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t tile_bar, acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_slot;

  const int tid = threadIdx.x, warp = uniform_warp_idx(), lane = tid & 31;
  const int h = blockIdx.y, b = blockIdx.z, t0 = blockIdx.x * a.TB;
  const int KH = a.KH[h], R = a.R[h], rows_out = a.rows_out[h];
  const int TBv = min(a.TB, a.Tf - t0);
  const int Wt = a.TB + 6;
  const int cols_in = min(Wt, a.T1 - t0);
  const int rows_in = rows_out + KH - 1;  // 23 (wrapped) or 12
  const uint32_t GPpos = head_tail_group_positions(Wt), GP = GPpos * 16;
  const int n_anchor = rows_out * Wt;
  const int n_mb = (n_anchor + kHtStride - 1) / kHtStride;
  uint8_t* s_hi = smem;            // [g 4][GP]
  uint8_t* s_lo = smem + 4 * GP;
  uint8_t* s_w = smem + 8 * GP;    // 24 starts x 512 B
  float* s_ex = reinterpret_cast<float*>(smem + 8 * GP + 24 * 512);  // [6 phases][128]

  if (warp == 4) tmem_alloc(&tmem_slot, 32);
  if (tid == 0) {
    mbar_init(&tile_bar, 1);
    mbar_init(&acc_full[0], 1), mbar_init(&acc_full[1], 1);
    mbar_init(&acc_empty[0], 128), mbar_init(&acc_empty[1], 128);
    mbar_init_fence();
  }
  {
    // zero what the bulk copies leave untouched (rows beyond rows_in, the tail padding, the column gap of a narrow last tile)
    const uint4 z = make_uint4(0, 0, 0, 0);
    const int gap = Wt - cols_in;
    for (int pl = 0; pl < 8; ++pl) {
      uint4* base = reinterpret_cast<uint4*>(smem + (size_t)pl * GP);
      for (uint32_t i = (uint32_t)rows_in * Wt + tid; i < GPpos; i += blockDim.x) base[i] = z;
      if (gap > 0)
        for (int i = tid; i < rows_in * gap; i += blockDim.x) base[(uint32_t)(i / gap) * Wt + cols_in + i % gap] = z;
    }
    fence_proxy_async();
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_slot;

  if (warp == 4) {
    // ------------------------------------------------------------ loader + MMA issuer
    const uint32_t row_bytes = (uint32_t)cols_in * 16;
    const uint32_t w_bytes = (uint32_t)KH * 2 * 512;
    if (lane == 0) mbar_arrive_expect_tx(&tile_bar, w_bytes + 8u * rows_in * row_bytes);
    __syncwarp();
    if (lane == 0) bulk_g2s(s_w, a.wimg[h], w_bytes, &tile_bar);
    for (int idx = lane; idx < 4 * rows_in; idx += 32) {
      const int g = idx / rows_in, r = idx - g * rows_in;
      const long long src = ((((long long)b * a.G_total[h] + a.g0[h] + g) * R + r) * a.T1 + t0) * 8;
      bulk_g2s(s_hi + (size_t)g * GP + (size_t)r * Wt * 16, a.in_hi[h] + src, row_bytes, &tile_bar);
      bulk_g2s(s_lo + (size_t)g * GP + (size_t)r * Wt * 16, a.in_lo[h] + src, row_bytes, &tile_bar);
    }
    // converged warp, one elected lane issues (see umma.cuh: single-lane issue)
    __syncwarp();
    mbar_wait(&tile_bar, 0);
    const uint64_t A_DESC = desc_hi(GP);          // chunk 1 = the next channel group of the same position
    constexpr uint64_t B_DESC = desc_hi(16 * 16);  // chunk stride: 16 rows x 16 B
    constexpr uint32_t IDESC = idesc_f16(16);
    const uint32_t hi0 = smem_u32(s_hi), lo0 = smem_u32(s_lo), w0 = smem_u32(s_w);
    for (int m = 0; m < n_mb; ++m) {
      const int buf = m & 1;
      mbar_wait(&acc_empty[buf], ((m >> 1) & 1) ^ 1);
      fence_after_sync();
      const uint32_t d = tmem + buf * 16;
      if (elect_one()) {
        for (int dp = 0; dp < KH; ++dp) {
#pragma unroll
          for (int kh = 0; kh < 2; ++kh) {
            const uint32_t off = (uint32_t)(m * kHtStride + dp * Wt) * 16 + (uint32_t)(2 * kh) * GP;
            const uint64_t bd = make_desc(B_DESC, w0 + (uint32_t)(dp * 2 + kh) * 512);
            mma_f16(d, make_desc(A_DESC, hi0 + off), bd, IDESC, (dp | kh) ? 1u : 0u);
            mma_f16(d, make_desc(A_DESC, lo0 + off), bd, IDESC, 1u);
          }
        }
        commit(&acc_full[buf]);
      }
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------ epilogue: thread = TMEM lane = anchor row
    const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
    const float bias = __ldg(a.bias[h]);
    for (int m = 0; m < n_mb; ++m) {
      const int buf = m & 1;
      mbar_wait(&acc_full[buf], (m >> 1) & 1);
      fence_after_sync();
      float u[16];
      tmem_ld16(lane_base + buf * 16, u);
      fence_before_sync();
      mbar_arrive(&acc_empty[buf]);
      float v[7];
#pragma unroll
      for (int f = 0; f < 7; ++f) v[f] = u[f] + u[8 + f];
#pragma unroll
      for (int f = 1; f < 7; ++f) s_ex[(f - 1) * 128 + tid] = v[f];
      asm volatile("bar.sync 1, 128;" ::: "memory");
      const int anchor = m * kHtStride + tid;
      if (tid < kHtStride && anchor < n_anchor) {
        const int c = anchor / Wt, tl = anchor - c * Wt;
        if (tl < TBv) {
          float o = v[0];
#pragma unroll
          for (int f = 1; f < 7; ++f) o += s_ex[(f - 1) * 128 + tid + f];
          a.out[h][((long long)b * rows_out + c) * a.Tf + t0 + tl] = fmaf(o, 1.f / kWScale, bias);
        }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem, 32);
}

}  // namespace ake

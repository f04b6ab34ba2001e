// cqt.cu -- constant-Q front-end (librosa.cqt call of KeyDataset.py:490-491 + abs/log1p of :497-499).
//
// The reference delegates the arithmetic to librosa 0.9.2 (vqt recursion) + resampy 0.3.1
// (kaiser_fast), neither of which is vendored; this file restates their published algorithm
// (SURVEY.md section 8 a-1):
//   for octave i = 0 (top) .. n_oct-1:   y_i = decimate2(y_{i-1}) * sqrt(2)      (63-tap kaiser_fast FIR)
//       C_i[k, t] = sqrt(2^i) * sum_n K[k, n] * y_i[t*hop_i - n_fft/2 + n]        (zero padded frames)
//   C[bin, t] = C_i[k, t] / sqrt(length(bin)),  out = log(1 + |C|)
// where K is the dense time-domain image of librosa's sparsified FFT basis (identical for every
// octave because f_k / sr_i does not depend on i).
#include <cmath>
#include <algorithm>
#include <complex>
#include <vector>

#include "common.cuh"
#include "umma.cuh"

namespace ake {

constexpr int kHalfTaps = 32;          // h[0..31]; full filter has 63 taps (|j| <= 31)
constexpr double kPi = 3.14159265358979323846;
constexpr double kHannBandwidth = 1.50018310546875;  // librosa.filters.WINDOW_BANDWIDTHS['hann']
constexpr double kBwFastest = 0.85;                   // resampy kaiser_fast rolloff (librosa.audio.BW_FASTEST)
// tensor-core filter bank (cqt_bank_umma_kernel)
constexpr int kUKB = 64;        // samples of K per pipeline stage (4 MMA k-steps of 16)
constexpr int kUStages = 2;
constexpr float kXScale = 8.f;  // audio is scaled into fp16's comfortable range; the bank by kBankScale
constexpr float kBankScale = 16.f;
constexpr float kDecScale = 16.f;  // decimator taps are scaled into fp16's normal range too


static double bessel_i0(double x) {
  double sum = 1.0, term = 1.0;
  const double q = x * x / 4.0;
  for (int k = 1; k < 200; ++k) {
    term *= q / ((double)k * k);
    sum += term;
    if (term < 1e-20 * sum) break;
  }
  return sum;
}

// resampy.filters.sinc_window(num_zeros=16, precision=9, window=kaiser(beta), rolloff=0.85) sampled at the
// polyphase positions a ratio-1/2 resample visits (every 256th table entry), times the ratio 1/2 that
// resampy.core.resample applies to the filter when down-sampling.
static void kaiser_fast_half(double* h) {
  const double beta = 8.555504641634386, rolloff = 0.85;
  const int num_zeros = 16;
  const double i0b = bessel_i0(beta);
  for (int m = 0; m < kHalfTaps; ++m) {
    const double u = 0.5 * m;                 // position in zero crossings
    const double x = rolloff * u;
    const double sinc = (m == 0) ? 1.0 : std::sin(kPi * x) / (kPi * x);
    const double r = u / num_zeros;           // 0..1 across the half window
    const double taper = bessel_i0(beta * std::sqrt(std::max(0.0, 1.0 - r * r))) / i0b;
    h[m] = 0.5 * rolloff * sinc * taper;
  }
}

}  // namespace ake

using namespace ake;

struct ake_cqt {
  double sr, fmin, filter_scale, sparsity;
  int hop, n_bins, bpo, n_oct, n_fft;
  std::vector<double> dec_half;    // kaiser_fast half filter (without the sqrt(2) gain)
  std::vector<float> bank;         // (2*bpo, n_fft): row 2k = Re K_k, 2k+1 = Im K_k
  std::vector<float> out_scale;    // (n_oct, bpo): sqrt(2^i) / sqrt(length of the full-rate bin)
  float* d_scale = nullptr;
  // tensor-core path: fp16 (hi | lo) image of the bank in the shared-memory operand layout, one block per 64 samples of K
  __half* d_bank_img = nullptr;
  __half* d_dec_img = nullptr;   // Toeplitz image of the 63-tap decimator (cascade_umma_kernel)
  float* d_scale_umma = nullptr;
  int npad = 0;  // filters per MMA (2*bpo rounded up to 16), 0: tensor-core path not available for this shape
};

namespace ake {

static int two_factors(int x) {
  int n = 0;
  while (x > 0 && x % 2 == 0) x /= 2, ++n;
  return n;
}

static void build_cqt(ake_cqt* p) {
  const int bpo = p->bpo, n_bins = p->n_bins;
  if (p->sr <= 0 || p->hop <= 0 || n_bins <= 0 || bpo <= 0) fail(AKE_ERR_INVALID, "sr, hop_length, n_bins, bins_per_octave must be positive");
  if (n_bins % bpo) fail(AKE_ERR_UNSUPPORTED, "n_bins must be a multiple of bins_per_octave");
  if (!(p->sparsity >= 0.0 && p->sparsity < 1.0)) fail(AKE_ERR_INVALID, "sparsity must be in [0, 1)");
  if (p->fmin <= 0) p->fmin = 32.70319566257483;  // note_to_hz('C1')
  p->n_oct = n_bins / bpo;
  const double alpha = std::pow(2.0, 1.0 / bpo) - 1.0;
  const double Q = p->filter_scale / alpha;
  std::vector<double> freqs(n_bins);
  for (int k = 0; k < n_bins; ++k) freqs[k] = p->fmin * std::pow(2.0, (double)k / bpo);
  const double fmin_t = freqs[n_bins - bpo], fmax_t = freqs[n_bins - 1];
  const double nyquist = p->sr / 2.0;
  // librosa.filters.constant_q_lengths: ParameterError when the top filter passes Nyquist
  if (fmax_t * (1 + 0.5 * kHannBandwidth / Q) > nyquist)
    fail(AKE_ERR_INVALID, "filter pass-band lies beyond Nyquist (fmax %.1f Hz, sr %.1f)", fmax_t, p->sr);
  const double filter_cutoff = fmax_t * (1 + 0.5 * kHannBandwidth / Q);
  if (!(filter_cutoff < kBwFastest * nyquist))
    fail(AKE_ERR_UNSUPPORTED, "top octave would need kaiser_best resampling (cutoff %.1f Hz); only the kaiser_fast recursion is built", filter_cutoff);
  // librosa.core.constantq.__early_downsample_count
  const int c1 = std::max(0, (int)(std::ceil(std::log2(kBwFastest * nyquist / filter_cutoff)) - 1) - 1);
  const int num_twos = two_factors(p->hop);
  const int c2 = std::max(0, num_twos - p->n_oct + 1);
  if (std::min(c1, c2) > 0) fail(AKE_ERR_UNSUPPORTED, "this sr/hop would early-downsample in librosa; not built");
  if (num_twos < p->n_oct - 1)
    fail(AKE_ERR_INVALID, "hop_length must be a positive integer multiple of 2^%d for %d-octave CQT", p->n_oct - 1, p->n_oct);

  // ---- filters.constant_q for the top octave (lengths are identical for every octave)
  std::vector<double> lengths(bpo);
  double max_len = 0;
  for (int k = 0; k < bpo; ++k) lengths[k] = Q * p->sr / (fmin_t * std::pow(2.0, (double)k / bpo)), max_len = std::max(max_len, lengths[k]);
  p->n_fft = 1 << (int)std::ceil(std::log2(max_len));
  const int N = p->n_fft, NF = N / 2 + 1;
  std::vector<std::complex<double>> tw(N);
  for (int m = 0; m < N; ++m) tw[m] = std::polar(1.0, -2.0 * kPi * m / N);
  p->bank.assign((size_t)2 * bpo * N, 0.f);
  std::vector<std::complex<double>> sig, spec(NF);
  std::vector<double> mags(NF), sorted(NF);
  for (int k = 0; k < bpo; ++k) {
    const double ilen = lengths[k], freq = fmin_t * std::pow(2.0, (double)k / bpo);
    const double start = std::floor(-ilen / 2.0), stop = std::floor(ilen / 2.0);  // np.arange(-ilen//2, ilen//2)
    const int len = (int)std::ceil(stop - start);
    sig.assign(len, 0.0);
    double wsum = 0;
    for (int m = 0; m < len; ++m) {
      const double w = 0.5 - 0.5 * std::cos(2.0 * kPi * m / len);  // periodic hann
      sig[m] = std::polar(1.0, (start + m) * 2.0 * kPi * freq / p->sr) * w;
      wsum += std::abs(sig[m]);
    }
    const int lpad = (N - len) / 2;  // util.pad_center
    const double gain = (ilen / N) / wsum;  // L1 normalisation, then basis *= lengths / n_fft
    for (int f = 0; f < NF; ++f) {
      std::complex<double> acc = 0;
      for (int m = 0; m < len; ++m) acc += sig[m] * tw[(int)(((long long)f * (lpad + m)) % N)];
      spec[f] = acc * gain;
      mags[f] = std::abs(spec[f]);
    }
    // util.sparsify_rows(quantile=sparsity): drop the smallest bins holding < quantile of the L1 mass
    sorted = mags;
    std::sort(sorted.begin(), sorted.end());
    double norm = 0;
    for (double m : mags) norm += m;
    double cum = 0, thresh = sorted[0];
    for (int f = 0; f < NF; ++f) {
      cum += sorted[f] / norm;
      if (!(cum < p->sparsity)) {
        thresh = sorted[f];
        break;
      }
    }
    for (int f = 0; f < NF; ++f) {
      if (mags[f] >= thresh) spec[f] = std::complex<double>((float)spec[f].real(), (float)spec[f].imag());  // complex64 basis
      else spec[f] = 0;
    }
    // dense time-domain image: K[n] = sum_f B[f] * exp(-2 pi i f n / N)
    for (int n = 0; n < N; ++n) {
      std::complex<double> acc = 0;
      for (int f = 0; f < NF; ++f)
        if (spec[f] != 0.0) acc += spec[f] * tw[(int)(((long long)f * n) % N)];
      p->bank[(size_t)(2 * k) * N + n] = (float)acc.real();
      p->bank[(size_t)(2 * k + 1) * N + n] = (float)acc.imag();
    }
  }
  // fft_basis *= sqrt(2^i); V /= sqrt(constant_q_lengths at the full rate)
  p->out_scale.resize((size_t)p->n_oct * bpo);
  for (int i = 0; i < p->n_oct; ++i)
    for (int k = 0; k < bpo; ++k) {
      const int bin = n_bins - bpo * (i + 1) + k;
      const double full_len = Q * p->sr / freqs[bin];
      p->out_scale[(size_t)i * bpo + k] = (float)(std::sqrt(std::pow(2.0, i)) / std::sqrt(full_len));
    }
  p->dec_half.resize(kHalfTaps);
  kaiser_fast_half(p->dec_half.data());
}

// Device copies are made lazily so that plan creation (and the bank / tap getters) need no GPU.
static void ensure_device(ake_cqt* p) {
  if (p->d_scale) return;
  {
    // Toeplitz operand of the decimator (cascade_umma_kernel): H[n][k] = sqrt(2) h[|k - 2n - 32|], fp16 hi | lo.
    // (sqrt(2): resample(scale=True) divides by sqrt(ratio).)
    std::vector<__half> img((size_t)16 * 64 * 8, __float2half(0.f));
    for (int n = 0; n < 32; ++n)
      for (int k = 0; k < 128; ++k) {
        const int m = std::abs(k - 2 * n - 32);
        if (m >= kHalfTaps) continue;
        const float v = (float)(p->dec_half[m] * std::sqrt(2.0) * kDecScale);
        const __half hi = __float2half_rn(v);
        const __half lo = __float2half_rn(v - __half2float(hi));
        const size_t c = k / 8, e = k % 8;
        img[(c * 64 + n) * 8 + e] = hi;
        img[(c * 64 + 32 + n) * 8 + e] = lo;
      }
    AKE_CUDA(cudaMalloc(&p->d_dec_img, sizeof(__half) * img.size()));
    AKE_CUDA(cudaMemcpy(p->d_dec_img, img.data(), sizeof(__half) * img.size(), cudaMemcpyHostToDevice));
  }
  AKE_CUDA(cudaMalloc(&p->d_scale, sizeof(float) * p->out_scale.size()));
  AKE_CUDA(cudaMemcpy(p->d_scale, p->out_scale.data(), sizeof(float) * p->out_scale.size(), cudaMemcpyHostToDevice));
  // tensor-core operand image: per 64-sample block of K, 8 chunks x (2*npad) rows x 8 halves; rows [0,npad) = hi, [npad,2npad) = lo
  const int nf = 2 * p->bpo;
  const int npad = (nf + 15) / 16 * 16;
  if ((npad == 80 || npad == 32) && p->n_fft % kUKB == 0) {
    p->npad = npad;
    const int n_kb = p->n_fft / kUKB;
    std::vector<__half> img((size_t)n_kb * 8 * 2 * npad * 8, __float2half(0.f));
    for (int f = 0; f < nf; ++f)
      for (int k = 0; k < p->n_fft; ++k) {
        const float v = p->bank[(size_t)f * p->n_fft + k] * kBankScale;
        const __half hi = __float2half_rn(v);
        const __half lo = __float2half_rn(v - __half2float(hi));
        const size_t blk = (size_t)(k / kUKB) * 8 * 2 * npad * 8, c = (k % kUKB) / 8, e = k % 8;
        img[blk + (c * 2 * npad + f) * 8 + e] = hi;
        img[blk + (c * 2 * npad + npad + f) * 8 + e] = lo;
      }
    AKE_CUDA(cudaMalloc(&p->d_bank_img, sizeof(__half) * img.size()));
    AKE_CUDA(cudaMemcpy(p->d_bank_img, img.data(), sizeof(__half) * img.size(), cudaMemcpyHostToDevice));
    std::vector<float> sc(p->out_scale);
    for (float& v : sc) v /= (kXScale * kBankScale);
    AKE_CUDA(cudaMalloc(&p->d_scale_umma, sizeof(float) * sc.size()));
    AKE_CUDA(cudaMemcpy(p->d_scale_umma, sc.data(), sizeof(float) * sc.size(), cudaMemcpyHostToDevice));
  }
}

static inline long long len_at(long long n0, int i) { return (n0 + (1LL << i) - 1) >> i; }  // ceil(n0 / 2^i)

static int frames_for(const ake_cqt* p, long long n) {
  long long T = -1;
  for (int i = 0; i < p->n_oct; ++i) {
    const long long t = 1 + len_at(n, i) / (p->hop >> i);
    T = (T < 0 || t < T) ? t : T;
  }
  return (int)T;
}

// ------------------------------------------------------------------------------------------ kernels
// ---- resampling cascade on tensor cores (tcgen05) -------------------------------------------------------------------
// One octave step is  out[t] = sqrt(2) * sum_{|j|<=31} h[|j|] * in[2t + j]  (zero extended) for t < floor(n_in/2);
// librosa pads the result to ceil(n_in/2) samples with a zero.  As a GEMM: a "row" is 64 consecutive input samples and
// yields 32 outputs,
//     D[r, n] = sum_{k<128} X[r, k] * H[n, k],   X[r, k] = in[base + 64 r + k],   H[n, k] = h[|k - 2n - 32|]   (Toeplitz)
// with X ~= Xh + Xl and H ~= Hh + Hl in fp16 (fp32 accumulation in TMEM, ~22 significant bits as in the filter bank):
//     MMA 1: A = Xh, B = [Hh | Hl] (N = 64);   MMA 2: A = Xl, B = Hh (N = 32);   out[r, n] = D[r, n] + D[r, 32 + n].
// Shared-memory operand ("transposed chunk planes"): the 16-byte chunk q (samples 8q .. 8q+7) of the tile's input span
// lives at  plane[q % 8] + (q / 8) * 16,  so operand row r = 64 samples is one 16-byte slot per plane, chunks 0..7 of a
// row are the 8 planes (LBO = plane pitch) and chunks 8..15 the same planes one row further (start address + 16 B).
// One CTA tile fuses TWO octave steps: 129 input rows -> 128 rows of level p+1 (4096 samples, kept in shared memory as
// the next operand and written to global memory) -> 63 rows of level p+2 (2016 samples).  Each tile recomputes a halo of
// 32 + 2*32 input samples per side (2.4 %) instead of exchanging state with its neighbours.  Persistent CTAs, 2 per SM; the next tile's
// span is landed in shared memory by cp.async while the current tile runs its MMAs and epilogues.
constexpr int kCasRows2 = 63;                       // level p+2 rows (of 32 outputs) per tile
constexpr int kCasOwn2 = kCasRows2 * 32;            // 2016 level p+2 outputs owned by a tile
constexpr int kCasOwn1 = 2 * kCasOwn2;              // 4032 level p+1 outputs owned by a tile (rows 1..126 of 128)
constexpr int kCasP0Rows = 129, kCasP1Rows = 65;    // plane rows (16 B each) of the level p / level p+1 operands
constexpr uint32_t kCasLBO0 = kCasP0Rows * 16, kCasLBO1 = kCasP1Rows * 16;  // odd multiples of 16 B: conflict-free scatter
constexpr uint32_t kCasP0Bytes = 8 * kCasLBO0, kCasP1Bytes = 8 * kCasLBO1;
constexpr uint32_t kCasImgBytes = 16 * 64 * 16;     // Toeplitz image: [chunk 16][n 64 = Hh 32 | Hl 32][8 halves]
constexpr uint32_t kCasSmem = 2 * kCasP1Bytes + 2 * kCasP0Bytes + kCasImgBytes;

struct CascadeArgs {
  const float* in;      // level p, clip b at in + b * in_stride
  long long in_stride;
  float* out1;          // level p+1
  long long stride1;
  float* out2;          // level p+2 (unused when n_levels == 1)
  long long stride2;
  const long long* lengths;  // full-rate samples per clip, or NULL (= n_uniform)
  long long n_uniform;
  int level_in, n_levels, tiles_per_clip, n_tiles;
  const __half* img;
};

// (a, b) -> fp16 pairs hi, lo with a ~= hi + lo
__device__ __forceinline__ void cas_split2(uint64_t x, uint32_t& hi, uint32_t& lo) {
  float x0, x1;
  f2_unpack(x, x0, x1);
  const __half2 h = __floats2half2_rn(x0, x1);
  const float2 hf = __half22float2(h);
  float d0, d1;
  f2_unpack(f2_sub(x, f2_pack(hf.x, hf.y)), d0, d1);
  const __half2 l = __floats2half2_rn(d0, d1);
  hi = *reinterpret_cast<const uint32_t*>(&h), lo = *reinterpret_cast<const uint32_t*>(&l);
}
template <bool SCALE>
__device__ __forceinline__ void cas_store_split8(uint8_t* hi_dst, uint8_t* lo_dst, const float (&v)[8], float scale) {
  uint32_t h[4], l[4];
  const uint64_t ss = f2_pack(scale, scale);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    uint64_t x = f2_pack(v[2 * e], v[2 * e + 1]);
    if (SCALE) x = f2_mul(x, ss);
    cas_split2(x, h[e], l[e]);
  }
  *reinterpret_cast<uint4*>(hi_dst) = make_uint4(h[0], h[1], h[2], h[3]);
  *reinterpret_cast<uint4*>(lo_dst) = make_uint4(l[0], l[1], l[2], l[3]);
}

constexpr int kCasThreads = 256;  // warps w and w+4 share accumulator lanes 32 (w % 4) ..: each takes 16 of a row's 32 outputs
constexpr int kCasChunks = kCasP0Rows * 8;                 // 1032 chunks of 8 samples per tile span
constexpr int kCasSpan = kCasChunks * 8;                   // 8256 input samples per tile
constexpr int kCasPer = (kCasChunks + kCasThreads - 1) / kCasThreads;  // chunks per thread (the last round holds 8)
constexpr int kCasPer4 = (2 * kCasChunks + kCasThreads - 1) / kCasThreads;  // float4s per thread
constexpr uint32_t kCasStageBytes = kCasSpan * 4;          // fp32 landing buffer of the next tile's span (bulk async copy)
constexpr int kCasTPitch = 36;                             // floats per row of the store-transposition buffers (144 B: conflict-free)
constexpr uint32_t kCasT2Bytes = kCasRows2 * kCasTPitch * 4;  // level p+2 outputs on their way to coalesced global stores
constexpr uint32_t kCasSmemTotal = kCasSmem + kCasStageBytes + kCasT2Bytes;
static_assert(128 * kCasTPitch * 4 <= 2 * kCasP0Bytes, "the level p+1 transposition buffer aliases the level-p planes");

// Software pipeline of one persistent CTA (two per SM), tile i, next tile j:
//   wait MMA1(i) -> epilogue 1(i): level p+1 to global memory and, as fp16 hi/lo planes, to shared memory
//   issue MMA2(i)                     | meanwhile: convert the landed span of tile j into the level-p planes
//   issue MMA1(j), start the bulk copy of the tile after j | meanwhile: wait MMA2(i) -> epilogue 2(i): level p+2 to global memory
// Levels 1 and 2 accumulate in separate TMEM columns so MMA1(j) can overlap epilogue 2(i).
__global__ void __launch_bounds__(kCasThreads, 2) cascade_umma_kernel(const CascadeArgs a) {
  using namespace umma;
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t img_bar, bar1, bar2, stage_bar;
  __shared__ uint32_t tmem_slot;
  // Order matters: the level p+2 MMA runs M = 128 over 63 useful rows; its surplus rows read on into the next buffer,
  // which must hold finite fp16 data (their products only reach accumulator rows that are never read back).
  uint8_t* p1h = smem;
  uint8_t* p1l = smem + kCasP1Bytes;
  uint8_t* p0h = smem + 2 * kCasP1Bytes;
  uint8_t* p0l = p0h + kCasP0Bytes;
  uint8_t* img = p0l + kCasP0Bytes;
  float* stage = reinterpret_cast<float*>(img + kCasImgBytes);
  float* tbuf2 = reinterpret_cast<float*>(img + kCasImgBytes + kCasStageBytes);
  float* tbuf1 = reinterpret_cast<float*>(p0h);  // free between MMA1 of a tile and the conversion of the next one
  const int tid = threadIdx.x, warp = uniform_warp_idx(), lane = tid & 31;
  const int row = 32 * (warp & 3) + lane;  // accumulator lane = row of 32 outputs
  const int hf = warp >> 2;                // which 16 of them this thread handles
  const bool two = a.n_levels == 2;

  if (warp == 0) tmem_alloc(&tmem_slot, 128);
  if (tid == 0) {
    mbar_init(&img_bar, 1), mbar_init(&bar1, 1), mbar_init(&bar2, 1), mbar_init(&stage_bar, 1);
    mbar_init_fence();
  }
  for (uint32_t i = tid; i < (2 * kCasP1Bytes + 2 * kCasP0Bytes) / 16; i += kCasThreads) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  if (tid == 0) {
    mbar_arrive_expect_tx(&img_bar, kCasImgBytes);
    bulk_g2s(img, a.img, kCasImgBytes, &img_bar);
  }
  const uint32_t tmem = tmem_slot;
  const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16) + 16 * hf;
  uint32_t ph1 = 0, ph2 = 0, phs = 0;
  // accumulator -> sample: level p+1 carries kXScale * kDecScale, level p+2 one more kDecScale (the level p+1 operand is the
  // raw accumulator: no rescaling between the two steps)
  constexpr float kInv1 = 1.f / (kXScale * kDecScale), kInv2 = kInv1 / kDecScale;

  auto issue_level = [&](uint32_t d, uint32_t hi0, uint32_t lo0, uint32_t lbo, uint64_t* bar) {
    const uint64_t a_desc = desc_hi(lbo);
    constexpr uint64_t B_DESC = desc_hi(64 * 16);
    constexpr uint32_t IDESC64 = idesc_f16(64), IDESC32 = idesc_f16(32);
    const uint32_t w0 = smem_u32(img);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint32_t off = (uint32_t)((2 * j) & 7) * lbo + (j >= 4 ? 16u : 0u);  // chunks 8..15 = planes 0..7, one row on
      const uint64_t bd = make_desc(B_DESC, w0 + (uint32_t)j * 2048);
      mma_f16(d, make_desc(a_desc, hi0 + off), bd, IDESC64, j ? 1u : 0u);
      mma_f16(d, make_desc(a_desc, lo0 + off), bd, IDESC32, 1u);
    }
    commit(bar);
  };
  // this thread's 16 accumulator columns: out[n] = D[n] + D[32 + n]
  auto read_acc = [&](uint32_t col0, float (&o)[16]) {
    float w[16];
    tmem_ld16(lane_base + col0, o);
    tmem_ld16(lane_base + col0 + 32, w);
#pragma unroll
    for (int j = 0; j < 16; ++j) o[j] += w[j];
  };
  auto clip_len = [&](int b) {  // samples of level p in clip b
    const long long n0 = a.lengths ? a.lengths[b] : a.n_uniform;
    return (n0 + (1LL << a.level_in) - 1) >> a.level_in;
  };
  // first tile at or after `tile` (in this CTA's stride) that has data; tiles beyond a short clip's end write nothing
  auto next_tile = [&](int tile) {
    for (; tile < a.n_tiles; tile += gridDim.x) {
      const int b = tile / a.tiles_per_clip, t = tile - b * a.tiles_per_clip;
      if ((long long)kCasOwn1 * t < ((clip_len(b) + 1) >> 1)) break;
    }
    return tile;
  };
  struct Span {
    const float* src;  // first sample of the span (may lie before the clip: zero extension)
    int vlo, vhi;      // samples [vlo, vhi) of the span exist
    bool bulk;         // 16-byte aligned: landed by one bulk async copy
  };
  auto span_of = [&](int tile) {
    const int b = tile / a.tiles_per_clip, t = tile - b * a.tiles_per_clip;
    const long long base0 = 2 * ((long long)kCasOwn1 * t - 32) - 32;
    Span s;
    s.src = a.in + (long long)b * a.in_stride + base0;
    s.vlo = (int)max(0LL, -base0), s.vhi = (int)min((long long)kCasSpan, clip_len(b) - base0);
    s.bulk = (reinterpret_cast<uintptr_t>(s.src) & 15) == 0;
    return s;
  };
  // Start landing a tile's span in the stage (call when no thread reads the stage any more).  Aligned clips: thread 0 sends
  // one bulk copy of the existing samples (whole float4s) and patches the <= 3 tail samples; otherwise every thread fetches
  // the chunks it will convert itself.  Samples outside [vlo, vhi) are masked at conversion time, not written here.
  auto start_stage = [&](int tile) {
    const Span s = span_of(tile);
    if (s.bulk) {
      if (tid == 0) {
        const int n4 = (s.vhi - s.vlo) & ~3;
        for (int i = s.vlo + n4; i < s.vhi; ++i) stage[i] = __ldg(s.src + i);
        if (n4 > 0) {
          mbar_arrive_expect_tx(&stage_bar, (uint32_t)n4 * 4);
          bulk_g2s(stage + s.vlo, s.src + s.vlo, (uint32_t)n4 * 4, &stage_bar);
        } else {
          mbar_arrive(&stage_bar);
        }
      }
    } else {
      for (int i = 0; i < kCasPer; ++i) {
        const int q = tid + kCasThreads * i;
        if (q < kCasChunks)
          for (int e = 0; e < 8; ++e)
            if (8 * q + e >= s.vlo && 8 * q + e < s.vhi) stage[8 * q + e] = __ldg(s.src + 8 * q + e);
      }
      if (tid == 0) mbar_arrive(&stage_bar);
    }
  };
  // stage -> level-p operand planes (fp16 hi/lo, transposed chunk planes)
  auto convert = [&](int tile) {
    const Span s = span_of(tile);
    mbar_wait(&stage_bar, phs);
    phs ^= 1;
    const bool interior = s.vlo == 0 && s.vhi == kCasSpan;
    const uint64_t ss = f2_pack(kXScale, kXScale);
    // one float4 per thread and round: consecutive lanes read consecutive 16 B of the stage and write 8-byte halves of the
    // operand chunks (both conflict-free)
#pragma unroll
    for (int i = 0; i < kCasPer4; ++i) {
      const int f = tid + kCasThreads * i;
      if (f < 2 * kCasChunks) {
        const float4 v = *reinterpret_cast<const float4*>(stage + 4 * f);
        float x[4] = {v.x, v.y, v.z, v.w};
        if (!interior) {
#pragma unroll
          for (int e = 0; e < 4; ++e) x[e] = (4 * f + e >= s.vlo && 4 * f + e < s.vhi) ? x[e] : 0.f;
        }
        uint32_t h[2], l[2];
        cas_split2(f2_mul(f2_pack(x[0], x[1]), ss), h[0], l[0]);
        cas_split2(f2_mul(f2_pack(x[2], x[3]), ss), h[1], l[1]);
        const int q = f >> 1;
        const uint32_t off = (uint32_t)(q & 7) * kCasLBO0 + (uint32_t)(q >> 3) * 16 + (uint32_t)(f & 1) * 8;
        *reinterpret_cast<uint2*>(p0h + off) = make_uint2(h[0], h[1]);
        *reinterpret_cast<uint2*>(p0l + off) = make_uint2(l[0], l[1]);
      }
    }
  };

  int tile = next_tile(blockIdx.x);
  if (tile < a.n_tiles) {
    start_stage(tile);
    convert(tile);
    fence_proxy_async();
    __syncthreads();
    const int nxt = next_tile(tile + gridDim.x);
    if (nxt < a.n_tiles) start_stage(nxt);
    if (warp == 0) {  // converged warp, one elected lane issues (see umma.cuh: single-lane issue)
      mbar_wait(&img_bar, 0);
      fence_after_sync();
      if (elect_one()) issue_level(tmem, smem_u32(p0h), smem_u32(p0l), kCasLBO0, &bar1);
      __syncwarp();
    }
  } else if (warp == 0) {
    mbar_wait(&img_bar, 0);  // never leave with a bulk copy in flight
  }
  while (tile < a.n_tiles) {
    const int b = tile / a.tiles_per_clip, t = tile - b * a.tiles_per_clip;
    const long long n_p = clip_len(b);
    const long long n_half1 = n_p >> 1, n1 = (n_p + 1) >> 1, n_half2 = n1 >> 1, n2 = (n1 + 1) >> 1;
    const long long o_lo2 = (long long)kCasOwn2 * t, o_lo1 = (long long)kCasOwn1 * t - 32;
    const int nxt = next_tile(tile + gridDim.x);

    // ---- level p+1
    mbar_wait(&bar1, ph1);
    ph1 ^= 1;
    fence_after_sync();
    {
      float o[16];
      read_acc(0, o);
      const long long i1 = o_lo1 + 32LL * row + 16 * hf;  // level p+1 index of o[0]
      if (o_lo1 < 0 || o_lo1 + 128 * 32 > n_half1) {
        // samples that do not exist (index < 0 or >= floor(n_p / 2)) are zero
        const int zlo = (int)min(16LL, max(0LL, -i1)), zhi = (int)min(16LL, max(0LL, n_half1 - i1));
#pragma unroll
        for (int n = 0; n < 16; ++n) o[n] = (n >= zlo && n < zhi) ? o[n] : 0.f;
      }
      // fp32 outputs go through shared memory so that the global stores are row-contiguous (a thread owns a row: storing
      // straight from registers would touch 32 different lines per instruction)
#pragma unroll
      for (int q = 0; q < 4; ++q)
        *reinterpret_cast<float4*>(tbuf1 + row * kCasTPitch + 16 * hf + 4 * q) =
            make_float4(o[4 * q] * kInv1, o[4 * q + 1] * kInv1, o[4 * q + 2] * kInv1, o[4 * q + 3] * kInv1);
      if (two) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          float v[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] = o[8 * i + e];
          const uint32_t off = (uint32_t)(4 * (row & 1) + 2 * hf + i) * kCasLBO1 + (uint32_t)(row >> 1) * 16;
          cas_store_split8<false>(p1h + off, p1l + off, v, 1.f);
        }
      }
    }
    fence_before_sync();
    fence_proxy_async();
    __syncthreads();
    if (two && warp == 0) {
      fence_after_sync();
      if (elect_one()) issue_level(tmem + 64, smem_u32(p1h), smem_u32(p1l), kCasLBO1, &bar2);
      __syncwarp();
    }
    {
      // rows 1..126 are owned (0 and 127 are the halo the next level needs): 1008 float4s, 8 per row
      float* dst = a.out1 + (long long)b * a.stride1 + o_lo1;
      const long long lim = n1 - o_lo1;  // row-local indices < lim exist (the buffers are padded to whole rows of 32)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int idx = tid + kCasThreads * i;
        const int r = 1 + (idx >> 3), c4 = idx & 7;
        if (idx < 126 * 8 && 32LL * r < lim)
          *reinterpret_cast<float4*>(dst + 32 * r + 4 * c4) = *reinterpret_cast<const float4*>(tbuf1 + r * kCasTPitch + 4 * c4);
      }
    }
    __syncthreads();  // tbuf1 aliases the level-p planes the conversion below overwrites
    // ---- next tile's operand while MMA2 runs (the level-p planes are free: MMA1 of this tile has completed)
    if (nxt < a.n_tiles) convert(nxt);
    fence_proxy_async();
    __syncthreads();
    if (nxt < a.n_tiles) {
      const int nxt2 = next_tile(nxt + gridDim.x);
      if (nxt2 < a.n_tiles) start_stage(nxt2);  // every thread has consumed the stage
      if (warp == 0) {
        fence_after_sync();
        if (elect_one()) issue_level(tmem, smem_u32(p0h), smem_u32(p0l), kCasLBO0, &bar1);
        __syncwarp();
      }
    }
    // ---- level p+2 (overlaps MMA1 of the next tile)
    if (two) {
      mbar_wait(&bar2, ph2);
      ph2 ^= 1;
      fence_after_sync();
      if ((warp & 3) < 2) {  // rows 0..62 live in accumulator lanes 0..63 (tcgen05.ld is warp-collective: no per-thread predicate)
        float o[16];
        read_acc(64, o);
        const long long i2 = o_lo2 + 32LL * row + 16 * hf;
        const int zhi = (int)min(16LL, max(0LL, n_half2 - i2));
        if (row < kCasRows2) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float4 v;
            v.x = (4 * q + 0 < zhi) ? o[4 * q + 0] * kInv2 : 0.f;
            v.y = (4 * q + 1 < zhi) ? o[4 * q + 1] * kInv2 : 0.f;
            v.z = (4 * q + 2 < zhi) ? o[4 * q + 2] * kInv2 : 0.f;
            v.w = (4 * q + 3 < zhi) ? o[4 * q + 3] * kInv2 : 0.f;
            *reinterpret_cast<float4*>(tbuf2 + row * kCasTPitch + 16 * hf + 4 * q) = v;
          }
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");  // warps 0, 1, 4, 5
        float* dst = a.out2 + (long long)b * a.stride2 + o_lo2;
        const long long lim = n2 - o_lo2;
        const int t128 = (warp >> 2) * 64 + (warp & 1) * 32 + lane;  // 0..127 over the four warps
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int idx = t128 + 128 * i;
          const int r = idx >> 3, c4 = idx & 7;
          if (idx < kCasRows2 * 8 && 32LL * r < lim)
            *reinterpret_cast<float4*>(dst + 32 * r + 4 * c4) = *reinterpret_cast<const float4*>(tbuf2 + r * kCasTPitch + 4 * c4);
        }
      }
      fence_before_sync();  // ordered before the next MMA2 by the barrier that precedes its issue
    }
    tile = nxt;
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 128);
}

// ---- tensor-core filter bank (tcgen05): one CTA = 128 frames of one octave x all filters x all of K ----------------
//   D[frame, n] = sum_k x[frame, k] * K[n, k],  x ~= xh + xl, K ~= Kh + Kl (fp16 pairs, fp32 accumulation in TMEM)
//   MMA 1: A = xh, B = [Kh | Kl]  (N = 2*NPAD)  -> columns [0,NPAD) += xh*Kh, [NPAD,2*NPAD) += xh*Kl
//   MMA 2: A = xl, B =  Kh        (N =   NPAD)  -> columns [0,NPAD) += xl*Kh
// Warps 0-3: stage frames (fp32 -> fp16 hi/lo, operand layout of umma.cuh) and run the |.|, scale, log1p epilogue out
// of TMEM; warp 4 lane 0 issues the MMAs; the bank block of each stage arrives by one bulk async copy.

struct BankArgs {
  const float* level[16];
  long long stride[16];
  const long long* lengths;
  long long n_uniform;
  int n_oct, hop0, n_fft, B, T_max, n_bins, bpo, mode;
  const __half* bank_img;
  const float* scale;
  float* out;
};

constexpr uint32_t kBankLBO = 129 * 16;                  // chunk pitch of the frame operand: odd multiple of 16 B (conflict-free 8-byte scatter)
constexpr uint32_t kBankAHalf = (kUKB / 8) * kBankLBO;   // one of {hi, lo}: 8 chunks x 128 rows x 16 B (+ pad)

template <int NPAD>
__global__ void __launch_bounds__(160) cqt_bank_umma_kernel(const BankArgs a) {
  using namespace umma;
  constexpr uint32_t A_HALF = kBankAHalf;
  constexpr uint32_t B_BYTES = (kUKB / 8) * 2 * NPAD * 16;   // 8 chunks x 2*NPAD rows x 16 B
  constexpr uint32_t STAGE = 2 * A_HALF + B_BYTES;
  constexpr uint32_t TMEM_COLS = (2 * NPAD <= 64) ? 64 : ((2 * NPAD <= 128) ? 128 : 256);
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t full_bar[kUStages], empty_bar[kUStages], done_bar;
  __shared__ uint32_t tmem_slot;
  __shared__ long long s_g0[128];   // index (into the octave's level array) of the first sample of frame row r
  __shared__ int2 s_valid[128];     // samples [x, y) of that frame exist (the rest is the zero padding of centred frames)

  const int tid = threadIdx.x, warp = uniform_warp_idx(), lane = tid & 31;
  const int octave = blockIdx.y;
  const long long m0 = (long long)blockIdx.x * 128;
  const int n_kb = a.n_fft / kUKB;

  if (warp == 4) tmem_alloc(&tmem_slot, TMEM_COLS);
  if (tid == 0) {
    for (int s = 0; s < kUStages; ++s) mbar_init(&full_bar[s], 128), mbar_init(&empty_bar[s], 1);
    mbar_init(&done_bar, 1);
    mbar_init_fence();
  }
  // ---- frame row `tid`: which clip / frame, where its samples live
  bool in_range = false, real_frame = false;
  int b = 0, t = 0;
  if (warp < 4) {
    const long long m = m0 + tid;
    in_range = m < (long long)a.B * a.T_max;
    b = in_range ? (int)(m / a.T_max) : 0, t = in_range ? (int)(m % a.T_max) : 0;
    const long long n0 = a.lengths ? a.lengths[b] : a.n_uniform;
    long long T = -1;
    for (int i = 0; i < a.n_oct; ++i) {
      const long long ti = 1 + ((n0 + (1LL << i) - 1) >> i) / (a.hop0 >> i);
      T = (T < 0 || ti < T) ? ti : T;
    }
    real_frame = in_range && t < T;
    const long long len = (n0 + (1LL << octave) - 1) >> octave;
    const long long first = (long long)t * (a.hop0 >> octave) - a.n_fft / 2;  // centred frame, zero padded (pad_mode='constant')
    s_g0[tid] = (long long)b * a.stride[octave] + first;
    s_valid[tid] = real_frame ? make_int2((int)max(0LL, -first), (int)max(0LL, min((long long)a.n_fft, len - first))) : make_int2(0, 0);
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_slot;

  if (warp < 4) {
    // ---------------------------------------------------------------- producer
    // A warp stages 32 frame rows; one instruction covers two rows x 64 samples: lanes 0-15 read the 16 float4s of one row,
    // lanes 16-31 those of the next (coalesced 256-byte runs), convert to fp16 hi/lo and scatter 8-byte halves of the
    // operand chunks (chunk c of row r at c * kBankLBO + r * 16).
    const float* level = a.level[octave];
    const bool base_aligned = (reinterpret_cast<uintptr_t>(level) & 15) == 0;
    const int f = lane & 15;
    const uint64_t ss = f2_pack(kXScale, kXScale);
    for (int kb = 0; kb < n_kb; ++kb) {
      const int s = kb % kUStages;
      const uint32_t phase = (kb / kUStages) & 1;
      mbar_wait(&empty_bar[s], phase ^ 1);
      uint8_t* stage = smem + (size_t)s * STAGE;
      if (tid == 0) {
        mbar_arrive_expect_tx(&full_bar[s], B_BYTES);
        bulk_g2s(stage + 2 * A_HALF, reinterpret_cast<const uint8_t*>(a.bank_img) + (size_t)kb * B_BYTES, B_BYTES, &full_bar[s]);
      }
      const int i0 = kb * kUKB + 4 * f;  // frame-local index of this lane's first sample
      // all 16 loads are issued before the first conversion consumes one (the branch below must not serialise them)
      float x[16][4];
#pragma unroll
      for (int it = 0; it < 16; ++it) {
        const int r = warp * 32 + 2 * it + (lane >> 4);
        const long long g = s_g0[r] + i0;
        const int2 v = s_valid[r];
        if (i0 >= v.x && i0 + 4 <= v.y && base_aligned && (g & 3) == 0) {
          const float4 q = __ldg(reinterpret_cast<const float4*>(level + g));
          x[it][0] = q.x, x[it][1] = q.y, x[it][2] = q.z, x[it][3] = q.w;
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e) x[it][e] = (i0 + e >= v.x && i0 + e < v.y) ? __ldg(level + g + e) : 0.f;
        }
      }
#pragma unroll
      for (int it = 0; it < 16; ++it) {
        const int r = warp * 32 + 2 * it + (lane >> 4);
        uint32_t h0, l0, h1, l1;
        cas_split2(f2_mul(f2_pack(x[it][0], x[it][1]), ss), h0, l0);
        cas_split2(f2_mul(f2_pack(x[it][2], x[it][3]), ss), h1, l1);
        const uint32_t off = (uint32_t)(f >> 1) * kBankLBO + (uint32_t)r * 16 + (uint32_t)(f & 1) * 8;
        *reinterpret_cast<uint2*>(stage + off) = make_uint2(h0, h1);
        *reinterpret_cast<uint2*>(stage + A_HALF + off) = make_uint2(l0, l1);
      }
      fence_proxy_async();
      if (tid != 0) mbar_arrive(&full_bar[s]);
    }
    // ---------------------------------------------------------------- epilogue: TMEM lane `tid` = frame row `tid`
    mbar_wait(&done_bar, 0);
    fence_after_sync();
    const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
    const int bpo = a.bpo;
#pragma unroll 1
    for (int c0 = 0; c0 < NPAD; c0 += 16) {
      float u[16], w[16];
      tmem_ld16(lane_base + c0, u);
      tmem_ld16(lane_base + NPAD + c0, w);
      if (!in_range) continue;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int k = c0 / 2 + j;
        if (k >= bpo) break;
        const int bin = a.n_bins - bpo * (octave + 1) + k;
        const float sc = a.scale[octave * bpo + k];
        float re = (u[2 * j] + w[2 * j]) * sc, im = (u[2 * j + 1] + w[2 * j + 1]) * sc;
        if (!real_frame) re = 0.f, im = 0.f;  // beyond the clip's frames: batch padding is zero (KeyDataset.py:242-254)
        if (a.mode == AKE_CQT_LOGMAG) {
          a.out[((long long)b * a.n_bins + bin) * a.T_max + t] = log1pf(sqrtf(re * re + im * im));
        } else {
          reinterpret_cast<float2*>(a.out)[((long long)b * a.n_bins + bin) * a.T_max + t] = make_float2(re, im);
        }
      }
    }
  } else {
    // ---------------------------------------------------------------- MMA issuer (converged warp, one elected lane issues)
    constexpr uint64_t A_DESC = desc_hi(kBankLBO);
    constexpr uint64_t B_DESC = desc_hi(2 * NPAD * 16);  // chunk stride = 2*NPAD rows x 16 B
    constexpr uint32_t IDESC_WIDE = idesc_f16(2 * NPAD), IDESC_NARROW = idesc_f16(NPAD);
    for (int kb = 0; kb < n_kb; ++kb) {
      const int s = kb % kUStages;
      mbar_wait(&full_bar[s], (kb / kUStages) & 1);
      fence_after_sync();
      const uint32_t base = smem_u32(smem + (size_t)s * STAGE);
      if (elect_one()) {
#pragma unroll
        for (int j = 0; j < kUKB / 16; ++j) {
          const uint64_t bd = make_desc(B_DESC, base + 2 * A_HALF + j * (2 * 2 * NPAD * 16));
          mma_f16(tmem, make_desc(A_DESC, base + j * 2 * kBankLBO), bd, IDESC_WIDE, (kb | j) ? 1u : 0u);
          mma_f16(tmem, make_desc(A_DESC, base + A_HALF + j * 2 * kBankLBO), bd, IDESC_NARROW, 1u);
        }
        commit(&empty_bar[s]);
        if (kb == n_kb - 1) commit(&done_bar);
      }
      __syncwarp();
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem, TMEM_COLS);
}

__global__ void cqt_seqlen_kernel(const long long* __restrict__ lengths, long long n_uniform, int B, int n_oct, int hop0,
                                  int T_max, int* __restrict__ seq_len) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const long long n0 = lengths ? lengths[b] : n_uniform;
  long long T = -1;
  for (int i = 0; i < n_oct; ++i) {
    const long long ti = 1 + ((n0 + (1LL << i) - 1) >> i) / (hop0 >> i);
    T = (T < 0 || ti < T) ? ti : T;
  }
  seq_len[b] = (int)(T < T_max ? T : T_max);
}

struct CqtWs {
  long long* d_len;
  float* level[16];
  long long stride[16];
};

static CqtWs carve(const ake_cqt* p, Arena& ar, int B, long long n_max) {
  CqtWs w{};
  w.d_len = ar.take<long long>(B);
  for (int i = 1; i < p->n_oct; ++i) {
    w.stride[i] = (long long)align_up((size_t)len_at(n_max, i), 32);  // the cascade kernel stores whole rows of 32 samples
    w.level[i] = ar.take<float>((size_t)B * w.stride[i]);
  }
  return w;
}

static void run_cqt(ake_cqt* p, const float* audio, long long stride, const int64_t* lengths_host, int B, long long n_max,
                    int mode, float* out, int T_max, int* seq_len_out, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (p->n_oct > 15) fail(AKE_ERR_UNSUPPORTED, "too many octaves");
  ensure_device(p);
  Arena ar(ws, ws_bytes);
  CqtWs w = carve(p, ar, B, n_max);
  const long long* d_len = nullptr;
  if (lengths_host) {
    for (int b = 0; b < B; ++b)
      if (lengths_host[b] < 0 || lengths_host[b] > n_max) fail(AKE_ERR_INVALID, "lengths_host[%d]=%lld outside [0, n_max]", b, (long long)lengths_host[b]);
    AKE_CUDA(cudaMemcpyAsync(w.d_len, lengths_host, sizeof(long long) * B, cudaMemcpyHostToDevice, st));
    d_len = w.d_len;
  }
  w.level[0] = const_cast<float*>(audio);
  w.stride[0] = stride;
  if (p->n_oct > 1) {
    // resampling cascade: two octave steps per pass (level p -> p+1, p+2), persistent CTAs, 3 per SM
    ProfScope prof("cqt.decimate", st);
    static bool configured = false;
    static int n_sm = 0;
    if (!configured) {
      AKE_CUDA(cudaFuncSetAttribute(cascade_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCasSmemTotal));
      int dev = 0;
      AKE_CUDA(cudaGetDevice(&dev));
      AKE_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
      configured = true;
    }
    for (int lv = 0; lv < p->n_oct - 1; lv += 2) {
      CascadeArgs ca{};
      ca.in = w.level[lv], ca.in_stride = w.stride[lv];
      ca.out1 = w.level[lv + 1], ca.stride1 = w.stride[lv + 1];
      ca.n_levels = std::min(2, p->n_oct - 1 - lv);
      if (ca.n_levels == 2) ca.out2 = w.level[lv + 2], ca.stride2 = w.stride[lv + 2];
      ca.lengths = d_len, ca.n_uniform = n_max, ca.level_in = lv;
      ca.tiles_per_clip = (int)cdiv64(len_at(n_max, lv + 1), kCasOwn1);
      ca.n_tiles = ca.tiles_per_clip * B;
      ca.img = p->d_dec_img;
      const int grid = std::min(ca.n_tiles, 2 * n_sm);
      cascade_umma_kernel<<<grid, kCasThreads, kCasSmemTotal, st>>>(ca);
      AKE_LAUNCHED();
    }
  }
  const long long rows = (long long)B * T_max;
  if (p->npad) {
    // tensor cores: every octave in one launch (grid.y = octave)
    ProfScope prof("cqt.bank", st);
    BankArgs ba{};
    for (int i = 0; i < p->n_oct; ++i) ba.level[i] = w.level[i], ba.stride[i] = w.stride[i];
    ba.lengths = d_len, ba.n_uniform = n_max, ba.n_oct = p->n_oct, ba.hop0 = p->hop, ba.n_fft = p->n_fft, ba.B = B, ba.T_max = T_max;
    ba.n_bins = p->n_bins, ba.bpo = p->bpo, ba.mode = mode, ba.bank_img = p->d_bank_img, ba.scale = p->d_scale_umma, ba.out = out;
    dim3 grid((unsigned)cdiv64(rows, 128), p->n_oct);
    if (p->npad == 80) {
      constexpr size_t smem = kUStages * (2 * kBankAHalf + (kUKB / 8) * 2 * 80 * 16);
      static bool configured = false;
      if (!configured) {
        AKE_CUDA(cudaFuncSetAttribute(cqt_bank_umma_kernel<80>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
      }
      cqt_bank_umma_kernel<80><<<grid, 160, smem, st>>>(ba);
    } else {
      constexpr size_t smem = kUStages * (2 * kBankAHalf + (kUKB / 8) * 2 * 32 * 16);
      static bool configured = false;
      if (!configured) {
        AKE_CUDA(cudaFuncSetAttribute(cqt_bank_umma_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
      }
      cqt_bank_umma_kernel<32><<<grid, 160, smem, st>>>(ba);
    }
    AKE_LAUNCHED();
  } else {
    fail(AKE_ERR_UNSUPPORTED, "bins_per_octave %d / n_fft %d: the filter-bank kernel is built for 36 and 12 bins per octave, n_fft %% 64 == 0",
         p->bpo, p->n_fft);
  }
  if (seq_len_out) {
    cqt_seqlen_kernel<<<cdiv(B, 128), 128, 0, st>>>(d_len, n_max, B, p->n_oct, p->hop, T_max, seq_len_out);
    AKE_LAUNCHED();
  }
}

}  // namespace ake

extern "C" {

int ake_cqt_create(double sr, int hop_length, int n_bins, int bins_per_octave, double fmin, double filter_scale,
                   double sparsity, ake_cqt** out) {
  return guarded([&] {
    if (!out) fail(AKE_ERR_INVALID, "null argument");
    ake_cqt* p = new ake_cqt();
    p->sr = sr, p->hop = hop_length, p->n_bins = n_bins, p->bpo = bins_per_octave, p->fmin = fmin;
    p->filter_scale = filter_scale, p->sparsity = sparsity;
    try {
      build_cqt(p);
    } catch (...) {
      delete p;
      throw;
    }
    *out = p;
  });
}

void ake_cqt_destroy(ake_cqt* p) {
  if (!p) return;
  if (p->d_scale) cudaFree(p->d_scale);
  if (p->d_bank_img) cudaFree(p->d_bank_img);
  if (p->d_dec_img) cudaFree(p->d_dec_img);
  if (p->d_scale_umma) cudaFree(p->d_scale_umma);
  delete p;
}

int ake_cqt_n_fft(const ake_cqt* p) { return p ? p->n_fft : AKE_ERR_INVALID; }
int ake_cqt_n_bins(const ake_cqt* p) { return p ? p->n_bins : AKE_ERR_INVALID; }
int ake_cqt_frames(const ake_cqt* p, int64_t n) { return (p && n >= 0) ? frames_for(p, n) : AKE_ERR_INVALID; }

int ake_cqt_get_bank(const ake_cqt* p, float* bank_host, int64_t cap) {
  return guarded([&] {
    if (!p || !bank_host) fail(AKE_ERR_INVALID, "null argument");
    if (cap < (int64_t)p->bank.size()) fail(AKE_ERR_INVALID, "bank needs %zu floats", p->bank.size());
    std::copy(p->bank.begin(), p->bank.end(), bank_host);
  });
}

int ake_cqt_get_decimator(const ake_cqt* p, float* taps_host, int cap) {
  if (!p || !taps_host || cap < kHalfTaps) return AKE_ERR_INVALID;
  for (int m = 0; m < kHalfTaps; ++m) taps_host[m] = (float)p->dec_half[m];
  return kHalfTaps;
}

size_t ake_cqt_workspace_bytes(const ake_cqt* p, int B, int64_t n_max) {
  if (!p || B <= 0 || n_max <= 0) return 0;
  Arena ar(nullptr, 0);
  carve(p, ar, B, n_max);
  return ar.off + 256;
}

int ake_cqt_run_f32(ake_cqt* p, const float* audio_dev, int64_t stride, const int64_t* lengths_host, int B, int64_t n_max,
                    int mode, float* out_dev, int T_max, int32_t* seq_len_out_dev, void* ws_dev, size_t ws_bytes,
                    void* stream) {
  return guarded([&] {
    if (!p || !audio_dev || !out_dev || !ws_dev) fail(AKE_ERR_INVALID, "null argument");
    if (B <= 0 || n_max <= 0 || stride < n_max || T_max <= 0) fail(AKE_ERR_INVALID, "bad sizes");
    if (mode != AKE_CQT_LOGMAG && mode != AKE_CQT_COMPLEX) fail(AKE_ERR_INVALID, "bad mode");
    ProfScope prof("cqt.total", static_cast<cudaStream_t>(stream));
    run_cqt(p, audio_dev, stride, lengths_host, B, n_max, mode, out_dev, T_max, seq_len_out_dev, ws_dev, ws_bytes,
            static_cast<cudaStream_t>(stream));
  });
}

}  // extern "C"

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
#include <cstdlib>
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
  int hop, n_bins, bpo, n_oct, n_fft;  // n_fft: of the top octave
  int recursion = AKE_CQT_RECURSION_092;
  // Per octave (0 = top): which decimated level it reads, its hop and frame length there, and which filter bank it uses.
  // librosa 0.9.2 recursion: octave i reads level i with ONE shared bank (f_k / sr_i does not depend on i).  "halve while the hop
  // is even" recursion: once the hop is odd the level stays and the filters double in length octave by octave.
  int oct_level[16] = {}, oct_hop[16] = {}, oct_nfft[16] = {}, oct_bank[16] = {};
  int n_levels = 1;                // decimated levels the cascade produces, plus the input (level 0)
  std::vector<double> dec_half;    // kaiser_fast half filter (without the sqrt(2) gain)
  std::vector<std::vector<float>> banks;  // distinct banks: (2*bpo, n_fft_b), row 2k = Re K_k, 2k+1 = Im K_k
  std::vector<int> bank_nfft;
  std::vector<float> out_scale;    // (n_oct, bpo): sqrt(sr / sr_i) / sqrt(length of the full-rate bin)
  float* d_scale = nullptr;
  // tensor-core path: fp16 (hi | lo) image of each bank in the shared-memory operand layout, one block per 64 samples of K
  std::vector<__half*> d_bank_img;
  __half* d_dec_img = nullptr;   // Toeplitz image of the 63-tap decimator (cascade_umma_kernel)
  float* d_scale_umma = nullptr;
  int npad = 0;  // filters per MMA (2*bpo rounded up to 16), 0: tensor-core path not available for this shape
  int device = -1;  // device the operand images live on (the one current at the first run)
  // Amplitude contract (ake_cqt_set_peak): peak > 0 -- the caller guarantees |sample| <= peak and the kernels pre-scale by the
  // exact power of two 2^-ceil(log2 peak); peak == 0 -- unknown: one extra pass measures max |sample| per clip and scales each
  // clip by its own power of two.  Either way the fp16 hi/lo operands see |x| <= 1 and the results carry no scaling error.
  float peak = 1.f;
  // page-locked staging ring for the per-clip lengths of ake_cqt_run_f32: a pageable source would make the H2D copy
  // synchronise the stream (no host run-ahead); a slot is reused only after its copy has executed
  static constexpr int kLenSlots = 4;
  long long* h_len[kLenSlots] = {};
  size_t h_len_cap[kLenSlots] = {};
  cudaEvent_t h_len_done[kLenSlots] = {};
  int h_len_next = 0;
};

namespace ake {

static int two_factors(int x) {
  int n = 0;
  while (x > 0 && x % 2 == 0) x /= 2, ++n;
  return n;
}

// In-place radix-2 FFT (forward, e^{-2 pi i fn/N}); n is a power of two.
static void fft_pow2(std::vector<std::complex<double>>& x) {
  const size_t n = x.size();
  for (size_t i = 1, j = 0; i < n; ++i) {
    size_t bit = n >> 1;
    for (; j & bit; bit >>= 1) j ^= bit;
    j ^= bit;
    if (i < j) std::swap(x[i], x[j]);
  }
  for (size_t len = 2; len <= n; len <<= 1) {
    const double ang = -2.0 * kPi / (double)len;
    std::vector<std::complex<double>> w(len / 2);
    for (size_t k = 0; k < len / 2; ++k) w[k] = std::polar(1.0, ang * (double)k);
    for (size_t i = 0; i < n; i += len)
      for (size_t k = 0; k < len / 2; ++k) {
        const std::complex<double> u = x[i + k], v = x[i + k + len / 2] * w[k];
        x[i + k] = u + v, x[i + k + len / 2] = u - v;
      }
  }
}

// One octave's filter bank: librosa.filters.constant_q(sr_oct, fmin_oct, bpo filters) -> __cqt_filter_fft (complex64 basis times
// lengths / n_fft, FFT, sparsify_rows at `sparsity`) -> its dense real time-domain image K[k][n] = sum_f B[k][f] e^{-2 pi i fn / N}
// (the response basis . rFFT(frame) is linear in the frame, so it equals frame . K^T).  Returns n_fft.
static int make_bank(double sr_oct, double fmin_oct, int bpo, double Q, double sparsity, std::vector<float>& bank) {
  std::vector<double> lengths(bpo);
  double max_len = 0;
  for (int k = 0; k < bpo; ++k) lengths[k] = Q * sr_oct / (fmin_oct * std::pow(2.0, (double)k / bpo)), max_len = std::max(max_len, lengths[k]);
  const int N = 1 << (int)std::ceil(std::log2(max_len)), NF = N / 2 + 1;
  bank.assign((size_t)2 * bpo * N, 0.f);
  std::vector<std::complex<double>> buf(N);
  std::vector<double> mags(NF), sorted(NF);
  for (int k = 0; k < bpo; ++k) {
    const double ilen = lengths[k], freq = fmin_oct * std::pow(2.0, (double)k / bpo);
    const double start = std::floor(-ilen / 2.0), stop = std::floor(ilen / 2.0);  // np.arange(-ilen//2, ilen//2)
    const int len = (int)std::ceil(stop - start);
    const int lpad = (N - len) / 2;  // util.pad_center
    std::fill(buf.begin(), buf.end(), std::complex<double>(0.0, 0.0));
    double wsum = 0;
    for (int m = 0; m < len; ++m) {
      const double w = 0.5 - 0.5 * std::cos(2.0 * kPi * m / len);  // periodic hann
      buf[lpad + m] = std::polar(1.0, (start + m) * 2.0 * kPi * freq / sr_oct) * w;
      wsum += std::abs(buf[lpad + m]);
    }
    const double gain = (ilen / N) / wsum;  // L1 normalisation, then basis *= lengths / n_fft
    fft_pow2(buf);
    double norm = 0;
    for (int f = 0; f < NF; ++f) buf[f] *= gain, mags[f] = std::abs(buf[f]), norm += mags[f];
    // util.sparsify_rows(quantile=sparsity): drop the smallest bins holding < quantile of the L1 mass
    sorted = mags;
    std::sort(sorted.begin(), sorted.end());
    double cum = 0, thresh = sorted[0];
    for (int f = 0; f < NF; ++f) {
      cum += sorted[f] / norm;
      if (!(cum < sparsity)) {
        thresh = sorted[f];
        break;
      }
    }
    for (int f = 0; f < N; ++f) {
      if (f < NF && mags[f] >= thresh) buf[f] = std::complex<double>((float)buf[f].real(), (float)buf[f].imag());  // complex64 basis
      else buf[f] = 0;
    }
    fft_pow2(buf);  // dense time-domain image of the kept bins
    for (int n = 0; n < N; ++n) {
      bank[(size_t)(2 * k) * N + n] = (float)buf[n].real();
      bank[(size_t)(2 * k + 1) * N + n] = (float)buf[n].imag();
    }
  }
  return N;
}

static void build_cqt(ake_cqt* p) {
  const int bpo = p->bpo, n_bins = p->n_bins;
  if (p->sr <= 0 || p->hop <= 0 || n_bins <= 0 || bpo <= 0) fail(AKE_ERR_INVALID, "sr, hop_length, n_bins, bins_per_octave must be positive");
  if (n_bins % bpo) fail(AKE_ERR_UNSUPPORTED, "n_bins must be a multiple of bins_per_octave");
  if (!(p->sparsity >= 0.0 && p->sparsity < 1.0)) fail(AKE_ERR_INVALID, "sparsity must be in [0, 1)");
  if (p->recursion != AKE_CQT_RECURSION_092 && p->recursion != AKE_CQT_RECURSION_HALVE_WHILE_EVEN) fail(AKE_ERR_INVALID, "unknown recursion mode %d", p->recursion);
  if (p->fmin <= 0) p->fmin = 32.70319566257483;  // note_to_hz('C1')
  p->n_oct = n_bins / bpo;
  if (p->n_oct > 15) fail(AKE_ERR_UNSUPPORTED, "too many octaves");
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
  if (p->recursion == AKE_CQT_RECURSION_092 && num_twos < p->n_oct - 1)
    fail(AKE_ERR_INVALID, "hop_length must be a positive integer multiple of 2^%d for %d-octave CQT (librosa 0.9.2; "
         "AKE_CQT_RECURSION_HALVE_WHILE_EVEN lifts this)", p->n_oct - 1, p->n_oct);

  // ---- the recursion: which level / hop / filters every octave uses
  p->banks.clear(), p->bank_nfft.clear();
  int level = 0, hop = p->hop;
  for (int i = 0; i < p->n_oct; ++i) {
    p->oct_level[i] = level, p->oct_hop[i] = hop;
    if (level == i && i > 0) {
      p->oct_bank[i] = 0;  // f_k / sr_i as in the top octave: the same bank
    } else {
      std::vector<float> bank;
      const int nfft = make_bank(p->sr / std::pow(2.0, level), fmin_t / std::pow(2.0, i), bpo, Q, p->sparsity, bank);
      p->oct_bank[i] = (int)p->banks.size();
      p->banks.push_back(std::move(bank)), p->bank_nfft.push_back(nfft);
    }
    p->oct_nfft[i] = p->bank_nfft[p->oct_bank[i]];
    // 0.9.2 halves before every further octave (the hop was checked above); the other rule halves while the hop stays even
    if (i + 1 < p->n_oct && (p->recursion == AKE_CQT_RECURSION_092 || hop % 2 == 0)) ++level, hop /= 2;
  }
  p->n_levels = p->oct_level[p->n_oct - 1] + 1;
  p->n_fft = p->oct_nfft[0];
  // fft_basis *= sqrt(sr / sr_i); V /= sqrt(constant_q_lengths at the full rate)
  p->out_scale.resize((size_t)p->n_oct * bpo);
  for (int i = 0; i < p->n_oct; ++i)
    for (int k = 0; k < bpo; ++k) {
      const int bin = n_bins - bpo * (i + 1) + k;
      const double full_len = Q * p->sr / freqs[bin];
      p->out_scale[(size_t)i * bpo + k] = (float)(std::sqrt(std::pow(2.0, p->oct_level[i])) / std::sqrt(full_len));
    }
  p->dec_half.resize(kHalfTaps);
  kaiser_fast_half(p->dec_half.data());
}

static void free_device_state(ake_cqt* p) {
  void** ptrs[] = {(void**)&p->d_scale, (void**)&p->d_dec_img, (void**)&p->d_scale_umma};
  for (void** q : ptrs) {
    if (*q) cudaFree(*q);
    *q = nullptr;
  }
  for (__half* q : p->d_bank_img)
    if (q) cudaFree(q);
  p->d_bank_img.clear();
}

// Device copies are made lazily so that plan creation (and the bank / tap getters) need no GPU; they live on the device
// that is current when the plan first runs, and are rebuilt if the plan is later driven on another device.
static void ensure_device(ake_cqt* p) {
  const int dev = current_device();
  if (p->d_scale && p->device == dev) return;
  free_device_state(p);
  p->device = dev;
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
  bool ok = npad == 80 || npad == 32;
  for (int nfft : p->bank_nfft) ok = ok && nfft % kUKB == 0;
  if (ok) {
    p->npad = npad;
    for (size_t bi = 0; bi < p->banks.size(); ++bi) {
      const int nfft = p->bank_nfft[bi], n_kb = nfft / kUKB;
      const std::vector<float>& bank = p->banks[bi];
      std::vector<__half> img((size_t)n_kb * 8 * 2 * npad * 8, __float2half(0.f));
      for (int f = 0; f < nf; ++f)
        for (int k = 0; k < nfft; ++k) {
          const float v = bank[(size_t)f * nfft + k] * kBankScale;
          const __half hi = __float2half_rn(v);
          const __half lo = __float2half_rn(v - __half2float(hi));
          const size_t blk = (size_t)(k / kUKB) * 8 * 2 * npad * 8, c = (k % kUKB) / 8, e = k % 8;
          img[blk + (c * 2 * npad + f) * 8 + e] = hi;
          img[blk + (c * 2 * npad + npad + f) * 8 + e] = lo;
        }
      __half* d = nullptr;
      AKE_CUDA(cudaMalloc(&d, sizeof(__half) * img.size()));
      AKE_CUDA(cudaMemcpy(d, img.data(), sizeof(__half) * img.size(), cudaMemcpyHostToDevice));
      p->d_bank_img.push_back(d);
    }
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
    const long long t = 1 + len_at(n, p->oct_level[i]) / p->oct_hop[i];
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
// 32 + 2*32 input samples per side (2.4 %) instead of exchanging state with its neighbours.
constexpr int kCasRows2 = 63;                       // level p+2 rows (of 32 outputs) per tile
constexpr int kCasOwn2 = kCasRows2 * 32;            // 2016 level p+2 outputs owned by a tile
constexpr int kCasOwn1 = 2 * kCasOwn2;              // 4032 level p+1 outputs owned by a tile (rows 1..126 of 128)
constexpr int kCasP0Rows = 129, kCasP1Rows = 65;    // plane rows (16 B each) of the level p / level p+1 operands
constexpr uint32_t kCasLBO0 = kCasP0Rows * 16, kCasLBO1 = kCasP1Rows * 16;  // odd multiples of 16 B: conflict-free scatter
constexpr uint32_t kCasP0Bytes = 8 * kCasLBO0, kCasP1Bytes = 8 * kCasLBO1;
constexpr uint32_t kCasImgBytes = 16 * 64 * 16;     // Toeplitz image: [chunk 16][n 64 = Hh 32 | Hl 32][8 halves]

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
  int sparse_hop, sparse_nfft;  // > 0: level p+1 is only read by the filter bank (hop, n_fft at that level): store just those rows
  int stream_in;                // the input is read exactly once (caller's audio): load it with the L2 evict_first policy
  // Amplitude pre-scale of the FIRST pass (level 0 = the caller's audio): clip b is multiplied by the exact power of two xs[b]
  // (xs == NULL: xs_uniform for every clip) on its way into the fp16 hi/lo operands; the decimated levels are stored scaled and
  // the filter bank's epilogue divides it out again.  Later passes run with xs == NULL, xs_uniform == 1.
  const float* xs;
  float xs_uniform;
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

// Warp roles of the persistent CTA (one per SM); every role walks the same tile sequence and hands over through mbarriers:
//   warps 0-7   epilogue 1, two groups of four warps taking alternate tiles: level p+1 accumulators -> fp16 hi/lo planes
//               (operand of the level p+2 MMAs) and, through a transposition buffer, row-contiguous fp32 stores (only the
//               rows the filter bank reads when `sparse_hop` > 0)
//   warps 8-11  epilogue 2: level p+2 accumulators (63 rows) -> fp32 stores.  The level p+2 MMAs run with M = 64 (half the
//               operand bytes out of shared memory), whose accumulator rows 16 q .. 16 q + 15 sit in lanes 0..15 of TMEM
//               quadrant q (tools/m64_probe.cu): one warp per quadrant, its upper half-warp idles through the drain
//   warps 12-23 converter: landed fp32 span -> fp16 hi/lo transposed chunk planes (level-p operand) IN PLACE: a span of 8256 fp32
//               samples and its hi + lo planes are both 33,024 bytes, so a ring buffer is landing area first and MMA operand
//               afterwards (every converter thread holds its samples in registers across a barrier before the first write)
//   warp 24     loader: one bulk async copy per tile span into the ring (kCasRing buffers: all but the one being multiplied can
//               be in flight)
//   warps 25-26 MMA issuers (converged, elect.sync), one per level: MMA1(i+2) is issued as soon as epilogue 1 has drained
//               accumulator i, MMA2(i) as soon as its operand planes exist; two warps because one cannot issue MMAs of
//               N <= 64 as fast as the tensor pipe executes them
constexpr int kCasThreads = 864;
constexpr int kCasEpi2Warp = 8, kCasConvWarp = 12, kCasLoadWarp = 24, kCasIssueWarp = 25;
constexpr int kCasConvThreads = 384;
constexpr int kCasChunks = kCasP0Rows * 8;                 // 1032 chunks of 8 samples per tile span
constexpr int kCasSpan = kCasChunks * 8;                   // 8256 input samples per tile
constexpr uint32_t kCasStageBytes = kCasSpan * 4;          // fp32 landing buffer of one tile span (bulk async copy)
constexpr int kCasTPitch = 36;                             // floats per row of the store-transposition buffers (144 B: conflict-free)
constexpr uint32_t kCasT1Bytes = 128 * kCasTPitch * 4;     // level p+1 outputs on their way to coalesced global stores
constexpr uint32_t kCasT2Bytes = 64 * kCasTPitch * 4;      // level p+2
constexpr int kCasRing = 4;                                // ring buffers: landing area, then level-p operand planes (same bytes)
static_assert(kCasStageBytes == 2 * kCasP0Bytes, "a landed span and its hi + lo planes must be the same size (in-place conversion)");
constexpr uint32_t kCasSmemTotal = 2 * 2 * kCasP1Bytes + kCasRing * kCasStageBytes + kCasImgBytes + 2 * kCasT1Bytes + kCasT2Bytes;

__global__ void __launch_bounds__(kCasThreads, 1) cascade_umma_kernel(const CascadeArgs a) {
  using namespace umma;
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t img_bar, stage_full[kCasRing], p0_full[kCasRing], p0_empty[kCasRing], p1_full[2], p1_empty[2], acc1_full[2],
      acc1_empty[2], acc2_full[2], acc2_empty[2];
  __shared__ uint32_t tmem_slot;
  uint8_t* p1 = smem;                                  // [buf 2][hi | lo][kCasP1Bytes]
  uint8_t* ring = smem + 4 * kCasP1Bytes;              // [kCasRing][kCasStageBytes]: fp32 span, then [hi | lo][kCasP0Bytes] in place
  uint8_t* img = ring + kCasRing * kCasStageBytes;
  float* tbuf1 = reinterpret_cast<float*>(img + kCasImgBytes);  // [group 2]
  float* tbuf2 = reinterpret_cast<float*>(img + kCasImgBytes + 2 * kCasT1Bytes);
  const int tid = threadIdx.x, warp = uniform_warp_idx(), lane = tid & 31;
  const bool two = a.n_levels == 2;

  if (warp == kCasIssueWarp) tmem_alloc(&tmem_slot, 256);
  if (tid == 0) {
    mbar_init(&img_bar, 1);
    for (int i = 0; i < kCasRing; ++i) mbar_init(&stage_full[i], 1), mbar_init(&p0_full[i], kCasConvThreads), mbar_init(&p0_empty[i], 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&p1_full[i], 128), mbar_init(&p1_empty[i], 1);
      mbar_init(&acc1_full[i], 1), mbar_init(&acc1_empty[i], 128);
      mbar_init(&acc2_full[i], 1), mbar_init(&acc2_empty[i], 128);
    }
    mbar_init_fence();
  }
  for (uint32_t i = tid; i < (4 * kCasP1Bytes + kCasRing * kCasStageBytes) / 16; i += kCasThreads) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_slot;
  // accumulator -> sample: level p+1 carries kXScale * kDecScale, level p+2 one more kDecScale (the level p+1 operand is the
  // raw accumulator: no rescaling between the two steps)
  constexpr float kInv1 = 1.f / (kXScale * kDecScale), kInv2 = kInv1 / kDecScale;

  // tile -> clip by multiply-high: every role decodes every tile, and an integer division is ~50 dependent instructions
  // (x / d == umulhi(x, 2^32 / d + 1) while x * d < 2^32)
  const uint32_t tpc_magic = 0xFFFFFFFFu / (uint32_t)a.tiles_per_clip + 1;
  auto clip_of = [&](int tile) { return a.tiles_per_clip == 1 ? tile : (int)__umulhi((uint32_t)tile, tpc_magic); };
  // Every role decodes every tile between two hand-overs, so this scalar code sits on the pipeline's critical path: the
  // uniform-length case (no lengths array) takes no global load, no loop and no 64-bit shifts per tile.
  const bool ragged = a.lengths != nullptr;
  const int len_uniform = (int)((a.n_uniform + (1LL << a.level_in) - 1) >> a.level_in);
  auto clip_len = [&](int b) {  // samples of level p in clip b (< 2^31: checked on the host; all per-tile arithmetic is 32-bit)
    return ragged ? (int)((a.lengths[b] + (1LL << a.level_in) - 1) >> a.level_in) : len_uniform;
  };
  // first tile at or after `tile` (in this CTA's stride) that has data; tiles beyond a short clip's end write nothing
  auto next_tile = [&](int tile) {
    if (!ragged) return tile;  // tiles_per_clip covers exactly the tiles that have data
    for (; tile < a.n_tiles; tile += gridDim.x) {
      const int b = clip_of(tile), t = tile - b * a.tiles_per_clip;
      if (kCasOwn1 * t < ((clip_len(b) + 1) >> 1)) break;
    }
    return tile;
  };
  struct Span {
    const float* src;  // first sample of the span (may lie before the clip: zero extension)
    int vlo, vhi;      // samples [vlo, vhi) of the span exist
    bool bulk;         // 16-byte aligned: landed by one bulk async copy
  };
  auto span_of = [&](int tile) {
    const int b = clip_of(tile), t = tile - b * a.tiles_per_clip;
    const int base0 = 2 * (kCasOwn1 * t - 32) - 32;
    Span s;
    s.src = a.in + ((long long)b * a.in_stride + base0);
    s.vlo = max(0, -base0), s.vhi = min(kCasSpan, clip_len(b) - base0);
    s.bulk = (reinterpret_cast<uintptr_t>(s.src) & 15) == 0;
    return s;
  };

  if (warp == kCasLoadWarp) {
    // ------------------------------------------------------------------ loader
    if (lane == 0) {
      mbar_arrive_expect_tx(&img_bar, kCasImgBytes);
      bulk_g2s(img, a.img, kCasImgBytes, &img_bar);
    }
    const uint64_t pol = l2_policy_evict_first();
    int i = 0;
    for (int tile = next_tile(blockIdx.x); tile < a.n_tiles; tile = next_tile(tile + gridDim.x), ++i) {
      const int s = i % kCasRing;
      float* st = reinterpret_cast<float*>(ring + (size_t)s * kCasStageBytes);
      const Span sp = span_of(tile);
      mbar_wait_relaxed(&p0_empty[s], ((i / kCasRing) & 1) ^ 1);  // free once the level p+1 MMAs of the tile it held have read it
      if (sp.bulk) {
        // one bulk copy of the existing samples (whole float4s); lane 0 patches the <= 3 tail samples.  Samples outside
        // [vlo, vhi) are masked at conversion time, not written here.
        if (lane == 0) {
          const int n4 = (sp.vhi - sp.vlo) & ~3;
          for (int q = sp.vlo + n4; q < sp.vhi; ++q) st[q] = __ldg(sp.src + q);
          if (n4 > 0) {
            mbar_arrive_expect_tx(&stage_full[s], (uint32_t)n4 * 4);
            if (a.stream_in) bulk_g2s_hint(st + sp.vlo, sp.src + sp.vlo, (uint32_t)n4 * 4, &stage_full[s], pol);
            else bulk_g2s(st + sp.vlo, sp.src + sp.vlo, (uint32_t)n4 * 4, &stage_full[s]);
          } else {
            mbar_arrive(&stage_full[s]);
          }
        }
      } else {
        for (int q = sp.vlo + lane; q < sp.vhi; q += 32) st[q] = __ldg(sp.src + q);
        __syncwarp();
        if (lane == 0) mbar_arrive(&stage_full[s]);
      }
      __syncwarp();
    }
  } else if (warp >= kCasIssueWarp) {
    // ------------------------------------------------------------------ MMA issuers (first warp: level p+1, second: level p+2)
    int n_my = 0;
    for (int tile = next_tile(blockIdx.x); tile < a.n_tiles; tile = next_tile(tile + gridDim.x)) ++n_my;
    mbar_wait(&img_bar, 0);
    const uint32_t w0 = smem_u32(img);
    auto issue_level = [&](uint32_t d, uint32_t hi0, uint32_t lo0, uint32_t lbo, uint64_t* bar_acc, uint64_t* bar_planes, uint32_t idesc64,
                           uint32_t idesc32) {
      const uint64_t a_desc = desc_hi(lbo);
      constexpr uint64_t B_DESC = desc_hi(64 * 16);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint32_t off = (uint32_t)((2 * j) & 7) * lbo + (j >= 4 ? 16u : 0u);  // chunks 8..15 = planes 0..7, one row on
        const uint64_t bd = make_desc(B_DESC, w0 + (uint32_t)j * 2048);
        mma_f16(d, make_desc(a_desc, hi0 + off), bd, idesc64, j ? 1u : 0u);
        mma_f16(d, make_desc(a_desc, lo0 + off), bd, idesc32, 1u);
      }
      commit(bar_acc);
      commit(bar_planes);
    };
    auto issue1 = [&](int i) {
      const int bf = i & 1, s = i % kCasRing;
      const uint32_t ph = (i >> 1) & 1;
      mbar_wait(&p0_full[s], (i / kCasRing) & 1);
      mbar_wait(&acc1_empty[bf], ph ^ 1);
      fence_after_sync();
      const uint32_t hi0 = smem_u32(ring + (size_t)s * kCasStageBytes);
      if (elect_one()) issue_level(tmem + bf * 64, hi0, hi0 + kCasP0Bytes, kCasLBO0, &acc1_full[bf], &p0_empty[s], idesc_f16(64), idesc_f16(32));
      __syncwarp();
    };
    auto issue2 = [&](int i) {
      const int bf = i & 1;
      const uint32_t ph = (i >> 1) & 1;
      mbar_wait(&p1_full[bf], ph);
      mbar_wait(&acc2_empty[bf], ph ^ 1);
      fence_after_sync();
      const uint32_t hi0 = smem_u32(p1 + (size_t)bf * 2 * kCasP1Bytes);
      // 63 rows of level p+2 per tile: M = 64 (the operand read is what an MMA of this N costs; M = 128 would read 128 rows)
      if (elect_one())
        issue_level(tmem + 128 + bf * 64, hi0, hi0 + kCasP1Bytes, kCasLBO1, &acc2_full[bf], &p1_empty[bf], idesc_f16(64, 64), idesc_f16(32, 64));
      __syncwarp();
    };
    // One warp issues at most one MMA per ~65 cycles (tools/umma_rate.cu: 7 x N=64 MMAs per block take 65 cycles each from
    // one warp, 48 -- the shared-memory operand rate -- from two), so the two levels are issued by two warps: independent
    // streams that only meet through the plane / accumulator barriers.
    if (warp == kCasIssueWarp) {
      for (int i = 0; i < n_my; ++i) issue1(i);
    } else if (two) {
      for (int i = 0; i < n_my; ++i) issue2(i);
    }
  } else if (warp >= kCasConvWarp) {
    // ------------------------------------------------------------------ converter: stage -> level-p operand planes
    const int ct = tid - kCasConvWarp * 32;
    int i = 0;
    for (int tile = next_tile(blockIdx.x); tile < a.n_tiles; tile = next_tile(tile + gridDim.x), ++i) {
      const int s = i % kCasRing;
      const Span sp = span_of(tile);
      const float xsc = kXScale * (a.xs ? __ldg(a.xs + clip_of(tile)) : a.xs_uniform);  // issued before the wait below
      const uint64_t ss = f2_pack(xsc, xsc);
      const float* st = reinterpret_cast<const float*>(ring + (size_t)s * kCasStageBytes);
      uint8_t* p0h = ring + (size_t)s * kCasStageBytes;  // the planes overwrite the span they are made from
      uint8_t* p0l = p0h + kCasP0Bytes;
      mbar_wait_relaxed(&stage_full[s], (i / kCasRing) & 1);
      const bool interior = sp.vlo == 0 && sp.vhi == kCasSpan;
      // one float4 per thread and round: consecutive lanes read consecutive 16 B of the stage and write 8-byte halves of the
      // operand chunks (both conflict-free).  float4 f = ct + kCasConvThreads r is half (f & 1) of chunk q = f / 2, which lives
      // at plane (q & 7), row (q >> 3): the plane and the half are fixed per thread, the row advances by a constant per round.
      const uint32_t off0 = (uint32_t)((ct >> 1) & 7) * kCasLBO0 + (uint32_t)(ct >> 4) * 16 + (uint32_t)(ct & 1) * 8;
      auto convert4 = [&](int r, float4 v, bool masked) {
        const int f = ct + kCasConvThreads * r;
        float x[4] = {v.x, v.y, v.z, v.w};
        if (masked) {
#pragma unroll
          for (int e = 0; e < 4; ++e) x[e] = (4 * f + e >= sp.vlo && 4 * f + e < sp.vhi) ? x[e] : 0.f;
        }
        uint32_t h[2], l[2];
        cas_split2(f2_mul(f2_pack(x[0], x[1]), ss), h[0], l[0]);
        cas_split2(f2_mul(f2_pack(x[2], x[3]), ss), h[1], l[1]);
        const uint32_t off = off0 + (uint32_t)(kCasConvThreads / 16 * 16) * (uint32_t)r;  // kCasConvThreads / 16 operand rows per round
        *reinterpret_cast<uint2*>(p0h + off) = make_uint2(h[0], h[1]);
        *reinterpret_cast<uint2*>(p0l + off) = make_uint2(l[0], l[1]);
      };
      constexpr int kFull = 2 * kCasChunks / kCasConvThreads, kRest = 2 * kCasChunks - kFull * kCasConvThreads;
      static_assert(kCasConvThreads % 16 == 0 && kFull == 5 && kRest > 0, "5 full rounds + a partial one");
      // The stage reads of a tile are issued before the first conversion: the compiler cannot hoist a shared-memory load
      // above the plane stores of the previous round (it cannot prove they do not alias), and one LDS latency per float4
      // was 40 % of the converter's time.
      const float4* st4 = reinterpret_cast<const float4*>(st) + ct;
      float4 v[kFull + 1];
#pragma unroll
      for (int u = 0; u < kFull; ++u) v[u] = st4[kCasConvThreads * u];
      v[kFull] = ct < kRest ? st4[kCasConvThreads * kFull] : make_float4(0.f, 0.f, 0.f, 0.f);
      // in-place conversion: every converter thread must hold its samples before any plane byte is written
      asm volatile("bar.sync 4, %0;" ::"n"(kCasConvThreads) : "memory");
      if (interior) {  // all but the first and last tiles of a clip: no per-sample predicates
#pragma unroll
        for (int u = 0; u < kFull; ++u) convert4(u, v[u], false);
      } else {
#pragma unroll
        for (int u = 0; u < kFull; ++u) convert4(u, v[u], true);
      }
      if (ct < kRest) convert4(kFull, v[kFull], !interior);
      fence_proxy_async();
      mbar_arrive(&p0_full[s]);
    }
  } else if (warp < 8) {
    // ------------------------------------------------------------------ epilogue 1: accumulator lane = row of 32 level p+1 outputs
    const int grp = warp >> 2, row = tid & 127;  // group g drains the tiles i = g, g + 2, ... (accumulator / plane buffer g)
    const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16) + grp * 64;
    float* tb = tbuf1 + (size_t)grp * (kCasT1Bytes / 4);
    uint8_t* p1h = p1 + (size_t)grp * 2 * kCasP1Bytes;
    uint8_t* p1l = p1h + kCasP1Bytes;
    const uint64_t inv1 = f2_pack(kInv1, kInv1);
    int i = 0;
    for (int tile = next_tile(blockIdx.x); tile < a.n_tiles; tile = next_tile(tile + gridDim.x), ++i) {
      if ((i & 1) != grp) continue;
      const uint32_t ph = (i >> 1) & 1;
      const int b = clip_of(tile), t = tile - b * a.tiles_per_clip;
      const int n_p = clip_len(b);
      const int n_half1 = n_p >> 1, n1 = (n_p + 1) >> 1;
      const int o_lo1 = kCasOwn1 * t - 32;
      mbar_wait_relaxed(&acc1_full[grp], ph);
      fence_after_sync();
      uint64_t o[16];  // 32 outputs as fp32 pairs
      {
        uint32_t u0[16], u1[16], w0[16], w1[16];
        tmem_ld16_issue(lane_base, u0);
        tmem_ld16_issue(lane_base + 16, u1);
        tmem_ld16_issue(lane_base + 32, w0);
        tmem_ld16_issue(lane_base + 48, w1);
        tmem_ld_wait16(u0), tmem_ld_wait16(u1), tmem_ld_wait16(w0), tmem_ld_wait16(w1);
#pragma unroll
        for (int n = 0; n < 8; ++n) {
          o[n] = f2_add(f2_pack(__uint_as_float(u0[2 * n]), __uint_as_float(u0[2 * n + 1])), f2_pack(__uint_as_float(w0[2 * n]), __uint_as_float(w0[2 * n + 1])));
          o[8 + n] = f2_add(f2_pack(__uint_as_float(u1[2 * n]), __uint_as_float(u1[2 * n + 1])), f2_pack(__uint_as_float(w1[2 * n]), __uint_as_float(w1[2 * n + 1])));
        }
      }
      fence_before_sync();
      mbar_arrive(&acc1_empty[grp]);  // accumulator drained: MMA1 of tile i + 2 may start
      if (o_lo1 < 0 || o_lo1 + 128 * 32 > n_half1) {
        // samples that do not exist (index < 0 or >= floor(n_p / 2)) are zero
        const int i1 = o_lo1 + 32 * row;  // level p+1 index of the row's first output
        const int zlo = min(32, max(0, -i1)), zhi = min(32, max(0, n_half1 - i1));
#pragma unroll
        for (int n = 0; n < 16; ++n) {
          float x0, x1;
          f2_unpack(o[n], x0, x1);
          o[n] = f2_pack((2 * n >= zlo && 2 * n < zhi) ? x0 : 0.f, (2 * n + 1 >= zlo && 2 * n + 1 < zhi) ? x1 : 0.f);
        }
      }
      if (two) {
        // level p+1 as the next operand: row r = half of operand row r / 2, chunks 4 (r & 1) .. + 3
        mbar_wait(&p1_empty[grp], ph ^ 1);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t h[4], l[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) cas_split2(o[4 * c + e], h[e], l[e]);
          const uint32_t off = (uint32_t)(4 * (row & 1) + c) * kCasLBO1 + (uint32_t)(row >> 1) * 16;
          *reinterpret_cast<uint4*>(p1h + off) = make_uint4(h[0], h[1], h[2], h[3]);
          *reinterpret_cast<uint4*>(p1l + off) = make_uint4(l[0], l[1], l[2], l[3]);
        }
        fence_proxy_async();
        mbar_arrive(&p1_full[grp]);
      }
      // fp32 outputs go through shared memory so that the global stores are row-contiguous (a thread owns a row: storing
      // straight from registers would touch 32 different lines per instruction)
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        float4 v;
        f2_unpack(f2_mul(o[2 * q], inv1), v.x, v.y);
        f2_unpack(f2_mul(o[2 * q + 1], inv1), v.z, v.w);
        *reinterpret_cast<float4*>(tb + row * kCasTPitch + 4 * q) = v;
      }
      asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
      {
        // rows 1..126 are owned (0 and 127 are the halo the next level needs): 1008 float4s, 8 per row; thread `row` stores
        // float4 (row & 7) of the rows 1 + (row >> 3) + 16 k
        float* dst = a.out1 + ((long long)b * a.stride1 + (o_lo1 + 32 * (1 + (row >> 3)) + 4 * (row & 7)));
        const float* src = tb + (1 + (row >> 3)) * kCasTPitch + 4 * (row & 7);
        // rows with 32 r < lim exist (the buffers are padded to whole rows of 32)
        const int r_lim = min(127, max(0, (n1 - o_lo1 + 31) >> 5));
        const bool sparse = a.sparse_hop > 0;
        int u = 0;
        if (sparse) {
          u = (int)((unsigned)(o_lo1 + a.sparse_nfft / 2) % (unsigned)a.sparse_hop) + 32 * (1 + (row >> 3));  // o_lo1 + n_fft/2 > 0
          u -= (u >= a.sparse_hop) ? a.sparse_hop : 0;
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int r = 1 + (row >> 3) + 16 * k;
          bool need = r < r_lim;
          // sparse: the only reader of this level is the filter bank, frames [t hop - n_fft/2, t hop + n_fft/2): keep the rows
          // whose offset u into the hop period (advanced by 16 rows = 512 samples per round) falls into a frame
          if (sparse) need = need && (u < a.sparse_nfft || u + 31 >= a.sparse_hop);
          u += 512;
          u -= (u >= a.sparse_hop) ? a.sparse_hop : 0;
          if (need) *reinterpret_cast<float4*>(dst + 512 * k) = *reinterpret_cast<const float4*>(src + 16 * k * kCasTPitch);
        }
      }
      asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");  // the transposition buffer is reused by the group's next tile
    }
  } else if (two) {
    // ------------------------------------------------------------------ epilogue 2 (warps 8-11): rows 0..62 of level p+2
    const int quad = warp - kCasEpi2Warp;       // TMEM quadrant; rows 16 quad .. + 15 of the M = 64 accumulator in its lanes 0..15
    const int row = 16 * quad + (lane & 15);
    const bool has_row = lane < 16;
    const int tid2 = tid - kCasEpi2Warp * 32;   // 0..127
    const uint32_t lane_base = tmem + ((uint32_t)(quad * 32) << 16) + 128;
    const uint64_t inv2 = f2_pack(kInv2, kInv2);
    int i = 0;
    for (int tile = next_tile(blockIdx.x); tile < a.n_tiles; tile = next_tile(tile + gridDim.x), ++i) {
      const int bf = i & 1;
      const uint32_t ph = (i >> 1) & 1;
      const int b = clip_of(tile), t = tile - b * a.tiles_per_clip;
      const int n_p = clip_len(b);
      const int n1 = (n_p + 1) >> 1, n_half2 = n1 >> 1, n2 = (n1 + 1) >> 1;
      const int o_lo2 = kCasOwn2 * t;
      mbar_wait_relaxed(&acc2_full[bf], ph);
      fence_after_sync();
      uint64_t o[16];  // 32 outputs as fp32 pairs
      {
        uint32_t u0[16], u1[16], w0[16], w1[16];
        tmem_ld16_issue(lane_base + bf * 64, u0);
        tmem_ld16_issue(lane_base + bf * 64 + 16, u1);
        tmem_ld16_issue(lane_base + bf * 64 + 32, w0);
        tmem_ld16_issue(lane_base + bf * 64 + 48, w1);
        tmem_ld_wait16(u0), tmem_ld_wait16(u1), tmem_ld_wait16(w0), tmem_ld_wait16(w1);
#pragma unroll
        for (int n = 0; n < 8; ++n) {
          o[n] = f2_add(f2_pack(__uint_as_float(u0[2 * n]), __uint_as_float(u0[2 * n + 1])), f2_pack(__uint_as_float(w0[2 * n]), __uint_as_float(w0[2 * n + 1])));
          o[8 + n] = f2_add(f2_pack(__uint_as_float(u1[2 * n]), __uint_as_float(u1[2 * n + 1])), f2_pack(__uint_as_float(w1[2 * n]), __uint_as_float(w1[2 * n + 1])));
        }
      }
      fence_before_sync();
      mbar_arrive(&acc2_empty[bf]);
      if (o_lo2 + kCasRows2 * 32 > n_half2) {
        // samples at or beyond floor(n1 / 2) do not exist: zero
        const int zhi = min(32, max(0, n_half2 - (o_lo2 + 32 * row)));
#pragma unroll
        for (int n = 0; n < 16; ++n) {
          float x0, x1;
          f2_unpack(o[n], x0, x1);
          o[n] = f2_pack(2 * n < zhi ? x0 : 0.f, 2 * n + 1 < zhi ? x1 : 0.f);
        }
      }
      if (has_row && row < kCasRows2) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          float4 v;
          f2_unpack(f2_mul(o[2 * q], inv2), v.x, v.y);
          f2_unpack(f2_mul(o[2 * q + 1], inv2), v.z, v.w);
          *reinterpret_cast<float4*>(tbuf2 + row * kCasTPitch + 4 * q) = v;
        }
      }
      asm volatile("bar.sync 3, 128;" ::: "memory");
      {
        // 63 rows x 8 float4s; thread tid2 stores float4 (tid2 & 7) of the rows (tid2 >> 3) + 16 k
        float* dst = a.out2 + ((long long)b * a.stride2 + (o_lo2 + 32 * (tid2 >> 3) + 4 * (tid2 & 7)));
        const float* src = tbuf2 + (tid2 >> 3) * kCasTPitch + 4 * (tid2 & 7);
        const int r_lim = min(kCasRows2, max(0, (n2 - o_lo2 + 31) >> 5));  // rows with 32 r < n2 - o_lo2 exist
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if ((tid2 >> 3) + 16 * k < r_lim) *reinterpret_cast<float4*>(dst + 512 * k) = *reinterpret_cast<const float4*>(src + 16 * k * kCasTPitch);
      }
      asm volatile("bar.sync 3, 128;" ::: "memory");  // the transposition buffer is reused by the next tile
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == kCasIssueWarp) tmem_dealloc(tmem, 256);
}

// ---- tensor-core filter bank (tcgen05): one CTA = 128 frames of one octave x all filters x all of K ----------------
//   D[frame, n] = sum_k x[frame, k] * K[n, k],  x ~= xh + xl, K ~= Kh + Kl (fp16 pairs, fp32 accumulation in TMEM)
//   MMA 1: A = xh, B = [Kh | Kl]  (N = 2*NPAD)  -> columns [0,NPAD) += xh*Kh, [NPAD,2*NPAD) += xh*Kl
//   MMA 2: A = xl, B =  Kh        (N =   NPAD)  -> columns [0,NPAD) += xl*Kh
// Producer warps stage frames (fp32 -> fp16 hi/lo, operand layout of umma.cuh), four of every eight also run the |.|, scale,
// log1p epilogue out of TMEM; the last warp issues the MMAs; the bank block of each stage arrives by one bulk async copy.
// The kernel is bound by L2 -> SM traffic (frames 4 KB per row + the 327 KB bank image per CTA, ~7 TB/s in total): a CTA
// multiplies MB blocks of 128 frame rows by every bank block it fetches, so the image is read once per 128 * MB rows.

struct BankArgs {
  // per octave (0 = top): the decimated level it reads (pointer, clip stride, decimation count), its hop and frame length there,
  // and its filter-bank image
  const float* level[16];
  long long stride[16];
  const __half* bank_img[16];
  int shift[16], hop[16], nfft[16];
  const long long* lengths;
  long long n_uniform;
  int n_oct, B, T_max, n_bins, bpo, mode;
  const float* scale;
  float* out;
  const float* xs;   // per-clip amplitude pre-scale (exact powers of two) or NULL = xs_uniform; see CascadeArgs
  float xs_uniform;
};

constexpr uint32_t kBankLBO = 129 * 16;                  // chunk pitch of the frame operand: odd multiple of 16 B (conflict-free 8-byte scatter)
constexpr uint32_t kBankAHalf = (kUKB / 8) * kBankLBO;
constexpr int kBankMB = 2;                                // blocks of 128 frame rows per CTA (one TMEM accumulator each)
constexpr int kBankThreads = 32 * (8 * kBankMB + 1);      // 8 producer warps per row block (4 of them also run its epilogue) + the MMA-issuer warp
__host__ __device__ constexpr uint32_t bank_stage_bytes(int npad, int mb) { return mb * 2 * kBankAHalf + (kUKB / 8) * 2 * npad * 16; }

template <int NPAD, int MB>
__global__ void __launch_bounds__(32 * (8 * MB + 1), 1) cqt_bank_umma_kernel(const BankArgs a) {
  using namespace umma;
  constexpr uint32_t A_HALF = kBankAHalf;                    // one of {hi, lo}: 8 chunks x 128 rows x 16 B (+ pad)
  constexpr uint32_t B_BYTES = (kUKB / 8) * 2 * NPAD * 16;   // 8 chunks x 2*NPAD rows x 16 B
  constexpr uint32_t STAGE = bank_stage_bytes(NPAD, MB);     // [row block 0: hi, lo] ... [row block MB-1: hi, lo] [bank block]
  constexpr uint32_t B_OFF = MB * 2 * A_HALF;
  constexpr uint32_t TMEM_NEED = MB * 2 * NPAD;
  constexpr uint32_t TMEM_COLS = (TMEM_NEED <= 64) ? 64 : ((TMEM_NEED <= 128) ? 128 : ((TMEM_NEED <= 256) ? 256 : 512));
  static_assert(TMEM_NEED <= 512, "accumulators exceed TMEM");
  constexpr int ISSUER = 8 * MB;
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t full_bar[kUStages], empty_bar[kUStages], done_bar;
  __shared__ uint32_t tmem_slot;
  __shared__ long long s_g0[128 * MB];   // index (into the octave's level array) of the first sample of frame row r
  __shared__ int2 s_valid[128 * MB];     // samples [x, y) of that frame exist (the rest is the zero padding of centred frames)
  __shared__ float s_xs[128 * MB];       // operand scale of the row: kXScale, times the clip's amplitude pre-scale where the row is raw audio

  const int tid = threadIdx.x, warp = uniform_warp_idx(), lane = tid & 31;
  const int octave = blockIdx.y;
  const long long m0 = (long long)blockIdx.x * (128 * MB);
  const int n_fft = a.nfft[octave];
  const int n_kb = n_fft / kUKB;
  const int mb = warp >> 3;                                   // row block of this producer warp
  const int row_e = 128 * mb + 32 * (warp & 3) + lane;        // epilogue warps (warp & 7) < 4: TMEM lane 32 (warp & 3) + lane of accumulator mb

  if (warp == ISSUER) tmem_alloc(&tmem_slot, TMEM_COLS);
  if (tid == 0) {
    for (int s = 0; s < kUStages; ++s) mbar_init(&full_bar[s], 256 * MB), mbar_init(&empty_bar[s], 1);
    mbar_init(&done_bar, 1);
    mbar_init_fence();
  }
  // ---- frame row `row_e`: which clip / frame, where its samples live
  bool in_range = false, real_frame = false;
  int b = 0, t = 0;
  float inv_xs = 1.f;
  if (warp < ISSUER && (warp & 7) < 4) {
    const long long m = m0 + row_e;
    in_range = m < (long long)a.B * a.T_max;
    b = in_range ? (int)(m / a.T_max) : 0, t = in_range ? (int)(m % a.T_max) : 0;
    const long long n0 = a.lengths ? a.lengths[b] : a.n_uniform;
    long long T = -1;
    for (int i = 0; i < a.n_oct; ++i) {
      const long long ti = 1 + ((n0 + (1LL << a.shift[i]) - 1) >> a.shift[i]) / a.hop[i];
      T = (T < 0 || ti < T) ? ti : T;
    }
    real_frame = in_range && t < T;
    const int sh = a.shift[octave];
    const long long len = (n0 + (1LL << sh) - 1) >> sh;
    const long long first = (long long)t * a.hop[octave] - n_fft / 2;  // centred frame, zero padded (pad_mode='constant')
    s_g0[row_e] = (long long)b * a.stride[octave] + first;
    s_valid[row_e] = real_frame ? make_int2((int)max(0LL, -first), (int)max(0LL, min((long long)n_fft, len - first))) : make_int2(0, 0);
    // the clip's amplitude pre-scale: applied here when the octave reads the raw audio (level 0), already in the data otherwise
    const float xs = a.xs ? __ldg(a.xs + b) : a.xs_uniform;
    s_xs[row_e] = kXScale * (sh == 0 ? xs : 1.f);
    inv_xs = 1.f / xs;  // exact: xs is a power of two
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_slot;

  if (warp < ISSUER) {
    // ---------------------------------------------------------------- producer (8 warps per row block)
    // Warps w and w + 4 of a row block stage the same 32 frame rows, 8 of the 16 row pairs each; one instruction covers two rows x 64
    // samples: lanes 0-15 read the 16 float4s of one row, lanes 16-31 those of the next (coalesced 256-byte runs), convert to
    // fp16 hi/lo and scatter 8-byte halves of the operand chunks (chunk c of row r at c * kBankLBO + r * 16).
    const int pw = warp & 3, it0 = 8 * ((warp >> 2) & 1);
    const int rb = 128 * mb;            // first row of this warp's row block
    const uint32_t a_off = (uint32_t)mb * 2 * A_HALF;
    const float* level = a.level[octave];
    const int f = lane & 15;
    // The 16 rows this thread feeds are the same for every block of K: their sample pointers (frame-local index 4 f of the
    // first block) and a 2-bit class live in registers.  Class 0: the whole frame exists and the pointer is 16-byte aligned
    // (one vector load); 1: the whole frame exists (four scalar loads); 2: centred-frame zero padding or the clip's end
    // cuts the frame (per-sample predicates).
    const float* rowp[8];
    float rowsc[8];
    uint32_t cls = 0;
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int r = rb + pw * 32 + 2 * (it0 + it) + (lane >> 4);
      const int2 v = s_valid[r];
      rowp[it] = level + s_g0[r] + 4 * f;
      rowsc[it] = s_xs[r];
      const uint32_t c = (v.x == 0 && v.y == n_fft) ? (((reinterpret_cast<uintptr_t>(rowp[it]) & 15) == 0) ? 0u : 1u) : 2u;
      cls |= c << (2 * it);
    }
    for (int kb = 0; kb < n_kb; ++kb) {
      const int s = kb % kUStages;
      const uint32_t phase = (kb / kUStages) & 1;
      mbar_wait_relaxed(&empty_bar[s], phase ^ 1);
      uint8_t* stage = smem + (size_t)s * STAGE;
      if (tid == 0) {
        mbar_arrive_expect_tx(&full_bar[s], B_BYTES);
        bulk_g2s(stage + B_OFF, reinterpret_cast<const uint8_t*>(a.bank_img[octave]) + (size_t)kb * B_BYTES, B_BYTES, &full_bar[s]);
      }
      const int i0 = kb * kUKB + 4 * f;  // frame-local index of this lane's first sample
      // all 16 loads are issued before the first conversion consumes one
      float4 x[8];
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const float* src = rowp[it] + kb * kUKB;
        const uint32_t c = (cls >> (2 * it)) & 3u;
        if (c == 0) {
          x[it] = __ldg(reinterpret_cast<const float4*>(src));
        } else if (c == 1) {
          x[it] = make_float4(__ldg(src), __ldg(src + 1), __ldg(src + 2), __ldg(src + 3));
        } else {
          const int2 v = s_valid[rb + pw * 32 + 2 * (it0 + it) + (lane >> 4)];
          x[it].x = (i0 + 0 >= v.x && i0 + 0 < v.y) ? __ldg(src + 0) : 0.f;
          x[it].y = (i0 + 1 >= v.x && i0 + 1 < v.y) ? __ldg(src + 1) : 0.f;
          x[it].z = (i0 + 2 >= v.x && i0 + 2 < v.y) ? __ldg(src + 2) : 0.f;
          x[it].w = (i0 + 3 >= v.x && i0 + 3 < v.y) ? __ldg(src + 3) : 0.f;
        }
      }
      const uint32_t off0 = a_off + (uint32_t)(f >> 1) * kBankLBO + (uint32_t)(pw * 32 + 2 * it0 + (lane >> 4)) * 16 + (uint32_t)(f & 1) * 8;
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        uint32_t h0, l0, h1, l1;
        const uint64_t ss = f2_pack(rowsc[it], rowsc[it]);
        cas_split2(f2_mul(f2_pack(x[it].x, x[it].y), ss), h0, l0);
        cas_split2(f2_mul(f2_pack(x[it].z, x[it].w), ss), h1, l1);
        const uint32_t off = off0 + 32u * it;  // row r = pw * 32 + 2 (it0 + it) + (lane >> 4)
        *reinterpret_cast<uint2*>(stage + off) = make_uint2(h0, h1);
        *reinterpret_cast<uint2*>(stage + A_HALF + off) = make_uint2(l0, l1);
      }
      fence_proxy_async();
      if (tid != 0) mbar_arrive(&full_bar[s]);
    }
    // ---------------------------------------------------------------- epilogue (4 warps per row block): TMEM lane = frame row within the block
    if ((warp & 7) < 4) {
    mbar_wait_relaxed(&done_bar, 0);
    fence_after_sync();
    const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)mb * 2 * NPAD;
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
        const float sc = a.scale[octave * bpo + k] * inv_xs;
        float re = (u[2 * j] + w[2 * j]) * sc, im = (u[2 * j + 1] + w[2 * j + 1]) * sc;
        if (!real_frame) re = 0.f, im = 0.f;  // beyond the clip's frames: batch padding is zero (KeyDataset.py:242-254)
        if (a.mode == AKE_CQT_LOGMAG) {
          a.out[((long long)b * a.n_bins + bin) * a.T_max + t] = log1pf(sqrtf(re * re + im * im));
        } else {
          reinterpret_cast<float2*>(a.out)[((long long)b * a.n_bins + bin) * a.T_max + t] = make_float2(re, im);
        }
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
          const uint64_t bd = make_desc(B_DESC, base + B_OFF + j * (2 * 2 * NPAD * 16));
#pragma unroll
          for (int q = 0; q < MB; ++q) {
            const uint32_t ab = base + q * 2 * A_HALF + j * 2 * kBankLBO, d = tmem + q * 2 * NPAD;
            mma_f16(d, make_desc(A_DESC, ab), bd, IDESC_WIDE, (kb | j) ? 1u : 0u);
            mma_f16(d, make_desc(A_DESC, ab + A_HALF), bd, IDESC_NARROW, 1u);
          }
        }
        commit(&empty_bar[s]);
        if (kb == n_kb - 1) commit(&done_bar);
      }
      __syncwarp();
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == ISSUER) tmem_dealloc(tmem, TMEM_COLS);
}

struct OctTable {
  int n_oct, shift[16], hop[16];
};

__global__ void cqt_seqlen_kernel(const long long* __restrict__ lengths, long long n_uniform, int B, const OctTable oc, int T_max,
                                  int* __restrict__ seq_len) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const long long n0 = lengths ? lengths[b] : n_uniform;
  long long T = -1;
  for (int i = 0; i < oc.n_oct; ++i) {
    const long long ti = 1 + ((n0 + (1LL << oc.shift[i]) - 1) >> oc.shift[i]) / oc.hop[i];
    T = (T < 0 || ti < T) ? ti : T;
  }
  seq_len[b] = (int)(T < T_max ? T : T_max);
}

// ---- amplitude pre-scale (ake_cqt_set_peak(plan, 0): "measure") ------------------------------------------------------------
// max |x| per clip (bit pattern of a non-negative float orders like the float), then the exact power of two 2^-ceil(log2 max).
__global__ void __launch_bounds__(256) peak_max_kernel(const float* __restrict__ audio, long long stride, const long long* __restrict__ lengths,
                                                       long long n_uniform, unsigned int* __restrict__ peak_bits) {
  const int b = blockIdx.y;
  const long long n = lengths ? lengths[b] : n_uniform;
  const float* x = audio + (long long)b * stride;
  float m = 0.f;
  // vector loads over the 16-byte aligned middle, scalar head / tail
  const long long head = min(n, (long long)((4 - ((reinterpret_cast<uintptr_t>(x) >> 2) & 3)) & 3));
  const long long n4 = (n - head) >> 2;
  const float4* x4 = reinterpret_cast<const float4*>(x + head);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = __ldg(x4 + i);
    m = fmaxf(fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
  }
  if (blockIdx.x == 0) {
    for (long long i = threadIdx.x; i < head; i += blockDim.x) m = fmaxf(m, fabsf(__ldg(x + i)));
    for (long long i = head + 4 * n4 + threadIdx.x; i < n; i += blockDim.x) m = fmaxf(m, fabsf(__ldg(x + i)));
  }
  for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(peak_bits + b, __float_as_uint(m));  // NaN samples are ignored by fmaxf
}
__global__ void peak_scale_kernel(const unsigned int* __restrict__ peak_bits, int B, float* __restrict__ xs) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float m = __uint_as_float(peak_bits[b]);
  float sc = 1.f;
  if (m > 0.f && m < INFINITY) {
    int e;
    frexpf(m, &e);                        // m = f * 2^e, f in [0.5, 1)  ->  m * 2^-e in [0.5, 1)
    e = max(-100, min(100, e));           // keep 2^-e and its inverse finite normal floats
    sc = ldexpf(1.f, -e);
  }
  xs[b] = sc;
}

struct CqtWs {
  long long* d_len;
  float* d_xs;              // per-clip amplitude pre-scale (measure mode)
  unsigned int* d_peak;
  float* level[16];
  long long stride[16];
};

// d_len / d_xs: one entry per clip of the batch; level buffers: one group of G clips (reused by every group)
static CqtWs carve(const ake_cqt* p, Arena& ar, int B, int G, long long n_max) {
  CqtWs w{};
  w.d_len = ar.take<long long>(B);
  w.d_xs = ar.take<float>(B);
  w.d_peak = ar.take<unsigned int>(B);
  for (int i = 1; i < p->n_levels; ++i) {
    w.stride[i] = (long long)align_up((size_t)len_at(n_max, i), 32);  // the cascade kernel stores whole rows of 32 samples
    w.level[i] = ar.take<float>((size_t)G * w.stride[i]);
  }
  return w;
}

// lengths_host -> w.d_len through the plan's page-locked ring (asynchronous: the host never waits for earlier work on `st`)
static void stage_lengths(ake_cqt* p, const int64_t* lengths_host, int B, long long* d_len, cudaStream_t st) {
  const int s = p->h_len_next;
  p->h_len_next = (s + 1) % ake_cqt::kLenSlots;
  if (!p->h_len_done[s]) AKE_CUDA(cudaEventCreateWithFlags(&p->h_len_done[s], cudaEventDisableTiming));
  else AKE_CUDA(cudaEventSynchronize(p->h_len_done[s]));  // four calls ago: normally long complete
  if (p->h_len_cap[s] < (size_t)B) {
    if (p->h_len[s]) cudaFreeHost(p->h_len[s]);
    p->h_len[s] = nullptr, p->h_len_cap[s] = 0;
    AKE_CUDA(cudaMallocHost(&p->h_len[s], sizeof(long long) * (size_t)B));
    p->h_len_cap[s] = (size_t)B;
  }
  for (int b = 0; b < B; ++b) p->h_len[s][b] = lengths_host[b];
  AKE_CUDA(cudaMemcpyAsync(d_len, p->h_len[s], sizeof(long long) * B, cudaMemcpyHostToDevice, st));
  AKE_CUDA(cudaEventRecord(p->h_len_done[s], st));
}

// Clips per group of the front-end.  The level buffers can be reused group by group so that the decimated levels stay resident
// in L2 (DRAM then sees the audio once and the output once), but on B200 that is SLOWER than one pass over the whole batch
// (measured, 256 standard clips: 0.90 ms in one group, 1.15 ms in 4 groups of 64, 1.76 ms in 16 groups of 16 -- the persistent
// kernels' ramp-up and tail are paid per launch, and neither kernel is DRAM-bound), so the default is one group; AKE_CQT_GROUP=n
// keeps the mechanism available for experiments and for workspaces that must stay small.
static int cqt_group_clips(int B) {
  if (const char* e = getenv("AKE_CQT_GROUP")) {
    const int g = atoi(e);
    return g <= 0 ? B : std::min(B, g);
  }
  return B;
}

// lengths_host: validated and staged here; lengths_dev: already on the device (the host-buffer pipeline uploads every clip's
// length once per batch); at most one of the two.
void run_cqt(ake_cqt* p, const float* audio, long long stride, const int64_t* lengths_host, const long long* lengths_dev, int B,
             long long n_max, int mode, float* out, int T_max, int* seq_len_out, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (n_max >= (1LL << 31) - 65536) fail(AKE_ERR_UNSUPPORTED, "clips of 2^31 samples or more are not supported (32-bit sample indices in the cascade)");
  ensure_device(p);
  const int G = cqt_group_clips(B);
  Arena ar(ws, ws_bytes);
  CqtWs w = carve(p, ar, B, G, n_max);
  const long long* d_len = lengths_dev;
  if (lengths_host) {
    for (int b = 0; b < B; ++b)
      if (lengths_host[b] < 0 || lengths_host[b] > n_max) fail(AKE_ERR_INVALID, "lengths_host[%d]=%lld outside [0, n_max]", b, (long long)lengths_host[b]);
    stage_lengths(p, lengths_host, B, w.d_len, st);
    d_len = w.d_len;
  }
  if (!p->npad)
    fail(AKE_ERR_UNSUPPORTED, "bins_per_octave %d / n_fft %d: the filter-bank kernel is built for 36 and 12 bins per octave, n_fft %% 64 == 0",
         p->bpo, p->n_fft);
  // ---- amplitude pre-scale: one exact power of two per clip
  const float* d_xs = nullptr;
  float xs_uniform = 1.f;
  if (p->peak > 0.f) {
    int e;
    frexpf(p->peak, &e);  // peak = f * 2^e, f in [0.5, 1): |x| <= peak  ->  |x * 2^-e| <= 1
    if (std::ldexp(0.5f, e) == p->peak) --e;  // peak itself a power of two
    xs_uniform = std::ldexp(1.f, -std::max(-100, std::min(100, e)));
  } else {
    ProfScope prof("cqt.peak", st);
    AKE_CUDA(cudaMemsetAsync(w.d_peak, 0, sizeof(unsigned int) * B, st));
    const int bx = (int)std::max<long long>(1, std::min<long long>(cdiv64(n_max, 256 * 4 * 8), std::max(1, 4 * sm_count() / B)));
    peak_max_kernel<<<dim3(bx, B), 256, 0, st>>>(audio, stride, d_len, n_max, w.d_peak);
    AKE_LAUNCHED();
    peak_scale_kernel<<<cdiv(B, 128), 128, 0, st>>>(w.d_peak, B, w.d_xs);
    AKE_LAUNCHED();
    d_xs = w.d_xs;
  }
  w.stride[0] = stride;
  const int n_sm = sm_count();
  const bool grouped = G < B;
  const int d_max = p->n_levels - 1;  // deepest decimated level
  for (int g0 = 0; g0 < B; g0 += G) {
    const int nb = std::min(G, B - g0);
    const long long* g_len = d_len ? d_len + g0 : nullptr;
    const float* g_xs = d_xs ? d_xs + g0 : nullptr;
    w.level[0] = const_cast<float*>(audio) + (size_t)g0 * stride;
    if (d_max > 0) {
      // resampling cascade: two octave steps per pass (level p -> p+1, p+2), one persistent warp-specialised CTA per SM
      ProfScope prof("cqt.decimate", st);
      ensure_dyn_smem(cascade_umma_kernel, kCasSmemTotal);
      for (int lv = 0; lv < d_max; lv += 2) {
        CascadeArgs ca{};
        ca.in = w.level[lv], ca.in_stride = w.stride[lv];
        ca.out1 = w.level[lv + 1], ca.stride1 = w.stride[lv + 1];
        ca.n_levels = std::min(2, d_max - lv);
        if (ca.n_levels == 2) ca.out2 = w.level[lv + 2], ca.stride2 = w.stride[lv + 2];
        ca.lengths = g_len, ca.n_uniform = n_max, ca.level_in = lv;
        ca.tiles_per_clip = (int)cdiv64(len_at(n_max, lv + 1), kCasOwn1);
        ca.n_tiles = ca.tiles_per_clip * nb;
        // the kernel decodes tile -> clip by multiply-high, exact while n_tiles * tiles_per_clip < 2^32 (very many very long clips
        // exceed it): refuse instead of decoding wrongly
        if ((long long)ca.tiles_per_clip * nb * ca.tiles_per_clip >= (1LL << 32))
          fail(AKE_ERR_UNSUPPORTED, "%d clips x %d tiles exceed the cascade's index decode range; split the batch", nb, ca.tiles_per_clip);
        ca.img = p->d_dec_img;
        // level lv+1 < d_max is consumed by octave lv+1's filter bank only (the next pass reads level lv+2): when its frames do
        // not overlap, the samples between them are never written
        const int hop1 = p->oct_hop[lv + 1], nfft1 = p->oct_nfft[lv + 1];
        ca.sparse_hop = (ca.n_levels == 2 && hop1 >= nfft1 + 64) ? hop1 : 0, ca.sparse_nfft = nfft1;
        ca.stream_in = (grouped && lv == 0) ? 1 : 0;  // the audio is read once: do not let it displace the resident levels
        ca.xs = lv == 0 ? g_xs : nullptr, ca.xs_uniform = lv == 0 ? xs_uniform : 1.f;
        const int grid = std::min(ca.n_tiles, n_sm);  // persistent: one CTA per SM
        cascade_umma_kernel<<<grid, kCasThreads, kCasSmemTotal, st>>>(ca);
        AKE_LAUNCHED();
      }
    }
    {
      // tensor cores: every octave in one launch (grid.y = octave)
      ProfScope prof("cqt.bank", st);
      const long long rows = (long long)nb * T_max;
      BankArgs ba{};
      for (int i = 0; i < p->n_oct; ++i) {
        const int lv = p->oct_level[i];
        ba.level[i] = w.level[lv], ba.stride[i] = w.stride[lv], ba.shift[i] = lv, ba.hop[i] = p->oct_hop[i], ba.nfft[i] = p->oct_nfft[i];
        ba.bank_img[i] = p->d_bank_img[p->oct_bank[i]];
      }
      ba.lengths = g_len, ba.n_uniform = n_max, ba.n_oct = p->n_oct, ba.B = nb, ba.T_max = T_max;
      ba.n_bins = p->n_bins, ba.bpo = p->bpo, ba.mode = mode, ba.scale = p->d_scale_umma;
      ba.out = out + (size_t)g0 * p->n_bins * T_max * (mode == AKE_CQT_COMPLEX ? 2 : 1);
      ba.xs = g_xs, ba.xs_uniform = xs_uniform;
      dim3 grid((unsigned)cdiv64(rows, 128 * kBankMB), p->n_oct);
      if (p->npad == 80) {
        constexpr size_t smem = kUStages * bank_stage_bytes(80, kBankMB);
        ensure_dyn_smem(cqt_bank_umma_kernel<80, kBankMB>, smem);
        cqt_bank_umma_kernel<80, kBankMB><<<grid, kBankThreads, smem, st>>>(ba);
      } else {
        constexpr size_t smem = kUStages * bank_stage_bytes(32, kBankMB);
        ensure_dyn_smem(cqt_bank_umma_kernel<32, kBankMB>, smem);
        cqt_bank_umma_kernel<32, kBankMB><<<grid, kBankThreads, smem, st>>>(ba);
      }
      AKE_LAUNCHED();
    }
  }
  if (seq_len_out) {
    OctTable oc{};
    oc.n_oct = p->n_oct;
    for (int i = 0; i < p->n_oct; ++i) oc.shift[i] = p->oct_level[i], oc.hop[i] = p->oct_hop[i];
    cqt_seqlen_kernel<<<cdiv(B, 128), 128, 0, st>>>(d_len, n_max, B, oc, T_max, seq_len_out);
    AKE_LAUNCHED();
  }
}

}  // namespace ake

extern "C" {

int ake_cqt_create(double sr, int hop_length, int n_bins, int bins_per_octave, double fmin, double filter_scale,
                   double sparsity, ake_cqt** out) {
  return ake_cqt_create_ex(sr, hop_length, n_bins, bins_per_octave, fmin, filter_scale, sparsity, AKE_CQT_RECURSION_092, out);
}

int ake_cqt_set_peak(ake_cqt* p, float peak) {
  return guarded([&] {
    if (!p) fail(AKE_ERR_INVALID, "null argument");
    if (!(peak >= 0.f) || std::isinf(peak)) fail(AKE_ERR_INVALID, "peak must be a finite non-negative number (0 = measure per clip)");
    p->peak = peak;
  });
}

int ake_cqt_create_ex(double sr, int hop_length, int n_bins, int bins_per_octave, double fmin, double filter_scale,
                      double sparsity, int recursion, ake_cqt** out) {
  return guarded([&] {
    if (!out) fail(AKE_ERR_INVALID, "null argument");
    ake_cqt* p = new ake_cqt();
    p->sr = sr, p->hop = hop_length, p->n_bins = n_bins, p->bpo = bins_per_octave, p->fmin = fmin;
    p->filter_scale = filter_scale, p->sparsity = sparsity, p->recursion = recursion;
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
  free_device_state(p);
  for (int s = 0; s < ake_cqt::kLenSlots; ++s) {
    if (p->h_len[s]) cudaFreeHost(p->h_len[s]);
    if (p->h_len_done[s]) cudaEventDestroy(p->h_len_done[s]);
  }
  delete p;
}

int ake_cqt_n_fft(const ake_cqt* p) { return p ? p->n_fft : AKE_ERR_INVALID; }
int ake_cqt_n_bins(const ake_cqt* p) { return p ? p->n_bins : AKE_ERR_INVALID; }
int ake_cqt_frames(const ake_cqt* p, int64_t n) { return (p && n >= 0) ? frames_for(p, n) : AKE_ERR_INVALID; }

int ake_cqt_get_bank(const ake_cqt* p, float* bank_host, int64_t cap) {
  return guarded([&] {
    if (!p || !bank_host) fail(AKE_ERR_INVALID, "null argument");
    const std::vector<float>& bank = p->banks[0];  // the top octave's bank
    if (cap < (int64_t)bank.size()) fail(AKE_ERR_INVALID, "bank needs %zu floats", bank.size());
    std::copy(bank.begin(), bank.end(), bank_host);
  });
}

int ake_cqt_get_decimator(const ake_cqt* p, float* taps_host, int cap) {
  if (!p || !taps_host || cap < kHalfTaps) return AKE_ERR_INVALID;
  for (int m = 0; m < kHalfTaps; ++m) taps_host[m] = (float)p->dec_half[m];
  return kHalfTaps;
}

size_t ake_cqt_workspace_bytes(const ake_cqt* p, int B, int64_t n_max) {
  if (!p || B <= 0 || n_max <= 0) return 0;
  // sized for the whole batch in one group, so that the grouping policy (L2 size of the device that runs the plan, experiments)
  // never has to be known when the caller allocates
  Arena ar(nullptr, 0);
  carve(p, ar, B, B, n_max);
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
    run_cqt(p, audio_dev, stride, lengths_host, nullptr, B, n_max, mode, out_dev, T_max, seq_len_out_dev, ws_dev, ws_bytes,
            static_cast<cudaStream_t>(stream));
  });
}

}  // extern "C"

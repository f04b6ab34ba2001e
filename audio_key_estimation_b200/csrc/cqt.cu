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

__constant__ float c_dec_taps[kHalfTaps];  // h[|j|] * sqrt(2), identical for every plan (kaiser_fast is fixed)

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
  float taps[kHalfTaps];
  for (int m = 0; m < kHalfTaps; ++m) taps[m] = (float)(p->dec_half[m] * std::sqrt(2.0));  // resample(scale=True): / sqrt(1/2)
  AKE_CUDA(cudaMemcpyToSymbol(c_dec_taps, taps, sizeof taps));
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
// One octave step of the resampling cascade: out[t] = sqrt(2) * sum_{|j|<=31} h[|j|] * in[2t + j], zero extended,
// for t < floor(n_in/2); librosa pads the result to ceil(n_in/2) samples with a zero.
// Block: 256 threads x 8 outputs.  The input span is de-interleaved into even/odd phases in shared memory so
// each thread reads two contiguous windows with 128-bit loads and runs 504 FFMAs from registers.
constexpr int kDecOutPerThread = 8;
constexpr int kDecThreads = 256;
constexpr int kDecOutPerBlock = kDecOutPerThread * kDecThreads;

__global__ void __launch_bounds__(kDecThreads) decimate2_kernel(const float* __restrict__ in, long long in_stride,
                                                                  float* __restrict__ out, long long out_stride,
                                                                  const long long* __restrict__ lengths, long long n_uniform,
                                                                  int level /* input is level-1 */) {
  constexpr int SPAN = kDecOutPerBlock + 32;  // phase samples m in [t0-16, t0+OUT+16)
  __shared__ __align__(16) float se[SPAN];
  __shared__ __align__(16) float so[SPAN];
  const int b = blockIdx.y;
  const long long n0 = lengths ? lengths[b] : n_uniform;
  const long long n_in = (n0 + (1LL << (level - 1)) - 1) >> (level - 1);
  const long long n_half = n_in >> 1, n_out = (n_in + 1) >> 1;
  const long long t0 = (long long)blockIdx.x * kDecOutPerBlock;
  if (t0 >= n_out) return;
  const float* src = in + b * in_stride;
  for (int i = threadIdx.x; i < SPAN; i += kDecThreads) {
    const long long m = t0 - 16 + i;
    const long long s = 2 * m;
    se[i] = (s >= 0 && s < n_in) ? __ldg(src + s) : 0.f;
    so[i] = (s + 1 >= 0 && s + 1 < n_in) ? __ldg(src + s + 1) : 0.f;
  }
  __syncthreads();
  const int lt = threadIdx.x * kDecOutPerThread;  // local output index; phase index of output t is lt + 16
  float xe[kDecOutPerThread + 32], xo[kDecOutPerThread + 32];
#pragma unroll
  for (int q = 0; q < (kDecOutPerThread + 32) / 4; ++q) {
    const float4 a = *reinterpret_cast<const float4*>(se + lt + 4 * q);
    const float4 c = *reinterpret_cast<const float4*>(so + lt + 4 * q);
    xe[4 * q] = a.x, xe[4 * q + 1] = a.y, xe[4 * q + 2] = a.z, xe[4 * q + 3] = a.w;
    xo[4 * q] = c.x, xo[4 * q + 1] = c.y, xo[4 * q + 2] = c.z, xo[4 * q + 3] = c.w;
  }
  float acc[kDecOutPerThread];
#pragma unroll
  for (int r = 0; r < kDecOutPerThread; ++r) {
    // output t = t0 + lt + r sits at phase index r + 16: even taps j = 2e -> xe[r+16+e], odd j = 2o+1 -> xo[r+16+o]
    float s = c_dec_taps[0] * xe[r + 16];
#pragma unroll
    for (int e = 1; e <= 15; ++e) s = fmaf(c_dec_taps[2 * e], xe[r + 16 + e] + xe[r + 16 - e], s);
#pragma unroll
    for (int o = 0; o <= 15; ++o) s = fmaf(c_dec_taps[2 * o + 1], xo[r + 16 + o] + xo[r + 15 - o], s);
    acc[r] = s;
  }
  float* dst = out + b * out_stride;
#pragma unroll
  for (int r = 0; r < kDecOutPerThread; ++r) {
    const long long t = t0 + lt + r;
    if (t < n_out) dst[t] = t < n_half ? acc[r] : 0.f;
  }
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

template <int NPAD>
__global__ void __launch_bounds__(160) cqt_bank_umma_kernel(const BankArgs a) {
  using namespace umma;
  constexpr uint32_t A_HALF = 128 * kUKB * 2;                // 16 KB: 8 chunks x 128 rows x 16 B
  constexpr uint32_t B_BYTES = (kUKB / 8) * 2 * NPAD * 16;   // 8 chunks x 2*NPAD rows x 16 B
  constexpr uint32_t STAGE = 2 * A_HALF + B_BYTES;
  constexpr uint32_t TMEM_COLS = (2 * NPAD <= 64) ? 64 : ((2 * NPAD <= 128) ? 128 : 256);
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t full_bar[kUStages], empty_bar[kUStages], done_bar;
  __shared__ uint32_t tmem_slot;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int octave = blockIdx.y;
  const long long m0 = (long long)blockIdx.x * 128;
  const int n_kb = a.n_fft / kUKB;

  if (warp == 4) tmem_alloc(&tmem_slot, TMEM_COLS);
  if (tid == 0) {
    for (int s = 0; s < kUStages; ++s) mbar_init(&full_bar[s], 128), mbar_init(&empty_bar[s], 1);
    mbar_init(&done_bar, 1);
    mbar_init_fence();
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_slot;

  if (warp < 4) {
    // ---------------------------------------------------------------- producer: frame row `tid`
    const long long m = m0 + tid;
    const bool in_range = m < (long long)a.B * a.T_max;
    const int b = in_range ? (int)(m / a.T_max) : 0, t = in_range ? (int)(m % a.T_max) : 0;
    const long long n0 = a.lengths ? a.lengths[b] : a.n_uniform;
    long long T = -1;
    for (int i = 0; i < a.n_oct; ++i) {
      const long long ti = 1 + ((n0 + (1LL << i) - 1) >> i) / (a.hop0 >> i);
      T = (T < 0 || ti < T) ? ti : T;
    }
    const bool real_frame = in_range && t < T;
    const long long len = (n0 + (1LL << octave) - 1) >> octave;
    const float* src = a.level[octave] + (long long)b * a.stride[octave];
    const long long first = (long long)t * (a.hop0 >> octave) - a.n_fft / 2;  // centred frame, zero padded (pad_mode='constant')

    for (int kb = 0; kb < n_kb; ++kb) {
      const int s = kb % kUStages;
      const uint32_t phase = (kb / kUStages) & 1;
      mbar_wait(&empty_bar[s], phase ^ 1);
      uint8_t* stage = smem + (size_t)s * STAGE;
      if (tid == 0) {
        mbar_arrive_expect_tx(&full_bar[s], B_BYTES);
        bulk_g2s(stage + 2 * A_HALF, reinterpret_cast<const uint8_t*>(a.bank_img) + (size_t)kb * B_BYTES, B_BYTES, &full_bar[s]);
      }
      const long long i0 = first + (long long)kb * kUKB;
      float x[kUKB];
      const float* p = src + i0;
      if (real_frame && i0 >= 0 && i0 + kUKB <= len && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) {
#pragma unroll
        for (int q = 0; q < kUKB / 4; ++q) {
          const float4 v = __ldg(reinterpret_cast<const float4*>(p) + q);
          x[4 * q] = v.x, x[4 * q + 1] = v.y, x[4 * q + 2] = v.z, x[4 * q + 3] = v.w;
        }
      } else {
#pragma unroll
        for (int q = 0; q < kUKB; ++q) {
          const long long i = i0 + q;
          x[q] = (real_frame && i >= 0 && i < len) ? __ldg(src + i) : 0.f;
        }
      }
#pragma unroll
      for (int c = 0; c < kUKB / 8; ++c) {
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          __half h0, l0, h1, l1;
          split_f16(x[8 * c + 2 * e] * kXScale, h0, l0);
          split_f16(x[8 * c + 2 * e + 1] * kXScale, h1, l1);
          hi[e] = pack_h2(h0, h1), lo[e] = pack_h2(l0, l1);
        }
        *reinterpret_cast<uint4*>(stage + c * 2048 + tid * 16) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4*>(stage + A_HALF + c * 2048 + tid * 16) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
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
    (void)lane;
  } else if (lane == 0) {
    // ---------------------------------------------------------------- MMA issuer
    constexpr uint64_t A_DESC = desc_hi(128 * 16);       // chunk stride 2 KB (128 rows x 16 B)
    constexpr uint64_t B_DESC = desc_hi(2 * NPAD * 16);  // chunk stride = 2*NPAD rows x 16 B
    constexpr uint32_t IDESC_WIDE = idesc_f16(2 * NPAD), IDESC_NARROW = idesc_f16(NPAD);
    for (int kb = 0; kb < n_kb; ++kb) {
      const int s = kb % kUStages;
      mbar_wait(&full_bar[s], (kb / kUStages) & 1);
      fence_after_sync();
      const uint32_t base = smem_u32(smem + (size_t)s * STAGE);
#pragma unroll
      for (int j = 0; j < kUKB / 16; ++j) {
        const uint64_t bd = make_desc(B_DESC, base + 2 * A_HALF + j * (2 * 2 * NPAD * 16));
        mma_f16(tmem, make_desc(A_DESC, base + j * 4096), bd, IDESC_WIDE, (kb | j) ? 1u : 0u);
        mma_f16(tmem, make_desc(A_DESC, base + A_HALF + j * 4096), bd, IDESC_NARROW, 1u);
      }
      commit(&empty_bar[s]);
    }
    commit(&done_bar);
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
    w.stride[i] = (long long)align_up((size_t)len_at(n_max, i), 4);
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
  for (int i = 1; i < p->n_oct; ++i) {
    ProfScope prof("cqt.decimate", st);
    const long long n_out = len_at(n_max, i);
    dim3 grid((unsigned)cdiv64(n_out, kDecOutPerBlock), B);
    decimate2_kernel<<<grid, kDecThreads, 0, st>>>(w.level[i - 1], w.stride[i - 1], w.level[i], w.stride[i], d_len, n_max, i);
    AKE_LAUNCHED();
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
      constexpr size_t smem = kUStages * (2 * 128 * kUKB * 2 + (kUKB / 8) * 2 * 80 * 16);
      static bool configured = false;
      if (!configured) {
        AKE_CUDA(cudaFuncSetAttribute(cqt_bank_umma_kernel<80>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
      }
      cqt_bank_umma_kernel<80><<<grid, 160, smem, st>>>(ba);
    } else {
      constexpr size_t smem = kUStages * (2 * 128 * kUKB * 2 + (kUKB / 8) * 2 * 32 * 16);
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

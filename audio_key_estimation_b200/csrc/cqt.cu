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

namespace ake {

constexpr int kHalfTaps = 32;          // h[0..31]; full filter has 63 taps (|j| <= 31)
constexpr double kPi = 3.14159265358979323846;
constexpr double kHannBandwidth = 1.50018310546875;  // librosa.filters.WINDOW_BANDWIDTHS['hann']
constexpr double kBwFastest = 0.85;                   // resampy kaiser_fast rolloff (librosa.audio.BW_FASTEST)

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
  float* d_bank = nullptr;
  float* d_scale = nullptr;
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
  if (p->d_bank) return;
  float taps[kHalfTaps];
  for (int m = 0; m < kHalfTaps; ++m) taps[m] = (float)(p->dec_half[m] * std::sqrt(2.0));  // resample(scale=True): / sqrt(1/2)
  AKE_CUDA(cudaMemcpyToSymbol(c_dec_taps, taps, sizeof taps));
  AKE_CUDA(cudaMalloc(&p->d_scale, sizeof(float) * p->out_scale.size()));
  AKE_CUDA(cudaMemcpy(p->d_scale, p->out_scale.data(), sizeof(float) * p->out_scale.size(), cudaMemcpyHostToDevice));
  AKE_CUDA(cudaMalloc(&p->d_bank, sizeof(float) * p->bank.size()));
  AKE_CUDA(cudaMemcpy(p->d_bank, p->bank.data(), sizeof(float) * p->bank.size(), cudaMemcpyHostToDevice));
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

// Filter-bank contraction + magnitude / scale / log1p epilogue for one octave.
//   rows m = (clip, frame) flattened, 32 per block; 72 = 2*36 real filters; K = n_fft in chunks of 32.
// Thread tile: 8 frames x 4 filters (two complex bins), 72 threads per block (4 frame groups x 18 bin pairs).
constexpr int kBankRows = 32, kBankBK = 32, kBankPitchA = 36;

template <int NF2 /* 2*bpo real filters, multiple of 4 */>
__global__ void cqt_bank_kernel(const float* __restrict__ sig, long long sig_stride, const long long* __restrict__ lengths,
                                long long n_uniform, int octave, int hop_i, int n_fft, const float* __restrict__ bank,
                                const float* __restrict__ scale, int B, int T_max, int n_oct, int hop0, int n_bins, int mode,
                                float* __restrict__ out) {
  constexpr int NTB = NF2 / 4;  // threads along filters
  __shared__ __align__(16) float As[kBankBK][kBankPitchA];  // [k][row]
  __shared__ __align__(16) float Bs[kBankBK][NF2];          // [k][filter]
  __shared__ long long s_base[kBankRows];   // offset of the clip's sample 0 inside `sig`
  __shared__ long long s_first[kBankRows];  // clip-relative index of the frame's sample k = 0 (negative at the left edge)
  __shared__ long long s_len[kBankRows];    // valid samples of this clip at this octave
  __shared__ int s_T[kBankRows];            // 1: real frame, 0: batch padding (t >= frames of the clip), -1: no row
  const int tid = threadIdx.x;
  const long long m0 = (long long)blockIdx.x * kBankRows;
  for (int row = tid; row < kBankRows; row += blockDim.x) {
    const long long m = m0 + row;
    s_T[row] = -1, s_base[row] = 0, s_first[row] = 0, s_len[row] = 0;
    if (m < (long long)B * T_max) {
      const int b = m / T_max, t = m % T_max;
      const long long n0 = lengths ? lengths[b] : n_uniform;
      long long T = -1;
      for (int i = 0; i < n_oct; ++i) {
        const long long ti = 1 + ((n0 + (1LL << i) - 1) >> i) / (hop0 >> i);
        T = (T < 0 || ti < T) ? ti : T;
      }
      s_T[row] = t < T ? 1 : 0;
      s_base[row] = b * sig_stride;
      s_first[row] = (long long)t * hop_i - n_fft / 2;
      s_len[row] = (n0 + (1LL << octave) - 1) >> octave;
    }
  }
  __syncthreads();
  const int fg = tid / NTB, bg = tid % NTB;  // frame group (8 frames), filter group (4 filters)
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < n_fft; k0 += kBankBK) {
    __syncthreads();
    for (int i = tid; i < kBankRows * kBankBK; i += blockDim.x) {
      const int row = i / kBankBK, kk = i % kBankBK;
      float v = 0.f;
      if (s_T[row] == 1) {
        const long long idx = s_first[row] + k0 + kk;  // centred frame, zero padded (pad_mode='constant')
        if (idx >= 0 && idx < s_len[row]) v = __ldg(sig + s_base[row] + idx);
      }
      As[kk][row] = v;
    }
    for (int i = tid; i < kBankBK * NF2; i += blockDim.x) {
      const int kk = i / NF2, f = i % NF2;
      Bs[kk][f] = __ldg(bank + (long long)f * n_fft + k0 + kk);
    }
    __syncthreads();
#pragma unroll 8
    for (int kk = 0; kk < kBankBK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][fg * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][fg * 8 + 4]);
      const float4 w = *reinterpret_cast<const float4*>(&Bs[kk][bg * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
    }
  }
  // epilogue
  const int bpo = NF2 / 2;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = fg * 8 + i;
    if (s_T[row] < 0) continue;
    const long long m = m0 + row;
    const int b = m / T_max, t = m % T_max;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int k = bg * 2 + j;
      const int bin = n_bins - bpo * (octave + 1) + k;
      const float s = scale[octave * bpo + k];
      float re = acc[i][2 * j] * s, im = acc[i][2 * j + 1] * s;
      if (s_T[row] == 0) re = 0.f, im = 0.f;  // beyond the clip's frames: batch padding is zero (KeyDataset.py:242-254)
      if (mode == AKE_CQT_LOGMAG) {
        out[((long long)b * n_bins + bin) * T_max + t] = log1pf(sqrtf(re * re + im * im));
      } else {
        float2* o2 = reinterpret_cast<float2*>(out) + ((long long)b * n_bins + bin) * T_max + t;
        *o2 = make_float2(re, im);
      }
    }
  }
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
  for (int i = 0; i < p->n_oct; ++i) {
    ProfScope prof("cqt.bank", st);
    if (2 * p->bpo == 72) {
      cqt_bank_kernel<72><<<(unsigned)cdiv64(rows, kBankRows), 4 * 18, 0, st>>>(
          w.level[i], w.stride[i], d_len, n_max, i, p->hop >> i, p->n_fft, p->d_bank, p->d_scale, B, T_max, p->n_oct, p->hop,
          p->n_bins, mode, out);
    } else if (2 * p->bpo == 24) {
      cqt_bank_kernel<24><<<(unsigned)cdiv64(rows, kBankRows), 4 * 6, 0, st>>>(
          w.level[i], w.stride[i], d_len, n_max, i, p->hop >> i, p->n_fft, p->d_bank, p->d_scale, B, T_max, p->n_oct, p->hop,
          p->n_bins, mode, out);
    } else {
      fail(AKE_ERR_UNSUPPORTED, "bins_per_octave %d: 36 and 12 are built", p->bpo);
    }
    AKE_LAUNCHED();
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
  if (p->d_bank) cudaFree(p->d_bank);
  if (p->d_scale) cudaFree(p->d_scale);
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

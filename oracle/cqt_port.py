"""CPU restatement of the reference's constant-Q front-end -- TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED.  The reference computes its input features with one library call,
``librosa.cqt(y, sr=rate, hop_length=round(rate/frames), bins_per_octave=36, n_bins=36*octaves)``
followed by ``abs -> log(1 + x) -> reshape(1, n_bins, T) -> .double()``
(KeyDataset.py:485-509; the same call at equivariance_test.py:155-170).  The arithmetic lives in
librosa 0.9.2 + resampy 0.3.1 (requirements.txt:250, :245), which are neither vendored under the
reference tree nor installed here, and the reference holds no CQT fixture.  This module restates
the published algorithm of those two packages (function by function, names kept) in numpy:

    librosa.core.constantq.vqt (gamma = 0)          -> cqt()
    librosa.core.constantq.__cqt_filter_fft          -> cqt_filter_fft()
    librosa.filters.constant_q / constant_q_lengths  -> constant_q() / constant_q_lengths()
    librosa.util.sparsify_rows                       -> sparsify_rows()
    librosa.core.constantq.__cqt_response (+ stft, window='ones', center=True, pad 'constant')
                                                     -> cqt_response()
    librosa.core.audio.resample(orig_sr=2, target_sr=1, 'kaiser_fast', scale=True)
      = resampy.resample + fix_length + / sqrt(1/2)  -> resample_half()
    resampy.filters.sinc_window(16, 9, kaiser(beta), 0.85)  -> kaiser_fast_window()

It follows librosa's own formulation (sparse FFT-domain basis times the rFFT of rectangular
frames) -- deliberately NOT the dense time-domain bank the CUDA kernels contract with -- so that
the GPU path is checked against an independent derivation.  Anchors available without librosa:
analytic known-answer tests (a pure tone at a bin centre peaks at that bin with
|C| ~ A/2 * sqrt(length)), linearity and time-shift properties (tests/test_oracle_cqt.py).
"""
from __future__ import annotations

import math
from functools import lru_cache
from typing import Optional, Tuple

import numpy as np

C1_HZ = 32.70319566257483  # librosa.note_to_hz('C1')
HANN_BANDWIDTH = 1.50018310546875  # librosa.filters.WINDOW_BANDWIDTHS['hann']
BW_FASTEST = 0.85  # librosa.core.audio.BW_FASTEST (resampy kaiser_fast rolloff)


class ParameterError(ValueError):
    """Stands in for librosa.util.exceptions.ParameterError."""


# ------------------------------------------------------------------------------- resampy restated
@lru_cache(maxsize=None)
def kaiser_fast_window() -> Tuple[np.ndarray, int]:
    """resampy.filters.sinc_window(num_zeros=16, precision=9, rolloff=0.85, kaiser beta=8.5555...)."""
    num_zeros, precision, rolloff, beta = 16, 9, 0.85, 8.555504641634386
    num_bits = 2 ** precision
    n = num_bits * num_zeros
    sinc_win = rolloff * np.sinc(rolloff * np.linspace(0, num_zeros, num=n + 1, endpoint=True))
    taper = np.kaiser(2 * n + 1, beta)[n:]
    return taper * sinc_win, num_bits


def decimator_taps() -> np.ndarray:
    """The 63 taps h[j], |j| <= 31, that resampy's interpolation loop visits at ratio 1/2.

    resampy.interpn.resample_f: with sample_ratio 1/2, ``index_step = 256`` and every output sits on
    an input sample (``frac = 0``), so the left wing reads interp_win[256 i], i = 0..31 against
    x[2t - i] and the right wing interp_win[256 (k + 1)], k = 0..30 against x[2t + k + 1]; the
    window was pre-multiplied by the ratio (``interp_win *= sample_ratio``)."""
    win, num_bits = kaiser_fast_window()
    step = num_bits // 2
    half = 0.5 * win[::step]  # 33 entries; index 32 (the window's end, value 0) is never reached
    left = half[: (len(win)) // step]           # i_max = nwin // 256 = 32
    right = half[1: 1 + (len(win) - step) // step]  # k_max = (nwin - 256) // 256 = 31
    return np.concatenate([right[::-1], left])  # j = -31 .. 31  (symmetric)


def resample_half(y: np.ndarray) -> np.ndarray:
    """librosa.resample(y, orig_sr=2, target_sr=1, res_type='kaiser_fast', fix=True, scale=True).

    resampy produces int(n * 0.5) samples y_hat[t] = sum_j h[j] x[2t + j] (x zero-extended: the loop
    bounds ``i_max = min(n + 1, ..)`` / ``k_max = min(n_orig - n - 1, ..)`` simply stop at the
    signal's edges); librosa then zero-pads to ceil(n / 2) samples and divides by sqrt(ratio)."""
    n = y.shape[-1]
    h = decimator_taps().astype(y.dtype)
    n_out = n // 2
    xp = np.concatenate([np.zeros(31, y.dtype), y, np.zeros(32, y.dtype)])
    # out[t] = sum_{j=-31..31} h[j] * x[2t + j] = sum_m h[m - 31] * xp[2t + m]
    full = np.convolve(xp, h[::-1], mode="valid")  # full[s] = sum_m xp[s + m] h[m-31]
    out = full[0: 2 * n_out: 2][:n_out]
    res = np.zeros(int(math.ceil(n / 2)), y.dtype)
    res[:n_out] = out
    return (res / np.sqrt(0.5)).astype(y.dtype)


# ------------------------------------------------------------------------------- librosa.filters
def constant_q_lengths(sr, fmin, n_bins, bins_per_octave, filter_scale=1.0, gamma=0.0) -> np.ndarray:
    alpha = 2.0 ** (1.0 / bins_per_octave) - 1.0
    Q = float(filter_scale) / alpha
    freq = fmin * (2.0 ** (np.arange(n_bins, dtype=float) / bins_per_octave))
    if max(freq * (1 + 0.5 * HANN_BANDWIDTH / Q)) > sr / 2.0:
        raise ParameterError("Filter pass-band lies beyond Nyquist")
    return Q * sr / (freq + gamma / alpha)


def constant_q(sr, fmin, n_bins, bins_per_octave, filter_scale=1.0, norm=1) -> Tuple[np.ndarray, np.ndarray]:
    """librosa.filters.constant_q(..., window='hann', pad_fft=True): complex64 time-domain filters."""
    lengths = constant_q_lengths(sr, fmin, n_bins, bins_per_octave, filter_scale)
    freqs = fmin * (2.0 ** (np.arange(n_bins, dtype=float) / bins_per_octave))
    max_len = int(2.0 ** (np.ceil(np.log2(max(lengths)))))
    filters = np.zeros((n_bins, max_len), dtype=np.complex128)
    for k, (ilen, freq) in enumerate(zip(lengths, freqs)):
        sig = np.exp(np.arange(-ilen // 2, ilen // 2, dtype=float) * 1j * 2 * np.pi * freq / sr)
        n = len(sig)
        sig = sig * (0.5 - 0.5 * np.cos(2 * np.pi * np.arange(n) / n))  # scipy get_window('hann', n, fftbins=True)
        if norm == 1:
            sig = sig / np.sum(np.abs(sig))
        else:
            raise ParameterError("only norm=1 is restated")
        lpad = (max_len - n) // 2  # util.pad_center
        filters[k, lpad: lpad + n] = sig
    return filters.astype(np.complex64), np.asarray(lengths)


def sparsify_rows(x: np.ndarray, quantile: float) -> np.ndarray:
    """librosa.util.sparsify_rows, returned dense (zeros where librosa's CSR matrix has no entry)."""
    if not 0.0 <= quantile < 1:
        raise ParameterError("Invalid quantile")
    mags = np.abs(x)
    norms = np.sum(mags, axis=1, keepdims=True)
    mag_sort = np.sort(mags, axis=1)
    cumulative_mag = np.cumsum(mag_sort / norms, axis=1)
    threshold_idx = np.argmin(cumulative_mag < quantile, axis=1)
    out = np.zeros_like(x)
    for i, j in enumerate(threshold_idx):
        idx = mags[i] >= mag_sort[i, j]
        out[i, idx] = x[i, idx]
    return out


@lru_cache(maxsize=16)
def cqt_filter_fft(sr, fmin, n_bins, bins_per_octave, filter_scale, sparsity) -> Tuple[np.ndarray, int]:
    """librosa.core.constantq.__cqt_filter_fft (hop_length=None): complex64 (n_bins, n_fft/2+1)."""
    basis, lengths = constant_q(sr, fmin, n_bins, bins_per_octave, filter_scale)
    n_fft = basis.shape[1]
    basis = basis * (lengths[:, np.newaxis] / float(n_fft)).astype(np.float32)  # in-place complex64 multiply
    basis = basis.astype(np.complex64)
    fft_basis = np.fft.fft(basis.astype(np.complex128), n=n_fft, axis=1)[:, : (n_fft // 2) + 1].astype(np.complex64)
    return sparsify_rows(fft_basis, sparsity), n_fft


def cqt_response(y: np.ndarray, n_fft: int, hop_length: int, fft_basis: np.ndarray, cdtype) -> np.ndarray:
    """__cqt_response: stft(window='ones', center=True, pad_mode='constant') then basis.dot(D)."""
    yp = np.concatenate([np.zeros(n_fft // 2, y.dtype), y, np.zeros(n_fft // 2, y.dtype)])
    n_frames = 1 + (len(yp) - n_fft) // hop_length
    idx = np.arange(n_fft)[:, None] + hop_length * np.arange(n_frames)[None, :]
    D = np.fft.rfft(yp[idx], axis=0).astype(cdtype)
    return fft_basis.astype(cdtype).dot(D)


def _num_two_factors(x: int) -> int:
    if x <= 0:
        return 0
    n = 0
    while x % 2 == 0:
        n += 1
        x //= 2
    return n


def early_downsample_count(nyquist, filter_cutoff, hop_length, n_octaves) -> int:
    """librosa.core.constantq.__early_downsample_count."""
    downsample_count1 = max(0, int(np.ceil(np.log2(BW_FASTEST * nyquist / filter_cutoff)) - 1) - 1)
    downsample_count2 = max(0, _num_two_factors(hop_length) - n_octaves + 1)
    return min(downsample_count1, downsample_count2)


RECURSION_092 = "librosa-0.9.2"
RECURSION_HALVE_WHILE_EVEN = "halve-while-even"


def octave_plan(hop_length: int, n_octaves: int, recursion: str = RECURSION_092):
    """[(decimation count, hop at that rate)] per octave, top first.  0.9.2: octave i sits i halvings down.
    halve-while-even (the rule of later librosa releases' vqt loop, ``if my_hop % 2 == 0``): halve after an octave only
    while the hop is even; afterwards the rate stays and the filters grow."""
    plan, level, hop = [], 0, int(hop_length)
    for i in range(n_octaves):
        plan.append((level, hop))
        if i + 1 < n_octaves and (recursion == RECURSION_092 or hop % 2 == 0):
            level, hop = level + 1, hop // 2
    return plan


def cqt(y: np.ndarray, sr: float = 22050, hop_length: int = 512, fmin: Optional[float] = None, n_bins: int = 84,
        bins_per_octave: int = 12, filter_scale: float = 1.0, sparsity: float = 0.01,
        dtype=np.float64, recursion: str = RECURSION_092) -> np.ndarray:
    """librosa.cqt == vqt(gamma=0) with tuning=0, norm=1, hann, scale=True, pad_mode='constant', res_type=None.

    ``dtype`` is the real working precision: np.float32 mirrors librosa (which keeps the input's
    precision, complex64 output); np.float64 (default) is the higher-precision yardstick the CUDA
    path is compared with.

    ``recursion``: RECURSION_092 is librosa 0.9.2 (the release the reference pins).  RECURSION_HALVE_WHILE_EVEN
    restates the octave loop of later releases -- filter response first, then ``if my_hop % 2 == 0`` halve hop, rate and
    signal -- ON TOP OF the 0.9.2 filter design and the kaiser_fast resampler (those releases also changed the filter
    bandwidth and default to soxr_hq): a variant that lets the reference's own hop = round(rate / 5) run at 44.1 kHz and
    22.05 kHz (KeyDataset.py:485), not a statement of any particular librosa release."""
    y = np.asarray(y, dtype=dtype)
    cdtype = np.complex64 if dtype == np.float32 else np.complex128
    n_octaves = int(np.ceil(float(n_bins) / bins_per_octave))
    n_filters = min(bins_per_octave, n_bins)
    alpha = 2.0 ** (1.0 / bins_per_octave) - 1
    if fmin is None:
        fmin = C1_HZ
    freqs = (fmin * 2.0 ** (np.arange(n_bins, dtype=float) / bins_per_octave))[-bins_per_octave:]
    fmin_t, fmax_t = np.min(freqs), np.max(freqs)
    Q = float(filter_scale) / alpha
    filter_cutoff = fmax_t * (1 + 0.5 * HANN_BANDWIDTH / Q)
    nyquist = sr / 2.0
    if not filter_cutoff < BW_FASTEST * nyquist:
        raise NotImplementedError("top octave would use kaiser_best resampling; only the kaiser_fast recursion is restated")
    if early_downsample_count(nyquist, filter_cutoff, hop_length, n_octaves) > 0:
        raise NotImplementedError("early down-sampling branch of librosa is not restated")
    if recursion not in (RECURSION_092, RECURSION_HALVE_WHILE_EVEN):
        raise ParameterError("unknown recursion")
    if recursion == RECURSION_092 and _num_two_factors(hop_length) < n_octaves - 1:
        raise ParameterError("hop_length must be a positive integer multiple of 2^{0:d} for {1:d}-octave CQT/VQT"
                             .format(n_octaves - 1, n_octaves))
    resp = []
    my_y, level = y, 0
    for i, (lv, my_hop) in enumerate(octave_plan(hop_length, n_octaves, recursion)):
        if lv > level:  # the previous octave halved the rate
            if len(my_y) < 2:
                raise ParameterError("Input signal length={} is too short for {:d}-octave CQT/VQT".format(len(y), n_octaves))
            my_y = resample_half(my_y)
            level = lv
        my_sr = float(sr) / 2 ** lv
        fft_basis, n_fft = cqt_filter_fft(my_sr, float(fmin_t * 2.0 ** -i), n_filters, bins_per_octave, float(filter_scale),
                                          float(sparsity))
        fft_basis = fft_basis * np.float32(np.sqrt(2 ** lv))  # fft_basis *= sqrt(sr / my_sr)
        resp.append(cqt_response(my_y, n_fft, my_hop, fft_basis, cdtype))
    # __trim_stack
    max_col = min(c.shape[-1] for c in resp)
    out = np.empty((n_bins, max_col), dtype=cdtype)
    end = n_bins
    for c in resp:
        n_oct = c.shape[0]
        if end < n_oct:
            out[:end] = c[-end:, :max_col]
        else:
            out[end - n_oct: end] = c[:, :max_col]
        end -= n_oct
    lengths = constant_q_lengths(sr, fmin, n_bins, bins_per_octave, filter_scale)
    out /= np.sqrt(lengths[:, np.newaxis]).astype(dtype)
    return out


def cqt_logmag(y: np.ndarray, sr: float, frames: int = 5, octaves: int = 8, dtype=np.float64,
               recursion: str = RECURSION_092) -> np.ndarray:
    """DatasetLoader.get_all's feature (KeyDataset.py:485-509): (1, 36*octaves, T) float64."""
    hop = round(sr / frames)
    C = cqt(y, sr=sr, hop_length=hop, bins_per_octave=36, n_bins=36 * octaves, dtype=dtype, recursion=recursion)
    mel = np.log(1 + np.abs(C))
    return mel.reshape(1, mel.shape[0], mel.shape[1]).astype(np.float64)


def n_frames(n_samples: int, hop_length: int, n_octaves: int, recursion: str = RECURSION_092) -> int:
    """Frames librosa returns: the minimum over octaves of 1 + ceil(n / 2^level) // hop at that level."""
    return min(1 + int(math.ceil(n_samples / 2 ** lv)) // hop for lv, hop in octave_plan(hop_length, n_octaves, recursion))

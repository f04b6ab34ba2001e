"""oracle.cqt_port (PARITY UNPINNED: no librosa here, no CQT fixture in the reference): anchored on
analytic known-answer tests, structural properties and a committed regression fixture."""
import ctypes as C

import numpy as np
import pytest

from conftest import load_golden
from oracle import cqt_port as cp

SR, HOP = 48000, 9600


def test_decimator_taps_properties():
    h = cp.decimator_taps()
    assert h.shape == (63,)
    np.testing.assert_allclose(h, h[::-1], atol=0)           # linear phase
    assert abs(h[31] - 0.5 * 0.85) < 1e-15                   # centre tap = ratio * rolloff
    assert abs(h.sum() - 1.0) < 1e-4                         # unit DC gain
    # stop band: a tone above the new Nyquist is attenuated by > 60 dB, a low tone passes
    n = np.arange(8192)
    for f, lo, hi in ((1.0 / 16, 0.99, 1.01), (0.42, 0.0, 1e-3)):
        y = cp.resample_half(np.sin(2 * np.pi * f * n)) * np.sqrt(0.5)
        amp = np.abs(y[200:-200]).max()
        assert lo <= amp <= hi, (f, amp)


def test_resample_half_length_and_tail():
    for n in (1000, 1001):
        y = cp.resample_half(np.ones(n))
        assert y.shape[0] == (n + 1) // 2
        if n % 2:
            assert y[-1] == 0.0  # librosa fix_length pads resampy's int(n/2) samples with a zero


def test_frames_and_hop_rule():
    assert cp.n_frames(SR * 30, HOP, 8) == 151
    assert cp.n_frames(SR * 240, HOP, 8) == 1201
    with pytest.raises(cp.ParameterError):
        cp.cqt(np.zeros(44100), sr=44100, hop_length=8820, n_bins=288, bins_per_octave=36)  # 8820 = 2^2 * 2205


@pytest.mark.parametrize("bin_", [36, 100, 199, 280])
def test_pure_tone_known_answer(bin_):
    """A*sin at a bin centre: peak at that bin, |C| ~ A/2 * sqrt(filter length)."""
    n = SR * 4
    f0 = cp.C1_HZ * 2 ** (bin_ / 36)
    y = 0.3 * np.sin(2 * np.pi * f0 * np.arange(n) / SR)
    Cq = np.abs(cp.cqt(y, SR, HOP, None, 288, 36))
    col = Cq[:, 10]
    assert col.argmax() == bin_
    expect = 0.5 * 0.3 * np.sqrt(cp.constant_q_lengths(SR, cp.C1_HZ, 288, 36)[bin_])
    assert abs(col.max() / expect - 1) < 2e-3


def test_linearity_and_float32_mode():
    rng = np.random.default_rng(0)
    a, b = rng.standard_normal(SR), rng.standard_normal(SR)
    Ca, Cb, Cab = (cp.cqt(v, SR, HOP, None, 288, 36) for v in (a, b, 2 * a - 3 * b))
    np.testing.assert_allclose(Cab, 2 * Ca - 3 * Cb, atol=1e-9)
    C32 = cp.cqt(a.astype(np.float32), SR, HOP, None, 288, 36, dtype=np.float32)
    assert C32.dtype == np.complex64
    assert np.abs(C32 - Ca).max() < 2e-4 * np.abs(Ca).max()


def test_time_domain_bank_equals_fft_basis():
    """The dense real bank the CUDA kernels contract with == librosa's sparse FFT basis (host tables only)."""
    from audio_key_estimation_b200 import _lib
    L = _lib.lib()
    h = C.c_void_p()
    assert L.ake_cqt_create(float(SR), HOP, 288, 36, 0.0, 1.0, 0.01, C.byref(h)) == 0
    try:
        n_fft = L.ake_cqt_n_fft(h)
        assert n_fft == 1024
        bank = np.zeros((72, n_fft), np.float32)
        assert L.ake_cqt_get_bank(h, bank.ctypes.data, bank.size) == 0
        fb, n2 = cp.cqt_filter_fft(float(SR), float(cp.C1_HZ * 2 ** 7), 36, 36, 1.0, 0.01)
        assert n2 == n_fft
        f, n = np.arange(n_fft // 2 + 1), np.arange(n_fft)
        K = fb.astype(np.complex128) @ np.exp(-2j * np.pi * np.outer(f, n) / n_fft)
        assert np.abs(K.real - bank[0::2]).max() < 1e-6 and np.abs(K.imag - bank[1::2]).max() < 1e-6
        taps = (C.c_float * 32)()
        assert L.ake_cqt_get_decimator(h, taps, 32) == 32
        np.testing.assert_allclose(np.array(taps[:]), cp.decimator_taps()[31:], atol=1e-7)
        for n_samp in (0, 1, 9599, 9600, SR * 30, SR * 30 + 1, 12345677):
            assert L.ake_cqt_frames(h, n_samp) == cp.n_frames(n_samp, HOP, 8)
    finally:
        L.ake_cqt_destroy(h)


def test_regression_fixture():
    from audio_key_estimation_b200 import synth
    g = load_golden("cqt_port.npz")
    y = synth.synth_clip(int(g["clip_id"]), int(g["n_samples"]), SR).numpy()
    np.testing.assert_array_equal(y[:64], g["audio_head"])
    Cq = cp.cqt(y, SR, HOP, None, 288, 36)
    np.testing.assert_allclose(Cq.real, g["C_re"], atol=2e-5)
    np.testing.assert_allclose(Cq.imag, g["C_im"], atol=2e-5)
    mel = cp.cqt_logmag(y, SR)
    assert mel.shape == (1, 288, 16) and mel.dtype == np.float64
    np.testing.assert_allclose(mel[0], g["logmag"], atol=2e-5)


def test_halve_while_even_recursion_runs_the_reference_sample_rates():
    """KeyDataset.py:485 sets hop = round(rate / 5): 8820 at 44.1 kHz (GiantSteps) and 4410 at 22.05 kHz (GTZAN), which the
    pinned librosa 0.9.2 rejects (not multiples of 2^7).  The halve-while-even variant (see cqt_port.cqt) decimates while the
    hop is even and then lengthens the filters; anchors: octave plan, frame count, and the pure-tone known answer in octaves
    that are computed WITHOUT further decimation (their filters are 2..32 times longer)."""
    R = cp.RECURSION_HALVE_WHILE_EVEN
    assert cp.octave_plan(9600, 8, R) == cp.octave_plan(9600, 8) == [(i, 9600 >> i) for i in range(8)]
    assert cp.octave_plan(8820, 8, R) == [(0, 8820), (1, 4410)] + [(2, 2205)] * 6
    assert cp.octave_plan(4410, 8, R) == [(0, 4410)] + [(1, 2205)] * 7
    for sr, hop in ((44100, 8820), (22050, 4410)):
        n = sr * 5
        assert cp.n_frames(n, hop, 8, R) == 1 + n // hop
        t = np.arange(n) / sr
        for bin_ in (20, 130, 260):
            f0 = cp.C1_HZ * 2 ** (bin_ / 36)
            Cq = np.abs(cp.cqt(0.3 * np.sin(2 * np.pi * f0 * t), sr, hop, None, 288, 36, recursion=R))
            col = Cq[:, Cq.shape[1] // 2]
            assert col.argmax() == bin_
            expect = 0.5 * 0.3 * np.sqrt(cp.constant_q_lengths(sr, cp.C1_HZ, 288, 36)[bin_])
            assert abs(col.max() / expect - 1) < 3e-3
    # at a hop that IS a multiple of 2^7 the two recursions are the same computation
    y = np.random.default_rng(1).standard_normal(SR)
    np.testing.assert_array_equal(cp.cqt(y, SR, HOP, None, 288, 36, recursion=R), cp.cqt(y, SR, HOP, None, 288, 36))


def test_extended_plan_tables_and_long_filter_banks():
    """ake_cqt_create_ex(HALVE_WHILE_EVEN): frame counts follow the extended oracle; the default mode still raises."""
    from audio_key_estimation_b200 import _lib
    L = _lib.lib()
    for sr, hop in ((44100, 8820), (22050, 4410)):
        h = C.c_void_p()
        assert L.ake_cqt_create(float(sr), hop, 288, 36, 0.0, 1.0, 0.01, C.byref(h)) == _lib.AKE_ERR_INVALID
        assert L.ake_cqt_create_ex(float(sr), hop, 288, 36, 0.0, 1.0, 0.01, _lib.CQT_RECURSION_HALVE_WHILE_EVEN, C.byref(h)) == 0
        try:
            for n_samp in (0, 1, hop - 1, hop, sr * 30, sr * 30 + 1):
                assert L.ake_cqt_frames(h, n_samp) == cp.n_frames(n_samp, hop, 8, cp.RECURSION_HALVE_WHILE_EVEN)
            assert L.ake_cqt_set_peak(h, 0.0) == 0 and L.ake_cqt_set_peak(h, 32768.0) == 0
            assert L.ake_cqt_set_peak(h, -1.0) == _lib.AKE_ERR_INVALID
        finally:
            L.ake_cqt_destroy(h)
    h = C.c_void_p()
    assert L.ake_cqt_create_ex(48000.0, 9600, 288, 36, 0.0, 1.0, 0.01, 7, C.byref(h)) == _lib.AKE_ERR_INVALID


def test_port_matches_librosa_when_importable():
    """Pins the restatement the moment librosa is importable (it is not in the build container: requirements.txt:250 pins
    librosa 0.9.2 + resampy 0.3.1, neither vendored).  Same call as KeyDataset.py:490-491."""
    librosa = pytest.importorskip("librosa")
    from audio_key_estimation_b200 import synth
    y = synth.synth_clip(5, SR * 6 + 77, SR).numpy()
    kwargs = dict(sr=SR, hop_length=HOP, bins_per_octave=36, n_bins=288)
    major, minor = (int(x) for x in librosa.__version__.split(".")[:2])
    if (major, minor) >= (0, 10):
        kwargs["res_type"] = "kaiser_fast"   # the 0.9.2 default this port restates (0.10 switched to soxr_hq)
    want = librosa.cqt(y, **kwargs)
    got = cp.cqt(y, SR, HOP, None, 288, 36, dtype=np.float32)
    assert got.shape == want.shape
    assert np.abs(got - want).max() <= 2e-5 * np.abs(want).max()
    np.testing.assert_allclose(cp.cqt_logmag(y, SR)[0], np.log(1 + np.abs(want)), atol=2e-5)


def test_decimator_agrees_with_torchaudio_kaiser_resampler():
    """An EXTERNAL anchor for the decimation stage while librosa / resampy are absent: torchaudio's windowed-sinc resampler with the
    parameters its documentation gives as the equivalent of resampy's ``kaiser_fast`` (lowpass_filter_width 16, rolloff 0.85,
    beta 8.555504641634386) is an independent implementation of the same published filter family.  Its taper is stretched by
    1 / rolloff, so the two agree in the passband rather than tap by tap: on a band-limited signal the port's ``resample_half`` must
    equal torchaudio's output times sqrt(2) (librosa's ``scale=True`` energy normalisation, audio.py resample) -- that pins the
    port's passband gain, its zero-phase alignment / output sample positions and the sqrt(2) per level.  Measured 2.3e-5."""
    torchaudio = pytest.importorskip("torchaudio")
    import torch
    n = SR
    t = np.arange(n) / SR
    y = sum(a * np.sin(2 * np.pi * f * t + ph) for a, f, ph in [(0.5, 440.0, 0.1), (0.3, 1234.5, 1.0), (0.2, 3100.0, 2.0), (0.1, 55.0, 0.3)])
    got = cp.resample_half(y.astype(np.float64))
    want = torchaudio.functional.resample(torch.from_numpy(y)[None].double(), 2, 1, lowpass_filter_width=16, rolloff=0.85,
                                          resampling_method="sinc_interp_kaiser", beta=8.555504641634386)[0].numpy() * np.sqrt(2.0)
    assert got.shape == want.shape
    edge = 64  # the two implementations pad the clip's ends differently
    err = np.abs(got[edge:-edge] - want[edge:-edge]).max()
    assert err <= 1e-4 * np.abs(want).max(), err

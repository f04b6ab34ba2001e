"""GPU parity of the constant-Q front-end (through the C ABI) against oracle.cqt_port.
CQT parity is UNPINNED (librosa is not available; see oracle/cqt_port.py): the yardstick is this
repo's float64 restatement.  Tolerance: |C| error <= 2e-4 of max |C| (fp32 arithmetic through a
7-deep decimator cascade), log-magnitude max-abs <= 5e-4."""
import numpy as np
import pytest
import torch

import audio_key_estimation_b200 as ake
from audio_key_estimation_b200 import synth
from conftest import load_golden
from oracle import cqt_port as cp

pytestmark = pytest.mark.gpu
SR, HOP = 48000, 9600


def test_complex_cqt_matches_oracle():
    y = synth.synth_clip(3, SR * 5 + 123, SR)
    want = cp.cqt(y.numpy(), SR, HOP, None, 288, 36)
    got = ake.cqt(y.cuda(), sr=SR, hop_length=HOP, n_bins=288, bins_per_octave=36)
    assert got.dtype == torch.complex64 and tuple(got.shape) == want.shape
    err = np.abs(got.cpu().numpy() - want).max()
    assert err <= 2e-4 * np.abs(want).max(), err


def test_fixture_and_logmag():
    g = load_golden("cqt_port.npz")
    y = synth.synth_clip(int(g["clip_id"]), int(g["n_samples"]), SR)
    mel, seq = ake.cqt_logmag(y[None].cuda(), SR)
    assert tuple(mel.shape) == (1, 1, 288, 16) and seq.tolist() == [16]
    assert np.abs(mel[0, 0].cpu().numpy() - g["logmag"]).max() <= 5e-4


def test_ragged_batch_padding_and_seq_length():
    """KeyDataset.py:242-254: clips are zero-padded in time to the longest one; seq_length keeps the true frames."""
    lens = [SR * 4, SR * 3 + 4801, 9599, SR * 4 - 1]
    clips = [synth.synth_clip(20 + i, n, SR) for i, n in enumerate(lens)]
    mel, seq = ake.cqt_logmag([c.cuda() for c in clips], SR)
    T = [cp.n_frames(n, HOP, 8) for n in lens]
    assert seq.tolist() == T and mel.shape[-1] == max(T)
    for i, c in enumerate(clips):
        want = cp.cqt_logmag(c.numpy(), SR)[0]
        got = mel[i, 0].cpu().numpy()
        assert np.abs(got[:, : T[i]] - want).max() <= 5e-4
        assert not got[:, T[i]:].any()


def test_pure_tone_and_linearity_on_device():
    n = SR * 4
    t = torch.arange(n, dtype=torch.float64) / SR
    f0 = cp.C1_HZ * 2 ** (150 / 36)
    a = (0.3 * torch.sin(2 * np.pi * f0 * t)).float().cuda()
    Ca = ake.cqt(a, sr=SR, hop_length=HOP, n_bins=288, bins_per_octave=36)
    col = Ca[:, 10].abs()
    assert int(col.argmax()) == 150
    expect = 0.5 * 0.3 * np.sqrt(cp.constant_q_lengths(SR, cp.C1_HZ, 288, 36)[150])
    assert abs(float(col.max()) / expect - 1) < 2e-3
    b = synth.synth_clip(1, n, SR).cuda()
    Cb = ake.cqt(b, sr=SR, hop_length=HOP, n_bins=288, bins_per_octave=36)
    Cab = ake.cqt(0.5 * a - 0.25 * b, sr=SR, hop_length=HOP, n_bins=288, bins_per_octave=36)
    assert (Cab - (0.5 * Ca - 0.25 * Cb)).abs().max() <= 2e-4 * Cb.abs().max()


def test_full_size_properties():
    """BASELINE config sizes (30 s clips): frame count, batch independence, determinism."""
    n = SR * 30
    batch = synth.synth_batch(0, 4, n, SR).cuda()
    mel, seq = ake.cqt_logmag(batch, SR)
    assert tuple(mel.shape) == (4, 1, 288, 151) and seq.tolist() == [151] * 4
    mel2, _ = ake.cqt_logmag(batch, SR)
    assert torch.equal(mel, mel2)
    single, _ = ake.cqt_logmag(batch[2:3].clone(), SR)
    assert torch.equal(single[0], mel[2])
    want = cp.cqt_logmag(batch[1].cpu().numpy(), SR)[0]
    assert np.abs(mel[1, 0].cpu().numpy() - want).max() <= 5e-4


def test_long_clip_240s_matches_oracle():
    """BASELINE configs[3] clip shape: 240 s @ 48 kHz = 11,520,000 samples -> 1201 frames (many cascade tiles per clip,
    every octave's frames far from the clip edges), next to a shorter clip in the same batch."""
    n = SR * 240
    a = synth.synth_clip(77, n, SR)
    b = synth.synth_clip(78, SR * 100 + 4321, SR)
    mel, seq = ake.cqt_logmag([a.cuda(), b.cuda()], SR)
    assert tuple(mel.shape) == (2, 1, 288, 1201) and seq.tolist() == [1201, cp.n_frames(SR * 100 + 4321, HOP, 8)]
    for i, y in enumerate((a, b)):
        want = cp.cqt_logmag(y.numpy(), SR)[0]
        got = mel[i, 0, :, : want.shape[-1]].cpu().numpy()
        assert np.abs(got - want).max() <= 5e-4
        assert not mel[i, 0, :, want.shape[-1]:].any()


@pytest.mark.parametrize("sr,hop", [(44100, 8820), (22050, 4410)])
def test_reference_sample_rates_with_halve_while_even_recursion(sr, hop):
    """hop = round(rate / 5) (KeyDataset.py:485) at the sample rates of the reference's datasets (GiantSteps 44.1 kHz, GTZAN
    22.05 kHz): rejected by the pinned librosa 0.9.2 recursion, run by the halve-while-even variant against its oracle
    restatement (octaves below the last even hop keep their rate: filters of up to 32768 samples)."""
    with pytest.raises(ValueError):
        ake.cqt(torch.zeros(sr).cuda(), sr=sr, hop_length=hop, n_bins=288, bins_per_octave=36)
    lens = [sr * 7 + 13, sr * 4]
    clips = [synth.synth_clip(30 + i, n, sr) for i, n in enumerate(lens)]
    mel, seq = ake.cqt_logmag([c.cuda() for c in clips], sr, recursion="halve-while-even")
    T = [cp.n_frames(n, hop, 8, cp.RECURSION_HALVE_WHILE_EVEN) for n in lens]
    assert seq.tolist() == T and tuple(mel.shape) == (2, 1, 288, max(T))
    for i, c in enumerate(clips):
        want = cp.cqt_logmag(c.numpy(), sr, recursion=cp.RECURSION_HALVE_WHILE_EVEN)[0]
        got = mel[i, 0].cpu().numpy()
        assert np.abs(got[:, : T[i]] - want).max() <= 5e-4
        assert not got[:, T[i]:].any()
    Cc = ake.cqt(clips[0].cuda(), sr=sr, hop_length=hop, n_bins=288, bins_per_octave=36, recursion="halve-while-even")
    wantc = cp.cqt(clips[0].numpy(), sr, hop, None, 288, 36, recursion=cp.RECURSION_HALVE_WHILE_EVEN)
    assert np.abs(Cc.cpu().numpy() - wantc).max() <= 2e-4 * np.abs(wantc).max()


@pytest.mark.parametrize("gain", [1e-3, 1e2, 32768.0])
def test_any_amplitude_matches_oracle(gain):
    """librosa.cqt accepts any amplitude.  The fp16 hi/lo operands do not: each clip is pre-scaled by an exact power of two
    (measured per clip when the plan's peak is unknown, or given by the caller), so quiet clips keep their 22 bits and
    int16-scale clips do not overflow.  Tolerance relative to each clip's own max |C|, as everywhere."""
    y = synth.synth_clip(3, SR * 5 + 123, SR)
    batch = torch.stack([y * gain, y * (gain * 0.01)]).cuda()     # two clips of different loudness in one batch
    got = ake.cqt(batch, sr=SR, hop_length=HOP, n_bins=288, bins_per_octave=36)       # peak unknown: measured per clip
    for i, g in enumerate((gain, gain * 0.01)):
        want = cp.cqt((y * g).numpy().astype(np.float64), SR, HOP, None, 288, 36)
        err = np.abs(got[i].cpu().numpy() - want).max()
        assert np.isfinite(got[i].abs().max().item()) and err <= 2e-4 * np.abs(want).max(), (g, err)
    # the caller states the peak instead: same result without the measuring pass (exact powers of two either way)
    hinted = ake.cqt(batch[0], sr=SR, hop_length=HOP, n_bins=288, bins_per_octave=36, peak=float(gain))
    want = cp.cqt((y * gain).numpy().astype(np.float64), SR, HOP, None, 288, 36)
    assert np.abs(hinted.cpu().numpy() - want).max() <= 2e-4 * np.abs(want).max()
    # log-magnitude of the loud clip (what the network would see)
    mel, _ = ake.cqt_logmag(batch[:1], SR)
    assert np.abs(mel[0, 0].cpu().numpy() - np.log1p(np.abs(want))).max() <= 5e-4 * max(1.0, np.log1p(np.abs(want)).max())


def test_other_bank_shapes_and_errors():
    y = synth.synth_clip(9, 22050 * 3, 22050).cuda()
    got = ake.cqt(y, sr=22050, hop_length=512, n_bins=84, bins_per_octave=12)   # librosa's own defaults
    want = cp.cqt(y.cpu().numpy(), 22050, 512, None, 84, 12)
    assert np.abs(got.cpu().numpy() - want).max() <= 2e-4 * np.abs(want).max()
    with pytest.raises(ValueError):
        ake.cqt(y, sr=44100, hop_length=8820, n_bins=288, bins_per_octave=36)


def test_cache_writer_files_load_like_the_reference_expects(tmp_path):
    """torch.load(name).shape[1] == 288 and the tensor is the (1, 288, T) float64 log-CQT (KeyDataset.py:182-185, 509)."""
    import argparse

    import audio_key_estimation_b200 as ake
    from audio_key_estimation_b200 import cache, synth

    opt = argparse.Namespace(octaves=8, frames=5, only_semitones=False)
    sr = 48000
    lens = [sr * 3, sr * 2 + 1234]
    waves = [synth.synth_clip(40 + i, n, sr, "cpu") for i, n in enumerate(lens)]
    paths = [str(tmp_path / "a.wav"), str(tmp_path / "sub.dir.mp3")]
    with pytest.raises(RuntimeError):
        ake.write_cqt_cache(paths, waves, sr, opt)   # CQT parity vs librosa itself is unpinned: the caller must say so
    written = ake.write_cqt_cache(paths, waves, sr, opt, accept_unpinned_cqt=True)
    assert written == [cache.cache_name(p, opt) for p in paths]
    mel, seq = ake.cqt_logmag([w.cuda() for w in waves], sr=sr, frames=5, octaves=8)
    for i, name in enumerate(written):
        t = torch.load(name)
        assert t.dtype == torch.float64 and t.shape == (1, 288, int(seq[i])) and t.shape[1] == cache.expected_bins(opt)
        assert torch.equal(t, mel[i][:, :, : int(seq[i])].double().cpu())
    assert ake.write_cqt_cache(paths, waves, sr, opt, accept_unpinned_cqt=True) == []   # existing files are kept, as the reference does
    # opt.frames == 0 (KeyDataset.py:485-503): per-clip hop = w_length // window_size + 1, cropped to window_size frames
    opt0 = argparse.Namespace(octaves=8, frames=0, only_semitones=False, window_size=100)
    n0 = 127950                                          # hop = n0 // 100 + 1 = 1280 = 10 * 2^7: accepted by librosa 0.9.2
    w0 = synth.synth_clip(60, n0, sr, "cpu")
    p0 = str(tmp_path / "frames0.wav")
    name0 = ake.write_cqt_cache([p0], [w0], sr, opt0, overwrite=True, accept_unpinned_cqt=True)[0]
    t0 = torch.load(name0)
    want0 = np.log(1 + np.abs(cp.cqt(w0.numpy(), sr, 1280, None, 288, 36)))
    assert t0.shape == (1, 288, 100) and want0.shape[-1] == 100
    assert np.abs(t0[0].numpy() - want0[:, :100]).max() <= 5e-4
    with pytest.raises(ValueError):                      # hop 1281: librosa 0.9.2 raises ParameterError
        ake.write_cqt_cache([str(tmp_path / "bad.wav")], [synth.synth_clip(61, n0 + 100, sr, "cpu")], sr, opt0, accept_unpinned_cqt=True)

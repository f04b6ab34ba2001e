"""Host-side mirror of the reference interface (no GPU needed): constructor, state_dict contract,
error behaviour, synthetic generators, shard arithmetic."""
import numpy as np
import pytest
import torch

import audio_key_estimation_b200 as ake
from audio_key_estimation_b200 import distributed as akd, synth
from conftest import golden_state_dict


def test_state_dict_contract_strict_load():
    for genre in (False, True):
        net = ake.PitchClassNet(288, 12, 2, 7, opt=ake.default_opt(genre=genre), window_size=592, batch_size=8,
                                train_set=None, val_set=None)
        ref = golden_state_dict(genre)
        assert list(net.state_dict()) == list(ref)
        for k, v in net.state_dict().items():
            assert tuple(v.shape) == tuple(ref[k].shape) and v.dtype == ref[k].dtype, k
        net.load_state_dict(ref, strict=True)   # eval.py:113-115
        assert torch.equal(net.state_dict()["model.1.p2p.layer.3.weight"], ref["model.1.p2p.layer.3.weight"])
        assert sum(p.numel() for p in net.parameters()) == (167031 if genre else 162902)
        net.double()                            # train_model.py:105 calls .double() on the module
        assert net.state_dict()["key_classifier.0.conv2d.weight"].dtype == torch.float64
        assert len(list(net.modules())) > 10


def test_default_init_matches_torch_defaults():
    torch.manual_seed(0)
    net = ake.PitchClassNet(288, 12, 2, 7, opt=ake.default_opt())
    sd = net.state_dict()
    w = sd["model.1.p2p.layer.0.weight"]
    bound = 1 / np.sqrt(5 * 49)
    assert w.abs().max() <= bound and w.abs().max() > 0.9 * bound
    assert torch.all(sd["model.1.p2p.layer.1.weight"] == 1) and torch.all(sd["model.1.p2p.layer.1.bias"] == 0)
    assert torch.all(sd["model.1.p2p.layer.1.running_var"] == 1) and sd["model.1.p2p.layer.1.num_batches_tracked"] == 0


def test_constructor_errors_mirror_reference():
    with pytest.raises(AttributeError):
        ake.PitchClassNet(288, 12, 2, 7)  # opt=None: models.py:662 dereferences it
    # the one switch the reference itself cannot run, and combinations its channel plan does not cover
    for bad in (dict(only_semitones=True), dict(stay_sixth=True, pc2p_mem=True), dict(denseblock=True, resblock=True)):
        with pytest.raises(NotImplementedError):
            ake.PitchClassNet(288, 12, 2, 7, opt=ake.default_opt(**bad))
    with pytest.raises(ValueError):
        ake.PitchClassNet(288, 10, 2, 7, opt=ake.default_opt())


def test_no_cpu_fallback():
    net = ake.PitchClassNet(288, 12, 2, 7, opt=ake.default_opt())
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        net(torch.zeros(1, 1, 288, 64), None)
    with pytest.raises(ValueError):
        net(torch.zeros(1, 288, 64), None)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ake.cqt(torch.zeros(48000), sr=48000, hop_length=9600, n_bins=288, bins_per_octave=36)
    with pytest.raises(ValueError):
        ake.CQTPlan(44100, 8820, 288, 36)  # librosa 0.9.2 raises ParameterError for this hop


def test_bn_counts():
    net = ake.PitchClassNet(288, 12, 2, 7, opt=ake.default_opt(genre=True))
    counts = dict(zip(net._bn_sites, net._bn_counts(3, 61)))
    assert counts["model.0.pool_semi_b"] == 3 * 96 * 61
    assert counts["model.1.up_sixth_b"] == 3 * 36 * 61
    assert counts["model.1.p2p.layer.7"] == 3 * 288 * 61
    assert counts["model.1.pc2pc.layer.1"] == 3 * 12 * 61
    assert counts["key_classifier.1"] == 3 * 12 * (30 - 6)
    assert counts["genre_classifier.1"] == 3 * 12 * (30 - 6)


def test_synth_is_deterministic_and_tonal():
    a = synth.synth_clip(5, 48000, 48000)
    b = synth.synth_clip(5, 48000, 48000)
    assert torch.equal(a, b) and a.dtype == torch.float32 and a.abs().max() <= 1.0
    assert not torch.equal(a, synth.synth_clip(6, 48000, 48000))
    pat = synth.custom_cqt_pattern(360, 592)
    assert pat.shape == (360, 592) and pat[50, 330] == 20 and pat[120, 30] == 1  # equivariance_test.py:266-277


def test_shard_ranges():
    for n, w in ((256, 8), (10, 4), (3, 8), (0, 2), (16384, 8)):
        ranges = [akd.shard_range(n, r, w) for r in range(w)]
        assert ranges[0][0] == 0 and ranges[-1][1] == n
        assert all(ranges[i][1] == ranges[i + 1][0] for i in range(w - 1))
        assert max(b - a for a, b in ranges) - min(b - a for a, b in ranges) <= 1
        assert akd.shard_counts(n, w) == [b - a for a, b in ranges]
    with pytest.raises(ValueError):
        akd.shard_range(4, 2, 2)


def test_cqt_cache_names_follow_the_reference_table():
    """KeyDataset.py:154-180: file-name suffix and accepted bin count per (octaves, frames, only_semitones)."""
    import argparse

    import torch

    from audio_key_estimation_b200 import cache

    def opt(**kw):
        d = dict(octaves=8, frames=5, only_semitones=False)
        d.update(kw)
        return argparse.Namespace(**d)

    assert cache.cache_name("/d/song.wav", opt()) == "/d/song8oct.pt"
    assert cache.cache_name("/d/a.b/song.mp3", opt(octaves=7)) == "/d/a.b/song7oct.pt"
    assert cache.cache_name("s.wav", opt(octaves=5)) == "s5oct_5frames.pt"
    assert cache.cache_name("s.wav", opt(octaves=5, frames=10)) == "s5oct_10frames.pt"
    assert cache.cache_name("s.wav", opt(octaves=5, frames=20)) == "s5oct_20frames.pt"
    assert cache.cache_name("s.wav", opt(octaves=5, frames=3)) == "sfmin64.pt"
    assert cache.cache_name("s.wav", opt(only_semitones=True)) == "s8oct_no_semi.pt"
    with pytest.raises(ValueError):
        cache.cache_name("s.wav", opt(octaves=6))          # the reference leaves `name` unbound there
    assert [cache.expected_bins(opt(octaves=o)) for o in (5, 7, 8, 6)] == [180, 252, 288, 360]
    assert cache.expected_bins(opt(only_semitones=True)) == 96
    # entry format of KeyDataset.py:509: (1, n_bins, T) float64 on the CPU, batch padding removed
    mel = torch.arange(2 * 288 * 7, dtype=torch.float32).reshape(2, 1, 288, 7)
    e = cache.cache_entry(mel[1], 5)
    assert e.shape == (1, 288, 5) and e.dtype == torch.float64 and e.device.type == "cpu"
    assert torch.equal(e, mel[1][:, :, :5].double())


def test_numa_bind_is_best_effort():
    """distributed.bind_to_gpu_numa_node never raises and leaves the affinity alone when the topology is not exposed (no GPU here)."""
    import os
    from audio_key_estimation_b200 import distributed as akd
    before = os.sched_getaffinity(0)
    node = akd.bind_to_gpu_numa_node(0)
    assert node is None or isinstance(node, int)
    if node is None:
        assert os.sched_getaffinity(0) == before


def test_workload_mac_counts_match_the_survey_figures():
    """551,306,304 MAC per clip at T = 150 (554,584,320 with the genre head), probe-counted on the reference;
    bench.py's roofline bookkeeping (audio_key_estimation_b200.workload) derives them from the conv shapes alone."""
    from audio_key_estimation_b200 import workload
    from conftest import golden_state_dict
    for genre, want in ((False, 551306304), (True, 554584320)):
        shapes = {k: tuple(v.shape) for k, v in golden_state_dict(genre).items()}
        assert workload.pcn_macs(shapes, 288, 150) == want
    assert workload.p2p_macs(288, 150) == 288 * 150 * 8 * 21 * 49
    assert workload.cqt_algorithmic_bytes(1440000, 288, 151) == 5933952

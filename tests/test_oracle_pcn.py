"""oracle.pcn_port pinned against the reference: golden vectors (always) and the live reference
models.py (only where /root/reference exists, i.e. the build container)."""
import numpy as np
import pytest
import torch

from conftest import float_state_dict, golden_state_dict
from oracle import pcn_port, ref_import


@pytest.mark.parametrize("tag", ["default", "genre"])
def test_port_matches_golden_eval(fwd_golden, tag):
    g = fwd_golden
    sd = float_state_dict(golden_state_dict(tag == "genre", torch.float64))
    x = torch.from_numpy(g["mel"]).double()[:, None]
    seq = torch.from_numpy(g["seq_length"])
    for stag, sl in (("seq", seq), ("noseq", None)):
        res = pcn_port.pcn_forward(sd, x, sl)
        assert len(res) == (3 if tag == "genre" else 2)
        for name, r in zip(("key", "tonic", "genre"), res):
            np.testing.assert_allclose(r.numpy(), g[f"{tag}.eval.{stag}.{name}"], rtol=0, atol=1e-12)
    res = pcn_port.pcn_forward(sd, x, seq, max_pool=True)
    for name, r in zip(("key", "tonic", "genre"), res):
        np.testing.assert_allclose(r.numpy(), g[f"{tag}.eval.seq_maxpool.{name}"], rtol=0, atol=1e-12)
    taps = {}
    pcn_port.pcn_forward(sd, x, seq, taps=taps)
    np.testing.assert_allclose(taps["pc_final"].numpy(), g[f"{tag}.eval.pc_final"], rtol=0, atol=1e-5)


def test_port_matches_golden_train(fwd_golden):
    g = fwd_golden
    sd = float_state_dict(golden_state_dict(True, torch.float64))
    x = torch.from_numpy(g["mel"]).double()[:, None]
    stats = {}
    res = pcn_port.pcn_forward(sd, x, torch.from_numpy(g["seq_length"]), train=True, stats=stats)
    for name, r in zip(("key", "tonic", "genre"), res):
        np.testing.assert_allclose(r.numpy(), g[f"genre.train.seq.{name}"], rtol=0, atol=1e-11)
    # running buffers after one train-mode step: (1-m) * old + m * batch (unbiased variance), m = 0.1
    for site in ("model.1.p2p.layer.7", "key_classifier.1"):
        mean, var, n = stats[site]
        np.testing.assert_allclose((0.9 * sd[site + ".running_mean"] + 0.1 * mean).numpy(),
                                   g[f"genre.train.buf.{site}.running_mean"], rtol=0, atol=1e-12)
        np.testing.assert_allclose((0.9 * sd[site + ".running_var"] + 0.1 * var * n / (n - 1)).numpy(),
                                   g[f"genre.train.buf.{site}.running_var"], rtol=0, atol=1e-12)


def test_port_equivariance_matches_golden(eq_golden):
    from audio_key_estimation_b200 import synth
    sd = float_state_dict(golden_state_dict(False, torch.float64))
    pat = synth.custom_cqt_pattern(360, 592, with_border=False)
    shifts = eq_golden["shifts"]
    rows = []
    for s in shifts[:4]:
        m = torch.zeros_like(pat)
        if s >= 0:
            m[3 * s:] = pat[: 360 - 3 * s]
        else:
            m[: 360 + 3 * s] = pat[-3 * s:]
        k, _ = pcn_port.pcn_forward(sd, m.reshape(1, 1, 360, -1), torch.tensor([[592]]))
        rows.append(k[0].numpy())
    np.testing.assert_allclose(np.stack(rows), eq_golden["pattern.eval.key"][:4], rtol=0, atol=1e-12)


def test_golden_reference_is_exactly_equivariant(eq_golden):
    """The property config 3 asserts, as observed on the reference itself (eval: bit-exact)."""
    shifts = eq_golden["shifts"]
    for name in ("pattern", "padded_cqt"):
        for head in ("key", "tonic"):
            rows = eq_golden[f"{name}.eval.{head}"]
            for i, s in enumerate(shifts):
                assert np.array_equal(rows[i], np.roll(rows[0], s))
            rows = eq_golden[f"{name}.train.{head}"]
            for i, s in enumerate(shifts):
                np.testing.assert_allclose(rows[i], np.roll(rows[0], s), rtol=0, atol=1e-12)


def test_decode_and_mac_count():
    m = pcn_port.key_signature_map()
    assert m.shape == (21, 12) and set(m.sum(1).tolist()) == {7.0}
    key = m[[3, 9]].clone() * 0.8 + 0.1
    ids = pcn_port.decode(key, torch.eye(12)[[5, 2]], torch.eye(11)[[10, 0]])
    assert ids[0].tolist() == [3, 9] and ids[1].tolist() == [5, 2] and ids[2].tolist() == [10, 0]
    # 551,306,304 MAC per clip at T = 150 (554,584,320 with the genre head)
    assert pcn_port.count_macs(golden_state_dict(False), 288, 150) == 551306304
    assert pcn_port.count_macs(golden_state_dict(True), 288, 150) == 554584320


@pytest.mark.skipif(not ref_import.reference_available(), reason="reference tree not present (GPU box)")
@pytest.mark.parametrize("cfg", [dict(), dict(genre=True), dict(num_layers=1), dict(num_layers=3, n_filters=2),
                                 dict(conv_layers=2, head_layers=3)])
def test_port_matches_live_reference(cfg):
    """Random-init reference net (any of the supported architectures) vs the port on its state_dict,
    including intermediate activations taken with forward hooks."""
    from audio_key_estimation_b200 import synth
    torch.manual_seed(1)
    opt = ref_import.default_opt(**cfg)
    net = ref_import.build_reference_net(288, opt)
    sd = synth.randomise_state_dict(net.state_dict(), seed=3, dtype=torch.float64)
    net.load_state_dict(sd, strict=True)
    x = torch.rand(2, 1, 288, 72, dtype=torch.float64) * 3
    seq = torch.tensor([72, 61])
    hooks = {}
    layer = net.model[min(1, opt.num_layers - 1)]
    layer.pc2pc.register_forward_hook(lambda m, i, o: hooks.__setitem__("pc2pc", o))
    if opt.num_layers > 1:
        layer.p2p.register_forward_hook(lambda m, i, o: hooks.__setitem__("p2p", o))
    for train in (False, True):
        net.train(train)
        with torch.no_grad():
            ref = net(x, seq)
        taps = {}
        got = pcn_port.pcn_forward(float_state_dict(sd), x, seq, train=train, taps=taps)
        assert len(ref) == len(got)
        for r, o in zip(ref, got):
            np.testing.assert_allclose(o.numpy(), r.numpy(), rtol=0, atol=1e-10)
        L = min(1, opt.num_layers - 1)
        np.testing.assert_allclose(taps[f"l{L}.pc2pc{opt.conv_layers - 1}"].numpy(), hooks["pc2pc"].numpy(), atol=1e-10)
        if opt.num_layers > 1:
            np.testing.assert_allclose(taps[f"l1.p2p{opt.conv_layers - 1}"].numpy(), hooks["p2p"].numpy(), atol=1e-10)
        sd = {k: v.clone() for k, v in net.state_dict().items()}  # train step moved the buffers

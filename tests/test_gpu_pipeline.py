"""End-to-end parity: host audio -> ake_estimate_host_f32 -> predictions, against the oracle chain
(cqt_port -> pcn_port) on the same seeded clips and weights (BASELINE config 1 at B = 8 standard clips)."""
import numpy as np
import pytest
import torch

import audio_key_estimation_b200 as ake
from audio_key_estimation_b200 import synth
from conftest import float_state_dict, golden_state_dict
from oracle import cqt_port as cp, pcn_port

pytestmark = pytest.mark.gpu
SR = 48000


def _estimator(genre=True):
    net = ake.PitchClassNet(288, 12, 2, 7, opt=ake.default_opt(genre=genre))
    net.load_state_dict(golden_state_dict(genre))
    return ake.KeyEstimator(net.cuda().eval(), SR)


def test_config1_eight_standard_clips():
    est = _estimator(True)
    B, n = 8, SR * 30
    lens = [n] * 6 + [n - 7 * 9600 - 11, SR * 20]
    audio = synth.synth_batch(0, B, n, SR).pin_memory()
    for i, ln in enumerate(lens):
        audio[i, ln:] = 0
    out = est.estimate_host(audio, lens)
    sd = {k: v.double() for k, v in float_state_dict(golden_state_dict(True)).items()}
    mels, seq = [], []
    for i in range(B):
        m = cp.cqt_logmag(audio[i, : lens[i]].numpy(), SR)
        seq.append(m.shape[-1])
        mels.append(np.pad(m, ((0, 0), (0, 0), (0, 151 - m.shape[-1]))))
    want = pcn_port.pcn_forward(sd, torch.from_numpy(np.stack(mels)), torch.tensor(seq))
    tol = 1e-3   # north_star: logits max-abs <= 1e-3 relative to max |logit|
    worst = 0.0
    for name, w in zip(("key", "tonic", "genre"), want):
        err = (out[name].double() - w).abs().max().item()
        worst = max(worst, err / max(1.0, w.abs().max().item()))
    assert worst <= tol, worst
    ids = pcn_port.decode(*want)
    for j, w in enumerate(want):
        # argmax identical wherever the oracle's top-2 margin exceeds 10x the tolerance actually achieved
        scores = w if j else torch.nn.functional.cosine_similarity(w[:, None], pcn_port.key_signature_map(torch.float64)[None], dim=2)
        top2 = scores.topk(2, dim=1).values
        safe = (top2[:, 0] - top2[:, 1]) > 10 * max(worst, 1e-6)
        assert torch.equal(out["ids"][j].long()[safe], ids[j][safe])
    # device-resident path gives the same numbers as the host-buffer path
    dev = est.estimate_device(audio.cuda(), lens)
    for name, d in zip(("key", "tonic", "genre"), dev):
        assert torch.equal(d.cpu(), out[name])


def test_estimator_requires_eval_and_cuda():
    net = ake.PitchClassNet(288, 12, 2, 7, opt=ake.default_opt())
    with pytest.raises(RuntimeError):
        ake.KeyEstimator(net.cuda().train(), SR)
    with pytest.raises(RuntimeError):
        ake.KeyEstimator(net.cpu().eval(), SR)
    est = ake.KeyEstimator(net.cuda().eval(), SR)
    with pytest.raises(ValueError):
        est.estimate_host(torch.zeros(2, 48000).cuda())

"""MIREX key score on the device (ake_mirex_f32) against the reference golden vectors and the oracle port."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "mirex.npz")


def test_counters_categories_and_ratios_match_reference():
    import audio_key_estimation_b200 as ake
    from oracle import pcn_port

    g = np.load(GOLDEN)
    names = ("key_out", "tonic_out", "key_labels", "tonic_labels", "key_signature_id")
    cpu = [torch.from_numpy(g[k]) for k in names]
    dev = [x.cuda() for x in cpu]
    counters, sim, cat = ake.mirex_counters(*dev, return_details=True)
    want_cnt, want_sim = pcn_port.mirex_counters(*cpu)
    assert counters.tolist() == [want_cnt[k] for k in pcn_port.MIREX_COUNTERS]            # integer work: exact
    assert cat.cpu().numpy().tolist() == g["categories"].tolist()                          # category of every clip
    np.testing.assert_allclose(sim.cpu().numpy(), want_sim.numpy(), rtol=0, atol=2e-6)     # fp32 cosine similarity
    got = np.array([float(x) for x in ake.mirex_from_counters(counters)])
    np.testing.assert_allclose(got, g["ratios"], rtol=0, atol=1e-7)                        # the reference's 7 returned ratios


def test_24_wide_signature_ids_match_reference():
    """key_signature_id as the data layer delivers it: tf.one_hot(id, 24) (KeyDataset.py:366, 447), label ids up to 23."""
    import audio_key_estimation_b200 as ake

    g = np.load(GOLDEN)
    dev = [torch.from_numpy(g["w24_" + k]).cuda() for k in ("key_out", "tonic_out", "key_labels", "tonic_labels", "key_signature_id")]
    assert dev[4].shape == (96, 24)
    counters, _, cat = ake.mirex_counters(*dev, return_details=True)
    assert cat.cpu().numpy().tolist() == g["w24_categories"].tolist()
    got = np.array([float(x) for x in ake.mirex_from_counters(counters)])
    np.testing.assert_allclose(got, g["w24_ratios"], rtol=0, atol=1e-7)


def test_accumulation_over_batches_equals_one_batch():
    import audio_key_estimation_b200 as ake

    g = np.load(GOLDEN)
    dev = [torch.from_numpy(g[k]).cuda() for k in ("key_out", "tonic_out", "key_labels", "tonic_labels", "key_signature_id")]
    whole = ake.mirex_counters(*dev)
    acc = None
    for lo, hi in ((0, 1), (1, 33), (33, 96)):  # ragged batches, as the last batch of an epoch is
        acc = ake.mirex_counters(*[x[lo:hi] for x in dev], counters=acc)
    assert acc.tolist() == whole.tolist()
    with pytest.raises(RuntimeError):
        ake.mirex_counters(*[x.cpu() for x in dev])  # no CPU fallback

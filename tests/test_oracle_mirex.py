"""MIREX key score: the oracle port against the golden vectors of the unmodified reference
``mirex_score`` (models.py:1065-1116; tests/golden/mirex.npz from oracle/make_golden_mirex.py)."""
import os

import numpy as np
import torch

from oracle import pcn_port, ref_import

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "mirex.npz")


def _load(prefix=""):
    g = np.load(GOLDEN)
    return g, [torch.from_numpy(g[prefix + k]) for k in ("key_out", "tonic_out", "key_labels", "tonic_labels", "key_signature_id")]


def test_port_matches_reference_golden():
    g, t = _load()
    cnt, sim = pcn_port.mirex_counters(*t)
    got = np.array(pcn_port.mirex_from_counters(cnt))
    np.testing.assert_allclose(got, g["ratios"], rtol=0, atol=1e-7)  # the reference returns float32 ratios
    hist = np.bincount(g["categories"], minlength=5)
    assert [cnt[k] for k in ("correct", "fifths", "relative", "parallel", "other")] == hist.tolist()
    assert min(hist) > 0, "the golden batch must reach every category"
    assert cnt["samples"] == len(g["categories"]) and sim.shape == (len(g["categories"]),)


def test_port_matches_reference_golden_24_wide_ids():
    """The data layer's one-hot is 24 wide (KeyDataset.py:366, 447): label ids 21..23 lie beyond the 21-row table."""
    g, t = _load("w24_")
    assert t[4].shape[1] == 24 and int(t[4].argmax(1).max()) >= 21
    cnt, _ = pcn_port.mirex_counters(*t)
    np.testing.assert_allclose(np.array(pcn_port.mirex_from_counters(cnt)), g["w24_ratios"], rtol=0, atol=1e-7)
    hist = np.bincount(g["w24_categories"], minlength=5)
    assert [cnt[k] for k in ("correct", "fifths", "relative", "parallel", "other")] == hist.tolist()


def test_port_matches_live_reference_when_available():
    if not ref_import.reference_available():
        import pytest
        pytest.skip("reference checkout not present")
    ref = ref_import.load_reference_models()
    from oracle.make_golden_mirex import make_inputs
    arrs = [torch.from_numpy(a) for a in make_inputs(B=40, seed=5)]
    key_out, tonic_out, key_labels, tonic_labels, sig = arrs
    want = [float(x) for x in ref.PitchClassNet.mirex_score(None, key_labels, key_out, tonic_labels, tonic_out, sig)]
    cnt, _ = pcn_port.mirex_counters(*arrs)
    np.testing.assert_allclose(np.array(pcn_port.mirex_from_counters(cnt)), np.array(want), rtol=0, atol=1e-7)

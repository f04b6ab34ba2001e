"""Golden vectors for the MIREX key score: the UNMODIFIED reference ``mirex_score``
(models.py:1065-1116) run in the build container on seeded inputs that reach every category.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Usage:  python -m oracle.make_golden_mirex
Writes tests/golden/mirex.npz (inputs, the reference's 7 returned ratios, and the category of every clip obtained by
calling the reference on one clip at a time)."""
from __future__ import annotations

import os

import numpy as np
import torch

from oracle import pcn_port, ref_import

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def make_inputs(B: int = 96, seed: int = 11, width: int = 21):
    """width 21: one id per row of KEY_SIGNATURE_MAP; width 24: the one-hot the data layer really delivers
    (tf.one_hot(key_signature_id, 24), KeyDataset.py:366, 447) with label ids up to 23."""
    rng = np.random.Generator(np.random.PCG64(seed))
    table = pcn_port.key_signature_map(torch.float32).numpy()
    key_out = np.zeros((B, 12), np.float32)
    tonic_out = rng.normal(0, 1, (B, 12)).astype(np.float32)
    key_labels = np.zeros((B, 12), np.float32)
    tonic_labels = np.zeros((B, 12), np.float32)
    sig_id = np.zeros((B, width), np.float32)
    for i in range(B):
        s = int(rng.integers(0, 15)) if width == 21 else int(rng.integers(0, width))  # labelled signature
        key_labels[i] = table[min(s, 20)]
        sig_id[i, s] = 1.0
        mode = i % 6
        if width == 21:
            pred = s if mode in (0, 1) else (s + int(rng.choice([-1, 1]))) % 15 if mode in (2, 3) else int(rng.integers(0, 21))
        else:  # labels 21..23 have no table row: predictions next to them (20) exercise |pred - label| == 1 beyond the table
            pred = min(s, 20) if mode in (0, 1) else min(max(s + int(rng.choice([-1, 1])), 0), 20) if mode in (2, 3) else int(rng.integers(0, 21))
        # sigmoid-like outputs around the predicted signature, noisy enough that a few decode to a neighbour
        key_out[i] = np.clip(0.15 + 0.7 * table[pred] + rng.normal(0, 0.12, 12), 0.01, 0.99)
        t = int(rng.integers(0, 12))
        tonic_labels[i, t] = 1.0
        if mode % 2 == 0:
            tonic_out[i, t] = tonic_out[i].max() + 1.0     # correct tonic
    return key_out, tonic_out, key_labels, tonic_labels, sig_id


def main() -> None:
    ref = ref_import.load_reference_models()
    arrs = make_inputs()
    t = [torch.from_numpy(a) for a in arrs]
    key_out, tonic_out, key_labels, tonic_labels, sig_id = t
    res = ref.PitchClassNet.mirex_score(None, key_labels, key_out, tonic_labels, tonic_out, sig_id)
    ratios = np.array([float(r) for r in res], np.float64)
    # category of every clip: the reference on a batch of one returns a one-hot over (correct, fifths, relative, parallel, other)
    cats = np.zeros(len(key_out), np.int32)
    for i in range(len(key_out)):
        r = ref.PitchClassNet.mirex_score(None, key_labels[i:i + 1], key_out[i:i + 1], tonic_labels[i:i + 1], tonic_out[i:i + 1],
                                          sig_id[i:i + 1])
        cats[i] = int(np.argmax([float(x) for x in r[1:6]]))
    # second case: the 24-wide one-hot of the data layer, label ids 0..23
    arrs24 = make_inputs(B=96, seed=12, width=24)
    t24 = [torch.from_numpy(a) for a in arrs24]
    res24 = ref.PitchClassNet.mirex_score(None, t24[2], t24[0], t24[3], t24[1], t24[4])
    ratios24 = np.array([float(r) for r in res24], np.float64)
    cats24 = np.zeros(len(arrs24[0]), np.int32)
    for i in range(len(cats24)):
        r = ref.PitchClassNet.mirex_score(None, t24[2][i:i + 1], t24[0][i:i + 1], t24[3][i:i + 1], t24[1][i:i + 1], t24[4][i:i + 1])
        cats24[i] = int(np.argmax([float(x) for x in r[1:6]]))
    assert int(arrs24[4].argmax(1).max()) >= 21, "the 24-wide case must label ids beyond the 21-row table"
    np.savez_compressed(os.path.join(GOLDEN, "mirex.npz"), key_out=arrs[0], tonic_out=arrs[1], key_labels=arrs[2],
                        tonic_labels=arrs[3], key_signature_id=arrs[4], ratios=ratios, categories=cats,
                        w24_key_out=arrs24[0], w24_tonic_out=arrs24[1], w24_key_labels=arrs24[2], w24_tonic_labels=arrs24[3],
                        w24_key_signature_id=arrs24[4], w24_ratios=ratios24, w24_categories=cats24)
    print("ratios (mirex, correct, fifths, relative, parallel, other, accuracy):", ratios)
    print("category histogram:", np.bincount(cats, minlength=5))
    print("24-wide: ratios", ratios24, "histogram", np.bincount(cats24, minlength=5), "label ids >= 21:",
          int((arrs24[4].argmax(1) >= 21).sum()))


if __name__ == "__main__":
    main()

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


def make_inputs(B: int = 96, seed: int = 11):
    rng = np.random.Generator(np.random.PCG64(seed))
    table = pcn_port.key_signature_map(torch.float32).numpy()
    key_out = np.zeros((B, 12), np.float32)
    tonic_out = rng.normal(0, 1, (B, 12)).astype(np.float32)
    key_labels = np.zeros((B, 12), np.float32)
    tonic_labels = np.zeros((B, 12), np.float32)
    sig_id = np.zeros((B, 21), np.float32)
    for i in range(B):
        s = int(rng.integers(0, 15))                       # labelled signature (practical keys)
        key_labels[i] = table[s]
        sig_id[i, s] = 1.0
        mode = i % 6
        pred = s if mode in (0, 1) else (s + int(rng.choice([-1, 1]))) % 15 if mode in (2, 3) else int(rng.integers(0, 21))
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
    np.savez_compressed(os.path.join(GOLDEN, "mirex.npz"), key_out=arrs[0], tonic_out=arrs[1], key_labels=arrs[2],
                        tonic_labels=arrs[3], key_signature_id=arrs[4], ratios=ratios, categories=cats)
    print("ratios (mirex, correct, fifths, relative, parallel, other, accuracy):", ratios)
    print("category histogram:", np.bincount(cats, minlength=5))


if __name__ == "__main__":
    main()

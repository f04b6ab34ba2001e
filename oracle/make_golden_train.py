"""Generate tests/golden/train_step.npz by running the UNMODIFIED reference training step in the build container.

TEST INFRASTRUCTURE.  The reference's ``PitchClassNet.general_step`` (models.py:819-927: forward in train mode,
BCELoss + CrossEntropyLoss (+ genre_weight * CrossEntropyLoss on the labelled clips)) is called on a seeded
batch and ``loss.backward()`` gives the gradient of every parameter (float64, torch CPU autograd).  The GPU
training step (ake_pcn_forward_f32 bn_mode 2 + ake_loss_f32 + ake_pcn_backward_f32) is checked against these.
Re-run with:  python -m oracle.make_golden_train
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_import  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


class _Accuracy:
    """Stands in for torchmetrics.Accuracy (metrics only; the loss does not depend on it)."""

    def cuda(self):
        return self

    def __call__(self, pred, target):
        return (pred == target).float().mean()


def labels(B: int, seed: int = 5):
    rng = np.random.default_rng(seed)
    key = (rng.random((B, 12)) < 0.6).astype(np.float32)           # key_labels ~ Bernoulli(0.6)
    tonic_idx = rng.integers(0, 12, B).astype(np.int64)
    genre_idx = rng.integers(0, 11, B).astype(np.int64)
    genre_idx[1] = -1                                               # one clip without a genre label (all-zero one-hot row)
    return key, tonic_idx, genre_idx


def main() -> None:
    ref = ref_import.load_reference_models()
    ref.Accuracy = _Accuracy
    w = np.load(os.path.join(GOLDEN, "weights_seed0.npz"))
    g = np.load(os.path.join(GOLDEN, "pcn_fwd.npz"))
    mel = torch.from_numpy(g["mel"]).double()[:, None]
    seq = torch.from_numpy(g["seq_length"])
    B = mel.shape[0]
    key, tonic_idx, genre_idx = labels(B)
    out = {"key_labels": key, "tonic_idx": tonic_idx, "genre_idx": genre_idx}
    tonic_1h = torch.zeros(B, 12, dtype=torch.long)
    tonic_1h[torch.arange(B), torch.from_numpy(tonic_idx)] = 1
    genre_1h = torch.zeros(B, 11, dtype=torch.long)
    for b, gi in enumerate(genre_idx):
        if gi >= 0:
            genre_1h[b, gi] = 1
    for tag, genre in (("default", False), ("genre", True)):
        opt = ref_import.default_opt(genre=genre)
        net = ref_import.build_reference_net(288, opt)
        sd = {k: (torch.from_numpy(w[k]).double() if w[k].dtype.kind == "f" else torch.from_numpy(w[k]))
              for k in w.files if genre or not k.startswith("genre_classifier.")}
        net.load_state_dict(sd, strict=True)
        net.train()
        batch = dict(mel=mel, key_signature_id=torch.zeros(B, dtype=torch.long), key_labels=torch.from_numpy(key),
                     tonic_labels=tonic_1h, genre=genre_1h, seq_length=seq)
        res = net.general_step(batch, 0, "train")
        loss = res[0]
        loss.backward()
        out[f"{tag}.loss"] = np.float64(loss.item())
        for name, prm in net.named_parameters():
            out[f"{tag}.grad.{name}"] = prm.grad.numpy().astype(np.float32)
        print(tag, "loss", loss.item(), "sum|grad|", sum(float(p.grad.abs().sum()) for p in net.parameters()))
    np.savez_compressed(os.path.join(GOLDEN, "train_step.npz"), **out)
    print("train_step.npz", os.path.getsize(os.path.join(GOLDEN, "train_step.npz")))


if __name__ == "__main__":
    main()

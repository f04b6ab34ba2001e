"""Seeded synthetic workloads: tonal audio clips and random-init weights.

There is no dataset or checkpoint on the bench box, so the benchmark and the
parity tests run on synthetic mono clips of the reference's shape (float32 in
[-1, 1], first channel only -- KeyDataset.py:478-481) and on a seeded
``state_dict`` with the reference's key names.  All random
draws come from numpy's PCG64 (bit-stable across hosts); the per-sample noise
is a counter-based integer hash so CPU and GPU synthesis agree to the last bit
of the noise term.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, Optional

import numpy as np
import torch

C1_HZ = 32.70319566257483  # librosa.note_to_hz('C1'), the default fmin of librosa.cqt
MAJOR_SCALE = (0, 2, 4, 5, 7, 9, 11)


def _hash_noise(n: int, seed: int, device, amp: float) -> torch.Tensor:
    """Uniform noise in [-amp, amp) from a splitmix64-style hash of the sample index."""
    idx = torch.arange(n, dtype=torch.int64, device=device)
    z = idx * -7046029254386353131 + (seed * 0x632BE59BD9B4E019 % (1 << 63))  # 0x9E3779B97F4A7C15 as int64
    z = (z ^ ((z >> 30) & 0x3FFFFFFFF)) * -4658895280553007687          # 0xBF58476D1CE4E5B9
    z = (z ^ ((z >> 27) & 0x1FFFFFFFFF)) * -7723592293110705685         # 0x94D049BB133111EB
    z = z ^ ((z >> 31) & 0x1FFFFFFFF)
    u = ((z >> 40) & 0xFFFFFF).to(torch.float32) * (1.0 / 16777216.0)   # 24 random bits
    return (u * 2.0 - 1.0) * amp


def synth_clip(clip_id: int, n_samples: int, sr: int = 48000, device="cpu",
               seed_base: int = 1234) -> torch.Tensor:
    """One tonal clip: 12 notes from a random major key, 3 harmonics each, plus weak noise."""
    rng = np.random.Generator(np.random.PCG64(seed_base + clip_id))
    dur = n_samples / sr
    root = int(rng.integers(0, 12))
    y = _hash_noise(n_samples, seed_base + clip_id, device, 0.005 * math.sqrt(3.0))
    for _ in range(12):
        degree = MAJOR_SCALE[int(rng.integers(0, 7))]
        octave = int(rng.integers(2, 6))
        f0 = C1_HZ * 2.0 ** ((12 * (octave - 1) + root + degree) / 12.0)
        length = float(rng.uniform(1.0, 3.0))
        onset = float(rng.uniform(0.0, max(dur - 3.0, 0.0)))
        phases = rng.uniform(0.0, 2 * math.pi, size=3)
        a = int(onset * sr)
        b = min(n_samples, a + int(length * sr))
        if b <= a:
            continue
        t = torch.arange(b - a, dtype=torch.float64, device=device) / sr
        seg = torch.zeros(b - a, dtype=torch.float64, device=device)
        for h, amp in enumerate((1.0, 0.5, 0.25)):
            seg += 0.1 * amp * torch.sin(2 * math.pi * f0 * (h + 1) * t + float(phases[h]))
        y[a:b] += seg.to(torch.float32)
    return y.clamp_(-1.0, 1.0)


def synth_batch(first_clip: int, n_clips: int, n_samples: int, sr: int = 48000, device="cpu",
                out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """(n_clips, n_samples) float32; clip ids first_clip .. first_clip + n_clips - 1."""
    if out is None:
        out = torch.empty((n_clips, n_samples), dtype=torch.float32, device=device)
    for i in range(n_clips):
        out[i].copy_(synth_clip(first_clip + i, n_samples, sr, device))
    return out


def custom_cqt_pattern(pitches: int = 360, frames: int = 592, with_border: bool = True) -> torch.Tensor:
    """The reference's only data-free input generator (equivariance_test.py:266-277)."""
    mel = torch.zeros(pitches, frames, dtype=torch.float64)
    mel[100:150, 20:50] = 1.0
    if with_border:
        mel[30:40, 400] = 10.0
        mel[10:15, 200] = 8.0
    mel[50, 320:350] = 20.0
    return mel


def randomise_state_dict(template: Dict[str, torch.Tensor], seed: int = 0,
                         dtype=torch.float32) -> "OrderedDict[str, torch.Tensor]":
    """Seeded values for every entry of a reference-format ``state_dict``.

    Conv weights/biases ~ U(-1/sqrt(fan_in), 1/sqrt(fan_in)) (the scale of torch's default
    init, train_model.py:14-17 ``weights_init`` is never applied); BatchNorm affine and
    running statistics are randomised (weight U(0.5,1.5), bias N(0,0.1), running_mean
    N(0,0.1), running_var U(0.01,0.51)) because default buffers collapse the logits
    (SURVEY.md section 8c)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    out: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    fan_in = 1.0
    for name, t in template.items():
        shape = tuple(t.shape)
        if name.endswith("num_batches_tracked"):
            out[name] = torch.zeros(shape, dtype=torch.int64)
            continue
        if name.endswith("running_mean"):
            v = rng.normal(0.0, 0.1, size=shape)
        elif name.endswith("running_var"):
            v = rng.uniform(0.01, 0.51, size=shape)
        elif len(shape) == 4:
            if "up_sixth" in name:   # ConvTranspose2d weight is (Cin, Cout, kh, kw)
                fan_in = shape[1] * shape[2] * shape[3]
            else:
                fan_in = shape[1] * shape[2] * shape[3]
            v = rng.uniform(-1.0, 1.0, size=shape) / math.sqrt(fan_in)
        else:
            prev_is_conv = name.endswith(".bias") and (name[:-5] + ".weight") in template and \
                template[name[:-5] + ".weight"].dim() == 4
            if name.endswith(".weight"):       # BN gamma
                v = rng.uniform(0.5, 1.5, size=shape)
            elif prev_is_conv:                 # conv bias
                v = rng.uniform(-1.0, 1.0, size=shape) / math.sqrt(fan_in)
            else:                              # BN beta
                v = rng.normal(0.0, 0.1, size=shape)
        out[name] = torch.tensor(np.asarray(v), dtype=dtype).reshape(shape)
    return out

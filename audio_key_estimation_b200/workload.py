"""Algorithmic work of the hot path (roofline bookkeeping for bench.py).

Counts come from the architecture alone -- the convolution shapes of the ``PitchClassNet`` plan
(models.py:266-350, 694-742 in the reference) -- with no padding inflation: ``out_elems * Cin * kh * kw``
multiply-accumulates per convolution.  Nothing here touches ``oracle/``.
"""
from __future__ import annotations

from typing import Dict, Sequence


def cqt_algorithmic_bytes(n_samples: int, n_bins: int, frames: int) -> int:
    """Bytes the CQT stage must move per clip: read the fp32 audio once, write the fp32 log-CQT once."""
    return 4 * int(n_samples) + 4 * int(n_bins) * int(frames)


def pcn_macs(shapes: Dict[str, Sequence[int]], pitches: int, T: int, time_pool_size: int = 2) -> int:
    """MACs of one clip's forward.  ``shapes``: state_dict name -> shape (only the 4-D conv weights are read)."""
    conv = {k: tuple(int(x) for x in v) for k, v in shapes.items() if len(v) == 4}
    layers = sorted({int(k.split(".")[1]) for k in conv if k.startswith("model.")})
    total, t = 0, int(T)
    for L in layers:
        for k, (a, b, kh, kw) in conv.items():
            if not k.startswith(f"model.{L}."):
                continue
            if ".pool_semi." in k:          # stride-3 semitone conv: pitches / 3 output rows (models.py:313, 337)
                total += (pitches // 3) * t * a * b * kh * kw
            elif ".up_sixth." in k:         # ConvTranspose2d (3,1)/(3,1): every one of the 36 output rows takes ONE tap
                total += 36 * t * a * b
            elif ".p2p." in k:              # 7x7 circular convs on all pitches (models.py:228-234)
                total += pitches * t * a * b * kh * kw
            elif ".pc2pc." in k:            # equivariant convs on the 12 pitch classes (models.py:22-51)
                total += 12 * t * a * b * kh * kw
        if L > 0:
            t //= time_pool_size
    for head in ("tonic_classifier", "key_classifier", "genre_classifier"):
        tt = t
        for k in sorted((k for k in conv if k.startswith(head + ".")), key=lambda s: int(s.split(".")[1])):
            co, ci, kh, kw = conv[k]
            tt -= kw - 1                    # valid in time
            rows = 12 if ".conv2d." in k else 12 - kh + 1
            total += rows * tt * co * ci * kh * kw
    return total


def p2p_macs(pitches: int, T: int, out_p: int = 8, cin_first: int = 5, convs: int = 3, k: int = 7) -> int:
    """MACs of the layer-1 Pitch2Pitch stack (the dominant section: 64.6 % of the forward at the defaults)."""
    return pitches * T * out_p * (cin_first + (convs - 1) * out_p) * k * k

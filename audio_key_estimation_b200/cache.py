"""CQT cache in the reference's on-disk format (KeyDataset.py:150-192, 497-509).

The reference computes ``librosa.cqt`` once per audio file and ``torch.save``s the (1, n_bins, T) float64 log-magnitude
next to the audio as ``<path minus .wav/.mp3><suffix>.pt``; ``KeyDataset.__getitem__`` then loads it whenever its
``shape[1]`` matches the expected bin count.  ``write_cqt_cache`` fills those files from the GPU front-end in batches, so the
UNMODIFIED ``KeyDataset`` picks up B200-computed spectrograms."""
from __future__ import annotations

import os
from typing import List, Sequence

import torch

from .cqt import CQTPlan, cqt_logmag


def _get(opt, name, default):
    return getattr(opt, name, default) if opt is not None else default


def cache_suffix(opt) -> str:
    """The name table of KeyDataset.py:154-167 (later branches override earlier ones exactly as the if-chain does)."""
    octaves, frames, semi = int(_get(opt, "octaves", 8)), int(_get(opt, "frames", 5)), bool(_get(opt, "only_semitones", False))
    suffix = None
    if octaves == 5 and frames == 5 and not semi:
        suffix = "5oct_5frames.pt"
    elif octaves == 5 and frames == 10 and not semi:
        suffix = "5oct_10frames.pt"
    elif octaves == 5 and frames == 20 and not semi:
        suffix = "5oct_20frames.pt"
    elif octaves == 5 and not semi:
        suffix = "fmin64.pt"
    elif octaves == 7 and not semi:
        suffix = "7oct.pt"
    elif octaves == 8 and not semi:
        suffix = "8oct.pt"
    if semi and octaves == 8:
        suffix = "8oct_no_semi.pt"
    if suffix is None:
        # the reference leaves `name` unbound here (UnboundLocalError at KeyDataset.py:182)
        raise ValueError(f"the reference has no cache file name for octaves={octaves}, only_semitones={semi}")
    return suffix


def cache_name(audio_path: str, opt) -> str:
    """KeyDataset.py:154-167: ``filename.replace('.wav','').replace('.mp3','') + suffix``."""
    return audio_path.replace(".wav", "").replace(".mp3", "") + cache_suffix(opt)


def expected_bins(opt) -> int:
    """``shape`` of KeyDataset.py:169-180: a cached tensor is accepted when its dim 1 equals this."""
    octaves, semi = int(_get(opt, "octaves", 8)), bool(_get(opt, "only_semitones", False))
    if octaves == 5 and not semi:
        return 180
    if octaves == 7 and not semi:
        return 252
    if octaves == 8 and not semi:
        return 288
    if octaves == 8 and semi:
        return 96
    return 360


def cache_entry(mel: torch.Tensor, n_frames: int) -> torch.Tensor:
    """One clip of a ``cqt_logmag`` batch -> what ``load_data_from_filename`` returns and the reference saves
    (KeyDataset.py:509): (1, n_bins, T) float64 on the CPU, without the batch's time padding."""
    if mel.dim() != 3 or mel.shape[0] != 1:
        raise ValueError("expected one clip of shape (1, n_bins, T_max)")
    return mel[:, :, :int(n_frames)].detach().to(dtype=torch.float64, device="cpu").contiguous()


def write_cqt_cache(audio_paths: Sequence[str], waveforms: Sequence[torch.Tensor], sr: int, opt, batch: int = 64,
                    overwrite: bool = False, accept_unpinned_cqt: bool = False, recursion: str = "librosa-0.9.2") -> List[str]:
    """Compute the log-CQT of ``waveforms`` (mono float32 tensors, CUDA or pinned/pageable host) on the GPU in batches and
    save each as the reference's cache file.  Returns the written file names.  Raises where the B200 front-end does not
    cover the reference's configuration (``ake_cqt_create`` says so).

    ``accept_unpinned_cqt`` must be set: the unmodified ``KeyDataset`` PREFERS these files over calling ``librosa.cqt``
    (KeyDataset.py:169-185), and the B200 front-end's parity is checked against this repository's restatement of
    librosa 0.9.2 (oracle/cqt_port.py), not against librosa itself, which is absent from the build environment
    (tests/test_oracle_cqt.py::test_port_matches_librosa_when_importable pins it where librosa
    exists).  Writing features the reference's training and evaluation will silently consume is therefore an explicit
    decision of the caller.

    ``recursion="halve-while-even"`` is needed for 44.1 kHz / 22.05 kHz material (hop = round(rate / frames) is then not a
    multiple of 2^(octaves-1), which librosa 0.9.2 rejects); see ``CQTPlan``.  Clips of any amplitude are exact (their peak is
    measured on the device)."""
    if len(audio_paths) != len(waveforms):
        raise ValueError("one path per waveform")
    if not accept_unpinned_cqt:
        raise RuntimeError("write_cqt_cache writes files the reference's KeyDataset loads instead of calling librosa.cqt, and CQT "
                           "parity against librosa itself is unpinned in this build (see the docstring); pass accept_unpinned_cqt=True "
                           "after running tests/test_oracle_cqt.py::test_port_matches_librosa_when_importable where librosa is installed")
    octaves, frames = int(_get(opt, "octaves", 8)), int(_get(opt, "frames", 5))
    bpo = 12 if bool(_get(opt, "only_semitones", False)) else 36
    names = [cache_name(p, opt) for p in audio_paths]
    written = []
    todo = [i for i, n in enumerate(names) if overwrite or not os.path.exists(n)]
    if frames <= 0:
        # opt.frames == 0 (KeyDataset.py:485-503): the hop is per clip, w_length // window_size + 1, and the result is cropped to
        # window_size frames -- one front-end plan per distinct hop (ValueError where librosa 0.9.2 raises ParameterError:
        # the hop must be a multiple of 2^(octaves-1))
        window = int(_get(opt, "window_size", 592))
        for i in todo:
            clip = waveforms[i].reshape(-1).to(device="cuda", dtype=torch.float32)
            plan = CQTPlan.get(sr, int(clip.numel()) // window + 1, bpo * octaves, bpo, recursion=recursion)
            mel, seq = plan.run(clip[None])
            entry = cache_entry(mel[0], min(int(seq[0]), window))
            if entry.shape[1] != expected_bins(opt):
                raise RuntimeError(f"{names[i]}: {entry.shape[1]} bins, the reference expects {expected_bins(opt)}")
            torch.save(entry, names[i])
            written.append(names[i])
        return written
    for lo in range(0, len(todo), batch):
        idx = todo[lo: lo + batch]
        clips = [waveforms[i].reshape(-1).to(device="cuda", dtype=torch.float32, non_blocking=True) for i in idx]
        mel, seq = cqt_logmag(clips, sr=sr, frames=frames, octaves=octaves, bins_per_octave=bpo, recursion=recursion)
        seq = seq.tolist()
        for j, i in enumerate(idx):
            entry = cache_entry(mel[j], seq[j])
            if entry.shape[1] != expected_bins(opt):
                raise RuntimeError(f"{names[i]}: {entry.shape[1]} bins, the reference expects {expected_bins(opt)}")
            torch.save(entry, names[i])
            written.append(names[i])
    return written

"""Constant-Q front-end on the GPU: the ``librosa.cqt`` call + abs/log1p of KeyDataset.py:485-509.

``cqt(y, sr, hop_length, fmin, n_bins, bins_per_octave, ...)`` mirrors the argument names and
meaning of the librosa call the reference makes (KeyDataset.py:490-491, equivariance_test.py:161)
and returns the complex (n_bins, T) array; ``cqt_logmag`` returns the network input the data
layer builds from it -- ``log(1 + |C|)`` as (B, 1, n_bins, T_max) zero-padded in time plus the
per-clip ``seq_length`` (KeyDataset.py:242-254, 497-509).  Both call libake_b200.so
(``ake_cqt_run_f32``); CPU tensors raise -- there is no fallback.
"""
from __future__ import annotations

import ctypes as C
import threading
from typing import Optional, Sequence, Tuple, Union

import torch

from . import _lib
from ._lib import check
from .models import _Workspace


class CQTPlan:
    """Decimator taps + dense filter bank for one (sr, hop, n_bins, bins_per_octave, fmin) setting."""

    _cache: dict = {}

    def __init__(self, sr: float, hop_length: int, n_bins: int = 288, bins_per_octave: int = 36,
                 fmin: Optional[float] = None, filter_scale: float = 1.0, sparsity: float = 0.01,
                 recursion: str = "librosa-0.9.2", peak: Optional[float] = None):
        """``recursion``: "librosa-0.9.2" (the release the reference pins: the hop must be a multiple of 2^(octaves-1)) or
        "halve-while-even" (the octave loop of later librosa releases on the 0.9.2 filter design; include/ake_b200.h) -- the
        latter lets hop = round(rate / 5) (KeyDataset.py:485) run at 44.1 kHz and 22.05 kHz.
        ``peak``: the largest |sample| the caller guarantees (1.0 for torchaudio-normalised audio, KeyDataset.py:478-481);
        None = unknown, measured per clip by one extra pass (any amplitude is then exact, as with librosa)."""
        lib = _lib.lib()
        h = C.c_void_p()
        modes = {"librosa-0.9.2": _lib.CQT_RECURSION_092, "halve-while-even": _lib.CQT_RECURSION_HALVE_WHILE_EVEN}
        if recursion not in modes:
            raise ValueError(f"recursion must be one of {sorted(modes)}")
        # ValueError where librosa raises ParameterError, NotImplementedError for un-built branches
        check(lib.ake_cqt_create_ex(float(sr), int(hop_length), int(n_bins), int(bins_per_octave),
                                    float(fmin) if fmin else 0.0, float(filter_scale), float(sparsity), modes[recursion], C.byref(h)))
        self._h = h
        check(lib.ake_cqt_set_peak(h, 0.0 if peak is None else float(peak)))
        self.sr, self.hop_length, self.n_bins, self.bins_per_octave = float(sr), int(hop_length), int(n_bins), int(bins_per_octave)
        self.recursion, self.peak = recursion, peak
        self.n_fft = lib.ake_cqt_n_fft(h)

    @classmethod
    def get(cls, sr, hop_length, n_bins=288, bins_per_octave=36, fmin=None, filter_scale=1.0, sparsity=0.01,
            recursion="librosa-0.9.2", peak=None) -> "CQTPlan":
        # one plan per host thread: a plan carries per-call staging state and may only be driven by one thread at a time
        key = (float(sr), int(hop_length), int(n_bins), int(bins_per_octave), float(fmin or 0.0), float(filter_scale),
               float(sparsity), recursion, peak, torch.cuda.current_device() if torch.cuda.is_available() else -1,
               threading.get_ident())
        plan = cls._cache.get(key)
        if plan is None:
            plan = cls._cache[key] = cls(sr, hop_length, n_bins, bins_per_octave, fmin, filter_scale, sparsity, recursion, peak)
        return plan

    def frames(self, n_samples: int) -> int:
        return int(_lib.lib().ake_cqt_frames(self._h, int(n_samples)))

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                _lib.lib().ake_cqt_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def run(self, audio: torch.Tensor, lengths: Optional[Sequence[int]] = None, mode: int = _lib.CQT_LOGMAG,
            T_max: Optional[int] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """audio (B, n_max) fp32 CUDA (row stride free, unit sample stride); lengths: valid samples per clip."""
        if not audio.is_cuda:
            raise RuntimeError("the B200 CQT runs on CUDA tensors only; there is no CPU fallback")
        if audio.dim() != 2:
            raise ValueError("audio must be (B, n_samples)")
        if audio.dtype != torch.float32 or audio.stride(1) != 1:
            audio = audio.to(torch.float32).contiguous()
        lib = _lib.lib()
        B, n_max = int(audio.shape[0]), int(audio.shape[1])
        dev = audio.device
        len_arr = None
        longest = n_max
        if lengths is not None:
            lens = [int(v) for v in lengths]
            if len(lens) != B:
                raise ValueError("lengths must have one entry per clip")
            len_arr = (C.c_int64 * B)(*lens)
            longest = max(lens)
        if T_max is None:
            T_max = self.frames(longest)
        if T_max <= 0:
            raise ValueError("clip too short")
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            shape = (B, 1, self.n_bins, T_max) if mode == _lib.CQT_LOGMAG else (B, self.n_bins, T_max, 2)
            out = torch.empty(shape, dtype=torch.float32, device=dev)
            seq = torch.empty(B, dtype=torch.int32, device=dev)
            ws_bytes = lib.ake_cqt_workspace_bytes(self._h, B, n_max)
            ws = _Workspace.get(dev, ws_bytes, "cqt")
            check(lib.ake_cqt_run_f32(self._h, audio.data_ptr(), int(audio.stride(0)), len_arr, B, n_max, int(mode),
                                      out.data_ptr(), int(T_max), seq.data_ptr(), ws.data_ptr(), ws.numel(), stream))
        return out, seq


def cqt(y: torch.Tensor, sr: float = 22050, hop_length: int = 512, fmin: Optional[float] = None, n_bins: int = 84,
        bins_per_octave: int = 12, filter_scale: float = 1.0, sparsity: float = 0.01, recursion: str = "librosa-0.9.2",
        peak: Optional[float] = None) -> torch.Tensor:
    """``librosa.cqt`` for a mono CUDA signal: complex64 (n_bins, T) (batched input (B, N) -> (B, n_bins, T)).
    Any amplitude is accepted, as librosa does (``peak=None``: each clip's peak is measured); see ``CQTPlan``."""
    single = y.dim() == 1
    plan = CQTPlan.get(sr, hop_length, n_bins, bins_per_octave, fmin, filter_scale, sparsity, recursion, peak)
    out, _ = plan.run(y[None] if single else y, mode=_lib.CQT_COMPLEX)
    c = torch.view_as_complex(out)
    return c[0] if single else c


def cqt_logmag(audio: Union[torch.Tensor, Sequence[torch.Tensor]], sr: float, frames: int = 5, octaves: int = 8,
               lengths: Optional[Sequence[int]] = None, bins_per_octave: int = 36, recursion: str = "librosa-0.9.2",
               peak: Optional[float] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """The network input of KeyDataset.py:485-509 for a batch of clips, computed on the GPU.

    Returns ``mel`` (B, 1, 36*octaves, T_max) fp32, zero-padded beyond each clip's frames
    (KeyDataset.py:242-254), and ``seq_length`` (B,) int32.  ``frames``/``octaves`` are the
    reference's ``opt.frames`` / ``opt.octaves`` (train_model.py:166-237)."""
    if not isinstance(audio, torch.Tensor):
        lengths = [int(a.numel()) for a in audio]
        n_max = max(lengths)
        batch = torch.zeros((len(audio), n_max), dtype=torch.float32, device=audio[0].device)
        for i, a in enumerate(audio):
            batch[i, : a.numel()] = a.reshape(-1)
        audio = batch
    hop = round(sr / frames)
    plan = CQTPlan.get(sr, hop, bins_per_octave * octaves, bins_per_octave, recursion=recursion, peak=peak)  # 12 per octave: opt.only_semitones (KeyDataset.py:493)
    return plan.run(audio, lengths=lengths, mode=_lib.CQT_LOGMAG)

"""Host audio in -> key / tonic / genre predictions out, in one C-ABI call per batch.

``KeyEstimator`` chains the two hot-path stages the way the reference's eval loop does
(eval.py:118-129: KeyDataset.get_all -> PitchClassNet.forward -> argmax decode,
models.py:1083-1085, 1096, 923) through ``ake_estimate_host_f32``: H2D copy of the clips, CQT,
forward (eval-mode BatchNorm), decode, D2H copy of the results, all on one stream.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import check
from .cqt import CQTPlan
from .models import PitchClassNet, _Workspace


class KeyEstimator:
    def __init__(self, net: PitchClassNet, sr: float, frames: int = 5, device: Optional[torch.device] = None,
                 recursion: str = "librosa-0.9.2", peak: Optional[float] = 1.0):
        """``peak``: largest |sample| of the audio this estimator will see -- 1.0 for what ``torchaudio.load`` delivers
        (KeyDataset.py:478-481) and for 16-bit PCM input; None measures every clip (one extra pass over the audio).
        ``recursion``: see ``CQTPlan`` ("halve-while-even" for 44.1 kHz / 22.05 kHz material)."""
        if net.training:
            raise RuntimeError("KeyEstimator runs the eval-mode forward (eval.py:116); call net.eval() first")
        self.net = net
        self.device = torch.device(device) if device is not None else next(net.parameters()).device
        if self.device.type != "cuda":
            raise RuntimeError("KeyEstimator needs the network on a CUDA device; there is no CPU fallback")
        self.sr = float(sr)
        self.plan = CQTPlan.get(sr, round(sr / frames), net.pitches, 36, recursion=recursion, peak=peak)
        self.genre = bool(net._genre)

    def frames(self, n_samples: int) -> int:
        return self.plan.frames(n_samples)

    def estimate_host(self, audio: torch.Tensor, lengths: Optional[Sequence[int]] = None, out: Optional[dict] = None) -> dict:
        """audio: (B, n_max) fp32 HOST tensor (pin it for full PCIe rate), or int16 = 16-bit PCM as the .wav files hold it
        (normalised on the device exactly as torchaudio.load does, int16 / 32768, KeyDataset.py:478-481: bit-identical results
        for half the PCIe bytes).  Returns host tensors
        ``key`` (B,12) sigmoid, ``tonic`` (B,12), ``genre`` (B,11)|None and ``ids`` (3,B) int32
        (key-signature id, tonic id, genre id or -1).  Synchronous (the call ends with a stream sync)."""
        if audio.is_cuda or audio.dtype not in (torch.float32, torch.int16) or audio.dim() != 2 or audio.stride(1) != 1:
            raise ValueError("audio must be a (B, n_max) float32 or int16 (16-bit PCM) host tensor with unit sample stride")
        pcm16 = audio.dtype == torch.int16
        lib = _lib.lib()
        B, n_max = int(audio.shape[0]), int(audio.shape[1])
        if out is None:
            out = {"key": torch.empty((B, 12), dtype=torch.float32).pin_memory(),
                   "tonic": torch.empty((B, 12), dtype=torch.float32).pin_memory(),
                   "genre": torch.empty((B, 11), dtype=torch.float32).pin_memory() if self.genre else None,
                   "ids": torch.empty((3, B), dtype=torch.int32).pin_memory()}
        len_arr = None
        if lengths is not None:
            if len(lengths) != B:
                raise ValueError("lengths must have one entry per clip")
            len_arr = (C.c_int64 * B)(*[int(v) for v in lengths])
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream(self.device).cuda_stream
            self.net._sync_params(self.device, stream)
            size_fn = lib.ake_estimate_workspace_bytes_i16 if pcm16 else lib.ake_estimate_workspace_bytes
            ws_bytes = size_fn(self.plan._h, self.net._plan, B, n_max)
            if ws_bytes == 0:
                check(_lib.AKE_ERR_INVALID)
            ws = _Workspace.get(self.device, ws_bytes, "estimate")
            run = lib.ake_estimate_host_i16 if pcm16 else lib.ake_estimate_host_f32
            check(run(
                self.plan._h, self.net._plan, audio.data_ptr(), int(audio.stride(0)), len_arr, B, n_max,
                out["key"].data_ptr(), out["tonic"].data_ptr(),
                out["genre"].data_ptr() if self.genre else None, out["ids"].data_ptr(),
                ws.data_ptr(), ws.numel(), stream))
        return out

    @torch.no_grad()
    def estimate_device(self, audio: torch.Tensor, lengths: Optional[Sequence[int]] = None):
        """audio already resident on the GPU: (key, tonic[, genre]) device tensors, asynchronous."""
        mel, seq = self.plan.run(audio, lengths=lengths)
        return self.net(mel, seq)

    @torch.no_grad()
    def estimate_device_rows(self, audio: torch.Tensor, lengths: Optional[Sequence[int]] = None):
        """audio already resident on the GPU -> ``rows`` (B, 35) = [key | tonic | genre] and ``ids`` (3, B) int32, asynchronous
        (CQT, then forward + decode in one C-ABI call writing the result rows a multi-GPU job all-gathers)."""
        mel, seq = self.plan.run(audio, lengths=lengths)
        return self.net.forward_rows(mel, seq)

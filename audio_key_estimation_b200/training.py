"""Training step of the reference (train_model.py:122 -> models.py:952-961 -> general_step, models.py:819-896).

``criterion`` restates the objective with torch ops (autograd-friendly; pairs with ``PitchClassNet.forward``
in train mode, whose backward runs in CUDA); ``TrainStep`` is the fused path: forward that keeps its
activations -> ``ake_loss_f32`` -> ``ake_pcn_backward_f32`` -> ONE flat gradient buffer in the layout of
the parameter buffer, which a data-parallel job all-reduces as a single bucket (distributed.allreduce_gradients)
before the optimizer step.  ``FusedAdam`` is the optimizer of models.py:1017-1027 (Adam + ExponentialLR) as one launch over
that bucket; torch's own ``torch.optim.Adam`` keeps working on the per-parameter ``.grad`` views as well.
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import nn

from . import _lib
from ._lib import check
from .models import PitchClassNet


def _opt(opt, name, default):
    return getattr(opt, name, default) if opt is not None else default


def criterion(outputs, key_labels: torch.Tensor, tonic_labels: torch.Tensor, genre_labels: Optional[torch.Tensor] = None,
              opt=None) -> torch.Tensor:
    """models.py:855-896, global key estimation: key_weight * BCELoss + tonic_weight * CrossEntropyLoss
    [+ genre_weight * CrossEntropyLoss over the clips that have a genre label].  Labels as the data layer
    delivers them: key_labels (B,12) multi-hot, tonic_labels (B,12) one-hot, genre_labels (B,11) one-hot or all-zero."""
    key_out, tonic_out = outputs[0], outputs[1]
    loss = _opt(opt, "key_weight", 1.0) * nn.functional.binary_cross_entropy(key_out, key_labels.to(key_out.dtype))
    loss = loss + _opt(opt, "tonic_weight", 1.0) * nn.functional.cross_entropy(tonic_out, tonic_labels.argmax(dim=1))
    if len(outputs) > 2 and genre_labels is not None:
        mask = genre_labels.sum(dim=1) == 1
        if bool(mask.any()):
            loss = loss + _opt(opt, "genre_weight", 0.1) * nn.functional.cross_entropy(outputs[2][mask], genre_labels[mask].argmax(dim=1))
    return loss


class TrainStep:
    """Fused forward + loss + backward on the device.  ``step(...)`` returns the loss terms and leaves the
    gradients in ``self.flat_grads`` (and, as views of it, in every parameter's ``.grad``)."""

    def __init__(self, net: PitchClassNet, opt=None):
        self.net, self.opt = net, opt if opt is not None else net.opt
        self.flat_grads: Optional[torch.Tensor] = None

    def step(self, mel: torch.Tensor, seq_length, key_labels: torch.Tensor, tonic_labels: torch.Tensor,
             genre_labels: Optional[torch.Tensor] = None, assign_grads: bool = True) -> dict:
        net = self.net
        if not net.training:
            raise RuntimeError("TrainStep needs the network in train mode (batch-statistics BatchNorm)")
        if not mel.is_cuda:
            raise RuntimeError("the training step runs on CUDA tensors only; there is no CPU fallback")
        lib = _lib.lib()
        dev = mel.device
        B, T = int(mel.shape[0]), int(mel.shape[3])
        x = mel.detach().to(torch.float32).contiguous()
        seq = None
        if seq_length is not None:
            seq = torch.as_tensor(seq_length).reshape(-1).to(device=dev, dtype=torch.int32).contiguous()
            if seq.numel() == 1 and B > 1:
                seq = seq.expand(B).contiguous()
        keyl = key_labels.to(device=dev, dtype=torch.float32).contiguous()
        tonic_idx = tonic_labels.to(dev).argmax(dim=1).to(torch.int32).contiguous()
        genre_idx = None
        if net._genre:
            if genre_labels is None:
                genre_idx = torch.full((B,), -1, dtype=torch.int32, device=dev)
            else:
                gl = genre_labels.to(dev)
                genre_idx = torch.where(gl.sum(dim=1) == 1, gl.argmax(dim=1), torch.full((B,), -1, device=dev)).to(torch.int32).contiguous()
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            net._sync_params(dev, stream)
            ws_bytes = lib.ake_pcn_workspace_bytes(net._plan, B, T, 2)
            if ws_bytes == 0:
                check(_lib.AKE_ERR_UNSUPPORTED)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            key = torch.empty((B, 12), dtype=torch.float32, device=dev)
            tonic = torch.empty((B, 12), dtype=torch.float32, device=dev)
            genre = torch.empty((B, 11), dtype=torch.float32, device=dev) if net._genre else None
            stats = torch.empty(2 * sum(net._bn_channels), dtype=torch.float32, device=dev)
            check(lib.ake_pcn_forward_f32(net._plan, x.data_ptr(), B, T, seq.data_ptr() if seq is not None else None, 2,
                                          key.data_ptr(), tonic.data_ptr(), genre.data_ptr() if genre is not None else None,
                                          stats.data_ptr(), ws.data_ptr(), ws.numel(), stream))
            loss = torch.empty(4, dtype=torch.float32, device=dev)
            dk, dt = torch.empty_like(key), torch.empty_like(tonic)
            dg = torch.empty_like(genre) if genre is not None else None
            check(lib.ake_loss_f32(key.data_ptr(), tonic.data_ptr(), genre.data_ptr() if genre is not None else None,
                                   keyl.data_ptr(), tonic_idx.data_ptr(), genre_idx.data_ptr() if genre_idx is not None else None, B,
                                   float(_opt(self.opt, "key_weight", 1.0)), float(_opt(self.opt, "tonic_weight", 1.0)),
                                   float(_opt(self.opt, "genre_weight", 0.1)), loss.data_ptr(), dk.data_ptr(), dt.data_ptr(),
                                   dg.data_ptr() if dg is not None else None, stream))
            flat = torch.empty(lib.ake_pcn_param_floats(net._plan), dtype=torch.float32, device=dev)
            check(lib.ake_pcn_backward_f32(net._plan, dk.data_ptr(), dt.data_ptr(), dg.data_ptr() if dg is not None else None,
                                           flat.data_ptr(), flat.numel(), ws.data_ptr(), ws.numel(), stream))
        net._update_running_stats(stats, B, T)
        self.flat_grads = flat
        if assign_grads:
            self.assign_grads()
        return {"loss": loss[0], "bce": loss[1], "tonic": loss[2], "genre": loss[3], "key_out": key, "tonic_out": tonic,
                "genre_out": genre}

    def assign_grads(self) -> None:
        """Point every parameter's .grad at its slice of the flat buffer (call again after an all-reduce in place)."""
        for prm, g in zip(self.net._grad_params(), self.net._split_flat_grads(self.flat_grads)):
            prm.grad = g


class FusedAdam:
    """models.py:1017-1027 on the flat gradient bucket: ``torch.optim.Adam(params, betas=(0.9, 0.999), lr=opt.lr,
    weight_decay=opt.reg)`` + ``ExponentialLR(gamma=opt.gamma)`` (``epoch_end()`` = scheduler.step()), one kernel launch
    per step (``ake_adam_step_f32``) instead of torch's per-tensor loop.  ``grad_scale`` folds in the 1/accumulate_grad_batches
    of train_model.py:120 (or 1/world_size after a summing all-reduce)."""

    def __init__(self, net: PitchClassNet, lr: Optional[float] = None, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: Optional[float] = None, gamma: Optional[float] = None):
        self.net = net
        self.lr0 = float(_opt(net.opt, "lr", 3e-4) if lr is None else lr)
        self.weight_decay = float(_opt(net.opt, "reg", 0.0) if weight_decay is None else weight_decay)
        self.gamma = float(_opt(net.opt, "gamma", 0.96) if gamma is None else gamma)
        self.betas, self.eps = (float(betas[0]), float(betas[1])), float(eps)
        self.epoch, self.t = 0, 0
        self._m = self._v = self._ptrs = self._offsets = None

    @property
    def lr(self) -> float:
        return self.lr0 * self.gamma ** self.epoch

    def epoch_end(self) -> None:
        self.epoch += 1

    def _tables(self, dev):
        net = self.net
        ptrs, offs, off = [], [0], 0
        for n in net._tensor_names:
            t = net._lookup(n)
            if isinstance(t, nn.Parameter):
                if t.dtype != torch.float32 or not t.is_cuda or not t.is_contiguous():
                    raise RuntimeError("FusedAdam updates contiguous float32 CUDA parameters in place")
                ptrs.append(t.data_ptr())
            else:
                ptrs.append(0)
            off += t.numel()
            offs.append(off)
        self._ptrs = torch.tensor(ptrs, dtype=torch.int64, device=dev)
        self._offsets = torch.tensor(offs, dtype=torch.int64, device=dev)
        self._key = tuple(ptrs)
        return off

    def step(self, flat_grads: torch.Tensor, grad_scale: float = 1.0) -> None:
        if not flat_grads.is_cuda or flat_grads.dtype != torch.float32:
            raise RuntimeError("FusedAdam runs on the float32 CUDA gradient bucket; there is no CPU fallback")
        dev = flat_grads.device
        net = self.net
        cur = tuple(net._lookup(n).data_ptr() if isinstance(net._lookup(n), nn.Parameter) else 0 for n in net._tensor_names)
        if self._ptrs is None or cur != self._key:
            total = self._tables(dev)
            if self._m is None:
                self._m = torch.zeros(total, dtype=torch.float32, device=dev)
                self._v = torch.zeros(total, dtype=torch.float32, device=dev)
        if flat_grads.numel() != self._m.numel():
            raise ValueError("gradient bucket does not match the parameter buffer layout")
        self.t += 1
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            check(_lib.lib().ake_adam_step_f32(flat_grads.data_ptr(), self._m.data_ptr(), self._v.data_ptr(), self._ptrs.data_ptr(),
                                               self._offsets.data_ptr(), len(net._tensor_names), self._m.numel(), self.lr,
                                               self.betas[0], self.betas[1], self.eps, self.weight_decay, float(grad_scale), self.t,
                                               stream))
        net._param_key = None  # the kernel wrote the parameters behind torch's version counters: re-upload on the next forward

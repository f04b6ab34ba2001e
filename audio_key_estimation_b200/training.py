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
    gradients in ``self.flat_grads`` (and, as views of it, in every parameter's ``.grad``).

    ``graph=True`` captures the ~230 launches of forward + loss + backward into ONE CUDA graph per input shape (static
    input / output / workspace buffers; the batch is copied into the static inputs and the graph replayed): at the
    reference's batch size of 8 the step is launch-bound, not compute-bound."""

    def __init__(self, net: PitchClassNet, opt=None, graph: bool = False):
        self.net, self.opt = net, opt if opt is not None else net.opt
        self.flat_grads: Optional[torch.Tensor] = None
        self.use_graph = bool(graph)
        self._graphs = {}
        self.launches_per_step = 0  # graph mode: kernels in the captured step (the library's launch counter only sees the capture)

    def _buffers(self, dev, B: int, T: int, has_seq: bool):
        net, lib = self.net, _lib.lib()
        ws_bytes = lib.ake_pcn_workspace_bytes(net._plan, B, T, 2)
        if ws_bytes == 0:
            check(_lib.AKE_ERR_UNSUPPORTED)
        f32 = dict(dtype=torch.float32, device=dev)
        i32 = dict(dtype=torch.int32, device=dev)
        return {
            "mel": torch.empty((B, 1, net.pitches, T), **f32), "seq": torch.empty(B, **i32) if has_seq else None,
            "keyl": torch.empty((B, 12), **f32), "tonic_idx": torch.empty(B, **i32),
            "genre_idx": torch.empty(B, **i32) if net._genre else None,
            "ws": torch.empty(ws_bytes, dtype=torch.uint8, device=dev),
            "key": torch.empty((B, 12), **f32), "tonic": torch.empty((B, 12), **f32),
            "genre": torch.empty((B, 11), **f32) if net._genre else None,
            "stats": torch.empty(2 * sum(net._bn_channels), **f32), "loss": torch.empty(4, **f32),
            "dk": torch.empty((B, 12), **f32), "dt": torch.empty((B, 12), **f32),
            "dg": torch.empty((B, 11), **f32) if net._genre else None,
            "flat": torch.empty(lib.ake_pcn_param_floats(net._plan), **f32),
        }

    def _launch(self, bf, B: int, T: int) -> None:
        """forward (activations kept) -> loss + d(loss)/d(outputs) -> backward, on the current stream."""
        net, lib = self.net, _lib.lib()
        ptr = lambda t: t.data_ptr() if t is not None else None  # noqa: E731
        stream = torch.cuda.current_stream(bf["mel"].device).cuda_stream
        check(lib.ake_pcn_forward_f32(net._plan, ptr(bf["mel"]), B, T, ptr(bf["seq"]), 2, ptr(bf["key"]), ptr(bf["tonic"]),
                                      ptr(bf["genre"]), ptr(bf["stats"]), ptr(bf["ws"]), bf["ws"].numel(), stream))
        check(lib.ake_loss_f32(ptr(bf["key"]), ptr(bf["tonic"]), ptr(bf["genre"]), ptr(bf["keyl"]), ptr(bf["tonic_idx"]),
                               ptr(bf["genre_idx"]), B, float(_opt(self.opt, "key_weight", 1.0)),
                               float(_opt(self.opt, "tonic_weight", 1.0)), float(_opt(self.opt, "genre_weight", 0.1)),
                               ptr(bf["loss"]), ptr(bf["dk"]), ptr(bf["dt"]), ptr(bf["dg"]), stream))
        check(lib.ake_pcn_backward_f32(net._plan, ptr(bf["dk"]), ptr(bf["dt"]), ptr(bf["dg"]), ptr(bf["flat"]), bf["flat"].numel(),
                                       ptr(bf["ws"]), bf["ws"].numel(), stream))

    def step(self, mel: torch.Tensor, seq_length, key_labels: torch.Tensor, tonic_labels: torch.Tensor,
             genre_labels: Optional[torch.Tensor] = None, assign_grads: bool = True) -> dict:
        net = self.net
        if not net.training:
            raise RuntimeError("TrainStep needs the network in train mode (batch-statistics BatchNorm)")
        if not mel.is_cuda:
            raise RuntimeError("the training step runs on CUDA tensors only; there is no CPU fallback")
        dev = mel.device
        B, T = int(mel.shape[0]), int(mel.shape[3])
        seq = None
        if seq_length is not None:
            seq = torch.as_tensor(seq_length).reshape(-1).to(device=dev, dtype=torch.int32)
            if seq.numel() == 1 and B > 1:
                seq = seq.expand(B)
        tonic_idx = tonic_labels.to(dev).argmax(dim=1).to(torch.int32)
        genre_idx = None
        if net._genre:
            if genre_labels is None:
                genre_idx = torch.full((B,), -1, dtype=torch.int32, device=dev)
            else:
                gl = genre_labels.to(dev)
                genre_idx = torch.where(gl.sum(dim=1) == 1, gl.argmax(dim=1), torch.full((B,), -1, device=dev)).to(torch.int32)
        with torch.cuda.device(dev):
            net._sync_params(dev, torch.cuda.current_stream(dev).cuda_stream)
            sig = (str(dev), B, T, seq is not None)
            entry = self._graphs.get(sig) if self.use_graph else None
            bf = entry[0] if entry else self._buffers(dev, B, T, seq is not None)
            bf["mel"].copy_(mel.detach().reshape(bf["mel"].shape))
            bf["keyl"].copy_(key_labels.to(dev))
            bf["tonic_idx"].copy_(tonic_idx)
            if seq is not None:
                bf["seq"].copy_(seq)
            if genre_idx is not None:
                bf["genre_idx"].copy_(genre_idx)
            if not self.use_graph:
                self._launch(bf, B, T)
            else:
                if entry is None:
                    # warm-up on a side stream (lazy one-time initialisation must not happen under capture), then capture
                    side = torch.cuda.Stream(device=dev)
                    side.wait_stream(torch.cuda.current_stream(dev))
                    with torch.cuda.stream(side):
                        n0 = _lib.lib().ake_launch_count(0)
                        self._launch(bf, B, T)
                        self.launches_per_step = int(_lib.lib().ake_launch_count(0) - n0)  # kernels one replay launches
                    torch.cuda.current_stream(dev).wait_stream(side)
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):
                        self._launch(bf, B, T)
                    entry = (bf, g)
                    self._graphs[sig] = entry
                entry[1].replay()
        net._update_running_stats(bf["stats"], B, T)
        self.flat_grads = bf["flat"]
        if assign_grads:
            self.assign_grads()
        return {"loss": bf["loss"][0], "bce": bf["loss"][1], "tonic": bf["loss"][2], "genre": bf["loss"][3], "key_out": bf["key"],
                "tonic_out": bf["tonic"], "genre_out": bf["genre"]}

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

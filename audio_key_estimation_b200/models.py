"""``PitchClassNet`` -- drop-in for the reference's ``models.PitchClassNet`` forward path.

Same constructor arguments, same ``state_dict`` key names/shapes, same ``forward(mel, seq_length)``
contract and output tuple as models.py:651-817 (call sites train_model.py:105, eval.py:98,
equivariance_test.py:178), but the forward pass is one call into libake_b200.so
(``ake_pcn_forward_f32``, include/ake_b200.h) running hand-written sm_100a kernels.  There is no
PyTorch/CPU fallback: CPU tensors, missing CUDA library and unsupported architecture switches
all raise.

Differences a maintainer should know (also in INTEGRATION.md):
* arithmetic is fp32 on the device (the reference runs cuDNN float64); inputs of any float dtype
  are accepted and the outputs are returned in the input's dtype (``general_step`` feeds
  ``key_out`` to ``BCELoss`` against ``.double()`` labels, models.py:823, 878);
* training (train_model.py:122, models.py:952-961): in ``.train()`` mode with grad enabled the forward
  runs through ``ake_pcn_forward_f32(bn_mode=2)`` and ``loss.backward()`` through
  ``ake_pcn_backward_f32`` (a ``torch.autograd.Function``), for the train_model.py default
  architecture (num_layers 2, head_layers 2, ``max_pool`` off); ``training.TrainStep`` is the fused
  forward + loss + backward call;
* ``net.modules()`` works (the reference shadows it with a list, models.py:673).
"""
from __future__ import annotations

import ctypes as C
import math
import threading
from typing import Optional, Tuple

import torch
from torch import nn

from . import _lib
from ._lib import PcnConfig, check

# architecture switches of the reference's opt (models.py:260-350, 720-722): all but denseblock / only_semitones are built
_ARCH_FLAGS = ("resblock", "denseblock", "stay_sixth", "only_semitones", "p2pc_conv", "pc2p_mem", "local")


class _Workspace:
    """Grow-only device scratch buffer per (device, CUDA stream, tag).

    Keyed by the stream the work is launched on: calls on one stream are ordered, so they may share scratch; two host
    threads driving two plans on two streams (include/ake_b200.h: "different plans may be driven from different host
    threads / streams") get separate buffers."""

    _bufs: dict = {}
    _lock = threading.Lock()

    @classmethod
    def get(cls, device: torch.device, nbytes: int, tag: str = "pcn") -> torch.Tensor:
        with cls._lock:
            return cls._get(device, nbytes, tag)

    @classmethod
    def _get(cls, device: torch.device, nbytes: int, tag: str) -> torch.Tensor:
        index = device.index if device.index is not None else torch.cuda.current_device()
        key = (index, int(torch.cuda.current_stream(device).cuda_stream), tag)
        buf = cls._bufs.get(key)
        if buf is None or buf.numel() < nbytes:
            buf = None
            cls._bufs.pop(key, None)
            buf = torch.empty(int(nbytes * 1.05) + 256, dtype=torch.uint8, device=device)
            cls._bufs[key] = buf
        return buf


def _opt_get(opt, name, default):
    return getattr(opt, name, default) if opt is not None else default


class _Node(nn.Module):
    """Name-only container so parameters appear under the reference's dotted state_dict keys."""


class _KeptForward(torch.autograd.Function):
    """Train-mode forward that keeps its activations (bn_mode 2) + backward through ake_pcn_backward_f32."""

    @staticmethod
    def forward(ctx, net, x, seq, *params):
        lib = _lib.lib()
        device = x.device
        B, T = int(x.shape[0]), int(x.shape[3])
        with torch.cuda.device(device):
            stream = torch.cuda.current_stream(device).cuda_stream
            ws_bytes = lib.ake_pcn_workspace_bytes(net._plan, B, T, 2)
            if ws_bytes == 0:
                check(_lib.AKE_ERR_UNSUPPORTED if b"built for" in (lib.ake_last_error() or b"") else _lib.AKE_ERR_INVALID)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=device)  # private: it must survive until backward
            key = torch.empty((B, 12), dtype=torch.float32, device=device)
            tonic = torch.empty((B, 12), dtype=torch.float32, device=device)
            genre = torch.empty((B, 11), dtype=torch.float32, device=device) if net._genre else None
            stats = torch.empty(2 * sum(net._bn_channels), dtype=torch.float32, device=device)
            check(lib.ake_pcn_forward_f32(
                net._plan, x.data_ptr(), B, T, seq.data_ptr() if seq is not None else None, 2, key.data_ptr(),
                tonic.data_ptr(), genre.data_ptr() if genre is not None else None, stats.data_ptr(), ws.data_ptr(),
                ws.numel(), stream))
        net._update_running_stats(stats, B, T)
        ctx.net, ctx.ws, ctx.keep = net, ws, (x, seq, key)
        ctx.n_params = len(params)
        return (key, tonic, genre) if genre is not None else (key, tonic)

    @staticmethod
    def backward(ctx, *douts):
        net, ws = ctx.net, ctx.ws
        lib = _lib.lib()
        device = ws.device
        d = [g.contiguous().to(torch.float32) if g is not None else None for g in douts]
        while len(d) < 3:
            d.append(None)
        with torch.cuda.device(device):
            stream = torch.cuda.current_stream(device).cuda_stream
            flat = torch.empty(lib.ake_pcn_param_floats(net._plan), dtype=torch.float32, device=device)
            check(lib.ake_pcn_backward_f32(net._plan, *(g.data_ptr() if g is not None else None for g in d), flat.data_ptr(),
                                           flat.numel(), ws.data_ptr(), ws.numel(), stream))
        return (None, None, None) + tuple(net._split_flat_grads(flat))


class PitchClassNet(nn.Module):
    """Reference signature: models.py:653.  ``opt`` carries the train_model.py:160-242 flags."""

    def __init__(self, pitches, pitch_classes, num_layers, kernel_size, opt=None, window_size=23, batch_size=4,
                 train_set=None, val_set=None):
        super().__init__()
        if opt is None:
            # the reference dereferences opt.conv_layers unconditionally (models.py:662) -> AttributeError
            raise AttributeError("'NoneType' object has no attribute 'conv_layers'")
        self.pitches, self.pitch_classes = int(pitches), int(pitch_classes)
        self.num_layers, self.kernel_size = int(num_layers), int(kernel_size)
        self.batch_size, self.window_size, self.opt = batch_size, window_size, opt
        self.conv_layers, self.n_filters = int(opt.conv_layers), int(opt.n_filters)
        self.resblock, self.denseblock = bool(_opt_get(opt, "resblock", False)), bool(_opt_get(opt, "denseblock", False))
        self.best_mirex_score = 0
        self.data = {"train": train_set, "val": val_set}

        cfg = PcnConfig(
            pitches=self.pitches, pitch_classes=self.pitch_classes, num_layers=self.num_layers,
            kernel_size=self.kernel_size, conv_layers=self.conv_layers, n_filters=self.n_filters,
            head_layers=int(_opt_get(opt, "head_layers", 2)), time_pool_size=int(_opt_get(opt, "time_pool_size", 2)),
            genre=int(bool(_opt_get(opt, "genre", False))), max_pool=int(bool(_opt_get(opt, "max_pool", False))),
            frames=int(_opt_get(opt, "frames", 5)), loc_window_size=int(_opt_get(opt, "loc_window_size", 10)),
            **{f: int(bool(_opt_get(opt, f, False))) for f in _ARCH_FLAGS})
        self._cfg = cfg
        self._genre = bool(cfg.genre)
        self._local = bool(cfg.local)
        self._default_arch = not any(getattr(cfg, f) for f in _ARCH_FLAGS)
        lib = _lib.lib()
        handle = C.c_void_p()
        check(lib.ake_pcn_create(C.byref(cfg), C.byref(handle)))  # NotImplementedError for unsupported switches
        self._plan = handle
        self._plan_device: Optional[torch.device] = None
        self._param_key = None

        # ---- parameters / buffers under the reference's names
        self._tensor_names = []
        n = lib.ake_pcn_num_tensors(handle)
        shape4 = (C.c_int64 * 4)()
        for i in range(n):
            name = lib.ake_pcn_tensor_name(handle, i).decode()
            nd = lib.ake_pcn_tensor_shape(handle, i, C.byref(shape4))
            shape = tuple(int(shape4[k]) for k in range(nd))
            self._tensor_names.append(name)
            self._register(name, shape)
        self._bn_sites = [nm[: -len(".running_mean")] for nm in self._tensor_names if nm.endswith(".running_mean")]
        self._bn_channels = [self._lookup(s + ".running_mean").numel() for s in self._bn_sites]
        self.sig = nn.Sigmoid()

    # -------------------------------------------------------------------------------- module tree
    def _register(self, name: str, shape) -> None:
        parts = name.split(".")
        node = self
        for p in parts[:-1]:
            if p not in node._modules:
                node.add_module(p, _Node())
            node = node._modules[p]
        leaf = parts[-1]
        if leaf in ("running_mean", "running_var"):
            init = torch.zeros(shape) if leaf == "running_mean" else torch.ones(shape)
            node.register_buffer(leaf, init)
            if leaf == "running_var":
                node.register_buffer("num_batches_tracked", torch.tensor(0, dtype=torch.long))
            return
        is_bn = len(shape) == 1 and not hasattr(node, "weight") and leaf == "weight"
        if len(shape) == 4:
            # torch default conv init: kaiming_uniform(a=sqrt(5)) == U(-1/sqrt(fan_in), 1/sqrt(fan_in));
            # train_model.py:14-17 defines weights_init but never applies it.
            fan_in = shape[1] * shape[2] * shape[3]
            node._ake_fan_in = fan_in
            t = torch.empty(shape).uniform_(-1.0 / math.sqrt(fan_in), 1.0 / math.sqrt(fan_in))
        elif is_bn:
            t = torch.ones(shape)
        elif leaf == "bias" and hasattr(node, "_ake_fan_in"):
            b = 1.0 / math.sqrt(node._ake_fan_in)
            t = torch.empty(shape).uniform_(-b, b)
        else:
            t = torch.zeros(shape)  # BatchNorm bias
        node.register_parameter(leaf, nn.Parameter(t))

    def _lookup(self, name: str) -> torch.Tensor:
        node = self
        parts = name.split(".")
        for p in parts[:-1]:
            node = node._modules[p]
        leaf = parts[-1]
        return node._parameters[leaf] if leaf in node._parameters else node._buffers[leaf]

    def __del__(self):
        try:
            if getattr(self, "_plan", None):
                _lib.lib().ake_pcn_destroy(self._plan)
                self._plan = None
        except Exception:
            pass

    # ------------------------------------------------------------------------------- parameters
    def _sync_params(self, device: torch.device, stream_ptr: int) -> None:
        tensors = [self._lookup(n) for n in self._tensor_names]
        key = (str(device),) + tuple((t.data_ptr(), t._version) for t in tensors)
        if key == self._param_key:
            return
        for t in tensors:
            if t.device != device:
                raise RuntimeError(f"PitchClassNet parameters live on {t.device} but the input is on {device}; "
                                   "move the module with .cuda()/.to(device) (there is no CPU path)")
        flat = torch.cat([t.detach().reshape(-1).to(torch.float32) for t in tensors])
        lib = _lib.lib()
        check(lib.ake_pcn_set_params_f32(self._plan, flat.data_ptr(), flat.numel(), stream_ptr))
        self._param_key = key

    # ---------------------------------------------------------------------------------- forward
    def forward(self, mel: torch.Tensor, seq_length=None) -> Tuple[torch.Tensor, ...]:
        """models.py:747-817.  mel (B,1,pitches,T); seq_length None | int tensor (B,) | (1,1)."""
        if not isinstance(mel, torch.Tensor) or mel.dim() != 4 or mel.shape[1] != 1 or mel.shape[2] != self.pitches:
            raise ValueError(f"mel must be (B, 1, {self.pitches}, T), got {tuple(getattr(mel, 'shape', ()))}")
        if not mel.is_cuda:
            raise RuntimeError("PitchClassNet (B200) runs on CUDA tensors only; there is no CPU fallback")
        if not mel.is_floating_point():
            raise ValueError("mel must be a floating-point tensor")
        lib = _lib.lib()
        device, out_dtype = mel.device, mel.dtype
        B, T = int(mel.shape[0]), int(mel.shape[3])
        with torch.cuda.device(device):
            stream = torch.cuda.current_stream(device).cuda_stream
            self._sync_params(device, stream)
            x = mel.detach().to(torch.float32).contiguous()
            seq = None
            if seq_length is not None:
                seq = torch.as_tensor(seq_length).reshape(-1)
                if seq.numel() == 1 and B > 1:
                    seq = seq.expand(B)
                if seq.numel() != B:
                    raise ValueError(f"seq_length has {seq.numel()} entries for a batch of {B}")
                seq = seq.to(device=device, dtype=torch.int32).contiguous()
            train = bool(self.training)
            if train and torch.is_grad_enabled() and mel.requires_grad:
                raise NotImplementedError("gradients with respect to the input spectrogram are not computed by the B200 backward "
                                          "pass (the reference never asks for them: KeyDataset tensors carry no grad)")
            if train and torch.is_grad_enabled() and any(p.requires_grad for p in self._grad_params()):
                outs = _KeptForward.apply(self, x, seq, *self._grad_params())
                return tuple(o.to(out_dtype) for o in outs)
            ws_bytes = lib.ake_pcn_workspace_bytes(self._plan, B, T, int(train))
            if ws_bytes == 0:
                check(_lib.AKE_ERR_INVALID)
            ws = _Workspace.get(device, ws_bytes)
            if self._local:
                # opt.local (models.py:804-810): one value per pitch class and window; the reference hands the (B, 1, rows, T')
                # maps out RESHAPED (not permuted) to (B, T', rows) -- the same memory, viewed the same way here
                g_frames = C.c_int(0)
                Tl = lib.ake_pcn_local_frames(self._plan, T, C.byref(g_frames))
                if Tl < 0:
                    raise ValueError(f"T={T} is too short for the sliding-window heads of opt.local")
                key = torch.empty((B, Tl, 12), dtype=torch.float32, device=device)
                tonic = torch.empty((B, Tl, 12), dtype=torch.float32, device=device)
                genre = torch.empty((B, g_frames.value, 11), dtype=torch.float32, device=device) if self._genre else None
            else:
                key = torch.empty((B, 12), dtype=torch.float32, device=device)
                tonic = torch.empty((B, 12), dtype=torch.float32, device=device)
                genre = torch.empty((B, 11), dtype=torch.float32, device=device) if self._genre else None
            stats = None
            if train:
                stats = torch.empty(2 * sum(self._bn_channels), dtype=torch.float32, device=device)
            check(lib.ake_pcn_forward_f32(
                self._plan, x.data_ptr(), B, T, seq.data_ptr() if seq is not None else None, int(train),
                key.data_ptr(), tonic.data_ptr(), genre.data_ptr() if genre is not None else None,
                stats.data_ptr() if stats is not None else None, ws.data_ptr(), ws.numel(), stream))
            if train:
                self._update_running_stats(stats, B, T)
        outs = (key, tonic) + ((genre,) if self._genre else ())
        return tuple(o.to(out_dtype) for o in outs)

    @torch.no_grad()
    def forward_rows(self, mel: torch.Tensor, seq_length=None, decode: bool = True):
        """Eval-mode forward with the three outputs side by side: ``rows`` (B, 35) fp32 = [12 key probabilities | 12 tonic
        logits | 11 genre logits (zeros without a genre head)] and, with ``decode``, ``ids`` (3, B) int32 (key signature,
        tonic, genre or -1) -- ONE call (``ake_pcn_forward_rows_f32``); ``rows`` is the table a data-parallel job
        all-gathers.  Same arithmetic as ``forward`` + ``decode``."""
        if self.training:
            raise RuntimeError("forward_rows is the eval-mode path (eval.py:116); call .eval() first")
        if not isinstance(mel, torch.Tensor) or mel.dim() != 4 or mel.shape[1] != 1 or mel.shape[2] != self.pitches:
            raise ValueError(f"mel must be (B, 1, {self.pitches}, T), got {tuple(getattr(mel, 'shape', ()))}")
        if not mel.is_cuda:
            raise RuntimeError("PitchClassNet (B200) runs on CUDA tensors only; there is no CPU fallback")
        lib = _lib.lib()
        device = mel.device
        B, T = int(mel.shape[0]), int(mel.shape[3])
        with torch.cuda.device(device):
            stream = torch.cuda.current_stream(device).cuda_stream
            self._sync_params(device, stream)
            x = mel.detach().to(torch.float32).contiguous()
            seq = None
            if seq_length is not None:
                seq = torch.as_tensor(seq_length).reshape(-1)
                if seq.numel() == 1 and B > 1:
                    seq = seq.expand(B)
                if seq.numel() != B:
                    raise ValueError(f"seq_length has {seq.numel()} entries for a batch of {B}")
                seq = seq.to(device=device, dtype=torch.int32).contiguous()
            ws_bytes = lib.ake_pcn_workspace_bytes(self._plan, B, T, 0)
            if ws_bytes == 0:
                check(_lib.AKE_ERR_INVALID)
            ws = _Workspace.get(device, ws_bytes)
            rows = torch.empty((B, _lib.ROW_FLOATS), dtype=torch.float32, device=device)
            ids = torch.empty((3, B), dtype=torch.int32, device=device) if decode else None
            check(lib.ake_pcn_forward_rows_f32(
                self._plan, x.data_ptr(), B, T, seq.data_ptr() if seq is not None else None, rows.data_ptr(),
                ids.data_ptr() if ids is not None else None, ws.data_ptr(), ws.numel(), stream))
        return rows, ids

    # ------------------------------------------------------------------------------- gradients
    def _grad_params(self):
        """The nn.Parameters in flat-buffer order (running statistics are buffers and carry no gradient)."""
        return [t for t in (self._lookup(n) for n in self._tensor_names) if isinstance(t, nn.Parameter)]

    def _split_flat_grads(self, flat: torch.Tensor):
        """Views of the flat gradient buffer (layout of ake_pcn_set_params_f32), one per nn.Parameter."""
        out, off = [], 0
        for n in self._tensor_names:
            t = self._lookup(n)
            if isinstance(t, nn.Parameter):
                out.append(flat[off: off + t.numel()].view(t.shape).to(t.dtype))
            off += t.numel()
        return out

    def _bn_counts(self, B: int, T: int):
        """Elements per channel each BN site normalises over (B * rows * frames), in site order."""
        counts = []
        S = self.pitches // 3
        k, hl = self.kernel_size, int(_opt_get(self.opt, "head_layers", 2))
        Tn = T
        for name in self._bn_sites:
            if name.startswith("model."):
                L = int(name.split(".")[1])
                Tl = T // (2 ** max(L - 1, 0)) if L > 0 else T
                if ".pool_semi_b" in name:
                    counts.append(B * S * Tl)
                elif ".up_sixth_b" in name:
                    counts.append(B * 36 * Tl)
                elif ".p2p." in name:
                    counts.append(B * self.pitches * Tl)
                else:
                    counts.append(B * 12 * Tl)
                Tn = Tl // 2 if L > 0 else Tl
            else:
                i = int(name.split(".")[1]) // 3  # head conv index this BN follows
                counts.append(B * 12 * (Tn - (k - 1) * (i + 1)))
        return counts

    def _bn_counts_of_last_forward(self):
        """Elements per channel each BN site normalised over in the last train-mode forward (from the library: the non-default
        architectures move the sites around)."""
        n = len(self._bn_sites)
        arr = (C.c_int64 * n)()
        got = _lib.lib().ake_pcn_bn_counts(self._plan, arr, n)
        if got != n:
            raise RuntimeError("BatchNorm site table mismatch")
        return [int(v) for v in arr]

    @torch.no_grad()
    def _update_running_stats(self, stats: torch.Tensor, B: int, T: int) -> None:
        """nn.BatchNorm2d train-mode buffer update (momentum 0.1, unbiased variance)."""
        # one multi-tensor launch per update instead of seven tiny ones per site (15 sites: the step is launch-bound at batch 8)
        counts = tuple(self._bn_counts(B, T) if self._default_arch else self._bn_counts_of_last_forward())
        first = self._lookup(self._bn_sites[0] + ".running_mean")   # .cuda() / .double() replace the buffers: rebuild the lists
        key = (counts, stats.device, stats.dtype, id(first), first.device, first.dtype)
        cached = getattr(self, "_bn_update_cache", None)
        if cached is None or cached[0] != key:
            factor = []
            for Cn, n in zip(self._bn_channels, counts):
                factor += [1.0] * Cn + [n / max(n - 1, 1)] * Cn   # batch mean as is, biased -> unbiased variance
            rms = [self._lookup(s + ".running_mean") for s in self._bn_sites]
            rvs = [self._lookup(s + ".running_var") for s in self._bn_sites]
            nbs = [self._lookup(s + ".num_batches_tracked") for s in self._bn_sites]
            cached = (key, torch.tensor(factor, dtype=stats.dtype, device=stats.device), rms, rvs, nbs)
            self._bn_update_cache = cached
        _, factor, rms, rvs, nbs = cached
        scaled = (stats * factor).to(rms[0].dtype)
        means, unbiased, off = [], [], 0
        for Cn in self._bn_channels:
            means.append(scaled[off: off + Cn])
            unbiased.append(scaled[off + Cn: off + 2 * Cn])
            off += 2 * Cn
        torch._foreach_mul_(rms + rvs, 0.9)
        torch._foreach_add_(rms + rvs, means + unbiased, alpha=0.1)
        torch._foreach_add_(nbs, 1)

    # ---------------------------------------------------------------------- parity/debug helpers
    def tap(self, name: str) -> torch.Tensor:
        """Flat fp32 copy of a named intermediate of the last forward (oracle/pcn_port.py tap names)."""
        lib = _lib.lib()
        n = lib.ake_pcn_get_tap(self._plan, name.encode(), None, 0, None)
        if n < 0:
            check(int(n))
        dev = self._lookup(self._tensor_names[0]).device
        out = torch.empty(int(n), dtype=torch.float32, device=dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        n2 = lib.ake_pcn_get_tap(self._plan, name.encode(), out.data_ptr(), out.numel(), stream)
        if n2 < 0:
            check(int(n2))
        return out


class PitchClassNet_Multi(nn.Module):
    """Drop-in for the reference's ``PitchClassNet_Multi`` (models.py:1118-1189; call sites train_model.py:100-103, eval.py:93-96):
    two ``PitchClassNet`` (``model1`` on ``mel1``, ``model2`` on ``mel2``, both running the B200 kernels) whose outputs are
    averaged, or -- with ``opt.linear_reg_multi`` -- combined by the reference's per-class linear regression.  As in the
    reference the regression coefficients are plain random tensors (``torch.randn``), not parameters: they are not part of the
    ``state_dict``, whose keys are ``model1.*`` / ``model2.*``.  The 24-value combination is host-side glue on device tensors."""

    def __init__(self, pitches1, pitches2, pitch_classes, num_layers, kernel_size, opt=None, window_size=23, batch_size=4,
                 train_set=None, val_set=None):
        super().__init__()
        if opt is None:
            raise AttributeError("'NoneType' object has no attribute 'conv_layers'")
        self.pitches1, self.pitches2, self.pitch_classes = pitches1, pitches2, pitch_classes
        self.num_layers, self.kernel_size, self.opt = num_layers, kernel_size, opt
        self.batch_size, self.window_size = batch_size, window_size
        self.conv_layers, self.n_filters = opt.conv_layers, opt.n_filters
        self.data = {"train": train_set, "val": val_set}
        kw = dict(opt=opt, window_size=window_size, batch_size=batch_size, train_set=train_set, val_set=val_set)
        self.model1 = PitchClassNet(pitches1, pitch_classes, num_layers, kernel_size, **kw)
        self.model2 = PitchClassNet(pitches2, pitch_classes, num_layers, kernel_size, **kw)
        self._genre = bool(_opt_get(opt, "genre", False))
        self._linear = bool(_opt_get(opt, "linear_reg_multi", False))
        if self._linear:
            dev = "cuda" if torch.cuda.is_available() else "cpu"
            self.wk, self.wt = torch.randn(2, 12, device=dev), torch.randn(2, 12, device=dev)
            self.bk, self.bt = torch.randn(12, device=dev), torch.randn(12, device=dev)
            if self._genre:
                self.wg, self.bg = torch.randn(2, 12, device=dev), torch.randn(12, device=dev)

    def forward(self, mel1, mel2, seq_length):
        x1 = self.model1(mel1, seq_length)
        x2 = self.model2(mel2, seq_length)
        if self._linear:
            # models.py:1169-1175 (note: the key outputs are sigmoid probabilities already and go through another sigmoid)
            f = lambda w, b, a1, a2: w[0].to(a1) * a1 + w[1].to(a1) * a2 + b.to(a1)  # noqa: E731
            out = (torch.sigmoid(f(self.wk, self.bk, x1[0], x2[0])), f(self.wt, self.bt, x1[1], x2[1]))
            if self._genre:
                out = out + (f(self.wg, self.bg, x1[2], x2[2]),)
            return out
        return tuple((a + b) / 2 for a, b in zip(x1, x2))


def decode(key_out: torch.Tensor, tonic_out: torch.Tensor, genre_out: Optional[torch.Tensor] = None):
    """argmax key signature / tonic / genre ids on the device (models.py:1083-1085, 1096, 923)."""
    if not key_out.is_cuda:
        raise RuntimeError("decode runs on CUDA tensors only")
    lib = _lib.lib()
    B = key_out.shape[0]
    dev = key_out.device
    k = key_out.detach().to(torch.float32).contiguous()
    t = tonic_out.detach().to(torch.float32).contiguous()
    g = genre_out.detach().to(torch.float32).contiguous() if genre_out is not None else None
    ids = torch.empty((3, B), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev).cuda_stream
        check(lib.ake_decode_f32(k.data_ptr(), t.data_ptr(), g.data_ptr() if g is not None else None, B,
                                 ids[0].data_ptr(), ids[1].data_ptr(), ids[2].data_ptr(), stream))
    return (ids[0], ids[1]) + ((ids[2],) if g is not None else ())


MIREX_COUNTERS = ("samples", "correct", "fifths", "relative", "parallel", "other", "all_keys", "tonics", "key_bits")


def mirex_counters(key_out: torch.Tensor, tonic_out: torch.Tensor, key_labels: torch.Tensor, tonic_labels: torch.Tensor,
                   key_signature_id: torch.Tensor, counters: Optional[torch.Tensor] = None, return_details: bool = False):
    """Category counters of the reference's ``mirex_score`` loop (models.py:1065-1116), computed on the device.

    Returns an int64 tensor of 9 counters (``MIREX_COUNTERS``) on ``key_out.device``; pass ``counters`` to keep accumulating
    over batches, and sum it over ranks with ``distributed.reduce_counters`` before ``mirex_from_counters``.  With
    ``return_details`` also returns the per-clip cosine similarity (models.py:1094) and category ids."""
    if not key_out.is_cuda:
        raise RuntimeError("mirex_counters runs on CUDA tensors only")
    lib = _lib.lib()
    dev = key_out.device
    B = key_out.shape[0]
    f = lambda x: x.detach().to(device=dev, dtype=torch.float32).contiguous()  # noqa: E731
    k, t, kl, tl, sid = f(key_out), f(tonic_out), f(key_labels), f(tonic_labels), f(key_signature_id)
    if kl.shape != (B, 12) or tl.shape != (B, 12) or t.shape != (B, 12) or k.shape != (B, 12) or sid.dim() != 2 or sid.shape[0] != B:
        raise ValueError("expected key/tonic outputs and labels of shape (B, 12) and key_signature_id of shape (B, W) "
                         "(W = 24 from the data layer's one-hot, KeyDataset.py:366, 447)")
    if counters is None:
        counters = torch.zeros(len(MIREX_COUNTERS), dtype=torch.int64, device=dev)
    sim = torch.empty(B, dtype=torch.float32, device=dev) if return_details else None
    cat = torch.empty(B, dtype=torch.int32, device=dev) if return_details else None
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev).cuda_stream
        check(lib.ake_mirex_f32(k.data_ptr(), t.data_ptr(), kl.data_ptr(), tl.data_ptr(), sid.data_ptr(), int(sid.shape[1]), B, counters.data_ptr(),
                                sim.data_ptr() if sim is not None else None, cat.data_ptr() if cat is not None else None, stream))
    return (counters, sim, cat) if return_details else counters


def mirex_from_counters(counters: torch.Tensor):
    """(mirex, correct, fifths, relative, parallel, other, accuracy) exactly as models.py:1113-1115 returns them."""
    c = [int(v) for v in counters.tolist()]
    n = max(1, c[0])
    mirex = 1.0 * c[1] + 0.5 * c[2] + 0.3 * c[3] + 0.2 * c[4]
    return tuple(torch.tensor(v / n).float() for v in (mirex, c[1], c[2], c[3], c[4], c[5], c[6]))

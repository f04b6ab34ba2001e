"""Import the UNMODIFIED reference ``models.py`` from /root/reference on a CPU box.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Only usable where
/root/reference exists (the build container); it is used by
oracle/make_golden.py to generate tests/golden/*.npz and by the CPU tests that
pin oracle.pcn_port against the real reference.  Never used on the GPU box.

Why stubs: models.py:3,7-10,15 import tensorflow, pytorch_lightning and
torchmetrics at module scope, utils/key_signatures.py:19 builds its table with
TensorFlow, and models.py:199,237,739-742 hard-code ``.double().cuda()``.
None of those packages is installed, and there is no GPU here, so we register
inert stand-ins and make ``.cuda()`` an identity when CUDA is unavailable.
No reference arithmetic is replaced: every Conv/BN/pool call is the
reference's own code running on torch CPU.
"""
from __future__ import annotations

import argparse
import importlib
import importlib.machinery
import os
import sys
import types

import numpy as np
import torch
from torch import nn

REFERENCE_ROOT = os.environ.get("AKE_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "models.py"))


def _stub(name: str) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, None)
    m.__path__ = []  # behave like a package so sub-imports resolve
    sys.modules[name] = m
    return m


class _TFTensor:
    """Just enough of tf.Tensor for utils/key_signatures.py (``.numpy()``)."""

    def __init__(self, a):
        self._a = np.asarray(a, dtype=np.float32)

    def numpy(self):
        return self._a

    def __array__(self, dtype=None, copy=None):
        return self._a if dtype is None else self._a.astype(dtype)


def _install_stubs() -> None:
    if "tensorflow" not in sys.modules:
        tf = _stub("tensorflow")
        tf.float32 = np.float32
        tf.cast = lambda x, dtype=None: _TFTensor(x)
        tf.convert_to_tensor = lambda x, *a, **k: _TFTensor(
            [np.asarray(r) for r in x] if isinstance(x, (list, tuple)) else x)
        tf.zeros = lambda shape, *a, **k: _TFTensor(np.zeros(shape))
    if "pytorch_lightning" not in sys.modules:
        pl = _stub("pytorch_lightning")
        pl.LightningModule = nn.Module
        pl.Trainer = object
        lg = _stub("pytorch_lightning.loggers")
        lg.TensorBoardLogger = object
        cb = _stub("pytorch_lightning.callbacks")
        cb.ModelCheckpoint = object
        es = _stub("pytorch_lightning.callbacks.early_stopping")
        es.EarlyStopping = object
        cb.early_stopping = es
        pl.loggers, pl.callbacks = lg, cb
    if "torchmetrics" not in sys.modules:
        tm = _stub("torchmetrics")
        tm.Accuracy = object


def _neutralise_cuda() -> None:
    if torch.cuda.is_available():
        return
    nn.Module.cuda = lambda self, device=None: self
    torch.Tensor.cuda = lambda self, *a, **k: self


_REF = None


def load_reference_models():
    """Return the reference ``models`` module (cached)."""
    global _REF
    if _REF is not None:
        return _REF
    if not reference_available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    _install_stubs()
    _neutralise_cuda()
    saved = sys.modules.pop("models", None), sys.modules.pop("utils", None)
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        _REF = importlib.import_module("models")
    finally:
        sys.path.remove(REFERENCE_ROOT)
        # keep the reference reachable only through the returned handle
        sys.modules["_ake_reference_models"] = sys.modules.pop("models")
        sys.modules.pop("utils", None)
        sys.modules.pop("utils.key_signatures", None)
        if saved[0] is not None:
            sys.modules["models"] = saved[0]
        if saved[1] is not None:
            sys.modules["utils"] = saved[1]
    return _REF


def default_opt(**overrides) -> argparse.Namespace:
    """``opt`` exactly as train_model.py:160-242 builds it with no flags."""
    d = dict(batch_size=8, lr=3e-4, drop=0.0, reg=0, gamma=0.96, acc_grad=8,
             epochs=100, window_size=592, local=False, gpu=0, octaves=8,
             conv_layers=3, n_filters=4, num_layers=2, kernel_size=7,
             key_weight=1.0, tonic_weight=1.0, genre_weight=0.1, resblock=False,
             denseblock=False, frames=5, genre=False, stay_sixth=False,
             p2pc_conv=False, head_layers=2, loc_window_size=10,
             time_pool_size=2, only_semitones=False, multi_scale=False,
             no_test=False, debug=False, linear_reg_multi=False, use_cos=False,
             pc2p_mem=False, no_ckpt=False, max_pool=False)
    d.update(overrides)
    return argparse.Namespace(**d)


def build_reference_net(pitches=288, opt=None, dtype=torch.float64):
    """The call train_model.py:105 / eval.py:98 / equivariance_test.py:178 makes."""
    ref = load_reference_models()
    opt = opt or default_opt()
    net = ref.PitchClassNet(pitches, 12, opt.num_layers, opt.kernel_size, opt=opt,
                            window_size=opt.window_size, batch_size=opt.batch_size,
                            train_set=None, val_set=None)
    return net.to(dtype)

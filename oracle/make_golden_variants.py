"""Golden vectors of the reference's NON-DEFAULT architectures: the UNMODIFIED reference
``PitchClassNet`` (models.py:108-166, 245-454, 651-817) run in the build container with each architecture switch on.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Usage:  python -m oracle.make_golden_variants
Writes tests/golden/variants.npz: per variant the tensor table of the reference state_dict (the seeded weights themselves are
regenerated from it by synth.randomise_state_dict; a checksum is stored), a seeded log-CQT-like input, and the reference's float64 outputs in eval mode (ragged seq_length) and train mode (batch
statistics) plus the BatchNorm running buffers after that train-mode forward.  Widths are reduced (n_filters 2, conv_layers 2)
where that keeps the fixture small; every switch is also exercised once at the train_model.py widths through the tensor table."""
from __future__ import annotations

import json
import os

import numpy as np
import torch

from audio_key_estimation_b200 import synth
from oracle import ref_import

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

SMALL = dict(n_filters=2, conv_layers=2)
VARIANTS = {
    "stay_sixth": dict(stay_sixth=True, **SMALL),
    "p2pc_conv": dict(p2pc_conv=True, **SMALL),
    "pc2p_mem": dict(pc2p_mem=True, **SMALL),
    "resblock": dict(resblock=True, **SMALL),
    "local": dict(local=True, **SMALL),
    "local_genre": dict(local=True, genre=True, loc_window_size=4, **SMALL),
    "res_mem_pconv_l3": dict(resblock=True, pc2p_mem=True, p2pc_conv=True, num_layers=3, n_filters=1, conv_layers=1),
    "stay_sixth_genre_l3": dict(stay_sixth=True, genre=True, num_layers=3, n_filters=1, conv_layers=2),
    "denseblock": dict(denseblock=True, **SMALL),
    "dense_l3": dict(denseblock=True, num_layers=3, n_filters=1, conv_layers=1),
    "dense_pconv_local_genre": dict(denseblock=True, p2pc_conv=True, local=True, genre=True, loc_window_size=5, **SMALL),
}
# tensor tables at the train_model.py widths (names + shapes only)
TABLES = {f: {f: True} for f in ("stay_sixth", "p2pc_conv", "pc2p_mem", "resblock", "local", "denseblock")}


def main() -> None:
    out = {}
    meta = {"variants": {}, "tables": {}}
    rng = np.random.default_rng(5)
    B, T = 3, 70
    mel = np.log1p(rng.gamma(1.0, 1.0, (B, 288, T))).astype(np.float32)
    seq = np.array([T, 61, 58], dtype=np.int64)
    for b, s in enumerate(seq):
        mel[b, :, s:] = 0.0
    out["mel"], out["seq_length"] = mel, seq
    x = torch.from_numpy(mel).double()[:, None]
    for name, kw in VARIANTS.items():
        torch.manual_seed(11)
        opt = ref_import.default_opt(**kw)
        net = ref_import.build_reference_net(288, opt)
        sd = synth.randomise_state_dict(net.state_dict(), seed=3, dtype=torch.float32)
        net.load_state_dict({k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}, strict=True)
        meta["variants"][name] = {"opt": kw, "tensors": [[k, list(v.shape)] for k, v in sd.items()]}
        # the weights are NOT stored: synth.randomise_state_dict regenerates them from the tensor table alone (seed 3); the
        # checksum lets the tests notice a numpy whose random stream differs
        out[f"{name}.sd_checksum"] = np.array([sum(float(v.double().sum()) for v in sd.values() if v.is_floating_point()),
                                               sum(float((v.double() ** 2).sum()) for v in sd.values() if v.is_floating_point())])
        net.eval()
        with torch.no_grad():
            res = net(x, torch.from_numpy(seq))
        for nm, r in zip(("key", "tonic", "genre"), res):
            out[f"{name}.eval.{nm}"] = r.numpy()
        net.train()
        with torch.no_grad():
            res = net(x, torch.from_numpy(seq))
        for nm, r in zip(("key", "tonic", "genre"), res):
            out[f"{name}.train.{nm}"] = r.numpy()
        new_sd = net.state_dict()
        for k, v in new_sd.items():
            if k.endswith("running_var") or k.endswith("running_mean"):
                out[f"{name}.buf.{k}"] = v.numpy().astype(np.float32)
        print(name, len(sd), "tensors;", [tuple(r.shape) for r in res])
    # PitchClassNet_Multi (models.py:1118-1189): two networks, outputs averaged (opt.linear_reg_multi uses unseeded random
    # coefficients that are no parameters -- not golden material)
    ref = ref_import.load_reference_models()
    kw = dict(genre=True, **SMALL)
    opt = ref_import.default_opt(**kw)
    multi = ref.PitchClassNet_Multi(288, 288, 12, opt.num_layers, opt.kernel_size, opt=opt).double()
    sd = synth.randomise_state_dict(multi.state_dict(), seed=3, dtype=torch.float32)
    multi.load_state_dict({k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}, strict=True)
    multi.eval()
    mel2 = np.log1p(rng.gamma(1.5, 0.7, (B, 288, T))).astype(np.float32)
    out["mel2"] = mel2
    with torch.no_grad():
        res = multi(x, torch.from_numpy(mel2).double()[:, None], torch.from_numpy(seq))
    for nm, r in zip(("key", "tonic", "genre"), res):
        out[f"multi.eval.{nm}"] = r.numpy()
    meta["multi"] = {"opt": kw, "tensors": [[k, list(v.shape)] for k, v in sd.items()]}
    out["multi.sd_checksum"] = np.array([sum(float(v.double().sum()) for v in sd.values() if v.is_floating_point()),
                                         sum(float((v.double() ** 2).sum()) for v in sd.values() if v.is_floating_point())])
    for name, kw in TABLES.items():
        net = ref_import.build_reference_net(288, ref_import.default_opt(**kw))
        meta["tables"][name] = {"opt": kw, "tensors": [[k, list(v.shape)] for k, v in net.state_dict().items()]}
    out["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    path = os.path.join(GOLDEN, "variants.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path))


if __name__ == "__main__":
    main()

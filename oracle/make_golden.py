"""Generate tests/golden/*.npz by running the UNMODIFIED reference (models.py) in the build container.

TEST INFRASTRUCTURE.  /root/reference cannot travel to the GPU box, so its outputs on seeded
inputs are frozen here and committed; the GPU parity tests and oracle.pcn_port are checked
against them.  Re-run with:  python -m oracle.make_golden

Files
-----
weights_seed0.npz      seeded state_dict (genre architecture; the default architecture is the
                       subset without ``genre_classifier.*``), fp32-representable values.
pcn_fwd.npz            B=3 log-CQT clips (T=61), ragged seq_length: reference float64 outputs of
                       PitchClassNet(288,...) default and --genre, eval and train mode, with and
                       without seq_length; the final pitch-class feature map; updated BN buffers.
equivariance.npz       equivariance_test.py:172-205 restated on its --custom_cqt pattern: 25 shifted
                       inputs through PitchClassNet(360,...), key and tonic rows, eval and train mode.
cqt_port.npz           oracle.cqt_port output on a seeded 3 s clip -- a REGRESSION fixture of the
                       restatement itself (CQT parity is unpinned: no librosa here), not a reference output.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import cqt_port, ref_import  # noqa: E402
from audio_key_estimation_b200 import synth  # noqa: E402  (seeded generators only; no CUDA code is touched)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def shift_up(mel: torch.Tensor, s: int) -> torch.Tensor:
    """equivariance_test.py:122-133 mel_shifting_up: rows move up by 3*s bins, zero fill."""
    out = torch.zeros_like(mel)
    if s == 0:
        return mel.clone()
    out[3 * s:] = mel[: mel.shape[0] - 3 * s]
    return out


def shift_down(mel: torch.Tensor, s: int) -> torch.Tensor:
    """equivariance_test.py:135-146 mel_shifting_down."""
    out = torch.zeros_like(mel)
    if s == 0:
        return mel.clone()
    out[: mel.shape[0] - 3 * s] = mel[3 * s:]
    return out


def main() -> None:
    os.makedirs(GOLDEN, exist_ok=True)
    torch.manual_seed(0)
    torch.set_num_threads(max(1, os.cpu_count() or 1))

    # ---------------------------------------------------------------- weights
    net_g = ref_import.build_reference_net(288, ref_import.default_opt(genre=True))
    sd = synth.randomise_state_dict(net_g.state_dict(), seed=0, dtype=torch.float32)
    np.savez(os.path.join(GOLDEN, "weights_seed0.npz"), **{k: v.numpy() for k, v in sd.items()})
    sd64 = {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}

    # ---------------------------------------------------------------- forward goldens
    sr, n = 48000, 48000 * 12
    clips = [synth.synth_clip(100 + i, n, sr).numpy() for i in range(3)]
    mel = np.stack([cqt_port.cqt_logmag(c, sr)[0] for c in clips]).astype(np.float32)  # (3, 288, 61)
    T = mel.shape[-1]
    seq = np.array([T, 50, 37], dtype=np.int64)
    for b, s in enumerate(seq):  # KeyDataset.py:242-254 zero-pads clips beyond their own length
        mel[b, :, s:] = 0.0
    out = {"mel": mel, "seq_length": seq}
    x = torch.from_numpy(mel).double()[:, None]
    for tag, genre in (("default", False), ("genre", True)):
        net = ref_import.build_reference_net(288, ref_import.default_opt(genre=genre))
        sub = {k: v for k, v in sd64.items() if genre or not k.startswith("genre_classifier.")}
        net.load_state_dict(sub, strict=True)
        net.eval()
        with torch.no_grad():
            for stag, sl in (("seq", torch.from_numpy(seq)), ("noseq", None)):
                res = net(x, sl)
                for name, r in zip(("key", "tonic", "genre"), res):
                    out[f"{tag}.eval.{stag}.{name}"] = r.numpy()
            p, pc = net.model((x, None))
            out[f"{tag}.eval.pc_final"] = pc.numpy().astype(np.float32)
            # max_pool quirk (models.py:765-785): honoured for sample 0 only
            net.opt.max_pool = True
            res = net(x, torch.from_numpy(seq))
            for name, r in zip(("key", "tonic", "genre"), res):
                out[f"{tag}.eval.seq_maxpool.{name}"] = r.numpy()
            net.opt.max_pool = False
        net.train()
        with torch.no_grad():
            res = net(x, torch.from_numpy(seq))
        for name, r in zip(("key", "tonic", "genre"), res):
            out[f"{tag}.train.seq.{name}"] = r.numpy()
        new_sd = net.state_dict()
        for k in ("model.1.p2p.layer.7.running_mean", "model.1.p2p.layer.7.running_var",
                  "key_classifier.1.running_mean", "key_classifier.1.running_var",
                  "model.0.pool_semi_b.running_var"):
            out[f"{tag}.train.buf.{k}"] = new_sd[k].numpy()
    np.savez_compressed(os.path.join(GOLDEN, "pcn_fwd.npz"), **out)

    # ---------------------------------------------------------------- equivariance (config 3)
    # Inputs: (a) the --custom_cqt block pattern WITHOUT the border blocks (equivariance_test.py:266-277;
    # the border blocks are pushed out of the frame by the zero-fill shift, which is what that flag is
    # for), (b) a real log-CQT padded with 36 zero rows on both sides as equivariance_test.py:172-176 does.
    pad = torch.zeros(36, T, dtype=torch.float64)
    inputs = {"pattern": synth.custom_cqt_pattern(360, 592, with_border=False),
              "padded_cqt": torch.cat([pad, torch.from_numpy(mel[0]).double(), pad], dim=0)}
    eq = {"padded_cqt.input288": mel[0]}
    sub = {k: v for k, v in sd64.items() if not k.startswith("genre_classifier.")}
    shifts = list(range(0, 13)) + [-i for i in range(1, 13)]
    eq["shifts"] = np.array(shifts)
    for iname, pat in inputs.items():
        for mode in ("eval", "train"):
            net = ref_import.build_reference_net(360, ref_import.default_opt())
            net.load_state_dict(sub, strict=True)
            net.train(mode == "train")  # the reference script leaves the model in train mode (:178)
            keys, tonics = [], []
            with torch.no_grad():
                for s in shifts:
                    m = shift_up(pat, s) if s >= 0 else shift_down(pat, -s)
                    # equivariance_test.py:188: forward(x.reshape(1,1,360,T).double(), tensor(T).reshape(1,1))
                    k, t = net(m.reshape(1, 1, 360, -1).double(), torch.tensor(m.shape[1]).reshape(1, 1))
                    keys.append(k[0].numpy()), tonics.append(t[0].numpy())
            eq[f"{iname}.{mode}.key"], eq[f"{iname}.{mode}.tonic"] = np.stack(keys), np.stack(tonics)
            dev = max(np.abs(keys[i] - np.roll(keys[0], s)).max() for i, s in enumerate(shifts))
            print(f"reference equivariance {iname} {mode}: max |key(s) - roll(key(0), s)| = {dev:.3e}")
    np.savez_compressed(os.path.join(GOLDEN, "equivariance.npz"), **eq)

    # ---------------------------------------------------------------- CQT restatement fixture
    y = synth.synth_clip(7, 48000 * 3, 48000).numpy()
    Cq = cqt_port.cqt(y, sr=48000, hop_length=9600, n_bins=288, bins_per_octave=36)
    np.savez_compressed(os.path.join(GOLDEN, "cqt_port.npz"), clip_id=7, n_samples=48000 * 3,
                        C_re=Cq.real.astype(np.float32), C_im=Cq.imag.astype(np.float32),
                        logmag=np.log1p(np.abs(Cq)).astype(np.float32), audio_head=y[:64])
    for f in sorted(os.listdir(GOLDEN)):
        print(f, os.path.getsize(os.path.join(GOLDEN, f)))


if __name__ == "__main__":
    main()

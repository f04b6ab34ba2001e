"""CPU restatement of ``PitchClassNet.forward`` -- TEST INFRASTRUCTURE ONLY.

Functional torch-CPU port of the reference network (models.py:651-817) driven
purely by a reference-format ``state_dict`` (key names as dumped from the
reference).  It exists because /root/reference cannot
travel to the GPU box; it is pinned against the real reference in the build
container (tests/test_oracle_pcn.py) and against tests/golden/*.npz.

Only the ``train_model.py`` default architecture family is restated (plain
conv stacks; ``resblock/denseblock/stay_sixth/only_semitones/p2pc_conv/
pc2p_mem/local`` are out of scope).

Every function cites the reference lines it follows.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
LRELU_SLOPE = 0.01  # nn.LeakyReLU() default, models.py:197,234,315,...
BN_EPS = 1e-5       # nn.BatchNorm2d default


def _bn_act(x: Tensor, sd: Dict[str, Tensor], prefix: str, train: bool, act: bool = True,
            stats: Optional[dict] = None) -> Tensor:
    """BatchNorm2d (+LeakyReLU).  eval: running stats; train: biased batch stats."""
    w, b = sd[prefix + ".weight"], sd[prefix + ".bias"]
    if train:
        mean = x.mean(dim=(0, 2, 3))
        var = x.var(dim=(0, 2, 3), unbiased=False)
        if stats is not None:
            stats[prefix] = (mean.clone(), var.clone(), x.numel() // x.shape[1])
    else:
        mean, var = sd[prefix + ".running_mean"], sd[prefix + ".running_var"]
    y = (x - mean[None, :, None, None]) / torch.sqrt(var[None, :, None, None] + BN_EPS)
    y = y * w[None, :, None, None] + b[None, :, None, None]
    return F.leaky_relu(y, LRELU_SLOPE) if act else y


def _circ_pad_time(x: Tensor, n: int) -> Tensor:
    return torch.cat([x[..., -n:], x, x[..., :n]], dim=-1)


def semitone_conv(x: Tensor, w: Tensor, b: Tensor) -> Tensor:
    """models.py:313 / 337: Conv2d(C,C,3,stride=(3,1),padding=(0,1),circular).

    Non-overlapping bin triples -> semitones; the time axis wraps circularly."""
    return F.conv2d(_circ_pad_time(x, 1), w, b, stride=(3, 1))


def octave_maxpool(x: Tensor, pitch_classes: int = 12) -> Tensor:
    """models.py:82-106 Pitch2PitchClassPool: pc[c] = max_o x[c + 12 o]."""
    B, C, P, T = x.shape
    n_oct = math.ceil(P / pitch_classes)
    pad = n_oct * pitch_classes - P
    if pad:
        x = torch.cat([x, x.new_full((B, C, pad, T), float("-inf"))], dim=2)
    return x.reshape(B, C, n_oct, pitch_classes, T).amax(dim=2)


def equivariant_conv(x: Tensor, w: Tensor, b: Tensor, same_time: bool) -> Tensor:
    """models.py:22-51: wrap 12 -> 23 rows, Conv2d(Cin,Cout,(12,k)), zero-pad or valid in time."""
    pcs = w.shape[2]
    xw = torch.cat([x, x[:, :, : pcs - 1]], dim=2)
    return F.conv2d(xw, w, b, padding=(0, w.shape[3] // 2 if same_time else 0))


def pitch_conv(x: Tensor, w: Tensor, b: Tensor) -> Tensor:
    """models.py:230-232: Conv2d(k, padding=k//2, padding_mode='circular') (pitch AND time wrap)."""
    k = w.shape[2] // 2
    xp = torch.cat([x[:, :, -k:], x, x[:, :, :k]], dim=2)
    xp = _circ_pad_time(xp, w.shape[3] // 2)
    return F.conv2d(xp, w, b)


def upsample_sixth(pc: Tensor, w: Tensor, b: Tensor) -> Tensor:
    """models.py:325: ConvTranspose2d(C,C,(3,1),stride=(3,1)): out[co,3c+r] = b + sum_ci W[ci,co,r] pc[ci,c]."""
    return F.conv_transpose2d(pc, w, b, stride=(3, 1))


def tile_to_pitch(x: Tensor, pitches: int) -> Tensor:
    """models.py:135-143 PitchClass2Pitch."""
    reps = math.ceil(pitches / x.shape[2])
    return x.repeat(1, 1, reps, 1)[:, :, :pitches]


def _stack_indices(sd: Dict[str, Tensor], prefix: str, conv_suffix: str) -> List[int]:
    idx = sorted({int(k[len(prefix):].split(".")[0]) for k in sd
                  if k.startswith(prefix) and k.endswith(conv_suffix)})
    return idx


def pcn_forward(sd: Dict[str, Tensor], mel: Tensor, seq_length: Optional[Tensor] = None, *,
                train: bool = False, time_pool_size: int = 2, max_pool: bool = False,
                taps: Optional[dict] = None, stats: Optional[dict] = None
                ) -> Tuple[Tensor, ...]:
    """Restates models.py:747-817 (forward) over models.py:352-399 (layers).

    ``sd`` tensors must share ``mel``'s dtype.  ``taps`` (optional dict) receives
    named intermediates for layer-by-layer parity checks.  ``stats`` receives the
    train-mode batch statistics per BN site."""
    pitches = mel.shape[2]
    num_layers = 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("model."))
    genre = any(k.startswith("genre_classifier.") for k in sd)

    def tap(name, t):
        if taps is not None:
            taps[name] = t.detach().clone()

    p, pc = mel, None
    for L in range(num_layers):
        pre = f"model.{L}."
        if L == 0:
            # models.py:359-369
            s = semitone_conv(p, sd[pre + "pool_semi.weight"], sd[pre + "pool_semi.bias"])
            s = _bn_act(s, sd, pre + "pool_semi_b", train, stats=stats)
            tap("l0.semi", s)
            pc = octave_maxpool(s)
            tap("l0.pool", pc)
        else:
            # models.py:370-393
            u = upsample_sixth(pc, sd[pre + "up_sixth.weight"], sd[pre + "up_sixth.bias"])
            u = _bn_act(u, sd, pre + "up_sixth_b", train, stats=stats)
            tap(f"l{L}.up", u)
            p = torch.cat([p, tile_to_pitch(u, pitches)], dim=1)
            for i in _stack_indices(sd, pre + "p2p.layer.", ".weight"):
                if sd[pre + f"p2p.layer.{i}.weight"].dim() != 4:
                    continue  # BN weight
                p = pitch_conv(p, sd[pre + f"p2p.layer.{i}.weight"], sd[pre + f"p2p.layer.{i}.bias"])
                p = _bn_act(p, sd, pre + f"p2p.layer.{i + 1}", train, stats=stats)
                tap(f"l{L}.p2p{i // 3}", p)
            s = semitone_conv(p, sd[pre + "pool_semi.weight"], sd[pre + "pool_semi.bias"])
            s = _bn_act(s, sd, pre + "pool_semi_b", train, stats=stats)
            pc2 = octave_maxpool(s)
            tap(f"l{L}.pool", pc2)
            pc = torch.cat([pc, pc2], dim=1)
        for i in _stack_indices(sd, pre + "pc2pc.layer.", ".conv2d.weight"):
            pc = equivariant_conv(pc, sd[pre + f"pc2pc.layer.{i}.conv2d.weight"],
                                  sd[pre + f"pc2pc.layer.{i}.conv2d.bias"], same_time=True)
            pc = _bn_act(pc, sd, pre + f"pc2pc.layer.{i + 1}", train, stats=stats)
            tap(f"l{L}.pc2pc{i // 3}", pc)
        if L > 0:
            # models.py:394-396 (floor pooling)
            p = F.max_pool2d(p, (1, time_pool_size))
            pc = F.max_pool2d(pc, (1, time_pool_size))
    tap("pc_final", pc)

    def head(name: str) -> Tensor:
        x = pc
        idx = _stack_indices(sd, name + ".", ".conv2d.weight")
        for n, i in enumerate(idx):
            x = equivariant_conv(x, sd[f"{name}.{i}.conv2d.weight"], sd[f"{name}.{i}.conv2d.bias"],
                                 same_time=False)
            if n != len(idx) - 1:
                x = _bn_act(x, sd, f"{name}.{i + 1}", train, stats=stats)
        return x

    tonic, key = head("tonic_classifier"), head("key_classifier")
    n_head = len(_stack_indices(sd, "tonic_classifier.", ".conv2d.weight"))
    k = sd["tonic_classifier.0.conv2d.weight"].shape[3]
    g = None
    if genre:
        # models.py:724,733: plain Conv2d (1,k) [+BN+LReLU] ... Conv2d (2,k)
        x = pc
        idx = sorted({int(kk.split(".")[1]) for kk in sd
                      if kk.startswith("genre_classifier.") and kk.endswith(".weight")
                      and sd[kk].dim() == 4})
        for n, i in enumerate(idx):
            x = F.conv2d(x, sd[f"genre_classifier.{i}.weight"], sd[f"genre_classifier.{i}.bias"])
            if n != len(idx) - 1:
                x = _bn_act(x, sd, f"genre_classifier.{i + 1}", train, stats=stats)
        g = x
    tap("tonic_frames", tonic)
    tap("key_frames", key)

    def reduce(x: Tensor) -> Tensor:
        if seq_length is None:
            # models.py:786-797
            return x.amax(dim=-1) if max_pool else x.mean(dim=-1)
        # models.py:757-785: per-sample masked mean; max_pool honoured for sample 0 only
        L = seq_length.reshape(-1).to(torch.float64)
        for _ in range(num_layers - 1):
            L = torch.floor(L / time_pool_size)
        L = L.to(torch.int32) - (k - 1) * n_head
        if L.numel() == 1 and x.shape[0] > 1:
            L = L.expand(x.shape[0])
        rows = []
        for j in range(x.shape[0]):
            xs = x[j, :, :, : int(L[j])]
            rows.append(xs.amax(dim=-1) if (max_pool and j == 0) else xs.mean(dim=-1))
        return torch.stack(rows)

    tonic_out = reduce(tonic).flatten(1)
    key_out = torch.sigmoid(reduce(key).flatten(1))
    if genre:
        return key_out, tonic_out, reduce(g).flatten(1)
    return key_out, tonic_out


# --- key decode (models.py:1083-1085 with the table utils/key_signatures.py:19-42) -----------
# 21 rows x 12 pitch classes [C, C#, D, ..., B]; 15 practical + 6 theoretical signatures.
KEY_SIGNATURE_ROWS = (
    "010110101011", "010101101011", "110101101010", "110101011010", "101101011010",
    "101101010110", "101011010110", "101011010101", "101010110101", "011010110101",
    "011010101101", "010110101101", "010110101011", "010101101011", "110101101010",
    "011010110101", "010110101101", "011010101101", "101101011010", "110101011010",
    "101101010110",
)


def key_signature_map(dtype=torch.float32) -> Tensor:
    return torch.tensor([[int(c) for c in r] for r in KEY_SIGNATURE_ROWS], dtype=dtype)


def decode(key_out: Tensor, tonic_out: Tensor, genre_out: Optional[Tensor] = None):
    """pred_key_id = argmax_r cos(key_out, MAP[r]) (models.py:1083-1085); tonic/genre argmax (1096, 923)."""
    m = key_signature_map(key_out.dtype)
    cos = F.cosine_similarity(key_out[:, None, :], m[None], dim=2)
    out = [cos.argmax(dim=1), tonic_out.argmax(dim=1)]
    if genre_out is not None:
        out.append(genre_out.argmax(dim=1))
    return tuple(out)


# --- MIREX-weighted key score (models.py:1065-1116, mirex_score) -----------------------------------
MIREX_COUNTERS = ("samples", "correct", "fifths", "relative", "parallel", "other", "all_keys", "tonics", "key_bits")


def mirex_counters(key_out: Tensor, tonic_out: Tensor, key_labels: Tensor, tonic_labels: Tensor, key_signature_id: Tensor):
    """Per-batch category counts of the reference's per-sample loop (models.py:1070-1112) + per-clip cosine similarity
    (models.py:1093).  Returns (dict of ints keyed by MIREX_COUNTERS, similarity (B,))."""
    m = key_signature_map(key_out.dtype)
    cnt = dict.fromkeys(MIREX_COUNTERS, 0)
    sims = []
    for i in range(key_out.shape[0]):
        pred_id = int(F.cosine_similarity(key_out[i][None], m, dim=1).argmax())        # :1083-1084
        key_pred = m[pred_id]                                                           # :1085
        label_id = int(key_signature_id[i].argmax())                                   # :1086
        correct_keys = int((key_pred == key_labels[i]).sum())                          # :1090
        sims.append(F.cosine_similarity(key_out[i], key_labels[i], dim=0))             # :1094
        diff = abs(pred_id - label_id)                                                 # :1095
        correct_tonic = int(tonic_labels[i].argmax() == tonic_out[i].argmax())         # :1096
        cnt["samples"] += 1
        cnt["key_bits"] += correct_keys
        cnt["all_keys"] += int(correct_keys == 12)
        cnt["tonics"] += correct_tonic
        if diff == 1 and not (correct_tonic == 1 and correct_keys == 12):              # :1100-1111, first match wins
            cnt["fifths"] += 1
        elif correct_tonic == 1 and correct_keys == 12:
            cnt["correct"] += 1
        elif correct_keys == 12 and correct_tonic == 0:
            cnt["relative"] += 1
        elif correct_tonic == 1 and correct_keys != 12:
            cnt["parallel"] += 1
        else:
            cnt["other"] += 1
    return cnt, torch.stack(sims)


def mirex_from_counters(cnt) -> Tuple[float, ...]:
    """(mirex, correct, fifths, relative, parallel, other, accuracy) as models.py:1113-1115 returns them."""
    n = max(1, cnt["samples"])
    mirex = 1.0 * cnt["correct"] + 0.5 * cnt["fifths"] + 0.3 * cnt["relative"] + 0.2 * cnt["parallel"]
    return (mirex / n, cnt["correct"] / n, cnt["fifths"] / n, cnt["relative"] / n, cnt["parallel"] / n, cnt["other"] / n,
            cnt["all_keys"] / n)


def count_macs(sd: Dict[str, Tensor], pitches: int, T: int, time_pool_size: int = 2) -> int:
    """Algorithmic MACs of one clip's forward (out_elems * Cin * kh * kw per conv)."""
    total = 0
    num_layers = 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("model."))
    t = T
    for L in range(num_layers):
        pre = f"model.{L}."
        for k, w in sd.items():
            if not k.startswith(pre) or w.dim() != 4:
                continue
            co, ci, kh, kw = w.shape
            if "pool_semi" in k:
                total += (pitches // 3) * t * co * ci * kh * kw
            elif "up_sixth" in k:
                total += 36 * t * w.shape[0] * w.shape[1] * kh * kw // 3
            elif "p2p" in k:
                total += pitches * t * co * ci * kh * kw
            elif "pc2pc" in k:
                total += 12 * t * co * ci * kh * kw
        if L > 0:
            t //= time_pool_size
    for name in ("tonic_classifier", "key_classifier", "genre_classifier"):
        tt = t
        for k in sorted(k for k in sd if k.startswith(name) and sd[k].dim() == 4):
            co, ci, kh, kw = sd[k].shape
            tt = tt - kw + 1
            rows = 12 if "conv2d" in k else 12 - kh + 1
            total += rows * tt * co * ci * kh * kw
    return total

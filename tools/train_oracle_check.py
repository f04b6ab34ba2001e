"""dev tool (GPU box): training-step gradients against the float64 oracle port for a list of BxT shapes (prints the worst tensors)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import audio_key_estimation_b200 as ake
from oracle import pcn_port

for arg in sys.argv[1:]:
    B, T = (int(v) for v in arg.split("x"))
    torch.manual_seed(11)
    net = ake.PitchClassNet(288, 12, 2, 7, opt=ake.default_opt(genre=True))
    for m in net.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.weight.data.uniform_(0.5, 1.5), m.bias.data.uniform_(-0.3, 0.3)
    sd64 = {k: v.detach().clone().double().requires_grad_("running_" not in k) for k, v in net.state_dict().items() if v.is_floating_point()}
    g = torch.Generator().manual_seed(5)
    mel = torch.log1p(torch.rand((B, 1, 288, T), generator=g) * 4)
    seq = torch.randint(T // 2, T + 1, (B,), generator=g)
    key = (torch.rand((B, 12), generator=g) < 0.6).float()
    tonic = torch.nn.functional.one_hot(torch.randint(0, 12, (B,), generator=g), 12)
    genre = torch.nn.functional.one_hot(torch.randint(0, 11, (B,), generator=g), 11)
    ostats = {}
    out = pcn_port.pcn_forward(sd64, mel.double(), seq, train=True, stats=ostats)
    want = ake.criterion(out, key.double(), tonic, genre)
    want.backward()
    net = net.cuda().train()
    old = {k: v.detach().clone().cpu() for k, v in net.state_dict().items() if 'running_' in k}
    step = ake.TrainStep(net)
    res = step.step(mel.cuda(), seq.cuda(), key.cuda(), tonic.cuda(), genre.cuda())
    floor = 1e-5 * max(float(v.grad.abs().max()) for v in sd64.values() if v.grad is not None)
    errs = []
    for name, prm in net.named_parameters():
        ref = sd64[name].grad.numpy()
        got = prm.grad.detach().cpu().numpy()
        err, scale = np.abs(got - ref).max(), np.abs(ref).max()
        errs.append((float(max(0.0, err - floor) / max(scale, 1e-30)), name))
    errs.sort(reverse=True)
    print(f"{B}x{T}: loss err {abs(res['loss'].item() - want.item()) / abs(want.item()):.1e}; worst:", [(f"{e:.1e}", n) for e, n in errs[:3]], flush=True)
    new = {k: v.detach().cpu() for k, v in net.state_dict().items() if 'running_' in k}
    bad = []
    for prefix, (m, v, n) in ostats.items():
        bm = (new[prefix + ".running_mean"] - 0.9 * old[prefix + ".running_mean"]) / 0.1
        bv = (new[prefix + ".running_var"] - 0.9 * old[prefix + ".running_var"]) / 0.1 * (n - 1) / n
        em = float((bm.double() - m).abs().max() / (m.abs().max() + v.sqrt().max()))
        ev = float((bv.double() - v).abs().max() / v.abs().max())
        bad.append((max(em, ev), prefix, em, ev))
    bad.sort(reverse=True)
    print("   BN batch statistics vs oracle, worst sites:", [(f"{a:.1e}", pfx) for a, pfx, _, _ in bad[:4]], flush=True)

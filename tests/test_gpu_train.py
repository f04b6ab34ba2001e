"""GPU parity of the training step (BASELINE config 5) against the reference's own
general_step + loss.backward() frozen in tests/golden/train_step.npz.
Tolerance: loss 1e-5 relative; every parameter gradient max-abs <= 5e-4 of that tensor's max |grad| (+ a floor for the
tensors whose gradient is exactly zero): fp32 arithmetic (7x7 convolutions: fp16 hi/lo three-product tensor-core operands,
22 bits) with atomically accumulated weight gradients against a float64 reference; measured worst case 6e-5."""
import numpy as np
import pytest
import torch

import audio_key_estimation_b200 as ake
from audio_key_estimation_b200 import distributed as akd
from conftest import golden_state_dict, load_golden

pytestmark = pytest.mark.gpu


def _setup(genre, fwd_golden):
    g = load_golden("train_step.npz")
    net = ake.PitchClassNet(288, 12, 2, 7, opt=ake.default_opt(genre=genre))
    net.load_state_dict(golden_state_dict(genre), strict=True)
    net = net.cuda().train()
    mel = torch.from_numpy(fwd_golden["mel"])[:, None].cuda()
    seq = torch.from_numpy(fwd_golden["seq_length"]).cuda()
    B = mel.shape[0]
    key_labels = torch.from_numpy(g["key_labels"]).cuda()
    tonic_1h = torch.nn.functional.one_hot(torch.from_numpy(g["tonic_idx"]), 12).cuda()
    gi = torch.from_numpy(g["genre_idx"])
    genre_1h = torch.zeros(B, 11, dtype=torch.long)
    genre_1h[gi >= 0, gi[gi >= 0]] = 1
    return g, net, mel, seq, key_labels, tonic_1h, genre_1h.cuda()


def _check_grads(net, g, tag, tol=5e-4):
    worst = 0.0
    n = 0
    # conv biases in front of a train-mode BatchNorm have an exactly-zero gradient (reference: ~1e-15): the floor is set by
    # fp32 cancellation relative to the largest gradient of the network
    floor = 1e-5 * max(np.abs(g[k]).max() for k in g.files if k.startswith(f"{tag}.grad."))
    for name, prm in net.named_parameters():
        ref = g[f"{tag}.grad.{name}"]
        got = prm.grad.detach().cpu().numpy()
        assert got.shape == ref.shape, name
        err = np.abs(got - ref).max()
        scale = np.abs(ref).max()
        assert err <= tol * scale + floor, f"{name}: max-abs error {err:.3e} vs max |grad| {scale:.3e}"
        if scale > 100 * floor:  # (tensors whose gradient is exactly zero have no relative error)
            worst = max(worst, err / scale)
        n += 1
    return n, worst


@pytest.mark.parametrize("tag", ["default", "genre"])
def test_fused_train_step_matches_reference(tag, fwd_golden):
    genre = tag == "genre"
    g, net, mel, seq, key_labels, tonic_1h, genre_1h = _setup(genre, fwd_golden)
    step = ake.TrainStep(net)
    res = step.step(mel, seq, key_labels, tonic_1h, genre_1h if genre else None)
    want = float(g[f"{tag}.loss"])
    assert abs(res["loss"].item() - want) <= 1e-5 * abs(want)
    n, worst = _check_grads(net, g, tag)
    print(f"[{tag}] worst gradient error / max |grad| over {n} tensors: {worst:.3e}")
    assert n == (66 if genre else 60)
    # the flat buffer is what a data-parallel job all-reduces; world size 1 leaves it untouched
    before = step.flat_grads.clone()
    akd.allreduce_gradients(step.flat_grads)
    assert torch.equal(before, step.flat_grads)
    # running statistics were updated like nn.BatchNorm2d does in train mode
    assert int(net.model._modules["0"].pool_semi_b.num_batches_tracked) == 1


def test_autograd_path_matches_reference(fwd_golden):
    """forward() in train mode + torch loss + loss.backward(): the route the reference's Lightning loop takes."""
    g, net, mel, seq, key_labels, tonic_1h, genre_1h = _setup(True, fwd_golden)
    out = net(mel.double(), seq)
    assert out[0].dtype == torch.float64 and out[0].requires_grad
    loss = ake.criterion(out, key_labels.double(), tonic_1h, genre_1h, net.opt)
    assert abs(loss.item() - float(g["genre.loss"])) <= 1e-5 * abs(float(g["genre.loss"]))
    loss.backward()
    _check_grads(net, g, "genre")
    # an optimizer step with the reference's settings (models.py:1017-1027) runs on those gradients
    opt = torch.optim.Adam(net.parameters(), lr=3e-4, betas=(0.9, 0.999))
    w0 = net.model._modules["1"].p2p.layer._modules["0"].weight.detach().clone()
    opt.step()
    assert not torch.equal(w0, net.model._modules["1"].p2p.layer._modules["0"].weight)
    out2 = net(mel, seq)  # the changed parameters are re-uploaded
    assert torch.isfinite(out2[0]).all()


def test_unsupported_training_configs_fail_loudly(fwd_golden):
    net = ake.PitchClassNet(288, 12, 2, 7, opt=ake.default_opt(max_pool=True)).cuda().train()
    mel = torch.from_numpy(fwd_golden["mel"])[:, None].cuda()
    with pytest.raises((NotImplementedError, ValueError)):
        net(mel, None)


def test_fused_adam_matches_torch_adam():
    """models.py:1017-1027: Adam(betas=(0.9, 0.999), lr, weight_decay=reg) + ExponentialLR(gamma).  The reference's optimizer
    IS torch.optim.Adam, so the fused one-launch step is checked against it on the same gradients for several steps/epochs."""
    import audio_key_estimation_b200 as ake

    torch.manual_seed(3)
    opt = ake.default_opt(genre=True)
    opt.lr, opt.reg, opt.gamma = 3e-3, 1e-2, 0.9
    net = ake.PitchClassNet(288, 12, 2, 7, opt=opt).cuda().train()
    ref = ake.PitchClassNet(288, 12, 2, 7, opt=opt).cuda().train()
    ref.load_state_dict(net.state_dict(), strict=True)
    ref_params = ref._grad_params()
    adam = torch.optim.Adam(ref_params, betas=(0.9, 0.999), lr=opt.lr, weight_decay=opt.reg)
    sched = torch.optim.lr_scheduler.ExponentialLR(adam, gamma=opt.gamma)
    fused = ake.FusedAdam(net)
    n_flat = sum(net._lookup(n).numel() for n in net._tensor_names)
    for step in range(6):
        flat = torch.randn(n_flat, device="cuda") * 0.1
        for p, g in zip(ref_params, ref._split_flat_grads(flat)):
            p.grad = g.clone()
        adam.step()
        fused.step(flat)
        if step % 2 == 1:
            sched.step()
            fused.epoch_end()
    worst = 0.0
    for a, b in zip(net._grad_params(), ref_params):
        worst = max(worst, (a.detach() - b.detach()).abs().max().item() / max(1e-3, b.detach().abs().max().item()))
    assert worst <= 2e-6, worst  # same formula in fp32; only the rounding of a few fused multiply-adds differs
    # buffers (running statistics) are not parameters: untouched
    for (n1, b1), (n2, b2) in zip(net.named_buffers(), ref.named_buffers()):
        assert torch.equal(b1, b2), n1
    # the next forward sees the updated parameters (the plan re-uploads them)
    mel = torch.rand(2, 1, 288, 40, device="cuda")
    net.eval(), ref.eval()
    with torch.no_grad():
        o1, o2 = net(mel, None), ref(mel, None)
    for x, y in zip(o1, o2):
        assert (x - y).abs().max().item() <= 1e-4


def test_graph_replay_equals_eager_step():
    """TrainStep(graph=True): forward + loss + backward captured once per shape and replayed; gradients, loss and
    BatchNorm statistics must equal the eager launches (same kernels, same order; gradients to the rounding of their atomic
    accumulation), also after the inputs and the parameters change between replays."""
    import audio_key_estimation_b200 as ake

    torch.manual_seed(1)
    opt = ake.default_opt(genre=True)
    nets = []
    for _ in range(2):
        n = ake.PitchClassNet(288, 12, 2, 7, opt=opt).cuda().train()
        nets.append(n)
    nets[1].load_state_dict(nets[0].state_dict(), strict=True)
    eager, graph = ake.TrainStep(nets[0]), ake.TrainStep(nets[1], graph=True)
    B, T = 4, 61
    for it in range(3):
        mel = torch.rand(B, 1, 288, T, device="cuda") * 3
        seq = torch.randint(40, T + 1, (B,), device="cuda")
        keyl = (torch.rand(B, 12, device="cuda") < 0.6).float()
        tonic = torch.nn.functional.one_hot(torch.randint(0, 12, (B,)), 12).cuda()
        genre = torch.nn.functional.one_hot(torch.randint(0, 11, (B,)), 11).cuda()
        r0 = eager.step(mel, seq, keyl, tonic, genre)
        r1 = graph.step(mel, seq, keyl, tonic, genre)
        assert torch.equal(r0["loss"], r1["loss"]), it  # the forward is deterministic
        # the weight-gradient kernels accumulate with fp32 atomics: two runs agree to rounding, not bit for bit
        scale = eager.flat_grads.abs().max().item()
        assert (eager.flat_grads - graph.flat_grads).abs().max().item() <= 1e-5 * scale, it
        for (n0, b0), (n1, b1) in zip(nets[0].named_buffers(), nets[1].named_buffers()):
            assert torch.equal(b0, b1), n0
        with torch.no_grad():  # an optimizer step between replays: the graph must see the new parameters
            for p0, p1 in zip(nets[0].parameters(), nets[1].parameters()):
                upd = 0.01 * torch.sign(p0.grad)  # the same update on both copies
                p0.add_(upd)
                p1.add_(upd)
    assert len(graph._graphs) == 1


TOL_OTHER_SHAPES = 2e-2


@pytest.mark.parametrize("B,T", [(3, 57), (2, 200)])
def test_train_step_matches_oracle_at_other_shapes(B, T):
    """The training step against the float64 oracle port (oracle/pcn_port.py + torch autograd: the restatement that the golden tests
    pin to the unmodified reference) at shapes the goldens do not cover: a frame count that is not a multiple of 16 (zero-padded K
    blocks of the weight-gradient GEMM, odd time tiles), two time tiles per clip in the 7x7 kernels, ragged masks, random BatchNorm
    affine parameters.  Loss to 1e-5.  Gradients to 2e-2 of each tensor's largest entry, NOT the goldens' 5e-4: on random data the
    network's max-pools (octave, time) and LeakyReLU kinks make the gradient discontinuous in the activations -- a 1e-6 rounding
    difference in a conv output moves a pooling winner and with it O(1e-3) of a gradient tensor (measured: up to 8e-3 at T = 176 / 200
    for the tensor-core path AND 1e-3 for the fp32 FFMA path, exactly 0 beyond the floor at T = 61 / 64 / 73 / 151 / 160 for both;
    tools/train_oracle_check.py).  A wrong tap, halo or scale shows up as O(1)."""
    from oracle import pcn_port

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
    out = pcn_port.pcn_forward(sd64, mel.double(), seq, train=True)
    want_loss = ake.criterion(out, key.double(), tonic, genre)
    want_loss.backward()

    net = net.cuda().train()
    step = ake.TrainStep(net)
    res = step.step(mel.cuda(), seq.cuda(), key.cuda(), tonic.cuda(), genre.cuda())
    assert abs(res["loss"].item() - want_loss.item()) <= 1e-5 * abs(want_loss.item())
    floor = 1e-5 * max(float(v.grad.abs().max()) for v in sd64.values() if v.grad is not None)
    errs = []
    for name, prm in net.named_parameters():
        ref = sd64[name].grad.numpy()
        got = prm.grad.detach().cpu().numpy()
        err, scale = np.abs(got - ref).max(), np.abs(ref).max()
        errs.append((float(max(0.0, err - floor) / max(scale, 1e-30)), name, float(err), float(scale)))
    errs.sort(reverse=True)
    print(f"[B={B} T={T}] worst gradient errors / max |grad| vs the float64 oracle:", [(f"{e:.2e}", n) for e, n, _, _ in errs[:5]])
    assert errs[0][0] <= TOL_OTHER_SHAPES, errs[:5]

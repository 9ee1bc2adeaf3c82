"""EXTENSIONS -- NO REFERENCE PARITY. The north star's features that the reference does not have (command-conditioned
branched heads with CE / L1 / MSE loss, crop + colour jitter + normalise in the staging kernel, folded BatchNorm) against
their specification in oracle/ext_oracle.py (written in the reference's style; see its header)."""
import numpy as np
import pytest
import torch

from oracle import bc_oracle as O
from oracle import ext_oracle as X

pytestmark = pytest.mark.gpu


def _dev():
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device (no CPU fallback exists to run instead)")
    return torch.device("cuda", 0)


def _rel(got, ref):
    got, ref = got.detach().double().cpu(), ref.detach().double().cpu()
    return float((got - ref).abs().max() / ref.abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("kind,n_out,G", [("ce", 9, 4), ("l1", 3, 4), ("mse", 3, 3), ("l1", 3, 1)])
def test_branched_heads_match_the_extension_oracle(kind, n_out, G):
    """Outputs of the commanded branch, loss and the gradients of every branch parameter (rel 1e-5, f32 kernels), the
    trunk's conv gradients (through the routing-dependent backward: 5e-3), and the branch-select mask: a branch no sample
    of the batch was sent to receives exactly zero gradient."""
    from src.architectures.branched import ConvNet1Branched
    dev = _dev()
    hp = {"obs_size": 4, "n_actions": n_out, "n_branches": G, "branch_loss": kind}
    torch.manual_seed(12345)
    net = ConvNet1Branched(hp).to(dev)
    P = X.init_branched_params(12345, 4, n_out, G)
    sd = {k: v.detach().cpu() for k, v in net.state_dict().items()}
    assert list(sd) == list(P) and all(torch.equal(sd[k], P[k]) for k in P)
    B = 7
    frames, labels = O.synth_frames(40 + G, B + 4)
    x, y = O.sequential_samples(frames, labels)
    x = torch.from_numpy(x)
    gen = torch.Generator().manual_seed(5)
    command = torch.randint(0, max(G - 1, 1), (B,), generator=gen)           # the last branch (G > 1) gets no sample
    target = torch.from_numpy(y) if kind == "ce" else torch.randn((B, n_out), generator=gen)
    loss = net.loss(x.to(dev), command.to(dev), target.to(dev))
    loss.backward()
    torch.cuda.synchronize()
    net.engine().check_device_errors()
    ref_loss, ref_out, ref = X.branched_loss_and_grads(P, x, command, target, G, kind)
    out = net(x.to(dev), command.to(dev))
    assert _rel(out, ref_out) <= 1e-5
    assert abs(float(loss) - float(ref_loss)) <= 1e-5 * abs(float(ref_loss))
    got = {k: p.grad for k, p in net.named_parameters()}
    for k in ref:
        tol = 1e-5 if k.startswith("branches.") else 5e-3
        if float(ref[k].abs().max()) == 0.0:
            assert float(got[k].abs().max()) == 0.0, k                        # branch-select mask: untouched branch
        else:
            assert _rel(got[k], ref[k]) <= tol, (k, _rel(got[k], ref[k]))
    if G > 1:
        assert float(got[f"branches.{G - 1}.0.weight"].abs().max()) == 0.0


def test_branched_module_trains_and_both_arenas_move():
    """ImitationBranched (the Imitation hook contract) + MultiArenaAdam: 5 steps of L1 regression lower the loss on a fixed
    batch; the trunk arena and the branch arena both change; the unused trunk head does not; an out-of-range command is
    reported through the device flag."""
    from src.architectures.branched import ConvNet1Branched
    from src.models.imitation_branched import ImitationBranched
    dev = _dev()
    hp = {"obs_size": 4, "n_actions": 3, "n_branches": 4, "branch_loss": "l1", "precision": "bf16"}
    torch.manual_seed(12345)
    net = ConvNet1Branched(hp).to(dev)
    model = ImitationBranched(hp, net, {})
    opt = model.configure_optimizers()[0][0]
    frames, _ = O.synth_frames(3, 20)
    x = torch.from_numpy(O.sequential_samples(frames, np.zeros(20, np.int64))[0]).to(dev)
    gen = torch.Generator().manual_seed(1)
    command = torch.randint(0, 4, (16,), generator=gen).to(dev)
    target = torch.rand((16, 3), generator=gen).to(dev)
    trunk0, br0 = net._arena.clone(), net._barena.clone()
    fc_lo = min(p._bc_offset for p in net._ordered_params[8:])
    fc_hi = max(p._bc_offset + p.numel() for p in net._ordered_params[8:])
    losses = []
    for i in range(5):
        loss = model.training_step((x, command, target), i)
        opt.zero_grad()
        loss.backward()
        opt.step()
        losses.append(float(loss))
    torch.cuda.synchronize()
    net.engine().check_device_errors()
    assert losses[-1] < losses[0]
    assert not torch.equal(net._barena, br0) and not torch.equal(net._arena, trunk0)
    assert torch.equal(net._arena[fc_lo:fc_hi], trunk0[fc_lo:fc_hi])        # the trunk's own (unregistered) head is never touched
    bad = command.clone()
    bad[3] = 4
    net.loss(x, bad, target)
    with pytest.raises(RuntimeError):
        net.engine().check_device_errors()


@pytest.mark.parametrize("src_hw", [(256, 256), (288, 320)])
def test_staging_augmentation_is_bit_exact_against_the_specification(src_hw):
    """crop + brightness / contrast / saturation jitter + normalise inside the staging kernel == the f32 numpy specification
    bit for bit (every step one rounded f32 operation, same order), from a host-seeded per-frame table; the Toeplitz-ready
    bf16 planes are the bf16 rounding of the same values; identity parameters reproduce the reference's gray conversion."""
    from carla_imitation_learning_b200 import stage_augmented, stage_gray
    from carla_imitation_learning_b200.data import augment_table
    from tests.test_gpu_tc import _tp_reference
    dev = _dev()
    n = 6
    rng = np.random.Generator(np.random.PCG64(11))
    frames = rng.integers(0, 256, size=(n, src_hw[0], src_hw[1], 3), dtype=np.uint8)
    table = augment_table(7, n, src_hw, mean=0.45, std=0.22)
    assert np.array_equal(table, X.augment_params(7, n, src_hw, mean=0.45, std=0.22))       # the product's host table == the specification's
    ref = X.stage_augmented(frames, table)
    fr = torch.from_numpy(frames).to(dev)
    got = stage_augmented(fr, torch.from_numpy(table), layout="plain")
    assert np.array_equal(got.cpu().numpy().view(np.uint32), ref.view(np.uint32))
    tp = stage_augmented(fr, torch.from_numpy(table), layout="tp")
    want = _tp_reference(torch.from_numpy(ref).to(torch.bfloat16))
    assert torch.equal(tp.tp.cpu().view(torch.int16), want.view(torch.int16).reshape(n, -1))
    if src_hw == (256, 256):
        ident = np.zeros((n, 8), np.float32)
        ident[:, 2:5] = 1.0
        ident[:, 6] = 1.0
        g_id = stage_augmented(fr, torch.from_numpy(ident), layout="plain").cpu().numpy()
        g_ref = stage_gray(fr).cpu().numpy()
        assert np.abs(g_id - g_ref).max() <= 2.5e-7                             # same conversion, f32 chain instead of the exact quotient


def test_folded_batchnorm_equals_conv_bn_relu_pool():
    """conv -> BatchNorm2d(eval) -> ReLU -> pool through the fused conv kernels after ConvNet1.fold_batchnorm == torch's
    conv2d + batch_norm(eval) + relu + max_pool2d on the CPU (f32 kernels, rel 1e-5), also for negative BN scales."""
    from src.architectures.nets import ConvNet1
    dev = _dev()
    torch.manual_seed(12345)
    net = ConvNet1({"obs_size": 4, "n_actions": 9}).to(dev)
    P = {k: v.detach().cpu().double() for k, v in net.state_dict().items()}
    gen = torch.Generator().manual_seed(9)
    bn = []
    for li, c in enumerate(O.CONV_CHANNELS):
        gamma = torch.randn(c, generator=gen)                                  # includes negative scales
        beta, mean, var = 0.1 * torch.randn(c, generator=gen), 0.1 * torch.randn(c, generator=gen), 0.5 + torch.rand(c, generator=gen)
        bn.append((gamma, beta, mean, var))
        net.fold_batchnorm(li, gamma, beta, mean, var)
    frames, labels = O.synth_frames(8, 9)
    x = torch.from_numpy(O.sequential_samples(frames, labels)[0])
    got = net(x.to(dev)).detach().cpu()
    h = x.double()
    F = torch.nn.functional
    for li, (name, k, s, p) in enumerate(O.CONV_SPECS):
        gamma, beta, mean, var = (t.double() for t in bn[li])
        h = F.conv2d(h, P[f"{name}.weight"], P[f"{name}.bias"], stride=s)
        h = F.batch_norm(h, mean, var, gamma, beta, training=False, eps=1e-5)
        h = F.max_pool2d(F.relu(h), p)
        w2, b2 = X.fold_batchnorm(P[f"{name}.weight"], P[f"{name}.bias"], gamma, beta, mean, var)
        assert _rel(net.state_dict()[f"{name}.weight"], w2) <= 1e-6 and _rel(net.state_dict()[f"{name}.bias"], b2) <= 1e-5
    h = torch.flatten(h, 1)
    for i, (name, _fi, _fo) in enumerate(O.FC_SPECS):
        h = F.linear(h, P[f"{name}.weight"], P[f"{name}.bias"])
        if i < 2:
            h = F.relu(h)
    assert _rel(got, h) <= 2e-5

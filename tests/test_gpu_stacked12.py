"""BASELINE configs[3] on the tensor cores: obs_size 12 = 3 cameras x 4 frames stacked (channel = 3*frame + camera), the input a
zero-copy window view advancing by 3 planes. In bf16 mode conv1 runs as three 4-frame camera streams accumulated into one result
(csrc/conv1_tc.cu: conv1_tp_kernel<ADD_IN, RAW_OUT>), its weight gradient as three launches of conv1_wgrad3_kernel over the
cameras' planes. Reference: /root/reference/src/architectures/nets.py:11-20 (obs_size is the conv's input channel count)."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from oracle import bc_oracle as O

pytestmark = pytest.mark.gpu
REL_BF16 = 2e-2
HP = {"obs_size": 12, "n_actions": 9, "precision": "bf16"}


def _dev():
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device (no CPU fallback exists to run instead)")
    return torch.device("cuda", 0)


def _net(hp=HP, seed=7):
    from src.architectures.nets import ConvNet1
    torch.manual_seed(seed)
    return ConvNet1(dict(hp)).to(_dev())


def _frames(seed, B):
    """3 cameras interleaved frame by frame: 3 * (B + 4) frames give B samples of 12 planes advancing by 3 (+ the frames
    that would carry the last sample's label, as in the 4-frame window of imitation_dataset.py:117-131)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    frames = rng.integers(0, 256, size=(3 * (B + 4), 256, 256, 3), dtype=np.uint8)
    frames[::5, 40:90] //= 3                      # some structure: not every plane has the same statistics
    return frames, rng.integers(0, 9, size=B, dtype=np.int64)


def _rel(got, ref):
    got, ref = got.detach().double().cpu(), ref.detach().double().cpu()
    return float((got - ref).abs().max() / ref.abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("B", [1, 5, 9])
def test_conv1_obs12_on_tensor_cores_matches_bf16_rounded_oracle(B):
    """conv1 + ReLU + pool of the 12-channel network == f64 conv of the bf16-ROUNDED planes and weights (rel 1e-5, routing checked),
    for the stacked window view and for the same batch materialised (stride 12 planes)."""
    from carla_imitation_learning_b200 import _lib, sliding_window, stage_frames, stage_gray
    from tests.test_gpu_parity import _check_routing
    dev = _dev()
    net = _net()
    eng = net.engine()
    frames, _ = _frames(20 + B, B)
    fr = torch.from_numpy(frames).to(dev)
    staged = stage_frames(fr, frame_skip=12, step=3)
    assert staged.shape == (B, 12, 256, 256)
    s = torch.cuda.current_stream().cuda_stream
    res = []
    for x in (staged, sliding_window(stage_gray(fr, dtype=torch.bfloat16), 12, 3).contiguous()):
        bufs = eng.alloc(B, eng.check_input(x), None, False)
        assert B == 1 or (bufs.x_tp_strides[0] == 3 * bufs.x_tp_strides[1]) == (x is staged)     # (a one-sample batch is its own window)
        c = eng.ctx(bufs)
        _lib.check(eng.lib.bc_conv_relu_pool_fwd(C.byref(c), 0, s), "conv1 tc obs12")
        torch.cuda.synchronize()
        eng.check_device_errors()
        res.append((bufs.act[0].clone(), bufs.amax[0].clone(), bufs.act_bf16[0].clone()))
    gray = stage_gray(fr, dtype=torch.bfloat16)
    xw = sliding_window(gray, 12, 3).double().cpu()
    w = net.cnn_base[0].weight.detach().to(torch.bfloat16).double().cpu()
    bias = net.cnn_base[0].bias.detach().double().cpu()
    z = torch.nn.functional.conv2d(xw, w, bias, stride=3)
    ref = torch.nn.functional.max_pool2d(torch.relu(z), 3)
    for act, amax, _p8 in res:
        assert _rel(act, ref) <= 1e-5
        _check_routing(z, amax.cpu(), act.cpu().double(), 3)
    assert _rel(res[1][0], res[0][0]) <= 1e-6
    # the bf16 P8 copy conv2 reads = the f32 activation rounded, [b][c/8][pixel][8]
    p8 = res[0][0].reshape(B, 2, 8, 784).permute(0, 1, 3, 2).to(torch.bfloat16)
    assert torch.equal(p8, res[0][2])


def test_bf16_whole_step_obs12_matches_oracle_given_device_decisions():
    """The whole 12-channel step in bf16 mode (all convolutions on tcgen05): pooled activations, logits, loss and ALL 14 gradients
    against the f64 oracle on the f32 master weights, given the device's pool routing and ReLU masks, per tensor <= 2e-2."""
    from carla_imitation_learning_b200 import stage_frames
    from tests.test_gpu_parity import _check_routing
    dev = _dev()
    B = 6
    net = _net()
    eng = net.engine()
    frames, labels = _frames(3, B)
    y = torch.from_numpy(labels)
    bufs = eng.train_forward_backward(stage_frames(torch.from_numpy(frames).to(dev), frame_skip=12, step=3), y.to(dev))
    torch.cuda.synchronize()
    eng.check_device_errors()
    got = {k: eng.grads[p._bc_offset:p._bc_offset + p.numel()].view(p.shape).detach().cpu() for k, p in net.named_parameters()}
    P = {k: v.detach().cpu() for k, v in net.state_dict().items()}
    gray = O.gray_stack(frames)                                   # (n, 256, 256) f32, the reference's formula
    xb = torch.from_numpy(np.stack([gray[3 * i:3 * i + 12] for i in range(B)]))
    amax = [a.cpu().long() for a in bufs.amax]
    relu = {"pooled": [a.cpu() > 0 for a in bufs.act], "fc": [bufs.hid1.cpu() > 0, bufs.hid2.cpu() > 0]}
    torch.set_num_threads(os.cpu_count() or 1)
    loss, logits, ref, aux = O.explicit_backward(P, xb, y, dtype=torch.float64, argmax_override=amax, relu_override=relu)
    for li in range(4):
        assert _rel(bufs.act[li], aux["pooled"][li]) <= REL_BF16, (li, _rel(bufs.act[li], aux["pooled"][li]))
        _check_routing(aux["conv_out"][li], bufs.amax[li].cpu(), aux["pooled"][li], O.CONV_SPECS[li][3], tol=REL_BF16)
    assert _rel(bufs.logits, logits) <= REL_BF16
    assert abs(float(bufs.loss) - float(loss)) <= REL_BF16 * float(loss)
    worst = {k: _rel(got[k], ref[k]) for k in O.PARAM_ORDER}
    print("obs 12, bf16 step vs f64 oracle given routing + ReLU masks:", {k: f"{v:.2e}" for k, v in worst.items()})
    assert max(worst.values()) <= REL_BF16, worst
    # the same batch in the reference's own format, a materialised (B,12,256,256) f32 tensor: planes are not shared, the weight
    # gradient runs on the second-generation kernel (three camera launches as well) -- same loss and gradients
    from carla_imitation_learning_b200 import sliding_window, stage_gray
    g_view, l_view = eng.grads.clone(), float(bufs.loss)
    xm = sliding_window(stage_gray(torch.from_numpy(frames).to(dev)), 12, 3).contiguous()
    bm = eng.train_forward_backward(xm, y.to(dev))
    torch.cuda.synchronize()
    eng.check_device_errors()
    assert bm.x_tp_strides[0] == 12 * bm.x_tp_strides[1]
    assert abs(float(bm.loss) - l_view) <= 1e-6 * l_view
    assert _rel(eng.grads, g_view) <= 1e-4


def test_obs12_adam_keeps_the_three_camera_images_current_and_modes_agree():
    """FusedAdam's operand refresh covers the three per-camera Toeplitz images (after 2 steps the images == a fresh pack of the
    final weights, bitwise), and the bf16 step's loss is within 2e-2 of the exact-f32 kernels' on the same batch."""
    from carla_imitation_learning_b200 import FusedAdam, sliding_window, stage_frames, stage_gray
    dev = _dev()
    B = 4
    frames, labels = _frames(9, B)
    fr, y = torch.from_numpy(frames).to(dev), torch.from_numpy(labels).to(dev)
    net = _net()
    eng = net.engine()
    opt = FusedAdam(list(net.parameters()), lr=1e-3)
    opt.prepare()
    x = stage_frames(fr, frame_skip=12, step=3)
    b = eng.alloc(B, x, y, True)
    l16 = None
    for i in range(2):
        eng.enqueue_train(b)
        if i == 0:
            l16 = float(b.loss)
        opt.step_flat(eng.grads)
    torch.cuda.synchronize()
    eng.check_device_errors()
    kept = eng.w_packed.clone()
    eng.pack_weights()
    torch.cuda.synchronize()
    assert torch.equal(kept, eng.w_packed)
    net32 = _net(dict(HP, precision="fp32"))
    b32 = net32.engine().train_forward_backward(sliding_window(stage_gray(fr), 12, 3), y)
    torch.cuda.synchronize()
    assert abs(l16 - float(b32.loss)) <= REL_BF16 * float(b32.loss)

"""Which tcgen05 piece keeps the bf16 run on the ln 9 plateau ON THE GOLDEN DATA STREAM? 300 steps at B=8 per conv_mode mask."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from carla_imitation_learning_b200 import stage_gray, FusedAdam
from oracle import bc_oracle as O
from src.architectures.nets import ConvNet1
dev = torch.device("cuda", 0)
B, steps = 8, int(os.environ.get("STEPS", "300"))
frames, labels = O.synth_frames(11, 1000 * B + 4)
frames, labels = frames[:steps * B + 4], labels[:steps * B + 4]
lab = torch.from_numpy(labels).to(dev)
fr = torch.from_numpy(frames).to(dev)
gray16 = stage_gray(fr, dtype=torch.bfloat16)
gray32 = stage_gray(fr, dtype=torch.float32)


def run(name, mask, gray):
    torch.manual_seed(12345)
    net = ConvNet1({"obs_size": 4, "n_actions": 9}).to(dev)
    opt = FusedAdam(list(net.parameters()), lr=1e-3)
    eng = net.engine()
    eng.conv_mode = mask
    L = []
    alive = []
    for s in range(steps):
        x = gray.as_strided((B, 4, 256, 256), (65536, 65536, 256, 1), s * B * 65536)
        y = lab[s * B + 4: s * B + 4 + B]
        if mask:
            eng.pack_weights()
        b = eng.forward(x, y, backward=True)
        eng.backward(b)
        for p in net.parameters():
            p.grad = eng.grads[p._bc_offset:p._bc_offset + p.numel()].view(p.shape).clone()
        opt.step()
        L.append(float(b.loss))
        if s % 30 == 29:
            alive.append(f"{float((b.act[3] > 0).float().mean()):.2f}/{float((b.act[0] > 0).float().mean()):.2f}")
    eng.check_device_errors()
    print(f"{name:44s}", " ".join(f"{np.mean(L[i:i+30]):.3f}" for i in range(0, steps, 30)), "| alive conv4/conv1:", " ".join(alive), flush=True)


run("f32 kernels, f32 planes", 0, gray32)
run("f32 kernels, bf16 planes", 0, gray16)
run("mask 1  (tc forward only)", 1, gray16)
run("mask 3  (+ tc dgrad)", 3, gray16)
run("mask 5  (fwd + tc wgrad conv2-4)", 5, gray16)
run("mask 9  (fwd + tc wgrad conv1)", 9, gray16)
run("mask 15 (all)", 15, gray16)

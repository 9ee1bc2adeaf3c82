"""Along the exact-f32 trajectory on the golden stream, how far are the tensor-core gradients from the f32 ones?
Every 10 steps the same batch is differentiated with each conv_mode mask at the current weights."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from carla_imitation_learning_b200 import stage_gray, FusedAdam
from oracle import bc_oracle as O
from src.architectures.nets import ConvNet1
dev = torch.device("cuda", 0)
B, steps = 8, int(os.environ.get("STEPS", "160"))
frames, labels = O.synth_frames(11, 1000 * B + 4)
frames, labels = frames[:steps * B + 4], labels[:steps * B + 4]
lab = torch.from_numpy(labels).to(dev)
fr = torch.from_numpy(frames).to(dev)
gray16 = stage_gray(fr, dtype=torch.bfloat16)
gray32 = stage_gray(fr, dtype=torch.float32)
torch.manual_seed(12345)
net = ConvNet1({"obs_size": 4, "n_actions": 9}).to(dev)
opt = FusedAdam(list(net.parameters()), lr=1e-3)
eng = net.engine()
names = [n for n, _ in net.named_parameters()]
params = dict(net.named_parameters())


def grads(mask, gray, s):
    eng.conv_mode = mask
    x = gray.as_strided((B, 4, 256, 256), (65536, 65536, 256, 1), s * B * 65536)
    y = lab[s * B + 4: s * B + 4 + B]
    if mask:
        eng.pack_weights()
    b = eng.forward(x, y, backward=True)
    eng.backward(b)
    torch.cuda.synchronize()
    return eng.grads.clone().double(), float(b.loss)


cos = lambda a, b: float((a * b).sum() / (a.norm() * b.norm() + 1e-300))
for s in range(steps):
    g0, loss = grads(0, gray32, s)
    if s % 10 == 9 or s < 3:
        line = [f"step {s:3d} loss {loss:.4f}"]
        for mask, gray in ((0, gray16), (1, gray16), (3, gray16), (5, gray16), (9, gray16), (15, gray16)):
            g, _ = grads(mask, gray, s)
            worst = ("", 1.0)
            for n in names:
                p = params[n]
                a, b = g[p._bc_offset:p._bc_offset + p.numel()], g0[p._bc_offset:p._bc_offset + p.numel()]
                c = cos(a, b)
                if c < worst[1]:
                    worst = (n, c)
            line.append(f"m{mask}: cos {cos(g, g0):.5f} worst {worst[0]} {worst[1]:.4f}")
        print(" | ".join(line), flush=True)
        g0, loss = grads(0, gray32, s)      # restore the exact gradient in eng.grads
    for p in net.parameters():
        p.grad = eng.grads[p._bc_offset:p._bc_offset + p.numel()].view(p.shape).clone()
    opt.step()
eng.check_device_errors()

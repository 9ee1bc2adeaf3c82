import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from carla_imitation_learning_b200 import stage_gray, FusedAdam
from oracle import bc_oracle as O
from src.architectures.nets import ConvNet1
dev = torch.device("cuda", 0)
B, steps = 8, 300
frames, labels = O.synth_frames(11, steps * B + 4)
lab = torch.from_numpy(labels).to(dev)
fr = torch.from_numpy(frames).to(dev)
gray = stage_gray(fr, dtype=torch.bfloat16)
def run(name, variant):
    torch.manual_seed(12345)
    net = ConvNet1({"obs_size": 4, "n_actions": 9, "precision": "bf16"}).to(dev)
    opt = FusedAdam(list(net.parameters()), lr=1e-3)
    L = []
    for s in range(steps):
        x = gray.as_strided((B, 4, 256, 256), (65536, 65536, 256, 1), s * B * 65536)
        y = lab[s * B + 4: s * B + 4 + B]
        if variant in ("A", "B"):
            loss = net.loss(x, y)
            opt.zero_grad(); loss.backward()
            if variant == "A":
                opt.step()
            else:
                opt.step_flat(net._engine.grads)
            L.append(float(loss.detach()))
        else:
            eng = net.engine()
            b = eng.forward(x, y, backward=True)
            eng.backward(b)
            for p in net.parameters():
                p.grad = eng.grads[p._bc_offset:p._bc_offset + p.numel()].view(p.shape).clone()
            opt.step()
            L.append(float(b.loss))
    print(f"{name:40s}", " ".join(f"{np.mean(L[i:i+30]):.3f}" for i in range(0, steps, 30)))
run("A module fwd/bwd + opt.step()", "A")
run("B module fwd/bwd + step_flat(eng.grads)", "B")
run("C engine fwd/bwd + opt.step()", "C")

import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from carla_imitation_learning_b200 import stage_gray
from oracle import bc_oracle as O
from src.architectures.nets import ConvNet1
from src.models.imitation import Imitation
dev = torch.device("cuda", 0)
B, steps = 8, 260
frames, labels = O.synth_frames(11, steps * B + 4)
lab = torch.from_numpy(labels).to(dev)
out = {}
for prec in ("fp32", "bf16"):
    torch.manual_seed(12345)
    net = ConvNet1({"obs_size": 4, "n_actions": 9, "precision": prec}).to(dev)
    model = Imitation({}, net, {})
    opt = model.configure_optimizers()[0][0]
    gray = stage_gray(torch.from_numpy(frames).to(dev), dtype=torch.bfloat16 if prec == "bf16" else torch.float32)
    L, G = [], []
    for s in range(steps):
        x = gray.as_strided((B, 4, 256, 256), (65536, 65536, 256, 1), s * B * 65536)
        loss = model.training_step((x, lab[s * B + 4: s * B + 4 + B]), s)
        opt.zero_grad(); loss.backward()
        G.append(torch.cat([p.grad.reshape(-1) for p in net.parameters()]).clone())
        opt.step()
        L.append(float(loss.detach()))
    net.engine().check_device_errors()
    out[prec] = (L, G, torch.cat([p.detach().reshape(-1) for p in net.parameters()]).clone())
for s in (0, 1, 2, 5, 10, 20, 50, 100, 150, 200, 259):
    a, b = out["bf16"][1][s].double(), out["fp32"][1][s].double()
    print(s, "loss bf16 %.5f fp32 %.5f  grad cos %.4f norm %.3e/%.3e" % (out["bf16"][0][s], out["fp32"][0][s], float((a*b).sum()/(a.norm()*b.norm())), float(a.norm()), float(b.norm())))
print("param diff", float((out["bf16"][2]-out["fp32"][2]).abs().max()))

import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from carla_imitation_learning_b200 import stage_gray, sliding_window
from oracle import bc_oracle as O
from src.architectures.nets import ConvNet1
dev = torch.device("cuda", 0)
frames, labels = O.synth_frames(77, 36)
y = torch.from_numpy(labels[4:36]).to(dev)
fr = torch.from_numpy(frames).to(dev)
res = {}
for prec in ("fp32", "bf16"):
    torch.manual_seed(12345)
    net = ConvNet1({"obs_size": 4, "n_actions": 9, "precision": prec}).to(dev)
    x = sliding_window(stage_gray(fr, dtype=torch.bfloat16 if prec == "bf16" else torch.float32))
    # module path
    loss = net.loss(x, y); loss.backward()
    res[prec, "mod"] = {k: p.grad.detach().clone() for k, p in net.named_parameters()}
    # engine path
    eng = net.engine()
    b = eng.train_forward_backward(x, y)
    torch.cuda.synchronize()
    res[prec, "eng"] = {k: eng.grads[p._bc_offset:p._bc_offset + p.numel()].view(p.shape).clone() for k, p in net.named_parameters()}
    print(prec, "loss", float(loss), float(b.loss))
rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
cos = lambda a, b: float((a.double() * b.double()).sum() / (a.double().norm() * b.double().norm() + 1e-30))
for k in res["fp32", "eng"]:
    a, b = res["bf16", "eng"][k], res["fp32", "eng"][k]
    print(f"{k:20s} cos {cos(a, b):.5f}  norm ratio {float(a.norm() / b.norm()):.4f}")
for k in res["fp32", "eng"]:
    print(f"{k:20s} bf16eng/fp32eng {rel(res['bf16','eng'][k], res['fp32','eng'][k]):.3e}   bf16mod/fp32eng {rel(res['bf16','mod'][k], res['fp32','eng'][k]):.3e}   fp32mod/fp32eng {rel(res['fp32','mod'][k], res['fp32','eng'][k]):.3e}   |g|max {float(res['fp32','eng'][k].abs().max()):.3e}")

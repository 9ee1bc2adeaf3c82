import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from carla_imitation_learning_b200 import stage_gray, FusedAdam
from oracle import bc_oracle as O
from src.architectures.nets import ConvNet1
from src.models.imitation import Imitation
dev = torch.device("cuda", 0)
B, steps = 8, 300
frames, labels = O.synth_frames(11, steps * B + 4)
lab = torch.from_numpy(labels).to(dev)
fr = torch.from_numpy(frames).to(dev)
def run(name, prec, how):
    gray = stage_gray(fr, dtype=torch.bfloat16 if prec == "bf16" else torch.float32)
    torch.manual_seed(12345)
    net = ConvNet1({"obs_size": 4, "n_actions": 9, "precision": prec}).to(dev)
    model = Imitation({}, net, {})
    if how == "cfg":
        opt = model.configure_optimizers()[0][0]
    elif how == "direct_modelparams":
        opt = FusedAdam(list(model.parameters()), lr=1e-3)
    else:
        opt = FusedAdam(list(net.parameters()), lr=1e-3)
    L = []
    for s in range(steps):
        x = gray.as_strided((B, 4, 256, 256), (65536, 65536, 256, 1), s * B * 65536)
        y = lab[s * B + 4: s * B + 4 + B]
        loss = model.training_step((x, y), s)
        opt.zero_grad(); loss.backward(); opt.step()
        L.append(float(loss.detach()))
    print(f"{name:40s}", " ".join(f"{np.mean(L[i:i+30]):.3f}" for i in range(0, steps, 30)), "lr", opt.param_groups[0]["lr"], type(opt.param_groups[0]["lr"]))
run("bf16 Imitation + configure_optimizers", "bf16", "cfg")
run("bf16 Imitation + FusedAdam(model.params)", "bf16", "direct_modelparams")
run("bf16 Imitation + FusedAdam(net.params)", "bf16", "direct")
run("fp32 Imitation + configure_optimizers", "fp32", "cfg")

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
keep = []
for prec in ("fp32", "bf16"):
    torch.manual_seed(12345)
    net = ConvNet1({"obs_size": 4, "n_actions": 9, "precision": prec}).to(dev)
    x = sliding_window(stage_gray(fr, dtype=torch.bfloat16 if prec == "bf16" else torch.float32))
    eng = net.engine()
    print(prec, "conv_mode", eng.conv_mode, "x dtype", x.dtype)
    b = eng.train_forward_backward(x, y)
    torch.cuda.synchronize()
    eng.check_device_errors()
    keep.append(b)
    print("  act sums", [float(a.double().sum()) for a in b.act], "act_bf16", [float(a.double().sum()) for a in b.act_bf16], "loss", float(b.loss), "logits0", b.logits[0, :3].tolist())

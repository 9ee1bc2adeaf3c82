"""Serving forward at small batch, eagerly (for `ncu --metrics gpu__time_duration.sum`): staging + conv1 + conv2 + the one-launch tail,
then the layer-by-layer path, B = 1 and 8."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from carla_imitation_learning_b200 import stage_frames
from src.architectures.nets import ConvNet1
dev = torch.device("cuda", 0)
torch.manual_seed(12345)
net = ConvNet1({"obs_size": 4, "n_actions": 9, "precision": "bf16"}).to(dev)
eng = net.engine()
rng = np.random.Generator(np.random.PCG64(0))
for B in (1, 8):
    fr = torch.from_numpy(rng.integers(0, 256, size=(B + 4, 256, 256, 3), dtype=np.uint8)).to(dev)
    staged = stage_frames(fr)
    bufs = eng.alloc(B, staged, None, False)
    out = torch.empty(B, dtype=torch.int64, device=dev)
    for tail in (True, False):
        for _ in range(3):
            stage_frames(fr, out=staged)
            eng.forward_act(staged, out=out, bufs=bufs, tail=tail)
        torch.cuda.synchronize()
        print("B", B, "tail", tail, out.tolist(), flush=True)
eng.check_device_errors()

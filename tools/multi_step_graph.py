"""Timing experiment: one CUDA graph per step (trainer.TrainStep as benched) against ONE graph holding a whole cycle of NBUF steps
(programmatic dependent launch then also links Adam(i) -> stage(i+1), and the device sees NBUF times fewer graph launches)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from carla_imitation_learning_b200 import FusedAdam
from carla_imitation_learning_b200.trainer import TrainStep
from src.architectures.nets import ConvNet1
dev = torch.device("cuda", 0)
B, NBUF, STEPS = 256, 4, 400
rng = np.random.Generator(np.random.PCG64(0))
frames = [torch.from_numpy(rng.integers(0, 256, size=(B + 4, 256, 256, 3), dtype=np.uint8)).to(dev) for _ in range(NBUF)]
labels = [torch.from_numpy(rng.integers(0, 9, size=B)).to(dev) for _ in range(NBUF)]

def make():
    torch.manual_seed(12345)
    net = ConvNet1({"obs_size": 4, "n_actions": 9, "precision": "bf16"}).to(dev)
    opt = FusedAdam(list(net.parameters()), lr=1e-3)
    return net, opt, TrainStep(net, opt, B)

def timed(fn, n):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

net, opt, ts = make()
for i in range(3 * NBUF):
    ts.step(frames[i % NBUF], labels[i % NBUF])
torch.cuda.synchronize()
state = {"i": 0}
def one():
    i = state["i"]; state["i"] += 1
    ts.step(frames[i % NBUF], labels[i % NBUF])
ms1 = timed(one, STEPS)
print(f"one graph per step: {ms1 * 1e3:.2f} us/step")

net2, opt2, ts2 = make()
for i in range(2 * NBUF):
    ts2._enqueue(frames[i % NBUF], labels[i % NBUF])
torch.cuda.synchronize()
opt2.prepare()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for i in range(NBUF):
        ts2._enqueue(frames[i], labels[i])
def cycle():
    g.replay()
ms4 = timed(cycle, STEPS // NBUF) / NBUF
print(f"one graph per {NBUF} steps: {ms4 * 1e3:.2f} us/step")
ts2.check()

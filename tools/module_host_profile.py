"""Host-side cost of one module-path step (Imitation.training_step -> zero_grad -> backward -> FusedAdam.step): cProfile over
N steps with the GPU running asynchronously. Shows where the Python time goes once the device step is shorter than the host loop."""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from carla_imitation_learning_b200 import StagedBatch, _lib, stage_frames
from src.architectures.nets import ConvNet1
from src.models.imitation import Imitation
dev = torch.device("cuda", 0)
B = 256
hp = {"obs_size": 4, "n_actions": 9, "precision": "bf16", "cuda_graph": True}
torch.manual_seed(12345)
net = ConvNet1(hp).to(dev)
model = Imitation(hp, net, {})
opt = model.configure_optimizers()[0][0]
rng = np.random.Generator(np.random.PCG64(0))
frames = [torch.from_numpy(rng.integers(0, 256, size=(B + 4, 256, 256, 3), dtype=np.uint8)).to(dev) for _ in range(4)]
labels = [torch.from_numpy(rng.integers(0, 9, size=B)).to(dev) for _ in range(4)]
stg = [StagedBatch(torch.empty((B + 4, _lib.TP_PLANE_ELEMS), dtype=torch.bfloat16, device=dev), None, 4) for _ in range(2)]

def step(i):
    x = stage_frames(frames[i % 4], out=stg[i & 1])
    loss = model.training_step((x, labels[i % 4]), i)
    opt.zero_grad()
    loss.backward()
    opt.step()

for i in range(20):
    step(i)
torch.cuda.synchronize()
N = 2000
t0 = time.perf_counter()
for i in range(N):
    step(i)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host loop {1e6 * (t1 - t0) / N:.1f} us/step, with the final sync {1e6 * (t2 - t0) / N:.1f} us/step")
pr = cProfile.Profile()
pr.enable()
for i in range(N):
    step(i)
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(22)

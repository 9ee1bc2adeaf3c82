"""BASELINE configs[4]: inference-only policy forward + greedy action, batch 1..4096, bf16 tensor-core mode, one B200.
Per batch size: staging (u8 RGB -> Toeplitz-ready planes) + conv1..4 + head + argmax captured as one CUDA graph;
latency p50 / p99 over 200 replays (CUDA events), frames/s = B / p50. Prints one JSON line per batch size."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from carla_imitation_learning_b200 import stage_frames
from src.architectures.nets import ConvNet1

dev = torch.device("cuda", 0)
torch.manual_seed(12345)
net = ConvNet1({"obs_size": 4, "n_actions": 9, "precision": "bf16"}).to(dev)
eng = net.engine()
rng = np.random.Generator(np.random.PCG64(0))
for B in (1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096):
    frames = torch.from_numpy(rng.integers(0, 256, size=(B + 4, 256, 256, 3), dtype=np.uint8)).to(dev)
    staged = stage_frames(frames)
    bufs = eng.alloc(B, staged, None, False)
    actions = torch.empty(B, dtype=torch.int64, device=dev)

    def step():
        import ctypes as C
        from carla_imitation_learning_b200 import _lib
        stage_frames(frames, out=staged)
        c = eng.ctx(bufs)
        s = torch.cuda.current_stream().cuda_stream
        _lib.check(eng.lib.bc_forward(C.byref(c), s))
        _lib.check(eng.lib.bc_argmax(bufs.logits.data_ptr(), actions.data_ptr(), B, 9, s))

    side = torch.cuda.Stream(dev)
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            step()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        step()
    for _ in range(10):
        g.replay()
    torch.cuda.synchronize()
    ts = []
    for _ in range(200):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts = np.sort(np.asarray(ts))
    p50, p99 = float(ts[len(ts) // 2]), float(ts[int(len(ts) * 0.99) - 1])
    print(json.dumps({"batch": B, "latency_us_p50": round(p50, 1), "latency_us_p99": round(p99, 1), "frames_per_s": round(B / (p50 * 1e-6))}), flush=True)
eng.check_device_errors()

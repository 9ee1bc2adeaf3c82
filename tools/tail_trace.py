"""Timeline of CTAs of the one-launch serving tail (policy_tail_kernel; needs a library built with BC_NVCC_EXTRA=-DBC_TRACE).
    python tools/tail_trace.py [batch]  -> per traced CTA and warp: 0 CTA start, 1 weights in shared memory, 2 predecessor complete, 3 act2 in
    shared memory, 4 conv3 partials done, 5 cluster start barrier passed, 6 act3 exchanged (cluster barrier), 7 conv4 done, 8 features
    exchanged (cluster barrier), 9 head + argmax done; cycles since the CTA's start. The step is replayed from a CUDA graph
    (staging + conv1 + conv2 + tail under programmatic dependent launch), as the serving loop runs it."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from carla_imitation_learning_b200 import _lib, stage_frames
from src.architectures.nets import ConvNet1
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
dev = torch.device("cuda", 0)
torch.manual_seed(12345)
net = ConvNet1({"obs_size": 4, "n_actions": 9, "precision": "bf16"}).to(dev)
eng = net.engine()
rng = np.random.Generator(np.random.PCG64(0))
fr = torch.from_numpy(rng.integers(0, 256, size=(B + 4, 256, 256, 3), dtype=np.uint8)).to(dev)
staged = stage_frames(fr)
bufs = eng.alloc(B, staged, None, False)
out = torch.empty(B, dtype=torch.int64, device=dev)
l = C.CDLL(_lib.LIB_PATH)
l.bc_debug_tail_trace.argtypes = [C.c_void_p, C.c_int]


def enqueue():
    stage_frames(fr, out=staged)
    eng.forward_act(staged, out=out, bufs=bufs, tail=True)


side = torch.cuda.Stream(dev)
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    enqueue()
torch.cuda.current_stream().wait_stream(side)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    enqueue()
for cta in (0, 1, 7):
    l.bc_debug_tail_trace(None, cta)
    for _ in range(5):
        g.replay()
    torch.cuda.synchronize()
    buf = np.zeros((20, 1024), np.uint64)
    l.bc_debug_tail_trace(buf.ctypes.data, cta)
    for w in (0, 3, 7):
        ev = []
        for v in buf[w]:
            v = int(v)
            if v == 0xFFFFFFFFFFFFFFFF:
                break
            ev.append((v >> 56, v & 0xffffffffff))
        print(f"B {B} cta {cta} warp {w}: " + " ".join(f"{e}@{t}" for e, t in ev))

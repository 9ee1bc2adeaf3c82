"""Timeline of one CTA of the swapped-role conv1 weight gradient (needs a library built with BC_NVCC_EXTRA=-DBC_TRACE).
    python tools/c1wg3_trace.py [cta]   -> per warp the event list: 0 CTA start, 10 predecessor complete, 1 plane load issued (job), 3 issuer has its operands (job), 4 issuer done issuing (job),
    5 builder asks for a slot (build), 6 slot free, 7 build published, 8 builders done, 9 all MMAs complete, 11 warp done."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from carla_imitation_learning_b200 import _lib, stage_frames
from src.architectures.nets import ConvNet1
cta = int(sys.argv[1]) if len(sys.argv) > 1 else 0
dev = torch.device("cuda", 0)
torch.manual_seed(12345)
net = ConvNet1({"obs_size": 4, "n_actions": 9, "precision": "bf16"}).to(dev)
eng = net.engine()
B = 256
rng = np.random.Generator(np.random.PCG64(0))
fr = torch.from_numpy(rng.integers(0, 256, size=(B + 4, 256, 256, 3), dtype=np.uint8)).to(dev)
y = torch.from_numpy(rng.integers(0, 9, size=B)).to(dev)
bufs = eng.train_forward_backward(stage_frames(fr), y)
c = eng.ctx(bufs)
s = torch.cuda.current_stream().cuda_stream
l = C.CDLL(_lib.LIB_PATH)
l.bc_debug_c1wg3_trace.argtypes = [C.c_void_p, C.c_int]
l.bc_debug_c1wg3_trace(None, cta)
for _ in range(3):
    eng.lib.bc_conv_bwd_wgrad(C.byref(c), 0, s)
torch.cuda.synchronize()
out = np.zeros((20, 1024), np.uint64)
l.bc_debug_c1wg3_trace(out.ctypes.data, cta)
for w in range(20):
    ev = []
    for v in out[w]:
        v = int(v)
        if v == 0xFFFFFFFFFFFFFFFF:
            break
        ev.append((v >> 56, (v >> 40) & 0xffff, v & 0xffffffffff))
    if ev:
        print(f"warp {w}: " + " ".join(f"{e}:{i}@{t}" for e, i, t in ev))

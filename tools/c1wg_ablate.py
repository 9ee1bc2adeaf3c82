"""Ablation of the conv1 wgrad kernels (needs a library built with BC_NVCC_EXTRA=-DBC_ABLATE; results are wrong by construction).
    BC_C1WG_ABLATE=<bits> python tools/c1wg_ablate.py      bits: 1 no MMAs, 2 no plane loads, 4 no gradient stores, 8 no gradient loads, 16 no epilogue
Prints the kernel's time at B=256 (CUDA events, 20 launches)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from carla_imitation_learning_b200 import _lib, stage_frames
from src.architectures.nets import ConvNet1
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
for name, fn in (("conv1_wgrad", lambda: eng.lib.bc_conv_bwd_wgrad(C.byref(c), 0, s)), ("conv1_fwd", lambda: eng.lib.bc_conv_relu_pool_fwd(C.byref(c), 0, s))):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print(f"ablate {os.environ.get('BC_C1WG_ABLATE', '0'):>2s} gen {os.environ.get('BC_C1WG_GEN', '3')} {name}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us", flush=True)
eng.err_flag.zero_()

"""Time the conv1 forward kernels (tcgen05 bf16 vs exact f32) for several batch sizes."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from carla_imitation_learning_b200 import _lib, sliding_window
from src.architectures.nets import ConvNet1
dev = torch.device("cuda", 0)
torch.manual_seed(12345)
net = ConvNet1({"obs_size": 4, "n_actions": 9}).to(dev)
eng = net.engine()
for mode, dt in (("bf16", torch.bfloat16), ("fp32", torch.float32)):
    eng.set_mode(mode)
    eng.pack_weights()
    for B in (1, 11, 64, 256, 512, 1024):
        gray = torch.rand((B + 4, 256, 256), device=dev).to(dt)
        bufs = eng.alloc(B, sliding_window(gray), None, False)
        c = eng.ctx(bufs)
        s = torch.cuda.current_stream().cuda_stream
        for _ in range(3):
            _lib.check(eng.lib.bc_conv_relu_pool_fwd(C.byref(c), 0, s))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            _lib.check(eng.lib.bc_conv_relu_pool_fwd(C.byref(c), 0, s))
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print(f"{mode} B={B:5d}: {ms * 1e3:8.1f} us  ({B * 14} tiles, {ms * 1e3 / max(1, (B * 14 + 147) // 148):6.2f} us per tile-wave)  {44255232 * B / ms / 1e9:7.1f} TFLOP/s useful")
    eng.check_device_errors()

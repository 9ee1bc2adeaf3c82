"""Where does the time of conv1_wgrad_tp_kernel / conv1_tp_kernel go? Event-timed at B=256 with BC_C1WG_ABLATE / BC_C1FW_ABLATE
= 0 (full), 1 (no MMAs), 2 (no plane loads), 3 (neither: builders / epilogue + barriers only). One process per setting."""
import os, subprocess, sys
if len(sys.argv) > 1:
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import ctypes as C, numpy as np, torch
    from carla_imitation_learning_b200 import _lib, stage_frames
    from src.architectures.nets import ConvNet1
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    net = ConvNet1({"obs_size": 4, "n_actions": 9, "precision": "bf16"}).to(dev)
    eng = net.engine()
    rng = np.random.Generator(np.random.PCG64(0))
    B = 256
    x = stage_frames(torch.from_numpy(rng.integers(0, 256, size=(B + 4, 256, 256, 3), dtype=np.uint8)).to(dev))
    y = torch.from_numpy(rng.integers(0, 9, size=B)).to(dev)
    bufs = eng.train_forward_backward(x, y)
    c = eng.ctx(bufs); s = torch.cuda.current_stream().cuda_stream
    for _ in range(5):
        _lib.check(eng.lib.bc_conv_bwd_wgrad(C.byref(c), 0, s))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        _lib.check(eng.lib.bc_conv_bwd_wgrad(C.byref(c), 0, s))
    e1.record(); torch.cuda.synchronize()
    t_w = e0.elapsed_time(e1) / 20 * 1e3
    for _ in range(5):
        _lib.check(eng.lib.bc_conv_relu_pool_fwd(C.byref(c), 0, s))
    e0.record()
    for _ in range(20):
        _lib.check(eng.lib.bc_conv_relu_pool_fwd(C.byref(c), 0, s))
    e1.record(); torch.cuda.synchronize()
    print(f"ablate {os.environ.get('BC_C1WG_ABLATE', '0')}: wgrad {t_w:.1f} us, forward {e0.elapsed_time(e1) / 20 * 1e3:.1f} us", flush=True)
else:
    for a in ("0", "1", "2", "3"):
        subprocess.run([sys.executable, __file__, "run"], env={**os.environ, "BC_C1WG_ABLATE": a, "BC_C1FW_ABLATE": a})

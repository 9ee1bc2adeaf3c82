"""Timing experiment: the staging kernel of step i+1 (HBM-bound: 51 MB of u8 frames in, 45 MB of planes out) on a side stream
UNDER the kernels of step i (double-buffered planes), against the serial step. One CUDA graph holds a cycle of NBUF steps so
that the next step's frames are known at capture time. Variants: where the side stream forks (after conv layer k's forward of
step i, k = 0: at the step's start) and whether the capture stream has a higher priority than the staging stream."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from carla_imitation_learning_b200 import FusedAdam, _lib, stage_frames
from carla_imitation_learning_b200.engine import StagedBatch
from carla_imitation_learning_b200.trainer import TrainStep
from src.architectures.nets import ConvNet1
dev = torch.device("cuda", 0)
B, NBUF, CYCLES = 256, 4, 100
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


def run(fork_after, hi_main):
    """fork_after = None: serial. Otherwise the staging of step i+1 forks after `fork_after` forward layers of step i."""
    net, opt, ts = make()
    eng, L = ts.eng, ts.eng.lib
    for i in range(2 * NBUF):
        ts._enqueue(frames[i % NBUF], labels[i % NBUF])
    torch.cuda.synchronize()
    opt.prepare()
    planes = [ts.staged, StagedBatch(torch.empty_like(ts.staged.tp), None, 4)]
    main = torch.cuda.Stream(priority=-1 if hi_main else 0)
    side = torch.cuda.Stream(priority=0)
    g = torch.cuda.CUDAGraph()
    if fork_after is not None:
        stage_frames(frames[0], out=planes[0])          # the first step's planes: staged before the first replay
        torch.cuda.synchronize()
    with torch.cuda.graph(g, stream=main):
        for i in range(NBUF):
            b = ts.bufs
            if fork_after is None:
                stage_frames(frames[i], out=planes[0])
                b.x_tp = planes[0].tp
            else:
                b.x_tp = planes[i & 1].tp
            b.y = labels[i]
            c = eng.ctx(b)
            s = torch.cuda.current_stream().cuda_stream
            ev_f, ev_j = torch.cuda.Event(), torch.cuda.Event()
            for layer in range(4):
                if fork_after is not None and layer == fork_after:
                    ev_f.record()
                    side.wait_event(ev_f)
                    with torch.cuda.stream(side):
                        stage_frames(frames[(i + 1) % NBUF], out=planes[(i + 1) & 1])
                        ev_j.record()
                _lib.check(L.bc_conv_relu_pool_fwd(C.byref(c), layer, s), "fwd")
            _lib.check(L.bc_backward(C.byref(c), 1, s), "bwd")
            opt.step_flat(eng.grads)
            if fork_after is not None:
                torch.cuda.current_stream().wait_event(ev_j)
    ms = timed(g.replay, CYCLES) / NBUF
    ts.check()
    loss = float(ts.bufs.loss)
    return ms, loss, net._arena.clone()


ref = None
for name, fa, hi in [("serial", None, False), ("fork at start, same prio", 0, False), ("fork at start, main high", 0, True),
                     ("fork after conv1 fwd, same", 1, False), ("fork after conv2 fwd, same", 2, False), ("fork after conv2 fwd, main high", 2, True),
                     ("fork after conv3 fwd, same", 3, False), ("serial again", None, False)]:
    ms, loss, arena = run(fa, hi)
    if ref is None:
        ref = arena
    print(f"{name:34s} {ms * 1e3:8.2f} us/step  loss {loss:.6f}  arena == serial: {bool(torch.equal(arena, ref))}", flush=True)

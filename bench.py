"""bench.py -- BC train frames/s on N B200s (BASELINE.json metric), with roofline and CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl b200|reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...          (N > 1)

One "step" = one optimisation step of the BC policy on `batch` frames per GPU:
stage (u8 RGB -> gray planes) -> conv1..4 (+ReLU+pool) -> head + CE -> backward -> [allreduce] -> Adam.
`value` = frames/s with the u8 frames already resident in HBM; `e2e` = the same step driven
from pinned HOST buffers (H2D of the frames and labels, D2H of the loss, every step).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "bc_train_frames_per_sec"
UNIT = "frames/s"
FLOPS_FWD = (44255232, 14745600, 5308416, 589824)          # SURVEY 8(d), per frame, conv1..4
FLOPS_TRAIN = 150505152
FRAME_BYTES = 256 * 256 * 3
TRAFFIC_CONV1_TP = 45459712        # dram__bytes_read.sum + dram__bytes_write.sum of conv1_tp_kernel at B=256 (profiles/r1i_ncu_full_conv1_tp_raw.csv)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def window(self, t0, t1):
        """Keep only samples taken inside [t0, t1] (host clock around the timed region)."""
        inside = [r for (t, r) in self.rows if t0 <= t <= t1 + 0.15]
        self.rows = [(0, r) for r in (inside if inside else [r for (_t, r) in self.rows][-3:])]

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        rows = [r for (_t, r) in self.rows]
        sm = sorted(float(r[0]) for r in rows if r and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def bind_to_gpu_numa_node(index: int):
    """Run this rank (and first-touch its pinned buffers) on the CPUs of the NUMA node its GPU hangs off. With 8 ranks
    uploading 51 MB per step each, host pages on the wrong socket halve the H2D rate. Best effort: returns the node or None."""
    try:
        bus = subprocess.run(["nvidia-smi", f"--id={index}", "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True, timeout=10).stdout.strip().lower()
        if bus.startswith("00000000:"):
            bus = bus[4:]                                     # sysfs uses a 4-digit PCI domain
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if allowed:
            os.sched_setaffinity(0, allowed)
            return node
    except Exception:
        pass
    return None


def synth_host_frames(seed: int, n: int):
    """Uniform-noise CARLA-shaped u8 RGB frames + uniform labels (worst case for caches)."""
    import numpy as np
    rng = np.random.Generator(np.random.PCG64(seed))
    return rng.integers(0, 256, size=(n, 256, 256, 3), dtype=np.uint8), rng.integers(0, 9, size=n, dtype=np.int64)


# ----------------------------------------------------------------------------------------------
def run_reference(args):
    """The reference's own CPU path (torch-CPU restatement in oracle/, all host threads)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from oracle import bc_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    B = 32
    frames, labels = synth_host_frames(0, B + 4)
    x, y = O.sequential_samples(frames, labels)
    x, y = torch.from_numpy(x), torch.from_numpy(y)
    tr = O.OracleTrainer(O.init_params(12345))
    for _ in range(args.warmup):
        tr.step(x, y)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        tr.step(x, y)
    dt = time.perf_counter() - t0
    fps = B * args.steps / dt
    sample = f"{B}-frame slice of the {args.batch}-frame batch per step, {args.steps} steps, torch {torch.__version__} CPU f32"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"ConvNet1 BC train step, obs 4x256x256, 9 actions, batch {args.batch}/GPU", "cpu_batch": B},
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}), flush=True)


def cpu_baseline(batch):
    import torch
    from oracle import bc_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    B = 32
    frames, labels = synth_host_frames(0, B + 4)
    x, y = O.sequential_samples(frames, labels)
    x, y = torch.from_numpy(x), torch.from_numpy(y)
    tr = O.OracleTrainer(O.init_params(12345))
    for _ in range(3):
        tr.step(x, y)
    n, t0 = 0, time.perf_counter()
    while n < 200 and time.perf_counter() - t0 < 12.0:
        tr.step(x, y)
        n += 1
    dt = time.perf_counter() - t0
    return {"value": B * n / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"oracle port (torch CPU f32, {cores} threads), {n} steps of a {B}-frame slice of the {batch}-frame batch, {1e3 * dt / n:.1f} ms/step"}


# ----------------------------------------------------------------------------------------------
def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from carla_imitation_learning_b200 import FusedAdam, sliding_window, stage_frames, stage_gray
    from carla_imitation_learning_b200 import _lib
    from src.architectures.nets import ConvNet1

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch N>1 with torch.distributed.run (see module docstring)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_to_gpu_numa_node(local)      # pinned host buffers are first-touched on the GPU's own NUMA node (e2e H2D path)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    _lib.build()
    B = args.batch
    torch.manual_seed(12345)
    net = ConvNet1({"obs_size": 4, "n_actions": 9}).to(dev)
    eng = net.engine()
    opt = FusedAdam(list(net.parameters()), lr=1e-3)
    if args.mode == "bf16":
        args.staged = "bf16"
    staged_dtype = torch.bfloat16 if args.staged == "bf16" else torch.float32

    # ---- data: NBUF distinct frame windows per rank (rotated so inputs are never L2-resident) ----
    NBUF = args.nbuf
    host_frames, host_labels, dev_frames, dev_labels = [], [], [], []
    for i in range(NBUF):
        f, l = synth_host_frames(1000 * rank + i, B + 4)
        hf = torch.from_numpy(f).pin_memory()
        hl = torch.from_numpy(l[4:4 + B].copy()).pin_memory()
        host_frames.append(hf); host_labels.append(hl)
        dev_frames.append(hf.to(dev)); dev_labels.append(hl.to(dev))
    if args.mode == "bf16":
        eng.set_mode("bf16")           # before alloc(): bf16 mode adds the bf16 activation copies (P8 / P8B layouts)
        staged = stage_frames(dev_frames[0])          # Toeplitz-ready bf16 planes, rewritten in place every step
        bufs = eng.alloc(B, staged, dev_labels[0], True)

        def stage(frames_u8):
            stage_frames(frames_u8, out=staged)
    else:
        gray = torch.empty((B + 4, 256, 256), dtype=staged_dtype, device=dev)
        bufs = eng.alloc(B, sliding_window(gray), dev_labels[0], True)

        def stage(frames_u8):
            stage_gray(frames_u8, out=gray)

    from carla_imitation_learning_b200.parallel import DataParallelStep, PeerExchangeStep
    # N > 1: gradient exchange fused into the Adam kernel over NVLink peer memory (default), or the 2-bucket NCCL all-reduce
    dp = None if world == 1 else (PeerExchangeStep(eng, opt) if args.dp == "peer" else DataParallelStep(eng, opt))

    def train(b):
        if dp is not None:
            dp(b)
        else:
            eng.enqueue_train(b)
            opt.step_flat(eng.grads)

    def device_step(i):
        stage(dev_frames[i % NBUF])
        if args.mode == "bf16":
            eng.pack_weights()                 # f32 master weights -> bf16 MMA operand images
        bufs.y = dev_labels[i % NBUF]
        train(bufs)

    # CUDA graphs: one per input buffer. For N > 1 the step is captured as three graph segments with the
    # two NCCL all-reduces launched eagerly between them (capturing NCCL inside one graph hung on this
    # torch/NCCL build): [stage..reduce fc-conv2] -> allreduce b0 || [conv1 wgrad, reduce] -> allreduce b1 -> [Adam].
    graphs = None
    if not args.no_graph:
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for i in range(3):
                device_step(i)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graphs = []
        for i in range(NBUF):
            if dp is None:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    device_step(i)
                graphs.append(g)
            else:
                def pre(i=i):
                    stage(dev_frames[i % NBUF])
                    if args.mode == "bf16":
                        eng.pack_weights()
                    bufs.y = dev_labels[i % NBUF]
                graphs.append(dp.capture(bufs, pre))
    def step(i):
        if graphs is None:
            device_step(i)
        elif dp is None:
            graphs[i % NBUF].replay()
        else:
            dp.replay(graphs[i % NBUF])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for i in range(max(args.warmup, 3)):
        step(i)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_host0 = time.time()
    e0.record()
    for i in range(args.steps):
        step(i)
    e1.record()
    barrier()
    t_host1 = time.time()
    ms = e0.elapsed_time(e1)
    clocks = None
    if rank == 0:
        time.sleep(0.25)       # let the 100 ms sampler emit the rows that cover the end of the region
        sampler.window(t_host0, t_host1)
        clocks = sampler.stop()
    loss_dev = float(bufs.loss)

    # ---- e2e: host buffers in, loss out, every step -------------------------------------------
    # Two staging slots: while step i computes, the copy stream uploads step i+1's frames and labels
    # (every step's H2D is inside the timed region; it overlaps the previous step's compute). The loss
    # is copied back and read on the host every step.
    slots = [(torch.empty_like(dev_frames[0]), torch.empty_like(dev_labels[0])) for _ in range(2)]
    loss_host = torch.zeros((), dtype=torch.float32).pin_memory()
    copy_stream = torch.cuda.Stream(dev)
    up_done = [torch.cuda.Event(), torch.cuda.Event()]

    def upload(i):
        fr, lb = slots[i & 1]
        with torch.cuda.stream(copy_stream):
            fr.copy_(host_frames[i % NBUF], non_blocking=True)
            lb.copy_(host_labels[i % NBUF], non_blocking=True)
            up_done[i & 1].record(copy_stream)

    def slot_step(k):
        stage(slots[k][0])
        if args.mode == "bf16":
            eng.pack_weights()
        bufs.y = slots[k][1]
        train(bufs)

    slot_graphs = None
    if graphs is not None and world == 1:
        upload(0); upload(1)
        torch.cuda.synchronize()
        slot_graphs = []
        for k in range(2):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                slot_step(k)
            slot_graphs.append(g)

    def e2e_step(i):
        upload(i + 1)                                         # next step's inputs, overlapping this step's compute
        torch.cuda.current_stream().wait_event(up_done[i & 1])
        if slot_graphs is not None:
            slot_graphs[i & 1].replay()
        else:
            slot_step(i & 1)
        loss_host.copy_(bufs.loss, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return float(loss_host)

    upload(0)
    for i in range(3):
        e2e_step(i)
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for i in range(3, 3 + args.steps):
        e2e_step(i)
    f1.record()
    barrier()
    ms_e2e = f0.elapsed_time(f1)

    # ---- per-kernel times, measured live with CUDA events on the launching stream (no profiler) -------
    import ctypes as C
    c = eng.ctx(bufs)
    cref = C.byref(c)
    s = torch.cuda.current_stream().cuda_stream
    L = eng.lib
    ops = [("stage_gray", lambda: stage(dev_frames[0]))]
    if args.mode == "bf16":
        ops.append(("pack_weights", eng.pack_weights))
    ops += [(f"conv{l + 1}_fwd", (lambda l=l: _lib.check(L.bc_conv_relu_pool_fwd(cref, l, s)))) for l in range(4)]
    ops.append(("head_fwd_ce_bwd", lambda: _lib.check(L.bc_head(cref, 3, s))))
    for l in (3, 2, 1):
        ops.append((f"conv{l + 1}_wgrad", (lambda l=l: _lib.check(L.bc_conv_bwd_wgrad(cref, l, s)))))
        ops.append((f"conv{l + 1}_dgrad", (lambda l=l: _lib.check(L.bc_conv_bwd_dgrad(cref, l, s)))))
    ops.append(("conv1_wgrad", lambda: _lib.check(L.bc_conv_bwd_wgrad(cref, 0, s))))
    ops.append(("reduce_partials", lambda: _lib.check(L.bc_reduce_partials(cref, 1, s))))
    breakdown = {}
    for name, fn in ops:
        for _ in range(3):
            fn()
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0.record()
        for _ in range(10):
            fn()
        k1.record()
        torch.cuda.synchronize()
        breakdown[name] = round(k0.elapsed_time(k1) / 10 * 1e3, 1)
    k_ms = breakdown["conv1_fwd"] * 1e-3
    stage_ms = breakdown["stage_gray"] * 1e-3
    # algorithmic bytes of the staging kernel: u8 RGB in; gray planes out (bf16 mode: Toeplitz-ready bf16 planes)
    stage_bytes = (B + 4) * (FRAME_BYTES + (2 * _lib.TP_PLANE_ELEMS if args.mode == "bf16" else 65536 * 4))

    # a bounded mbarrier / peer-flag wait that expired would have left wrong results behind: never report such a run
    eng.check_device_errors()
    if dp is not None and hasattr(dp, "peer"):
        dp.peer.check()
    if not (loss_dev == loss_dev and 0.0 < loss_dev < 20.0):
        raise RuntimeError(f"training diverged or produced a non-finite loss ({loss_dev}): the timed run is invalid")
    times = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(times[0]), float(times[1])
    if rank == 0:
        peaks = load_peaks()
        frames = B * world * args.steps
        value = frames / (ms * 1e-3)
        achieved = FLOPS_FWD[0] * B / (k_ms * 1e-3) / 1e12
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": ("bf16" if args.mode == "bf16" else "f32"), "data": "synthetic",
            "config": {"workload": f"ConvNet1 BC train step (BASELINE configs[1]), obs 4x256x256, 9 actions, batch {B}/GPU, "
                                   f"u8 RGB frames staged to {args.staged} gray planes, sliding 4-frame window",
                       "global_batch": B * world, "parallelism": f"dp{world}", "cuda_graph": graphs is not None,
                       "host_numa_node": numa,
                       "exchange": (None if world == 1 else ("adam kernel reads peer gradient arenas over NVLink" if args.dp == "peer" else "nccl 2-bucket all-reduce")),
                       "l2": f"inputs rotate over {NBUF} x {(B + 4) * FRAME_BYTES / 1e6:.0f} MB device buffers (> 126 MB L2)",
                       "final_loss": loss_dev},
            "e2e": {"value": frames / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": (B + 4) * FRAME_BYTES + 8 * B, "d2h_bytes_per_step": 4},
            # per step: stage, [pack], 4 conv fwd, head, 3 x (wgrad, dgrad), conv1 wgrad, reduce, adam (tick folded in)
            "gpu_launches": (16 if args.mode == "bf16" else 15) * args.steps,
            "roofline": {"kernel": ("conv1_tp_kernel (tcgen05 Toeplitz implicit GEMM on Toeplitz-ready planes, bf16)" if args.mode == "bf16"
                                    else "conv_relu_pool_fwd_kernel<conv1> (exact-f32 FFMA variant)"),
                         "bound": "tensor", "achieved": achieved, "peak": peaks["tf_burst"], "unit": "TFLOP/s",
                         "frac": achieved / peaks["tf_burst"],
                         "traffic": (TRAFFIC_CONV1_TP if (args.mode == "bf16" and B == 256) else None),   # ncu dram read+write of this kernel, profiles/r1i
                         "peak_source": peaks["src"] + " (bf16 cuBLAS burst)",
                         "kernel_ms": k_ms, "flops_per_launch": FLOPS_FWD[0] * B,
                         "step_frac_of_dense_flops": FLOPS_TRAIN * B / (ms / args.steps * 1e-3) / 1e12 / peaks["tf_sust"]},
            "roofline_hbm": {"kernel": ("stage_gray_tp_kernel" if args.mode == "bf16" else "stage_gray_kernel"), "bound": "hbm", "achieved": stage_bytes / (stage_ms * 1e-3) / 1e9,
                             "peak": peaks["hbm"], "unit": "GB/s", "frac": stage_bytes / (stage_ms * 1e-3) / 1e9 / peaks["hbm"],
                             "bytes_per_launch": stage_bytes, "kernel_ms": stage_ms},
            # every conv kernel against the tensor roofline: algorithmic FLOPs (fwd = dgrad = wgrad per layer) / live time
            "roofline_conv_tflops": {k: round(FLOPS_FWD[int(k[4]) - 1] * B / (v * 1e-6) / 1e12, 1)
                                     for k, v in breakdown.items() if k.startswith("conv")},
            "breakdown_us": breakdown,
            "clocks": clocks,
        }
        if not args.no_cpu:
            out["cpu_baseline"] = cpu_baseline(B)
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--batch", type=int, default=256, help="frames per GPU per step")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--staged", default="f32", choices=["f32", "bf16"])
    ap.add_argument("--mode", default="bf16", choices=["fp32", "bf16"], help="fp32 = exact FFMA kernels; bf16 = tcgen05 kernels")
    ap.add_argument("--nbuf", type=int, default=4)
    ap.add_argument("--dp", default="peer", choices=["peer", "nccl"], help="N>1 gradient exchange: fused peer-memory Adam, or NCCL buckets")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()

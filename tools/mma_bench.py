"""Cycles per tcgen05.mma vs N and vs synchronisation pattern (run on the GPU box): python tools/mma_bench.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from carla_imitation_learning_b200 import _lib
lib = _lib.lib()
dev = torch.device("cuda", 0)
cyc = torch.zeros(2, dtype=torch.int64, device=dev)
err = torch.zeros(1, dtype=torch.int32, device=dev)
def run(N, reps, mode, grid=148):
    _lib.check(lib.bc_tc_mma_bench(N, reps, mode, grid, cyc.data_ptr(), err.data_ptr(), torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    c = cyc.cpu().tolist()
    return c[0] / reps, c[1] / reps, int(err[0])
for N in (16, 64, 128, 256):
    print(f"N {N:3d} back-to-back, one commit      : issue %.1f complete %.1f cyc/mma err %d" % run(N, 512, 0))
print("N  64 commit after every MMA          : issue %.1f complete %.1f cyc/mma err %d" % run(64, 512, 2))
for ns in (1, 2, 4, 7, 8, 14):
    print(f"N  64 ring of {ns:2d} stages (wait/commit per MMA): issue %.1f complete %.1f cyc/mma err %d" % run(64, 448, 2 + ns))
for N in (16, 64, 128, 256):
    print(f"N {N:3d} unrolled x8 groups, one commit   : issue %.1f complete %.1f cyc/mma err %d" % run(N, 512, 1))
for th in (512, 64, 128, 256, 512):
    print(f"N  64 back-to-back, CTA of {th:3d} threads: issue %.1f complete %.1f cyc/mma err %d" % run(64, 512, (th << 8) | 0))
for N in (32, 64):
    print(f"M=64 N {N:3d} unrolled x8 groups, one commit: issue %.1f complete %.1f cyc/mma err %d" % run(N, 512, 1 | (1 << 20)))
for N in (16, 32, 64):
    print(f"MN-major M=128 N {N:3d} unrolled x8 groups: issue %.1f complete %.1f cyc/mma err %d" % run(N, 512, 1 | (1 << 21)))
for off in (0, 1, 2, 5, 7):
    print(f"K-major  N  64 unrolled x8, A start +{16 * off:3d} B : issue %.1f complete %.1f cyc/mma err %d" % run(64, 512, 1 | (off << 22)))
for off in (0, 2, 5):
    print(f"K-major  N 256 unrolled x8, A start +{16 * off:3d} B : issue %.1f complete %.1f cyc/mma err %d" % run(256, 512, 1 | (off << 22)))
for sbo in (0, 1):
    for off in (0, 2):
        print(f"MN-major N  64 unrolled x8, A core-group stride {2016 if sbo else 2048}, start +{16 * off} B: issue %.1f complete %.1f cyc/mma err %d" % run(64, 512, 1 | (1 << 21) | (sbo << 25) | (off << 22)))

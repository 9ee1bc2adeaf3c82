"""N>1 host logic on CPU: world_size-2 gloo. The exchange + mean-folding must reproduce the
single-process global-batch step (reference DDP semantics)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import bc_oracle as O


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _flat_arena(named, offsets, total):
    flat = torch.zeros(total, dtype=torch.float64)
    for i, k in enumerate(O.PARAM_ORDER):
        t = named[k].reshape(-1).double()
        flat[offsets[i]:offsets[i] + t.numel()] = t
    return flat


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from carla_imitation_learning_b200 import _lib
        from carla_imitation_learning_b200.parallel import GradExchange, shard_range
        torch.set_num_threads(2)
        total, offsets, _ = _lib.arena_layout(4, 9)
        frames, labels = O.synth_frames(0, 8)
        x, y = O.sequential_samples(frames, labels)            # global batch of 4
        lo, hi = shard_range(4, rank, world)
        params = O.init_params(12345)
        loss, _, g = O.loss_and_grads(params, torch.from_numpy(x[lo:hi]), torch.from_numpy(y[lo:hi]), torch.float64)
        flat = _flat_arena(g, offsets, total)                  # local-mean gradients in arena order
        ex = GradExchange(4, 9)
        assert ex.world == world and ex.buckets[0][0] == 0 and ex.buckets[-1][1] == total
        ex.start(flat, 0)
        ex.start(flat, 1)
        ex.finish()
        flat *= ex.grad_scale                                   # what FusedAdam folds into its read
        if rank == 0:
            np.save(out, flat.numpy())
    finally:
        dist.destroy_process_group()


def test_two_rank_exchange_equals_global_batch_gradient(tmp_path):
    out = str(tmp_path / "g.npy")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    from carla_imitation_learning_b200 import _lib
    total, offsets, _ = _lib.arena_layout(4, 9)
    frames, labels = O.synth_frames(0, 8)
    x, y = O.sequential_samples(frames, labels)
    _, _, g = O.loss_and_grads(O.init_params(12345), torch.from_numpy(x), torch.from_numpy(y), torch.float64)
    ref = _flat_arena(g, offsets, total).numpy()
    got = np.load(out)
    assert np.abs(got - ref).max() <= 1e-12 * np.abs(ref).max() + 1e-18


def test_buckets_follow_reverse_completion_order():
    from carla_imitation_learning_b200 import _lib
    from carla_imitation_learning_b200.parallel import grad_buckets, shard_range
    total, offsets, sizes = _lib.arena_layout(4, 9)
    (a0, a1), (b0, b1) = grad_buckets(4, 9)
    assert a0 == 0 and a1 == b0 and b1 == total
    assert b0 == offsets[0]                                   # conv1 weight opens the last bucket
    assert all(offsets[i] < b0 for i in range(2, 14))         # everything else precedes it
    assert b1 - b0 == 3168 and sizes[0] == 3136               # 3,152 conv1 params padded to 32
    # arena order = reverse of backward completion: fc.4 first, conv1 last
    assert offsets[12] < offsets[10] < offsets[8] < offsets[6] < offsets[4] < offsets[2] < offsets[0]
    assert shard_range(2048, 3, 8) == (768, 1024)
    with pytest.raises(ValueError):
        shard_range(10, 0, 4)

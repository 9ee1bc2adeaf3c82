"""Does torch symmetric memory give peer-mapped pointers on this box? (run under torchrun, 2 ranks)"""
import os, torch, torch.distributed as dist
rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
import torch.distributed._symmetric_memory as symm
print(rank, "torch", torch.__version__, "symm attrs", [a for a in dir(symm) if not a.startswith("_")][:30], flush=True)
t = symm.empty(1024, dtype=torch.float32, device=dev)
t.fill_(float(rank + 1))
hdl = symm.rendezvous(t, dist.group.WORLD.group_name if hasattr(dist.group.WORLD, "group_name") else dist.group.WORLD)
print(rank, "handle", type(hdl).__name__, [a for a in dir(hdl) if not a.startswith("_")], flush=True)
print(rank, "buffer_ptrs", [hex(p) for p in hdl.buffer_ptrs], "signal_pad_ptrs", [hex(p) for p in hdl.signal_pad_ptrs], "pad size", getattr(hdl, "signal_pad_size", None), flush=True)
torch.cuda.synchronize(); dist.barrier()
peer = hdl.get_buffer((rank + 1) % world, (1024,), torch.float32)
print(rank, "peer value", float(peer[0]), flush=True)
torch.cuda.synchronize(); dist.barrier()
dist.destroy_process_group()

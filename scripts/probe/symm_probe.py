"""Probe: does torch symmetric memory give usable peer pointers on this box?  (torchrun, 2+ GPUs)"""
import os, torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm
rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
t = symm.empty((1024, 256), dtype=torch.float32, device=torch.device("cuda", local))
hdl = symm.rendezvous(t, dist.group.WORLD)
print(rank, "rank/world", hdl.rank, hdl.world_size, "ptrs", [hex(p) for p in hdl.buffer_ptrs][:4], flush=True)
t.fill_(float(rank))
hdl.barrier()
peer = hdl.get_buffer((rank + 1) % world, t.shape, t.dtype)
print(rank, "peer value", float(peer[0, 0]), flush=True)
peer[1].fill_(100.0 + rank)   # P2P store
hdl.barrier()
print(rank, "my row1 after peer store", float(t[1, 0]), flush=True)
dist.destroy_process_group()

"""Probe: torch symmetric memory over NVLink on this box (rendezvous, peer views, device-side barrier, timing)."""
import os, sys, time
import torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
t = symm_mem.empty((8192, 1536), dtype=torch.float16, device=dev)
hdl = symm_mem.rendezvous(t, dist.group.WORLD)
print(rank, "rendezvous ok", type(hdl).__name__, "world", hdl.world_size, "ptrs", len(hdl.buffer_ptrs), flush=True)
b_loc = 8192 // world
src = torch.full((b_loc, 1536), float(rank + 1), dtype=torch.float16, device=dev)
peers = [hdl.get_buffer(p, t.shape, t.dtype) for p in range(world)]
hdl.barrier(channel=0)
for p in range(world):
    peers[p][rank * b_loc:(rank + 1) * b_loc].copy_(src)
hdl.barrier(channel=0)
torch.cuda.synchronize()
ok = all(float(t[p * b_loc, 0]) == p + 1 for p in range(world))
print(rank, "peer writes visible:", ok, flush=True)
# timing: barrier alone, NCCL all-gather of the same payload
g = torch.empty_like(t)
for _ in range(5):
    hdl.barrier(channel=0); dist.all_gather_into_tensor(g, src)
torch.cuda.synchronize()
e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
e[0].record()
for _ in range(50): hdl.barrier(channel=0)
e[1].record()
for _ in range(50): dist.all_gather_into_tensor(g, src)
e[2].record()
for _ in range(50):
    for p in range(world):
        peers[p][rank * b_loc:(rank + 1) * b_loc].copy_(src)
    hdl.barrier(channel=0)
e[3].record()
torch.cuda.synchronize()
if rank == 0:
    print("barrier us", e[0].elapsed_time(e[1]) * 20, "nccl all_gather us", e[1].elapsed_time(e[2]) * 20,
          "peer copies + barrier us", e[2].elapsed_time(e[3]) * 20, flush=True)
small = torch.zeros((3, 3, 8192), dtype=torch.float32, device=dev)
for _ in range(5): dist.all_reduce(small)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50): dist.all_reduce(small)
e1.record(); torch.cuda.synchronize()
if rank == 0: print("nccl all_reduce 295KB us", e0.elapsed_time(e1) * 20, flush=True)
dist.destroy_process_group()

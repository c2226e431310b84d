"""2-rank probe of torch symmetric memory (the plumbing of the fused gradient reduction):
torchrun --nproc-per-node 2 tools/symm_probe.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm

rank = int(os.environ["RANK"]); ws = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
n = 1_187_000
t = symm.empty(n, dtype=torch.float32, device=dev)
hdl = symm.rendezvous(t, dist.group.WORLD)
print(rank, "rendezvous ok", hdl.rank, hdl.world_size, "multicast", hdl.has_multicast_support, hex(hdl.multicast_ptr) if hdl.has_multicast_support else None,
      [hex(p) for p in hdl.buffer_ptrs], flush=True)
t.fill_(float(rank + 1))
hdl.barrier(channel=0)
peer = hdl.get_buffer((rank + 1) % ws, (n,), torch.float32)
s = float(peer[:10].sum()) ; print(rank, "peer sum of 10", s, flush=True)
hdl.barrier(channel=0)
# timing of barrier pairs
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(100):
    hdl.barrier(channel=0)
e1.record(); torch.cuda.synchronize()
print(rank, "barrier us", e0.elapsed_time(e1) * 10, flush=True)
dist.barrier(); dist.destroy_process_group()

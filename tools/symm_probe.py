import os, time, torch, torch.distributed as dist
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
import torch.distributed._symmetric_memory as symm_mem
n = 3_200_000
t = symm_mem.empty(n, dtype=torch.float64, device=dev)
hdl = symm_mem.rendezvous(t, group=dist.group.WORLD)
t.fill_(rank + 1.0)
hdl.barrier(channel=0)
peer = (rank + 1) % world
pb = hdl.get_buffer(peer, (n,), torch.float64)
local = torch.empty(n, dtype=torch.float64, device=dev)
torch.cuda.synchronize()
for nn in (n, n // 8):
    for _ in range(3):
        local[:nn].copy_(pb[:nn])
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(20):
        local[:nn].copy_(pb[:nn])
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 20
    if rank == 0:
        print("pull %.1f MB: %.1f us -> %.0f GB/s" % (nn * 8 / 1e6, ms * 1e3, nn * 8 / ms / 1e6), "ok", float(local[0].item()) == peer + 1.0)
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); s.record()
for _ in range(20):
    hdl.barrier(channel=0)
e.record(); torch.cuda.synchronize()
if rank == 0:
    print("symm barrier: %.1f us" % (s.elapsed_time(e) / 20 * 1e3))
dist.destroy_process_group()

#!/usr/bin/env python
"""Platform floor of the end-to-end call: raw pinned-memory copies of the e2e byte counts
(168 MB host->device, 654 MB device->host per rank), one rank alone and all ranks together.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29513 tools/pcie_floor.py
"""
import os, time
import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
H2D, D2H = 167_772_160, 654_325_040
h_in = torch.empty(H2D, dtype=torch.uint8, pin_memory=True)
h_out = torch.empty(D2H, dtype=torch.uint8, pin_memory=True)
d_in = torch.empty(H2D, dtype=torch.uint8, device="cuda")
d_out = torch.empty(D2H, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def one(both=True):
    with torch.cuda.stream(s1):
        h_out.copy_(d_out, non_blocking=True)
    if both:
        with torch.cuda.stream(s2):
            d_in.copy_(h_in, non_blocking=True)
    s1.synchronize()
    s2.synchronize()


def timed(reps=8, both=True):
    one(both)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        one(both)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    t = torch.tensor([dt], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


# all ranks together
t_all = timed()
t_all_d2h = timed(both=False)
# one rank at a time
t_solo = None
for r in range(world):
    if world > 1:
        dist.barrier()
    if r == rank:
        one()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(8):
            one()
        torch.cuda.synchronize()
        t_solo = (time.perf_counter() - t0) / 8
    if world > 1:
        dist.barrier()
ts = torch.tensor([t_solo], device="cuda", dtype=torch.float64)
if world > 1:
    parts = [torch.zeros_like(ts) for _ in range(world)]
    dist.all_gather(parts, ts)
    solos = [float(p.item()) for p in parts]
else:
    solos = [t_solo]
if rank == 0:
    print(f"ranks {world}: D2H 654 MB + H2D 168 MB per rank, pinned, both directions at once")
    print(f"  one rank at a time : {', '.join(f'{1e3*t:.2f}' for t in solos)} ms  ({D2H/min(solos)/1e9:.1f} GB/s D2H best)")
    print(f"  all ranks together : {1e3*t_all:.2f} ms (max over ranks)  -> {world*D2H/t_all/1e9:.1f} GB/s D2H aggregate, "
          f"{D2H/t_all/1e9:.1f} GB/s per rank")
    print(f"  all ranks, D2H only: {1e3*t_all_d2h:.2f} ms  -> {world*D2H/t_all_d2h/1e9:.1f} GB/s aggregate")
    try:
        import subprocess
        print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout[:1500])
        print(subprocess.run(["bash", "-c", "lscpu | egrep 'Model name|Socket|NUMA|^CPU\\(s\\)'; free -g | head -2"], capture_output=True, text=True).stdout)
    except Exception as e:
        print(e)
if world > 1:
    dist.destroy_process_group()

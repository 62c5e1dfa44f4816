"""Copy-only probe of the box's host<->device ceiling: every rank moves page-locked host memory to
its GPU (and back) with plain asynchronous copies and no kernel, all ranks at once.  What the
host-buffer legs of bench.py (e2e_loglike) can at best reach.
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/h2d_probe.py
One JSON object on rank 0."""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, ROOT)

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
bound = None
if world > 1:
    from mbb_emcee_b200.sharding import bind_rank_cpus
    pr = torch.cuda.get_device_properties(local)
    if os.environ.get("PROBE_NO_BIND") is None:
        bound = bind_rank_cpus("%08x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id), local, world)
    dist.init_process_group("nccl", device_id=dev)
GB = 1 << 30
host = torch.empty(GB // 8, dtype=torch.float64).pin_memory()
host.fill_(1.0)
devb = torch.empty(GB // 8, dtype=torch.float64, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def sync():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def run(mode, seconds=1.5, chunk=64 << 20):
    n = chunk // 8
    nchunks = host.numel() // n
    sync()
    t0 = time.perf_counter()
    moved = 0
    while time.perf_counter() - t0 < seconds:
        for k in range(nchunks):
            sl = slice(k * n, (k + 1) * n)
            if mode in ("h2d", "both"):
                with torch.cuda.stream(s1):
                    devb[sl].copy_(host[sl], non_blocking=True)
            if mode in ("d2h", "both"):
                with torch.cuda.stream(s2):
                    host[sl].copy_(devb[sl], non_blocking=True)
        s1.synchronize()
        s2.synchronize()
        moved += GB * (2 if mode == "both" else 1)
    dt = time.perf_counter() - t0
    t = torch.tensor([moved / dt / 1e9], dtype=torch.float64, device=dev)
    mn = t.clone()
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        dist.all_reduce(mn, op=dist.ReduceOp.MIN)
    return {"aggregate_gb_per_s": float(t.item()), "slowest_rank_gb_per_s": float(mn.item())}


out = {"ranks": world, "cpu_binding": ("%d CPUs per rank" % len(bound)) if bound else "none",
       "buffer": "1 GiB page-locked per rank, 64 MiB copies"}
for mode in ("h2d", "d2h", "both"):
    out[mode] = run(mode)
if rank == 0:
    print(json.dumps(out), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()

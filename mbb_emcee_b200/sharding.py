"""Sharding of independent work across one-process-per-GPU ranks.

Every (source, walker) evaluation and every chain sample is independent
(SURVEY.md 8e), so multi-GPU use is: cut the unit range into contiguous
per-rank pieces, run the single-GPU path on each, and gather once at the end.
Nothing on the evaluation path communicates.  These helpers are the whole of
the host-side logic; they work with any ``torch.distributed`` backend (NCCL on
the GPU box, gloo in the CPU tests).
"""
import os

import numpy as np

__all__ = ["rank_world", "shard_range", "shard_sources", "gather_concat", "shared_array", "bind_to_gpu_numa_node",
           "bind_rank_cpus", "parse_cpulist"]


def rank_world():
    """(rank, world_size) from the torchrun environment (0, 1 when absent)."""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def parse_cpulist(text):
    """'0-15,32-47' -> [0, ..., 15, 32, ..., 47]  (sysfs cpulist format)."""
    cpus = []
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.extend(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_to_gpu_numa_node(pci_bus_id, sysfs="/sys"):
    """Restrict this process to the CPUs of the NUMA node the GPU hangs off, so that the
    pinned host buffers of the host-memory (MBB_HOST) calls are allocated next to the GPU's
    PCIe root port.  With one process per GPU on a multi-socket box this keeps every rank's
    H2D/D2H traffic off the inter-socket link.  `pci_bus_id` as CUDA reports it
    ('00000000:1B:00.0'); returns the CPU list it bound to, or None when the topology is not
    exposed (single-node hosts, containers without sysfs) -- never raises."""
    try:
        dom, bus, rest = pci_bus_id.lower().split(":")
        dev = "%s:%s:%s" % (dom[-4:], bus, rest)
        node = int(open(os.path.join(sysfs, "bus/pci/devices", dev, "numa_node")).read())
        if node < 0:
            return None
        cpus = parse_cpulist(open(os.path.join(sysfs, "devices/system/node/node%d/cpulist" % node)).read())
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return allowed
    except Exception:
        return None


def bind_rank_cpus(pci_bus_id, local_rank, local_world, sysfs="/sys"):
    """CPU placement of one rank of a one-process-per-GPU launch: the GPU's NUMA node where the
    topology is exposed (bind_to_gpu_numa_node); otherwise -- single-node VMs, containers that hide
    sysfs -- the process's CPU set is cut into ``local_world`` disjoint contiguous pieces and the
    rank takes its own, so that the staging copies and driver threads of different ranks never
    share a core.  Returns the CPU list bound to, or None (never raises)."""
    got = bind_to_gpu_numa_node(pci_bus_id, sysfs=sysfs)
    if got:
        return got
    try:
        cpus = sorted(os.sched_getaffinity(0))
        if local_world <= 1 or len(cpus) < local_world:
            return None
        lo, hi = shard_range(len(cpus), local_rank, local_world)
        mine = cpus[lo:hi]
        os.sched_setaffinity(0, mine)
        return mine
    except Exception:
        return None


def shard_range(n, rank, world):
    """Contiguous, balanced [lo, hi) of ``n`` units for ``rank`` (sizes differ by <= 1)."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError("bad rank/world")
    base, extra = divmod(int(n), world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_sources(nsrc, nwalkers, rank, world):
    """Evaluation index range owned by ``rank`` when sources (never the walkers
    of one source) are split across ranks: (src_lo, src_hi, eval_lo, eval_hi)."""
    lo, hi = shard_range(nsrc, rank, world)
    return lo, hi, lo * nwalkers, hi * nwalkers


class shared_array(object):
    """One array that every rank of a one-process-per-GPU launch on ONE node maps and fills in
    disjoint slices: the "final gather" without a collective.  Rank 0 creates a file
    in /dev/shm, everyone maps it and page-locks its mapping (cudaHostRegister), so that each GPU's
    device-to-host copy lands directly in the shared pages; a barrier after the last copy is all the
    synchronisation there is.  ``barrier`` is any callable that synchronises the ranks
    (torch.distributed.barrier); without one (single process) the array is private.

        sh = shared_array("lir", (nwalkers, nsteps), rank, world, barrier=dist.barrier)
        ctx.chain_post_into(chain[lo:hi], 2, lir=sh.array[lo:hi], ...)
        sh.sync()            # every rank's slice is complete and visible
        ... sh.array ...     # the whole result, on every rank
        sh.close()
    """

    def __init__(self, tag, shape, rank=0, world=1, dtype=np.float64, barrier=None, register=True,
                 directory="/dev/shm"):
        self._barrier = barrier if world > 1 else None
        self._rank = rank
        nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        self._path = os.path.join(directory, "mbb_b200_%s_%s" % (os.environ.get("MASTER_PORT", str(os.getppid())), tag))
        if rank == 0:
            with open(self._path, "wb") as fh:
                fh.truncate(max(nbytes, 1))
        if self._barrier:
            self._barrier()
        self.array = np.memmap(self._path, dtype=dtype, mode="r+", shape=tuple(shape))
        self._unregister = None
        if register:
            from . import _native
            self._unregister = _native.host_register(self.array) or None

    def sync(self):
        if self._barrier:
            self._barrier()

    def close(self):
        if self._unregister:
            self._unregister()
            self._unregister = None
        self.sync()
        if self._rank == 0 and os.path.exists(self._path):
            os.unlink(self._path)


def gather_concat(local, group=None):
    """The one final gather: every rank's 1-D/2-D float64 array, concatenated
    along axis 0 in rank order, on every rank.  Shards may differ in length."""
    import torch
    import torch.distributed as dist
    local = np.ascontiguousarray(local, dtype=np.float64)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    n = torch.tensor([local.shape[0]], dtype=torch.int64, device=dev)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    width = int(np.prod(local.shape[1:])) if local.ndim > 1 else 1
    pad = torch.zeros((max(sizes), width), dtype=torch.float64, device=dev)
    pad[:local.shape[0]] = torch.from_numpy(local.reshape(local.shape[0], width)).to(dev)
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    out = np.concatenate([p[:s].cpu().numpy() for p, s in zip(parts, sizes)], axis=0)
    return out.reshape((-1,) + local.shape[1:]) if local.ndim > 1 else out.ravel()

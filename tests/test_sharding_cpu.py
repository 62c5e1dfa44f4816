"""CPU, world_size 2 over gloo: the N>1 path (shard -> evaluate -> one gather).

The per-rank evaluation is stood in for by the host emulation of the device
code (there is no GPU here); what is under test is the host-side logic the
multi-GPU runs rely on: balanced contiguous shards, walkers of a source never
split, rank-ordered gather equal to the single-process result."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT


def test_shard_range_properties():
    from mbb_emcee_b200.sharding import shard_range, shard_sources
    for n in (0, 1, 7, 100000, 100003):
        for world in (1, 2, 3, 8):
            pieces = [shard_range(n, r, world) for r in range(world)]
            assert pieces[0][0] == 0 and pieces[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(pieces, pieces[1:]))
            sizes = [hi - lo for lo, hi in pieces]
            assert max(sizes) - min(sizes) <= 1
    assert shard_sources(10, 512, 1, 4) == (3, 6, 1536, 3072)
    with pytest.raises(ValueError):
        shard_range(5, 2, 2)


def _worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    import hostemu_lib as emu
    from mbb_emcee_b200.sharding import gather_concat, rank_world, shard_sources
    dist.init_process_group("gloo", rank=rank, world_size=world)
    assert rank_world() == (rank, world)
    d = np.load(os.path.join(tmp, "problem.npz"))
    nsrc, nw = d["flux"].shape[0], int(d["nw"])
    s_lo, s_hi, e_lo, e_hi = shard_sources(nsrc, nw, rank, world)
    ep = emu.priors_struct([1, 0.1, 1, 0.1, 1e-3], [0, 1, 1, 1, 0, 0],
                           [np.inf, 20, 1500, 20, np.inf, np.inf], [0] * 6, [0] * 6, [1] * 6)
    off = np.arange(7, dtype=np.int32)
    ll, st = emu.loglike(True, True, True, d["P"][e_lo:e_hi], 500.0, ep, off, d["waves"], np.ones(6),
                         np.zeros(6), d["flux"][s_lo:s_hi], ivar=d["ivar"][s_lo:s_hi], wps=nw)
    full = gather_concat(ll)
    both = gather_concat(np.stack([ll, st.astype(np.float64)], axis=1))
    if rank == 0:
        np.savez(os.path.join(tmp, "gathered.npz"), ll=full, both=both)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_matches_single_process(tmp_path):
    import torch.multiprocessing as mp
    import hostemu_lib as emu
    rng = np.random.RandomState(5)
    nsrc, nw = 7, 16                      # odd source count: shards of 4 and 3 sources
    waves = np.array([70.0, 100.0, 160.0, 250.0, 350.0, 500.0])
    flux = rng.uniform(5, 80, (nsrc, 6))
    ivar = 1.0 / rng.uniform(1, 6, (nsrc, 6))**2
    P = np.column_stack([rng.uniform(8, 25, nsrc * nw), rng.uniform(1.2, 2.4, nsrc * nw),
                         np.full(nsrc * nw, 1300.0), np.full(nsrc * nw, 4.0),
                         rng.uniform(5, 100, nsrc * nw)])
    np.savez(tmp_path / "problem.npz", P=P, flux=flux, ivar=ivar, waves=waves, nw=nw)
    emu.lib()                             # build once before forking workers
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    ep = emu.priors_struct([1, 0.1, 1, 0.1, 1e-3], [0, 1, 1, 1, 0, 0],
                           [np.inf, 20, 1500, 20, np.inf, np.inf], [0] * 6, [0] * 6, [1] * 6)
    want, st = emu.loglike(True, True, True, P, 500.0, ep, np.arange(7, dtype=np.int32), waves,
                           np.ones(6), np.zeros(6), flux, ivar=ivar, wps=nw)
    got = np.load(tmp_path / "gathered.npz")
    assert np.array_equal(got["ll"], want)
    assert np.array_equal(got["both"][:, 0], want) and np.array_equal(got["both"][:, 1], st)


def _chain_worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch.distributed as dist
    import mbb_oracle as oracle
    from mbb_emcee_b200.sharding import gather_concat, shard_range, shared_array
    dist.init_process_group("gloo", rank=rank, world_size=world)
    chain = np.load(os.path.join(tmp, "chain.npy"))
    nw, ns = chain.shape[:2]
    lo, hi = shard_range(nw, rank, world)
    # the final gather without a collective: every rank fills its rows of one shared mapping
    sh = shared_array("peak", (nw, ns), rank, world, barrier=dist.barrier, register=False)
    sh.array[lo:hi] = oracle.map_chain(chain[lo:hi], oracle.peaklambda_step)
    sh.sync()
    whole = np.array(sh.array)
    # ... and the same through the collective
    coll = gather_concat(np.ascontiguousarray(sh.array[lo:hi]))
    assert np.array_equal(whole, coll)
    if rank == 1:                         # any rank sees the whole result
        np.save(os.path.join(tmp, "peak_sharded.npy"), whole)
    sh.close()
    dist.destroy_process_group()


def test_chain_rows_shard_with_local_repeat_scan(tmp_path):
    """configs[3] over ranks: walker rows are cut contiguously, so the sequential np.allclose
    repeat scan of _map_chain (reference results.py:553-566) never crosses a shard; the shards land
    in disjoint rows of one shared array (no collective) and equal the single-process result."""
    import torch.multiprocessing as mp
    import mbb_oracle as oracle
    from mbb_emcee_b200 import synthetic
    chain = synthetic.random_walk_chain((14.0, 1.8, 400.0, 3.0, 30.0), 7, 40, np.random.RandomState(3))
    np.save(tmp_path / "chain.npy", chain)
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_chain_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    want = oracle.map_chain(chain, oracle.peaklambda_step)
    got = np.load(tmp_path / "peak_sharded.npy")
    assert np.array_equal(got, want)
    assert (np.diff(want, axis=1) == 0).any()         # the chain does contain repeats


def test_rank_cpu_sets_are_disjoint():
    """Where sysfs hides the NUMA topology every rank still gets its own CPUs."""
    from mbb_emcee_b200.sharding import bind_rank_cpus
    before = sorted(os.sched_getaffinity(0))
    if len(before) < 2:
        pytest.skip("one CPU")
    try:
        sets = []
        for r in range(2):
            os.sched_setaffinity(0, before)
            sets.append(bind_rank_cpus("00000000:ff:00.0", r, 2, sysfs="/nonexistent"))
        assert sets[0] and sets[1] and not set(sets[0]) & set(sets[1])
        assert sorted(sets[0] + sets[1]) == before
    finally:
        os.sched_setaffinity(0, before)


def test_numa_binding_helper(tmp_path):
    """bind_to_gpu_numa_node reads the GPU's NUMA node and that node's CPU list from sysfs
    and never raises when the topology is not exposed."""
    import os
    from mbb_emcee_b200.sharding import bind_to_gpu_numa_node, parse_cpulist
    assert parse_cpulist("0-3,8,10-11\n") == [0, 1, 2, 3, 8, 10, 11]
    assert bind_to_gpu_numa_node("00000000:1b:00.0", sysfs=str(tmp_path)) is None      # nothing there
    before = sorted(os.sched_getaffinity(0))
    dev = tmp_path / "bus/pci/devices/0000:1b:00.0"
    dev.mkdir(parents=True)
    (dev / "numa_node").write_text("1\n")
    node = tmp_path / "devices/system/node/node1"
    node.mkdir(parents=True)
    (node / "cpulist").write_text("%d\n" % before[0])
    try:
        assert bind_to_gpu_numa_node("00000000:1B:00.0", sysfs=str(tmp_path)) == [before[0]]
        assert sorted(os.sched_getaffinity(0)) == [before[0]]
    finally:
        os.sched_setaffinity(0, before)
    (dev / "numa_node").write_text("-1\n")
    assert bind_to_gpu_numa_node("00000000:1b:00.0", sysfs=str(tmp_path)) is None

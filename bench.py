#!/usr/bin/env python
"""Benchmark of the modified-blackbody log-likelihood hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                    [--workload cfg5|cfg2]

Metric (BASELINE.json): MBB log-likelihood evals/sec (walker-SEDs/s).  One eval
= one likelihood.__call__-equivalent (5 parameters in, 1 log-probability out,
incl. per-walker setup, all bands x nodes, chi-square, limits and priors).

Workload (default): BASELINE configs[4] -- batch fit of 1e5 synthetic sources x
512 walkers on the cfg1 band set (6 delta bands 70..500 um, optically thin, no
alpha).  One step = one pass of the likelihood over all 5.12e7 walker positions
of this rank's sources.  Under torchrun every rank owns its own 1e5 sources
(weak scaling, no data-path collective).  `--workload cfg2` runs the tabulated
passband configuration (1688 nodes, optically thick + alpha) instead.

One JSON line on stdout (rank 0).  `value` is device-resident throughput;
`e2e` goes through the host-buffer C-ABI call (pinned host arrays in and out,
H2D/D2H inside the timed region).  `--impl reference` times the CPU oracle
(the reference's algorithm; its own compiled fnu.pyx when oracle/_ref was
built) on all host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "mbb_loglike_evals_per_sec"
UNIT = "evals/s"

# Algorithmic FP64 work per evaluation: SURVEY.md 8(d) -- op counts of the
# REFERENCE formulation with CUDA-12.9 libdevice on sm_100a (pow 125, expm1 35,
# exp 31, log 47, div 20, FMA = 2 flops).
FLOP_PER_EVAL = {"cfg5": 6 * 182 + 225 + 18 + 40,            # = 1375 (cfg1 band set)
                 "cfg2": 1688 * 366 + 3600 + 58,             # = 621466
                 "cfg5p": 1688 * 184 + 225 + 18 + 40}        # = 310875 (thin, no alpha, cfg2 band set)
BYTES_PER_EVAL = 48                                          # 40 B parameters in + 8 B out


def ncu_fact(workload, key):
    """Figures taken from the committed `ncu --set full` capture of the dominant kernel
    (profiles/ncu_facts.json, written by tools/ncu_summary.py facts): not measured live."""
    try:
        facts = json.load(open(os.path.join(ROOT, "profiles", "ncu_facts.json")))[workload]
    except Exception:
        return None
    return facts if key is None else facts.get(key)


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# --------------------------------------------------------------------------- clocks
class ClockSampler(object):
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(mx)) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------- workload setup
def build_workload(name, rank, nsrc_override=None):
    """Returns dict(ctx, n, nw, nsrc, P_host (pinned torch), like-or-None, spec pieces)."""
    import torch
    from mbb_emcee_b200 import _native, likelihood, synthetic
    cfg5 = synthetic.CONFIGS["cfg5"]
    dev = torch.device("cuda", torch.cuda.current_device())
    if name == "cfg5":
        cfg = cfg5
        nsrc = nsrc_override or cfg["nsources"]
        nw = cfg["nwalkers"]
        truth = np.array([12.0, 1.8, 1300.0, 4.0, 30.0])
    elif name == "cfg5p":
        cfg = synthetic.CONFIGS["cfg5p"]
        nsrc = nsrc_override or cfg["nsources"]
        nw = cfg["nwalkers"]
        truth = np.array(cfg["truth"])
    else:
        cfg = synthetic.CONFIGS["cfg2"]
        nsrc = nsrc_override or 2048
        nw = 512
        truth = np.array(cfg["truth"])
    rng = np.random.RandomState(cfg["seed"] + 1000 * rank)
    ctx = _native.Context(dev.index)
    ctx.set_model(cfg["wavenorm"], cfg["opthin"], cfg["noalpha"])
    # one likelihood object is used only for its host-side table builder
    like = likelihood(wavenorm=cfg["wavenorm"], noalpha=cfg["noalpha"], opthin=cfg["opthin"],
                      response=cfg["response"], device=dev.index)
    nb = len(cfg["bands"])
    like.set_phot(cfg["bands"], np.ones(nb), np.ones(nb))
    off, wave, weight, scalar = like.band_tables()
    ctx.set_bands(off, wave, weight, scalar)
    # per-source truths: T~U(8,25), beta~U(1.2,2.4), fnorm~logU(5,100);
    # model photometry of the truth through the product itself (single-source band fluxes)
    T = rng.uniform(8.0, 25.0, nsrc)
    beta = rng.uniform(1.2, 2.4, nsrc)
    fnorm = np.exp(rng.uniform(np.log(5.0), np.log(100.0), nsrc))
    truths = np.stack([T, beta, np.full(nsrc, truth[2]), np.full(nsrc, truth[3]), fnorm], axis=1)
    # band fluxes of the truths: evaluate lnlike machinery with zero data and unit ivar is not
    # invertible per band, so use the SED kernel on the band nodes and the host weights
    ctx.set_model(cfg["wavenorm"], cfg["opthin"], cfg["noalpha"])
    model = np.empty((nsrc, nb))
    freq = 299792458e-3 / wave
    step = 16384
    for i0 in range(0, nsrc, step):
        f, st = ctx.fnu(truths[i0:i0 + step], freq)
        for b in range(nb):
            model[i0:i0 + step, b] = (f[:, off[b]:off[b + 1]] * weight[off[b]:off[b + 1]]).sum(axis=1)
    flux, unc = synthetic.noisy_photometry(model, rng)
    ctx.set_data(flux, ivar=1.0 / unc**2)
    ctx.set_priors(like.lowlims, like.has_uplims, like.uplims, like.has_gpriors,
                   like.gprior_means, like.gprior_ivars)
    # walker positions ~ N(truth_s, sigma) clipped inside the limits, generated on the device
    n = nsrc * nw
    g = torch.Generator(device=dev)
    g.manual_seed(cfg["seed"] + 17 * rank)
    tr = torch.as_tensor(truths, dtype=torch.float64, device=dev).repeat_interleave(nw, dim=0)
    sig = torch.as_tensor(synthetic.P0_SIGMA, dtype=torch.float64, device=dev)
    P = tr + sig * torch.randn((n, 5), dtype=torch.float64, device=dev, generator=g)
    low = torch.as_tensor(np.asarray(like.lowlims) * 1.01, dtype=torch.float64, device=dev)
    P = torch.maximum(P, low)
    up = torch.as_tensor(np.where(like.has_uplims[:5], like.uplims[:5], np.inf) * 0.99,
                         dtype=torch.float64, device=dev)
    P = torch.minimum(P, up).contiguous()
    del tr
    return dict(ctx=ctx, n=n, nw=nw, nsrc=nsrc, P=P, like=like, cfg=cfg, flux=flux, unc=unc,
                truths=truths)


# ---------------------------------------------------------------------- CPU oracle
def _cpu_worker(args):
    (cfgname, flux, unc, P) = args
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import mbb_oracle as oracle
    from mbb_emcee_b200 import synthetic
    from mbb_emcee_b200.response import response_set
    cfg = synthetic.CONFIGS[cfgname if cfgname in synthetic.CONFIGS else "cfg2"]
    spec = oracle.LikeSpec(cfg["wavenorm"], cfg["noalpha"], cfg["opthin"])
    if cfg["response"]:
        wheel = response_set()
        spec.set_phot([oracle.band_from_response(wheel[nm]) for nm in cfg["bands"]], flux, unc)
        spec.auto_lambda0_uplim(max(wheel[nm].effective_wavelength for nm in cfg["bands"]))
    else:
        spec.set_phot(cfg["bands"], flux, unc)
        spec.auto_lambda0_uplim(max(cfg["bands"]))
    oracle.loglike(spec, P[0])                       # warm-up (imports, lazy loads)
    t0 = time.perf_counter()
    out = oracle.loglike_batch(spec, P)
    return time.perf_counter() - t0, len(P), float(np.sum(out[np.isfinite(out)]))


def cpu_baseline(cfgname, flux0, unc0, P, per_worker):
    """The oracle on every host core (mirrors emcee's threads= pool,
    reference mbb_fit.py:80-81): each worker loops the likelihood over its slice."""
    import multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import mbb_oracle as oracle
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count()
    P = np.asarray(P[:cores * per_worker], dtype=np.float64)
    jobs = [(cfgname, flux0, unc0, P[i * per_worker:(i + 1) * per_worker]) for i in range(cores)]
    jobs = [j for j in jobs if len(j[3])]
    ctxmp = mp.get_context("spawn")
    t0 = time.perf_counter()
    with ctxmp.Pool(len(jobs)) as pool:
        res = pool.map(_cpu_worker, jobs)
    wall = time.perf_counter() - t0
    slowest = max(r[0] for r in res)
    total = sum(r[1] for r in res)
    return {"value": total / slowest, "unit": UNIT, "cores": len(jobs),
            "kind": "reference" if oracle.native_kind() == "reference" else "port",
            "sample": "%d evals of the same workload (source 0 photometry, first %d walker rows), "
                      "%d per worker process, timed inside the workers (%.1f s incl. spawn)"
                      % (total, total, per_worker, wall),
            "us_per_eval_per_core": 1e6 * slowest / per_worker}


# ----------------------------------------------------------------------- reference arm
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from mbb_emcee_b200 import synthetic
    name = args.workload
    cfg = synthetic.CONFIGS[name]
    rng = np.random.RandomState(cfg["seed"])
    nb = len(cfg["bands"])
    truth = np.array([12.0, 1.8, 1300.0, 4.0, 30.0]) if name == "cfg5" else np.array(cfg["truth"])
    flux = rng.uniform(10.0, 60.0, nb)
    unc = np.maximum(0.1 * flux, 1.0)
    low = np.array([1, 0.1, 1, 0.1, 1e-3])
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count()
    per_worker = 20000 if name == "cfg5" else 1500
    times, vals, last = [], [], None
    for it in range(args.warmup + args.steps):
        P = synthetic.walker_cloud(truth, cores * per_worker, rng, low)
        cb = cpu_baseline(name, flux, unc, P, per_worker)
        if it >= args.warmup:
            vals.append(cb["value"])
            times.append(cores * per_worker / cb["value"])
        last = cb
    value = float(np.mean(vals))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * float(np.mean(times)), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": main_config(name),
            "cpu_baseline": dict(last, value=value),
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_config(name, n_per_rank, nsrc, nw):
    if name == "cfg5":
        return {"workload": "BASELINE configs[4]: batch fit of %d synthetic sources x %d walkers per "
                            "GPU, 6 delta bands (70,100,160,250,350,500 um), optically thin, no alpha, "
                            "wavenorm 500; one step = one likelihood pass over all walker positions"
                            % (nsrc, nw),
                "evals_per_step_per_gpu": int(n_per_rank), "bands": 6, "nodes": 6,
                "l2_policy": "inputs (%.2f GB of parameters per step) exceed the 126 MB L2"
                             % (n_per_rank * 40 / 1e9)}
    if name == "cfg5p":
        return {"workload": "BASELINE configs[4] on the tabulated band set: %d sources x %d walkers per GPU, "
                            "passband integration over PACS_100/160, SPIRE_250/350/500, SCUBA2_850 (1688 "
                            "nodes), optically thin, no alpha" % (nsrc, nw),
                "evals_per_step_per_gpu": int(n_per_rank), "bands": 6, "nodes": 1688,
                "l2_policy": "L2 flushed between timed steps by writing a 256 MB buffer"}
    return {"workload": "BASELINE configs[1]-style: %d sources x %d walkers per GPU, passband "
                        "integration over PACS_100/160, SPIRE_250/350/500, SCUBA2_850 (1688 nodes), "
                        "optically thick + alpha join" % (nsrc, nw),
            "evals_per_step_per_gpu": int(n_per_rank), "bands": 6, "nodes": 1688,
            "l2_policy": "L2 flushed between timed steps by writing a 256 MB buffer"}


def main_config(name, nsrc=None):
    """The `config` both arms report (the reference arm times a bounded sample of it)."""
    from mbb_emcee_b200 import synthetic
    if name == "cfg5":
        nsrc = nsrc or synthetic.CONFIGS["cfg5"]["nsources"]
        nw = synthetic.CONFIGS["cfg5"]["nwalkers"]
    elif name == "cfg5p":
        nsrc = nsrc or synthetic.CONFIGS["cfg5p"]["nsources"]
        nw = synthetic.CONFIGS["cfg5p"]["nwalkers"]
    else:
        nsrc, nw = nsrc or 2048, 512
    return workload_config(name, nsrc * nw, nsrc, nw)


# ------------------------------------------------------------------------- main arm
def run_b200(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # one process per GPU: stay on the GPU's NUMA node (pinned host buffers next to its PCIe port)
    all_cpus = sorted(os.sched_getaffinity(0))
    bound = None
    if world > 1:
        from mbb_emcee_b200.sharding import bind_to_gpu_numa_node
        pr = torch.cuda.get_device_properties(local)
        bound = bind_to_gpu_numa_node("%08x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id))
        dist.init_process_group("nccl", device_id=dev)
    name = args.workload
    W = build_workload(name, rank, args.nsrc)
    ctx, n, nw, P = W["ctx"], W["n"], W["nw"], W["P"]
    MODES = {"faithful": 0, "fast": 1, "gauss": 2}
    ctx.set_math_mode(MODES[args.math])
    out = torch.empty(n, dtype=torch.float64, device=dev)
    st = torch.empty(n, dtype=torch.int32, device=dev)
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev) if name != "cfg5" else None

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def device_leg(ctx, n, nw, P, out, st, flush, steps, warmup):
        """K timed likelihood passes with inputs resident in HBM; CUDA events on the stream the
        library launches on (torch only sees its own current stream, so wrap the handle)."""
        lstream = torch.cuda.ExternalStream(ctx.stream_handle(), device=dev)

        def step():
            ctx.loglike_device(n, P.data_ptr(), out.data_ptr(), st.data_ptr(), walkers_per_source=nw)

        for _ in range(warmup):
            step()
        ctx.sync()
        barrier()
        launches0 = ctx.launch_count()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
              for _ in range(steps)]
        barrier()
        t0 = time.perf_counter()
        for k in range(steps):
            if flush is not None:
                with torch.cuda.stream(lstream):
                    flush.fill_(1.0)
            ev[k][0].record(lstream)
            step()
            ev[k][1].record(lstream)
        ctx.sync()
        barrier()
        wall = time.perf_counter() - t0
        launches = ctx.launch_count() - launches0
        kernel_ms = [a.elapsed_time(b) for a, b in ev]
        step_ms = float(np.mean(kernel_ms))
        t = torch.tensor([step_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return dict(step_ms=step_ms, step_ms_max=float(t.item()), kernel_ms=kernel_ms, wall=wall,
                    launches=launches, lstream=lstream)

    # ---- device-resident throughput ------------------------------------------
    peak_tf = ctx.fp64_peak(20000)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    leg = device_leg(ctx, n, nw, P, out, st, flush, args.steps, args.warmup)
    lstream, launches, wall = leg["lstream"], leg["launches"], leg["wall"]
    kernel_ms, step_ms, step_ms_max = leg["kernel_ms"], leg["step_ms"], leg["step_ms_max"]
    nbad = int((st > 1).sum().item())
    nneg = int(torch.isneginf(out).sum().item())

    # ---- end-to-end through the host-buffer C-ABI call ------------------------
    e2e = None
    if not args.no_e2e:
        P_host = torch.empty((n, 5), dtype=torch.float64).pin_memory()
        out_host = torch.empty(n, dtype=torch.float64).pin_memory()
        P_host.copy_(P)
        torch.cuda.synchronize()
        Pn, on = P_host.numpy(), out_host.numpy()
        import ctypes
        lib = ctx._lib

        def step_host():
            rc = lib.mbb_loglike(ctx._h, n, ctypes.c_void_p(Pn.ctypes.data), 0, None, nw,
                                 ctypes.c_void_p(on.ctypes.data), None, 0)
            if rc != 0:
                raise RuntimeError(lib.mbb_last_error().decode())

        e2e_steps = max(1, min(args.steps, 5))
        l0 = ctx.launch_count()
        step_host()
        e2e_launches = ctx.launch_count() - l0
        barrier()
        e2e_ms = []
        for _ in range(e2e_steps):
            t1 = time.perf_counter()
            step_host()                      # synchronous: returns with results in host memory
            e2e_ms.append(1e3 * (time.perf_counter() - t1))
        barrier()
        same = bool(np.array_equal(on[:100000], out[:100000].cpu().numpy()))
        te = torch.tensor([float(np.mean(e2e_ms))], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e_ms_max = float(te.item())
        e2e = {"value": n * world / (e2e_ms_max * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": int(n * 40), "d2h_bytes_per_step": int(n * 8),
               "ms_per_step": e2e_ms_max, "steps": e2e_steps,
               "gpu_launches_per_step": int(e2e_launches),
               "api": "mbb_loglike(MBB_HOST) with pinned host arrays; pipelined H2D/kernel/D2H",
               "matches_device_path": same}
    # ---- batch fit: the device-resident ensemble sampler on the same workload ----
    batch = None
    if not args.no_e2e:
        import ctypes
        K = 10
        lnp = torch.empty(n, dtype=torch.float64, device=dev)
        nacc = torch.zeros(n, dtype=torch.int32, device=dev)
        Pw = P.clone()
        ctx.ensemble_run_device(W["nsrc"], nw, 2, Pw.data_ptr(), lnp.data_ptr(), False, seed=7,
                                naccept_ptr=nacc.data_ptr())          # warm-up (+ initial lnprob)
        ctx.sync()
        barrier()
        l0 = ctx.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(lstream)
        ctx.ensemble_run_device(W["nsrc"], nw, K, Pw.data_ptr(), lnp.data_ptr(), True, seed=7, step0=2,
                                naccept_ptr=nacc.data_ptr())
        e1.record(lstream)
        ctx.sync()
        barrier()
        dev_ms = e0.elapsed_time(e1)
        samp_launches = ctx.launch_count() - l0
        acc_frac = float(nacc.double().mean().item()) / K
        # same through host buffers: positions up, K iterations, positions + lnprob down
        Ph = torch.empty((n, 5), dtype=torch.float64).pin_memory()
        Lh = torch.empty(n, dtype=torch.float64).pin_memory()
        Ph.copy_(P)
        Phn, Lhn = Ph.numpy(), Lh.numpy()
        torch.cuda.synchronize()

        def fit_host(k):
            rc = ctx._lib.mbb_ensemble_run(ctx._h, W["nsrc"], nw, k, 2.0, 7, 0,
                                           ctypes.c_void_p(Phn.ctypes.data), ctypes.c_void_p(Lhn.ctypes.data),
                                           0, None, None, None, None, 1, 0)
            if rc != 0:
                raise RuntimeError(ctx._lib.mbb_last_error().decode())

        fit_host(1)
        barrier()
        t1 = time.perf_counter()
        fit_host(K)
        host_ms = 1e3 * (time.perf_counter() - t1)
        barrier()
        tb = torch.tensor([dev_ms, host_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tb, op=dist.ReduceOp.MAX)
        dev_ms, host_ms = float(tb[0].item()), float(tb[1].item())
        batch = {"api": "batch_fitter / mbb_ensemble_run: stretch-move sampler resident on the device "
                        "(Philox draws, proposal, likelihood, accept/reject fused per half-step)",
                 "iterations_per_call": K, "evals_per_iteration": int(n * world),
                 "device_resident": {"value": n * world * K / (dev_ms * 1e-3), "unit": UNIT,
                                     "ms_per_iteration": dev_ms / K, "gpu_launches": int(samp_launches)},
                 "e2e": {"value": n * world * (K + 1) / (host_ms * 1e-3), "unit": UNIT,
                         "ms_per_call": host_ms, "h2d_bytes_per_call": int(n * 40),
                         "d2h_bytes_per_call": int(n * 48),
                         "note": "host call: walker positions uploaded, initial log-probability + K "
                                 "iterations, positions and log-probabilities downloaded"},
                 "mean_acceptance_fraction": acc_frac}
    # ---- tabulated passband sets: configs[1] (thick + alpha) and configs[4]-secondary (thin) ----
    def passband_leg(wname):
        W2 = build_workload(wname, rank, None)
        ctx2, n2, P2 = W2["ctx"], W2["n"], W2["P"]
        ctx2.set_math_mode(MODES[args.math])
        out2 = torch.empty(n2, dtype=torch.float64, device=dev)
        st2 = torch.empty(n2, dtype=torch.int32, device=dev)
        flush2 = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
        k = max(3, min(args.steps, 5))
        leg2 = device_leg(ctx2, n2, W2["nw"], P2, out2, st2, flush2, k, 3)
        f2 = FLOP_PER_EVAL[wname]
        # the same launches with the tabulated bands' 32-point Gauss rules (MBB_MATH_FAST_GAUSS)
        gauss = None
        if args.math == "fast":
            ref_out = out2.clone()
            ctx2.set_math_mode(MODES["gauss"])
            leg3 = device_leg(ctx2, n2, W2["nw"], P2, out2, st2, flush2, k, 3)
            fin = torch.isfinite(ref_out)
            rel = ((out2[fin] - ref_out[fin]).abs() / ref_out[fin].abs()).max().item()
            gauss = {"value": n2 * world / (leg3["step_ms_max"] * 1e-3), "unit": UNIT,
                     "ms_per_step": leg3["step_ms_max"],
                     "max_rel_diff_vs_full_tables": rel,
                     "note": "math mode gauss: per (walker, band) the band's 32-point Gauss rule where a "
                             "per-walker bound shows it agrees with the full table to rounding"}
        res = {"workload": main_config(wname, W2["nsrc"])["workload"],
               "value": n2 * world / (leg2["step_ms_max"] * 1e-3), "unit": UNIT,
               "ms_per_step": leg2["step_ms_max"], "evals_per_step_per_gpu": int(n2),
               "roofline_frac": n2 * f2 / (leg2["step_ms"] * 1e-3) / 1e12 / peak_tf,
               "flop_per_eval": f2, "status_errors": int((st2 > 1).sum().item()),
               "gauss_rules": gauss,
               "l2_policy": "L2 flushed between timed steps by writing a 256 MB buffer"}
        del W2, ctx2, P2, out2, st2, flush2
        torch.cuda.empty_cache()
        return res

    passband = passband_batch = None
    per_worker = 20000 if name == "cfg5" else 1500
    cores = len(all_cpus)
    Pc = P[:cores * per_worker].cpu().numpy() if rank == 0 else None
    if name == "cfg5" and not args.no_passband:
        del P, out, st
        torch.cuda.empty_cache()
        passband = passband_leg("cfg2")
        passband_batch = passband_leg("cfg5p")
    clocks = sampler.stop()

    if rank == 0:
        total = n * world
        value = total / (step_ms_max * 1e-3)
        flop = FLOP_PER_EVAL[name]
        achieved_tf = n * flop / (step_ms * 1e-3) / 1e12
        hbm_peak = None
        try:
            hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
        except Exception:
            hbm_peak = 6650.0
        achieved_gbs = n * BYTES_PER_EVAL / (step_ms * 1e-3) / 1e9
        cb = None
        try:
            if args.no_cpu_baseline:
                raise RuntimeError("skipped (--no-cpu-baseline)")
            os.sched_setaffinity(0, all_cpus)      # the CPU leg uses every host core again
            cb = cpu_baseline(name, W["flux"][0], W["unc"][0], Pc, per_worker)
        except Exception as exc:           # the baseline is informational; never hide the GPU line
            cb = {"value": None, "unit": UNIT, "cores": 0, "kind": "port",
                  "sample": "failed: %r" % (exc,)}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": step_ms_max, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": dict(main_config(name, W["nsrc"]), parallelism="sources sharded, "
                           "%d rank(s), no data-path collective" % world,
                           numa_binding=("rank 0 bound to %d CPUs of its GPU's NUMA node" % len(bound))
                           if bound else "none",
                           math_mode=args.math),
            "roofline": {"bound": "fp64", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                         "frac": achieved_tf / peak_tf, "traffic": ncu_fact(name, "dram_bytes_per_launch"),
                         "ncu": ncu_fact(name, None),
                         "note": "achieved = %d algorithmic FP64 flop/eval (reference formulation, "
                                 "SURVEY 8d) x evals per launch / CUDA-event kernel time; peak = DFMA "
                                 "rate measured live by mbb_fp64_peak on this GPU (no tensor cores: "
                                 "exp/pow-bound FP64)" % flop,
                         "hbm": {"achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s",
                                 "frac": achieved_gbs / hbm_peak, "peak_source": "measured"
                                 if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json"))
                                 else "fallback"}},
            "cpu_baseline": cb,
            "e2e": e2e,
            "batch_fit": batch,
            "passband": passband,
            "passband_batch": passband_batch,
            "gpu_launches": int(launches),
            "clocks": clocks,
            "check": {"status_errors": nbad, "neg_inf": nneg, "wall_s_timed_region": wall,
                      "kernel_ms_each": [round(x, 4) for x in kernel_ms]},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg5", choices=["cfg5", "cfg2", "cfg5p"])
    ap.add_argument("--nsrc", type=int, default=None, help="sources per GPU (default: workload's)")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (profiling runs)")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU oracle leg")
    ap.add_argument("--no-passband", action="store_true", help="skip the tabulated-passband leg")
    ap.add_argument("--math", default="fast", choices=["fast", "faithful", "gauss"],
                    help="arithmetic mode: fast (default), faithful (reference order, libdevice), "
                         "gauss (fast + 32-point Gauss rules for tabulated passbands)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()

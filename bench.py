#!/usr/bin/env python
"""Benchmark of the modified-blackbody log-likelihood hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                    [--workload cfg5|cfg2|cfg5p|cfg3|default_model] [--scaling weak|strong]

Metric (BASELINE.json): MBB log-likelihood evals/sec (walker-SEDs/s).  One eval
= one likelihood.__call__-equivalent (5 parameters in, 1 log-probability out,
incl. per-walker setup, all bands x nodes, chi-square, limits and priors).

Workload (default): BASELINE configs[4] -- batch fit of 1e5 synthetic sources x
512 walkers on the cfg1 band set (6 delta bands 70..500 um, optically thin, no
alpha).  One step = one pass of the likelihood over all 5.12e7 walker positions
of this rank's sources.  Under torchrun every rank owns its own 1e5 sources
(`--scaling weak`, the default) or 1e5/N of them (`--scaling strong`); no
data-path collective either way.

One JSON line on stdout (rank 0):
  value      device-resident likelihood throughput (inputs in HBM), CUDA events;
  roofline   executed FP64 work / measured DFMA peak for that kernel (instruction
             counts from the committed ncu capture of the same build, time live);
  e2e        the batch FIT through the host-buffer C-ABI call: photometry and
             starting ensembles up from pinned host memory, burn-in + main run
             of the device-resident sampler, final ensembles and per-source
             posterior summaries back down -- every copy inside the timed region;
  e2e_loglike  one likelihood pass through host buffers (PCIe-bound);
  batch_fit / default_model / passband / passband_batch / cfg3 / cfg4 / fits:
             the other BASELINE configurations, each with its own numbers;
  cpu_baseline  the reference's own likelihood.__call__ (baseline/_ref or
             /root/reference; else the oracle port) on the box's host cores.
`--impl reference` times that CPU arm alone on the same `config`.
"""
import argparse
import ctypes
import hashlib
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
MODES = {"faithful": 0, "fast": 1, "gauss": 2}

# Algorithmic FP64 work per evaluation: SURVEY.md 8(d) -- op counts of the
# REFERENCE formulation with CUDA-12.9 libdevice on sm_100a (pow 125, expm1 35,
# exp 31, log 47, div 20, FMA = 2 flops).  Reported as `algorithmic`, beside the
# executed-work roofline fraction.
FLOP_PER_EVAL = {"cfg5": 6 * 182 + 225 + 18 + 40,            # = 1375 (cfg1 band set)
                 "cfg2": 1688 * 366 + 3600 + 58,             # = 621466
                 "cfg5p": 1688 * 184 + 225 + 18 + 40,        # = 310875 (thin, no alpha, cfg2 band set)
                 "cfg3": 1342 * 366 + 3600 + 144 + 40 + 2 * 8000,    # = 510956
                 "default_model": 6 * 366 + 3600 + 18 + 40}  # = 5854 (thick + alpha, 6 delta bands)
BYTES_PER_EVAL = 48                                          # 40 B parameters in + 8 B out


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ----------------------------------------------------------------------- ncu facts
def csrc_sha():
    """Hash of every source the CUDA library is built from: what an ncu capture is valid for."""
    h = hashlib.sha256()
    d = os.path.join(ROOT, "mbb_emcee_b200", "csrc")
    names = sorted(f for f in os.listdir(d) if f.endswith((".cu", ".cuh", ".h", ".inc")))
    for f in names + [os.path.join("..", "..", "include", "mbb_b200.h")]:
        h.update(f.encode())
        h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()[:16]


_FACTS = None


def ncu_facts(key):
    """(facts, stale): figures of the committed `ncu --set full` capture of one kernel
    (profiles/ncu_facts.json, written on the GPU box by tools/ncu_capture.py together with the hash
    of the sources it was built from); stale = the library has changed since."""
    global _FACTS
    if _FACTS is None:
        try:
            _FACTS = json.load(open(os.path.join(ROOT, "profiles", "ncu_facts.json")))
        except Exception:
            _FACTS = {}
    k = _FACTS.get("kernels", {}).get(key)
    return k, (_FACTS.get("csrc_sha") != csrc_sha())


def executed_roofline(key, units, step_ms, peak_tf, flop_per_unit=None):
    """Roofline of one kernel on EXECUTED work: FP64 warp instructions per unit (ncu capture) x
    units x 64 flop (every FP64 instruction counted as a 32-lane FMA) / live CUDA-event time,
    against the DFMA rate measured live (mbb_fp64_peak)."""
    f, stale = ncu_facts(key)
    out = {"bound": "fp64", "peak": peak_tf, "unit": "TFLOP/s", "ncu_stale": bool(stale) if f else None}
    if f and f.get("units_per_launch"):
        per_unit = f["fp64_warp_instructions"] / f["units_per_launch"]
        ach = per_unit * units * 64.0 / (step_ms * 1e-3) / 1e12
        out.update(achieved=ach, frac=ach / peak_tf,
                   fp64_warp_instructions_per_unit=per_unit,
                   traffic=f["dram_bytes_per_launch"] / f["units_per_launch"] * units,
                   ncu={k: f.get(k) for k in ("kernel", "gpu_time_ms", "fp64_pipe_pct", "issue_slot_pct",
                                              "dram_pct", "warp_instructions", "units_per_launch",
                                              "issue_bound_frac", "registers", "local_ld", "local_st",
                                              "long_scoreboard", "barrier")})
    else:
        out.update(achieved=None, frac=None, traffic=None, ncu=None,
                   note="no ncu capture for this kernel in profiles/ncu_facts.json")
    if flop_per_unit:
        alg = units * flop_per_unit / (step_ms * 1e-3) / 1e12
        out["algorithmic"] = {"flop_per_unit": flop_per_unit, "achieved": alg, "ratio": alg / peak_tf,
                              "note": "FP64 flops of the REFERENCE formulation (libdevice op weights, SURVEY "
                                      "8d) per unit / time: exceeds 1 where FAST removes work algebraically"}
    return out


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
def workload_cfg(name):
    from mbb_emcee_b200 import synthetic
    if name == "default_model":
        cfg = dict(synthetic.CONFIGS["cfg5"], opthin=False, noalpha=False, truth=(14.0, 1.8, 400.0, 3.0, 30.0))
        return cfg, cfg["nsources"], cfg["nwalkers"]
    cfg = synthetic.CONFIGS[name]
    if name in ("cfg5", "cfg5p"):
        return cfg, cfg["nsources"], cfg["nwalkers"]
    return cfg, (1024 if name == "cfg3" else 2048), 512


def build_workload(name, rank, nsrc_override=None, with_walkers=True):
    """One rank's synthetic problem on the device: context staged with bands, per-source photometry
    (truth photometry through the product's own SED kernel + seeded noise), limits/priors, and the
    walker positions P[n][5] ~ N(truth_s, sigma) inside the limits (generated on the device)."""
    import torch
    from mbb_emcee_b200 import _native, likelihood, synthetic
    cfg, nsrc, nw = workload_cfg(name)
    nsrc = nsrc_override or nsrc
    dev = torch.device("cuda", torch.cuda.current_device())
    truth = np.array(cfg.get("truth", (12.0, 1.8, 1300.0, 4.0, 30.0)))
    rng = np.random.RandomState(cfg["seed"] + 1000 * rank)
    ctx = _native.Context(dev.index)
    ctx.set_model(cfg["wavenorm"], cfg["opthin"], cfg["noalpha"])
    # one likelihood object is used only for its host-side table builder and its limit bookkeeping
    like = likelihood(wavenorm=cfg["wavenorm"], noalpha=cfg["noalpha"], opthin=cfg["opthin"],
                      response=cfg["response"], device=dev.index)
    nb = len(cfg["bands"])
    like.set_phot(cfg["bands"], np.ones(nb), np.ones(nb))
    for nm, v in cfg.get("uplims", []):
        like.set_uplim(nm, v)
    for nm, m, s in cfg.get("gpriors", []):
        like.set_gaussian_prior(nm, m, s)
    off, wave, weight, scalar = like.band_tables()
    ctx.set_bands(off, wave, weight, scalar)
    # per-source truths: T~U(8,25), beta~U(1.2,2.4), fnorm~logU(5,100)
    T = rng.uniform(8.0, 25.0, nsrc)
    beta = rng.uniform(1.2, 2.4, nsrc)
    fnorm = np.exp(rng.uniform(np.log(5.0), np.log(100.0), nsrc))
    truths = np.stack([T, beta, np.full(nsrc, truth[2]), np.full(nsrc, truth[3]), fnorm], axis=1)
    model = np.empty((nsrc, nb))
    freq = 299792458e-3 / wave
    step = 16384
    for i0 in range(0, nsrc, step):
        f, st = ctx.fnu(truths[i0:i0 + step], freq)
        for b in range(nb):
            model[i0:i0 + step, b] = (f[:, off[b]:off[b + 1]] * weight[off[b]:off[b + 1]]).sum(axis=1)
    flux, unc = synthetic.noisy_photometry(model, rng)
    cinv = None
    if name == "cfg3":
        # C = diag(sigma^2) + 0.05^2 f f^T + rho-block(SPIRE, rho = 0.3), per source
        cov = np.stack([synthetic.cfg3_covariance(flux[s], unc[s]) for s in range(nsrc)])
        cinv = np.linalg.inv(cov)
        ctx.set_data(flux, cinv=cinv)
    else:
        ctx.set_data(flux, ivar=1.0 / unc**2)
    ctx.set_priors(like.lowlims, like.has_uplims, like.uplims, like.has_gpriors,
                   like.gprior_means, like.gprior_ivars)
    n = nsrc * nw
    P = None
    if with_walkers:
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
    return dict(ctx=ctx, n=n, nw=nw, nsrc=nsrc, P=P, like=like, cfg=cfg, flux=flux, unc=unc, cinv=cinv,
                truths=truths)


def workload_config(name, nsrc, nw):
    n = nsrc * nw
    if name == "cfg5":
        return {"workload": "BASELINE configs[4]: batch fit of %d synthetic sources x %d walkers per "
                            "GPU, 6 delta bands (70,100,160,250,350,500 um), optically thin, no alpha, "
                            "wavenorm 500; one step = one likelihood pass over all walker positions"
                            % (nsrc, nw),
                "evals_per_step_per_gpu": int(n), "bands": 6, "nodes": 6,
                "l2_policy": "inputs (%.2f GB of parameters per step) exceed the 126 MB L2" % (n * 40 / 1e9)}
    if name == "default_model":
        return {"workload": "BASELINE configs[4] with the reference's default model: %d sources x %d walkers "
                            "per GPU, 6 delta bands, optically thick + alpha join" % (nsrc, nw),
                "evals_per_step_per_gpu": int(n), "bands": 6, "nodes": 6,
                "l2_policy": "inputs (%.2f GB of parameters per step) exceed the 126 MB L2" % (n * 40 / 1e9)}
    if name == "cfg5p":
        return {"workload": "BASELINE configs[4] on the tabulated band set: %d sources x %d walkers per GPU, "
                            "passband integration over PACS_100/160, SPIRE_250/350/500, SCUBA2_850 (1688 "
                            "nodes), optically thin, no alpha" % (nsrc, nw),
                "evals_per_step_per_gpu": int(n), "bands": 6, "nodes": 1688,
                "l2_policy": "L2 flushed between timed steps by writing a 256 MB buffer"}
    if name == "cfg3":
        return {"workload": "BASELINE configs[2]-style: %d sources x %d walkers per GPU, SPIRE_250/350/500, "
                            "SCUBA2_850, ALMA_alma_345/230, SMA_dsb_230_8_2, PdBI_box_135_3.6 (1342 nodes), "
                            "full 8x8 covariance per source, T and lambda_peak soft limits, beta and "
                            "lambda_peak Gaussian priors, optically thick + alpha" % (nsrc, nw),
                "evals_per_step_per_gpu": int(n), "bands": 8, "nodes": 1342,
                "l2_policy": "L2 flushed between timed steps by writing a 256 MB buffer"}
    return {"workload": "BASELINE configs[1]-style: %d sources x %d walkers per GPU, passband "
                        "integration over PACS_100/160, SPIRE_250/350/500, SCUBA2_850 (1688 nodes), "
                        "optically thick + alpha join" % (nsrc, nw),
            "evals_per_step_per_gpu": int(n), "bands": 6, "nodes": 1688,
            "l2_policy": "L2 flushed between timed steps by writing a 256 MB buffer"}


def main_config(name, nsrc=None):
    """The `config` both arms report (the reference arm times a bounded sample of it)."""
    _, n0, nw = workload_cfg(name)
    return workload_config(name, nsrc or n0, nw)


# ------------------------------------------------------------------ CPU arms (rank 0)
def _cpu_worker(args):
    """One worker process: the reference's own likelihood.__call__ when the reference can be
    imported here (its tree, or its pip install under baseline/_ref), else the oracle port."""
    (name, rows, want_ref) = args
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from mbb_emcee_b200 import synthetic
    cfg, flux, unc, cov, _ = synthetic.sample_problem(name, 1)
    # rows: an array, or (count, seed) -- then this worker draws its own slice of the sample
    P = rows if isinstance(rows, np.ndarray) else synthetic.sample_problem(name, rows[0], rows[1])[4]
    kind, where, like = "port", None, None
    if want_ref:
        import ref_harness
        ref, where = ref_harness.import_any_reference()
        if ref is not None:
            like = ref.likelihood(wavenorm=cfg["wavenorm"], noalpha=cfg["noalpha"], opthin=cfg["opthin"],
                                  response=cfg["response"])
            like.set_phot(cfg["bands"], flux, unc)
            if cov is not None:
                like.set_cov(cov)
            for nm, v in cfg.get("uplims", []):
                like.set_uplim(nm, v)
            for nm, m, s in cfg.get("gpriors", []):
                like.set_gaussian_prior(nm, m, s)
            kind = "reference"
    if like is None:
        import mbb_oracle as oracle
        from mbb_emcee_b200.response import response_set
        spec = oracle.LikeSpec(cfg["wavenorm"], cfg["noalpha"], cfg["opthin"])
        if cfg["response"]:
            wheel = response_set()
            for nm in cfg["bands"]:
                if nm not in wheel:
                    wheel.add_special(nm)
            spec.set_phot([oracle.band_from_response(wheel[nm]) for nm in cfg["bands"]], flux, unc)
            spec.auto_lambda0_uplim(max(wheel[nm].effective_wavelength for nm in cfg["bands"]))
        else:
            spec.set_phot(cfg["bands"], flux, unc)
            spec.auto_lambda0_uplim(max(cfg["bands"]))
        if cov is not None:
            spec.set_cov(cov)
        like = lambda p: oracle.loglike(spec, p)          # noqa: E731
        where = "oracle/mbb_oracle.py (restatement; node loops in %s)" % oracle.native_kind()
    like(P[0])                                           # warm-up (imports, lazy loads)
    out = np.empty(len(P))
    t0 = time.perf_counter()
    for i in range(len(P)):
        out[i] = like(P[i])                              # one call per walker, as emcee makes them
    return time.perf_counter() - t0, out, kind, where


def cpu_arm(name, per_worker):
    """The CPU likelihood on every host core (mirrors emcee's threads= pool, reference
    mbb_fit.py:80-81): each worker process loops likelihood.__call__ over its slice of the
    bounded sample (synthetic.sample_problem: same data and rows in every arm)."""
    import multiprocessing as mp
    from mbb_emcee_b200 import synthetic
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count()
    cfg = synthetic.sample_problem(name, 1)[0]
    jobs = [(name, (per_worker, cfg["seed"] + 1000 + i), True) for i in range(cores)]
    ctxmp = mp.get_context("spawn")
    t0 = time.perf_counter()
    with ctxmp.Pool(len(jobs)) as pool:
        res = pool.map(_cpu_worker, jobs)
    wall = time.perf_counter() - t0
    slowest = max(r[0] for r in res)
    total = sum(len(r[1]) for r in res)
    P = synthetic.sample_problem(name, per_worker, cfg["seed"] + 1000)[4]       # worker 0's rows
    out = res[0][1]
    return {"value": total / slowest, "unit": UNIT, "cores": len(jobs), "kind": res[0][2],
            "implementation": res[0][3],
            "sample": "%d evals: synthetic.sample_problem(%r) -- one representative source, %d walker rows "
                      "per worker process, one likelihood call per row, timed inside the workers "
                      "(%.1f s incl. spawn)" % (total, name, per_worker, wall),
            "us_per_eval_per_core": 1e6 * slowest / per_worker, "evals": total}, P, out


# rows per worker process of one CPU step: ~5 s of work per core
PER_WORKER = {"cfg5": 400000, "default_model": 100000, "cfg2": 30000, "cfg5p": 50000, "cfg3": 25000}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    name = args.workload
    per_worker = PER_WORKER[name]
    vals, times, last = [], [], None
    for it in range(args.warmup + args.steps):
        cb, P, _ = cpu_arm(name, per_worker)
        if it >= args.warmup:
            vals.append(cb["value"])
            times.append(cb["evals"] / cb["value"])
        last = cb
    value = float(np.mean(vals))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * float(np.mean(times)), "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": main_config(name, args.nsrc),
            "cpu_baseline": dict(last, value=value),
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------- main arm
def run_b200(args):
    import torch
    import torch.distributed as dist
    from mbb_emcee_b200 import _native, synthetic
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # one process per GPU: stay on the GPU's NUMA node where the topology is exposed, else spread the
    # ranks over disjoint CPU sets (pinned staging and the copy threads of different ranks apart)
    all_cpus = sorted(os.sched_getaffinity(0))
    bound = None
    if world > 1:
        from mbb_emcee_b200.sharding import bind_rank_cpus
        pr = torch.cuda.get_device_properties(local)
        bound = bind_rank_cpus("%08x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id),
                               local, world)
        dist.init_process_group("nccl", device_id=dev)
    name = args.workload
    _, nsrc_cfg, _ = workload_cfg(name)
    nsrc_rank = args.nsrc or (nsrc_cfg if args.scaling == "weak" else max(nsrc_cfg // world, 1))
    t_setup = time.perf_counter()
    W = build_workload(name, rank, nsrc_rank)
    ctx, n, nw, P = W["ctx"], W["n"], W["nw"], W["P"]
    ctx.set_math_mode(MODES[args.math])
    out = torch.empty(n, dtype=torch.float64, device=dev)
    st = torch.empty(n, dtype=torch.int32, device=dev)
    delta = name in ("cfg5", "default_model")
    flush = None if delta else torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
    log("[bench] rank %d: workload %s built in %.1f s" % (rank, name, time.perf_counter() - t_setup))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(*vals):
        t = torch.tensor(list(vals), dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(x) for x in t.tolist()]

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
        return dict(step_ms=step_ms, step_ms_max=allmax(step_ms)[0], kernel_ms=kernel_ms, wall=wall,
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
    lib = ctx._lib
    vp = ctypes.c_void_p

    # ---- one likelihood pass through host buffers (PCIe-bound) -----------------
    e2e_loglike = None
    if not args.no_e2e:
        P_host = torch.empty((n, 5), dtype=torch.float64).pin_memory()
        out_host = torch.empty(n, dtype=torch.float64).pin_memory()
        P_host.copy_(P)
        torch.cuda.synchronize()
        Pn, on = P_host.numpy(), out_host.numpy()

        def step_host():
            rc = lib.mbb_loglike(ctx._h, n, vp(Pn.ctypes.data), 0, None, nw, vp(on.ctypes.data), None, 0)
            if rc != 0:
                raise RuntimeError(lib.mbb_last_error().decode())

        k2 = max(1, min(args.steps, 3))
        l0 = ctx.launch_count()
        step_host()
        e2e_launches = ctx.launch_count() - l0
        barrier()
        ms = []
        for _ in range(k2):
            t1 = time.perf_counter()
            step_host()                      # synchronous: returns with results in host memory
            ms.append(1e3 * (time.perf_counter() - t1))
        barrier()
        same = bool(np.array_equal(on[:100000], out[:100000].cpu().numpy()))
        ms_max = allmax(float(np.mean(ms)))[0]
        e2e_loglike = {"value": n * world / (ms_max * 1e-3), "unit": UNIT,
                       "h2d_bytes_per_step": int(n * 40), "d2h_bytes_per_step": int(n * 8),
                       "ms_per_step": ms_max, "steps": k2, "gpu_launches_per_step": int(e2e_launches),
                       "h2d_gb_per_s_per_gpu": n * 40 / (ms_max * 1e-3) / 1e9,
                       "api": "mbb_loglike(MBB_HOST) with pinned host arrays; pipelined H2D/kernel/D2H",
                       "matches_device_path": same}
        # the same pass with the fixed parameters declared (mbb_set_fixed_params; reference
        # mbb_fit.fix_param, e.g. alpha under --noalpha and lambda0 for an optically thin fit:
        # one value in every walker): SoA host block, fixed columns never cross PCIe
        if W["cfg"]["opthin"] and W["cfg"]["noalpha"]:
            fixed, fvals = (0, 0, 1, 1, 0), (0.0, 0.0, 1300.0, 4.0, 0.0)
            P2 = P.clone()
            P2[:, 2], P2[:, 3] = fvals[2], fvals[3]
            out2 = torch.empty_like(out)
            ctx.loglike_device(n, P2.data_ptr(), out2.data_ptr(), 0, walkers_per_source=nw)
            ctx.sync()
            T_host = torch.empty((5, n), dtype=torch.float64).pin_memory()
            T_host.copy_(P2.t())
            T_host[2].fill_(float("nan"))          # fixed columns of the host block are never read
            T_host[3].fill_(float("nan"))
            torch.cuda.synchronize()
            Tn = T_host.numpy()
            ctx.set_fixed_params(fixed, fvals)

            def step_fixed():
                rc = lib.mbb_loglike(ctx._h, n, vp(Tn.ctypes.data), 1, None, nw, vp(on.ctypes.data), None, 0)
                if rc != 0:
                    raise RuntimeError(lib.mbb_last_error().decode())

            step_fixed()
            barrier()
            msf = []
            for _ in range(k2):
                t1 = time.perf_counter()
                step_fixed()
                msf.append(1e3 * (time.perf_counter() - t1))
            barrier()
            ctx.set_fixed_params(None)
            same2 = bool(np.array_equal(on[:100000], out2[:100000].cpu().numpy()))
            msf_max = allmax(float(np.mean(msf)))[0]
            e2e_loglike["fixed_columns"] = {
                "value": n * world / (msf_max * 1e-3), "unit": UNIT, "ms_per_step": msf_max,
                "h2d_bytes_per_step": int(n * 24), "d2h_bytes_per_step": int(n * 8),
                "h2d_gb_per_s_per_gpu": n * 24 / (msf_max * 1e-3) / 1e9,
                "api": "mbb_set_fixed_params(lambda0, alpha) + mbb_loglike(MBB_HOST, MBB_SOA): 24 instead of 40 "
                       "bytes per evaluation host-to-device",
                "matches_device_path": same2}
            del P2, out2, T_host, Tn
        del P_host, out_host, Pn, on

    # ---- the batch fit: device-resident sampler, resident in HBM and through host buffers --------
    batch, e2e = None, None
    if delta and not args.no_e2e:
        nsrc = W["nsrc"]
        nburn, nmain, thin = args.fit_burn, args.fit_steps, args.fit_thin
        K = nburn + nmain
        evals_fit = n * (K + 1)                               # starting ensembles + K iterations
        lnp = torch.empty(n, dtype=torch.float64, device=dev)
        nacc = torch.zeros(n, dtype=torch.int32, device=dev)
        stats = torch.zeros((nsrc, _native.FIT_NSTATS), dtype=torch.float64, device=dev)
        Pw = P.clone()
        ctx.ensemble_fit_device(nsrc, nw, 0, 2, Pw.data_ptr(), lnp.data_ptr(), False, seed=7,
                                naccept_ptr=nacc.data_ptr())          # warm-up
        ctx.sync()
        Pw.copy_(P)
        barrier()
        l0 = ctx.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(lstream)
        ctx.ensemble_fit_device(nsrc, nw, nburn, nmain, Pw.data_ptr(), lnp.data_ptr(), False, seed=7,
                                src0=rank * nsrc, naccept_ptr=nacc.data_ptr(), stats_ptr=stats.data_ptr(),
                                thin=thin)
        e1.record(lstream)
        ctx.sync()
        barrier()
        dev_ms = allmax(e0.elapsed_time(e1))[0]
        samp_launches = ctx.launch_count() - l0
        acc_frac = float(stats[:, _native.FS_ACC].mean().item())
        stats_dev = stats.cpu().numpy()
        roof_s = executed_roofline("ens", n * K, dev_ms, peak_tf, FLOP_PER_EVAL[name])
        batch = {"api": "mbb_ensemble_fit(MBB_DEVICE): source-resident stretch-move sampler (ensembles in "
                        "shared memory for all iterations of the launch; Philox draws, proposal, likelihood, "
                        "accept/reject and posterior summaries in one kernel)",
                 "burn_in": nburn, "iterations": nmain, "thin": thin,
                 "value": evals_fit * world / (dev_ms * 1e-3), "unit": UNIT,
                 "ms_per_iteration": dev_ms / K, "ms_per_fit": dev_ms, "gpu_launches": int(samp_launches),
                 "sources_per_s": nsrc * world / (dev_ms * 1e-3),
                 "mean_acceptance_fraction": acc_frac, "roofline": roof_s}
        del Pw, lnp, nacc, stats
        # the same fit as a user of the drop-in runs it: host arrays in, host arrays out
        pin = _native.pinned_empty
        h_pos, h_p0 = pin((nsrc, nw, 5)), pin((nsrc, nw, 5))
        h_lnp, h_nacc, h_st = pin((nsrc, nw)), pin((nsrc, nw), np.int32), pin((nsrc, nw), np.int32)
        h_stats = pin((nsrc, _native.FIT_NSTATS))
        torch.from_numpy(h_p0).copy_(P.view(nsrc, nw, 5))
        torch.cuda.synchronize()
        flux_h, ivar_h = np.ascontiguousarray(W["flux"]), np.ascontiguousarray(1.0 / W["unc"]**2)

        def fit_host():
            h_pos[...] = h_p0                            # (host-side copy: the call updates pos in place)
            t1 = time.perf_counter()
            ctx.set_data(flux_h, ivar=ivar_h)            # this step's photometry goes up too
            ctx.ensemble_fit_into(h_pos, h_lnp, nburn, nmain, naccept=h_nacc, status=h_st, stats=h_stats,
                                  seed=7, src0=rank * nsrc, thin=thin)
            return 1e3 * (time.perf_counter() - t1)

        fit_host()
        barrier()
        k3 = max(1, min(args.steps, 3))
        l0 = ctx.launch_count()
        ms = []
        for _ in range(k3):
            ms.append(fit_host())
        fit_launches = (ctx.launch_count() - l0) // k3
        barrier()
        host_ms = allmax(float(np.mean(ms)))[0]
        h2d = int(n * 40 + flux_h.nbytes + ivar_h.nbytes)
        d2h = int(n * 40 + n * 8 + n * 4 + n * 4 + h_stats.nbytes)
        same_fit = bool(np.array_equal(h_stats, stats_dev))
        e2e = {"value": evals_fit * world / (host_ms * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": host_ms, "steps": k3,
               "what": "one step = one batch fit of this rank's %d sources: burn-in %d + %d iterations of %d "
                       "walkers (thin %d), %d evals; mbb_set_data + mbb_ensemble_fit(MBB_HOST) with page-locked "
                       "host arrays: photometry and starting ensembles up, final ensembles, log-probabilities, "
                       "acceptance counts, status and per-source posterior summaries down; source chunks "
                       "pipelined through three streams" % (nsrc, nburn, nmain, nw, thin, evals_fit),
               "gpu_launches_per_step": int(fit_launches),
               "fraction_of_device_resident_fit": dev_ms / host_ms,
               "sources_per_s": nsrc * world / (host_ms * 1e-3),
               "matches_device_path": same_fit}
        del h_pos, h_p0, h_lnp, h_nacc, h_st, h_stats
    elif e2e_loglike is not None:
        e2e = dict(e2e_loglike, what="one likelihood pass through host buffers (no delta-band batch fit for "
                                     "this workload)")

    # ---- other configurations, each a sub-leg with its own numbers -------------------------------
    def eval_leg(wname, key, gauss_too):
        W2 = build_workload(wname, rank, None)
        ctx2, n2, P2 = W2["ctx"], W2["n"], W2["P"]
        ctx2.set_math_mode(MODES[args.math])
        out2 = torch.empty(n2, dtype=torch.float64, device=dev)
        st2 = torch.empty(n2, dtype=torch.int32, device=dev)
        dl = wname == "default_model"
        flush2 = None if dl else torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
        k = max(3, min(args.steps, 5))
        leg2 = device_leg(ctx2, n2, W2["nw"], P2, out2, st2, flush2, k, 3)
        f2 = FLOP_PER_EVAL[wname]
        gauss = None
        if gauss_too and args.math == "fast":
            ref_out = out2.clone()
            ctx2.set_math_mode(MODES["gauss"])
            leg3 = device_leg(ctx2, n2, W2["nw"], P2, out2, st2, flush2, k, 3)
            fin = torch.isfinite(ref_out)
            rel = ((out2[fin] - ref_out[fin]).abs() / ref_out[fin].abs()).max().item()
            gauss = {"value": n2 * world / (leg3["step_ms_max"] * 1e-3), "unit": UNIT,
                     "ms_per_step": leg3["step_ms_max"], "max_rel_diff_vs_full_tables": rel,
                     "roofline": executed_roofline(key + "_gauss", n2, leg3["step_ms"], peak_tf, f2),
                     "note": "math mode gauss: per (walker, band) the band's 32-point Gauss rule where a "
                             "per-walker bound shows it agrees with the full table to rounding"}
        res = {"workload": workload_config(wname, W2["nsrc"], W2["nw"])["workload"],
               "value": n2 * world / (leg2["step_ms_max"] * 1e-3), "unit": UNIT,
               "ms_per_step": leg2["step_ms_max"], "evals_per_step_per_gpu": int(n2),
               "gpu_launches_per_step": int(leg2["launches"] // k),
               "roofline": executed_roofline(key, n2, leg2["step_ms"], peak_tf, f2),
               "status_errors": int((st2 > 1).sum().item()),
               "l2_policy": "inputs exceed L2" if dl else "L2 flushed between timed steps by writing a 256 MB buffer"}
        if gauss is not None:
            res["gauss_rules"] = gauss
        del W2, ctx2, P2, out2, st2, flush2
        torch.cuda.empty_cache()
        return res

    sub = {}
    if name == "cfg5" and not args.no_sublegs:
        del P, out, st
        torch.cuda.empty_cache()
        for wname, key, g2 in (("default_model", "default_model", False), ("cfg2", "cfg2", True),
                               ("cfg5p", "cfg5p", True), ("cfg3", "cfg3", True)):
            t1 = time.perf_counter()
            sub[wname] = eval_leg(wname, key, g2)
            log("[bench] rank %d: leg %s %.1f s" % (rank, wname, time.perf_counter() - t1))
        t1 = time.perf_counter()
        sub["cfg4"] = chain_leg(dev, rank, world, barrier, allmax, peak_tf)
        log("[bench] rank %d: leg cfg4 %.1f s" % (rank, time.perf_counter() - t1))
        if rank == 0:
            t1 = time.perf_counter()
            sub["fits"] = fits_leg(dev)
            log("[bench] rank 0: leg fits %.1f s" % (time.perf_counter() - t1))
    clocks = sampler.stop()

    if rank == 0:
        total = n * world
        value = total / (step_ms_max * 1e-3)
        try:
            hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
            hbm_src = "measured (MEASURED_PEAKS.json)"
        except Exception:
            hbm_peak, hbm_src = 6650.0, "fallback"
        achieved_gbs = n * BYTES_PER_EVAL / (step_ms * 1e-3) / 1e9
        roof = executed_roofline(name, n, step_ms, peak_tf, FLOP_PER_EVAL[name])
        roof["note"] = ("achieved = FP64 warp instructions per evaluation (ncu capture of this build) x evals "
                        "per launch x 64 flop / CUDA-event kernel time; peak = DFMA rate measured live by "
                        "mbb_fp64_peak on this GPU (no tensor cores: exp/pow-bound FP64 elementwise math)")
        roof["hbm"] = {"achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": achieved_gbs / hbm_peak,
                       "peak_source": hbm_src, "algorithmic_bytes_per_eval": BYTES_PER_EVAL}
        cb, check = None, None
        try:
            if args.no_cpu_baseline:
                raise RuntimeError("skipped (--no-cpu-baseline)")
            os.sched_setaffinity(0, all_cpus)      # the CPU leg uses every host core again
            cb, Ps, want = cpu_arm(name, PER_WORKER[name])
            # the same rows through the GPU path: the CPU arm's numbers are the reference's, so this is a
            # live parity check of the bench's own workload
            cfg_s, flux_s, unc_s, cov_s, _ = synthetic.sample_problem(name, 1)
            Wc = W["like"]
            c2 = _native.Context(dev.index)
            c2.set_model(cfg_s["wavenorm"], cfg_s["opthin"], cfg_s["noalpha"])
            c2.set_math_mode(MODES[args.math])
            c2.set_bands(*Wc.band_tables())
            if cov_s is not None:
                c2.set_data(flux_s[None], cinv=np.linalg.inv(cov_s)[None])
            else:
                c2.set_data(flux_s[None], ivar=(1.0 / unc_s**2)[None])
            c2.set_priors(Wc.lowlims, Wc.has_uplims, Wc.uplims, Wc.has_gpriors, Wc.gprior_means,
                          Wc.gprior_ivars)
            got, _ = c2.loglike(Ps)
            fin = np.isfinite(want)
            check = {"rows": int(len(Ps)),
                     "max_rel_err_gpu_vs_cpu_arm": float(np.max(np.abs(got[fin] - want[fin]) / np.abs(want[fin]))),
                     "neg_inf_agree": bool(np.array_equal(np.isneginf(got), np.isneginf(want)))}
        except Exception as exc:           # the baseline is informational; never hide the GPU line
            if cb is None:
                cb = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": "failed: %r" % (exc,)}
            else:
                check = {"failed": repr(exc)}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": step_ms_max, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": main_config(name, args.nsrc),
            "run": {"parallelism": "sources sharded over %d rank(s), no data-path collective; %d sources on "
                                   "each rank" % (world, W["nsrc"]),
                    "cpu_binding": ("rank 0 bound to %d CPUs" % len(bound)) if bound else "none",
                    "math_mode": args.math, "csrc_sha": csrc_sha()},
            "roofline": roof,
            "cpu_baseline": cb,
            "cpu_vs_gpu_check": check,
            "e2e": e2e,
            "e2e_loglike": e2e_loglike,
            "batch_fit": batch,
            "default_model": sub.get("default_model"),
            "passband": sub.get("cfg2"),
            "passband_batch": sub.get("cfg5p"),
            "cfg3": sub.get("cfg3"),
            "cfg4": sub.get("cfg4"),
            "fits": sub.get("fits"),
            "gpu_launches": int(launches),
            "clocks": clocks,
            "check": {"status_errors": nbad, "neg_inf": nneg, "wall_s_timed_region": wall,
                      "kernel_ms_each": [round(x, 4) for x in kernel_ms]},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ------------------------------------------------------------- configs[3]: chain post-processing
def chain_leg(dev, rank, world, barrier, allmax, peak_tf):
    """L_IR / dust mass / peak wavelength of one 1e7-sample chain (500 walkers x 20000 steps).  The
    walker rows are sharded over the ranks (the per-walker dedupe scan stays local), every rank
    post-processes its rows through the host-buffer call, and the one final gather of the outputs
    is inside the timed region."""
    import torch
    from mbb_emcee_b200 import _native, synthetic
    from mbb_emcee_b200.sharding import gather_concat, shard_range, shared_array
    cfg = synthetic.CONFIGS["cfg4"]
    nw, ns = 500, 20000
    chain = synthetic.random_walk_chain(cfg["truth"], nw, ns, np.random.RandomState(cfg["seed"]))
    lo, hi = shard_range(nw, rank, world)
    mine = _native.pinned_empty((hi - lo, ns, 5))          # page-locked, like the other host-buffer legs
    mine[...] = chain[lo:hi]
    uniq = 1.0 - float(np.all(chain[:, 1:] == chain[:, :-1], axis=2).mean())
    ctx = _native.Context(dev.index)
    lib = ctx._lib
    vp = ctypes.c_void_p
    ch = torch.as_tensor(mine, device=dev)
    o = torch.empty((hi - lo, ns), dtype=torch.float64, device=dev)
    s = torch.empty((hi - lo, ns), dtype=torch.int32, device=dev)
    res = {"workload": "BASELINE configs[3]: chain 500 walkers x 20000 steps = 1e7 samples, %.0f%% of steps "
                       "new; thick + alpha; z=2, dl=1.6e4 Mpc; walker rows sharded over %d rank(s)"
                       % (100 * uniq, world),
           "samples": nw * ns, "unit": "samples/s"}
    for key, which, wavenorm, method, fkey, flop in (("peak_lambda", 1, 500.0, "quadpack", "chain_unique", None),
                                                     ("L_IR_quadpack_replay", 2, 500.0, "quadpack", "chain_lir_qags",
                                                      357 * 366),
                                                     ("L_IR_gauss", 2, 500.0, "gauss", "chain_lir", 128 * 366),
                                                     ("dust_mass", 4, cfg["wavenorm"], "quadpack", None, None)):
        ctx.set_model(wavenorm, False, False)
        ctx.set_lir_method(method)
        a = [ctx._h, hi - lo, ns, vp(ch.data_ptr()), which, cfg["z"], cfg["lumdist"], 8.0, 1000.0,
             cfg["kappa"], cfg["kappa_wave"], vp(o.data_ptr()) if which == 1 else None,
             vp(o.data_ptr()) if which == 2 else None, vp(o.data_ptr()) if which == 4 else None,
             vp(s.data_ptr()), 1]
        ms = []
        for it in range(3):
            barrier()
            assert lib.mbb_chain_post(*a) == 0, lib.mbb_last_error()
            ctx.sync()
            ms.append(ctx.last_kernel_ms())
        t = allmax(float(np.mean(ms[1:])))[0]
        res[key] = {"ms": t, "samples_per_s": nw * ns / (t * 1e-3),
                    "finite": bool(torch.isfinite(o).all().item())}
        if fkey:
            f, stale = ncu_facts(fkey)
            if f and f.get("units_per_launch"):
                per = f["fp64_warp_instructions"] / f["units_per_launch"]
                # the dominant kernel's share of the call, from the capture; executed work over the live time
                res[key]["roofline"] = {"kernel": f["kernel"], "frac_of_fp64_peak_in_kernel": f["fp64_pipe_pct"] / 100.0,
                                        "fp64_warp_instructions_per_unique_sample": per, "ncu_stale": bool(stale)}
    del ch, o, s
    # end to end: host chain rows in, all three quantities out; the "final gather" is every rank's
    # device-to-host copy landing in its rows of one shared page-locked array, then a barrier
    import torch.distributed as dist
    bar = dist.barrier if world > 1 else None
    shared = [shared_array(t, (nw, ns), rank, world, barrier=bar) for t in ("peak", "lir", "dust")]
    stt = _native.pinned_empty((hi - lo, ns), np.int32)
    ctx.set_model(500.0, False, False)
    ctx.set_lir_method("quadpack")
    # warm-up at full size (device buffers of the context are allocated on first use)
    ctx.chain_post_into(mine, 7, peak=shared[0].array[lo:hi], lir=shared[1].array[lo:hi],
                        dustmass=shared[2].array[lo:hi], status=stt, z=cfg["z"], dl_mpc=cfg["lumdist"],
                        kappa=cfg["kappa"], kappa_wave=cfg["kappa_wave"])
    for sh in shared:
        sh.array[lo:hi] = np.nan
    barrier()
    t0 = time.perf_counter()
    ctx.chain_post_into(mine, 7, peak=shared[0].array[lo:hi], lir=shared[1].array[lo:hi],
                        dustmass=shared[2].array[lo:hi], status=stt, z=cfg["z"], dl_mpc=cfg["lumdist"],
                        kappa=cfg["kappa"], kappa_wave=cfg["kappa_wave"])
    t_local = time.perf_counter() - t0
    for sh in shared:
        sh.sync()
    t_all = allmax(time.perf_counter() - t0)[0]
    ok = bool(all(np.isfinite(sh.array).all() for sh in shared))
    # the same gather through the collective, for comparison (outside the e2e time)
    t1 = time.perf_counter()
    coll = gather_concat(np.ascontiguousarray(shared[1].array[lo:hi]))
    t_coll = allmax(time.perf_counter() - t1)[0]
    same = bool(np.array_equal(coll, shared[1].array))
    res["e2e"] = {"value": nw * ns / t_all, "unit": "samples/s", "s_per_chain": t_all,
                  "s_local_post_processing": allmax(t_local)[0],
                  "h2d_bytes": int(mine.nbytes), "d2h_bytes": int(3 * mine.shape[0] * ns * 8 + mine.shape[0] * ns * 4),
                  "gathered_shape": list(shared[1].array.shape),
                  "what": "mbb_chain_post(MBB_HOST), page-locked chain in: peak wavelength + L_IR (QUADPACK replay) + "
                          "dust mass of this rank's walker rows, written by the device-to-host copies straight into this rank's rows "
                          "of arrays shared by all ranks (page-locked /dev/shm mappings), then a barrier: the final "
                          "gather, inside the time, without a collective",
                  "all_finite": ok,
                  "nccl_all_gather_of_one_output_s": t_coll, "nccl_gather_matches": same}
    for sh in shared:
        sh.close()
    return res


# --------------------------------------- configs[0-2]: single-source fits as a user runs them
def fits_leg(dev):
    """run_mbb_emcee-style fits (reference run_mbb_emcee.py:285-294 -> mbb_fit.py:481-563): 250 walkers,
    burn-in 50, 1000 steps, the emcee-2.2 stretch move on the host with every half-ensemble (125
    walkers) evaluated by ONE host-buffer call -- replayed as a CUDA graph.  Wall time of
    mbb_fitter.run and the latency of one 125-walker call."""
    from mbb_emcee_b200 import mbb_fitter, synthetic
    out = {"walkers": 250, "burn_in": 50, "steps": 1000,
           "api": "mbb_fitter.run (host stretch move, emcee 2.2 semantics) over likelihood.__call__ -> "
                  "mbb_loglike(MBB_HOST), 125 rows per call, one CUDA graph launch per call"}
    for name in ("cfg1", "cfg2", "cfg3"):
        cfg, flux, unc, cov, _ = synthetic.sample_problem(name, 1)
        fit = mbb_fitter(nwalkers=250, wavenorm=cfg["wavenorm"], noalpha=cfg["noalpha"], opthin=cfg["opthin"],
                         response=cfg["response"], device=dev.index)
        fit.set_data(cfg["bands"], flux, unc, covmatrix=cov)
        for nm, v in cfg.get("uplims", []):
            fit.set_uplim(nm, v)
        for nm, m, s in cfg.get("gpriors", []):
            fit.set_gaussian_prior(nm, m, s)
        if cfg["noalpha"]:
            fit.fix_param('alpha')
        np.random.seed(cfg["seed"])
        truth = np.array(cfg["truth"])
        p0 = fit.generate_initial_values(truth, synthetic.P0_SIGMA)
        fit.sampler.random_state = np.random.RandomState(cfg["seed"]).get_state()
        fit.run(5, 5, p0)                                   # warm-up: buffers, graph capture
        t0 = time.perf_counter()
        fit.run(50, 1000, p0)
        wall = time.perf_counter() - t0
        # latency of the call emcee makes 2100 times
        rows = np.ascontiguousarray(p0[:125])
        for _ in range(20):
            fit.like(rows)
        t1 = time.perf_counter()
        for _ in range(300):
            fit.like(rows)
        call_us = 1e6 * (time.perf_counter() - t1) / 300
        ctx = fit.like.context
        kern_us = 1e3 * ctx.last_kernel_ms()
        out[name] = {"wall_s": wall, "evals": 250 * 1051, "evals_per_s": 250 * 1051 / wall,
                     "us_per_125_walker_call": call_us, "device_us_per_call": kern_us,
                     "mean_acceptance": float(np.mean(fit.sampler.acceptance_fraction))}
        # the same fit with the stretch move inside the library (mbb_fitter(sampler="device")):
        # burn-in and main run are two mbb_ensemble_fit(MBB_HOST) calls, the chain comes back once
        dfit = mbb_fitter(nwalkers=250, wavenorm=cfg["wavenorm"], noalpha=cfg["noalpha"], opthin=cfg["opthin"],
                          response=cfg["response"], device=dev.index, sampler="device", seed=cfg["seed"])
        dfit.set_data(cfg["bands"], flux, unc, covmatrix=cov)
        for nm, v in cfg.get("uplims", []):
            dfit.set_uplim(nm, v)
        for nm, m, s in cfg.get("gpriors", []):
            dfit.set_gaussian_prior(nm, m, s)
        dfit.run(50, 1000, p0)                             # warm-up at full size: page-locked staging, device buffers
        dfit.sampler.random_state = cfg["seed"]
        l0 = dfit.like.context.launch_count()
        t0 = time.perf_counter()
        dfit.run(50, 1000, p0)
        dwall = time.perf_counter() - t0
        dch = dfit.sampler.chain
        hch = fit.sampler.chain
        out[name]["device_sampler"] = {
            "wall_s": dwall, "evals_per_s": 250 * 1052 / dwall,
            "gpu_launches": int(dfit.like.context.launch_count() - l0),
            "mean_acceptance": float(np.mean(dfit.sampler.acceptance_fraction)),
            "chain_shape": list(dch.shape),
            # same posterior as the host sampler's chain (different random streams): T and beta
            "posterior_mean_T_beta": [float(dch[:, 200:, 0].mean()), float(dch[:, 200:, 1].mean())],
            "host_sampler_mean_T_beta": [float(hch[:, 200:, 0].mean()), float(hch[:, 200:, 1].mean())],
            "speedup_vs_host_sampler": wall / dwall}
        del dfit
    # the same fits on the host: the CPU likelihood timed on one core (the reference's default nthreads=1),
    # times the 250 x 1051 calls emcee makes -- sampler overhead not included
    try:
        import multiprocessing as mp
        with mp.get_context("spawn").Pool(3) as pool:
            jobs = [(nm, synthetic.sample_problem(nm, 600 if nm == "cfg1" else 150)[4], True)
                    for nm in ("cfg1", "cfg2", "cfg3")]
            res = pool.map(_cpu_worker, jobs)
        for nm, job, r in zip(("cfg1", "cfg2", "cfg3"), jobs, res):
            us = 1e6 * r[0] / len(job[1])
            out[nm]["cpu_us_per_eval_one_core"] = us
            out[nm]["cpu_kind"] = r[2]
            out[nm]["cpu_fit_estimate_s_one_core"] = us * 1e-6 * 250 * 1051
            out[nm]["speedup_vs_one_core"] = out[nm]["cpu_fit_estimate_s_one_core"] / out[nm]["wall_s"]
            out[nm]["device_sampler"]["speedup_vs_one_core"] = \
                out[nm]["cpu_fit_estimate_s_one_core"] / out[nm]["device_sampler"]["wall_s"]
    except Exception as exc:
        out["cpu_estimate_failed"] = repr(exc)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg5", choices=["cfg5", "cfg2", "cfg5p", "cfg3", "default_model"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: every rank its own 1e5 sources; strong: the workload's sources divided over the ranks")
    ap.add_argument("--nsrc", type=int, default=None, help="sources per GPU (default: workload's)")
    ap.add_argument("--fit-burn", type=int, default=50, help="burn-in iterations of the batch-fit legs")
    ap.add_argument("--fit-steps", type=int, default=150, help="main-run iterations of the batch-fit legs")
    ap.add_argument("--fit-thin", type=int, default=10)
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer and batch-fit legs (profiling runs)")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU leg")
    ap.add_argument("--no-sublegs", "--no-passband", dest="no_sublegs", action="store_true",
                    help="skip the other configurations' legs")
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

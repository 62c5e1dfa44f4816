"""One leg of the bench with nothing around it, for ncu: builds the workload, launches its dominant
kernel a few times, writes how many units one launch processes.
    python tools/profile_leg.py <key> [out.json]
keys: cfg5 default_model cfg2 cfg2_gauss cfg5p cfg5p_gauss cfg3 cfg3_gauss ens chain"""
import ctypes
import json
import os
import sys

import numpy as np

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from mbb_emcee_b200 import _native, synthetic  # noqa: E402

key = sys.argv[1]
outp = sys.argv[2] if len(sys.argv) > 2 else None
torch.cuda.set_device(0)
dev = torch.device("cuda:0")
units = {}
NSRC = {"cfg5": 20000, "default_model": 20000, "cfg2": 1024, "cfg5p": 2048, "cfg3": 512}
if key == "chain":
    cfg = synthetic.CONFIGS["cfg4"]
    nw, ns = 500, 20000
    chain = synthetic.random_walk_chain(cfg["truth"], nw, ns, np.random.RandomState(cfg["seed"]))
    uniq = int(nw + np.any(chain[:, 1:] != chain[:, :-1], axis=2).sum())
    ctx = _native.Context(0)
    ch = torch.as_tensor(chain, device=dev)
    o = [torch.empty((nw, ns), dtype=torch.float64, device=dev) for _ in range(3)]
    s = torch.empty((nw, ns), dtype=torch.int32, device=dev)
    vp = ctypes.c_void_p
    ctx.set_model(500.0, False, False)
    for method in ("quadpack", "gauss"):
        ctx.set_lir_method(method)
        for _ in range(2):
            rc = ctx._lib.mbb_chain_post(ctx._h, nw, ns, vp(ch.data_ptr()), 7, cfg["z"], cfg["lumdist"], 8.0, 1000.0,
                                         cfg["kappa"], cfg["kappa_wave"], vp(o[0].data_ptr()), vp(o[1].data_ptr()),
                                         vp(o[2].data_ptr()), vp(s.data_ptr()), 1)
            assert rc == 0
            ctx.sync()
    units = {"chain_dedupe": nw * ns, "chain_unique": uniq, "chain_lir_qags": uniq, "chain_lir": uniq,
             "chain_fill": nw * ns}
elif key == "ens":
    W = bench.build_workload("cfg5", 0, 20000)
    ctx, n, nw, nsrc = W["ctx"], W["n"], W["nw"], W["nsrc"]
    Pw = W["P"].clone()
    lnp = torch.empty(n, dtype=torch.float64, device=dev)
    nacc = torch.zeros(n, dtype=torch.int32, device=dev)
    stats = torch.zeros((nsrc, _native.FIT_NSTATS), dtype=torch.float64, device=dev)
    K = 20
    for rep in range(3):
        ctx.ensemble_fit_device(nsrc, nw, 0, K, Pw.data_ptr(), lnp.data_ptr(), rep > 0, seed=7, step0=rep * K,
                                naccept_ptr=nacc.data_ptr(), stats_ptr=stats.data_ptr(), thin=10)
        ctx.sync()
    units = {"ens": n * K}
else:
    wname = key.replace("_gauss", "")
    W = bench.build_workload(wname, 0, NSRC[wname])
    ctx, n, nw = W["ctx"], W["n"], W["nw"]
    ctx.set_math_mode(2 if key.endswith("_gauss") else 1)
    out = torch.empty(n, dtype=torch.float64, device=dev)
    st = torch.empty(n, dtype=torch.int32, device=dev)
    for _ in range(3):
        ctx.loglike_device(n, W["P"].data_ptr(), out.data_ptr(), st.data_ptr(), walkers_per_source=nw)
        ctx.sync()
    assert int((st > 1).sum().item()) == 0
    units = {key: n}
if outp:
    json.dump(units, open(outp, "w"))
print(json.dumps(units))

"""Time the device-resident sampler alone on the cfg5 workload (1e5 sources x 512 walkers):
ms per iteration (two half-steps) and proposals/s, with and without the posterior summaries.
usage: python tools/sampler_probe.py [iterations] [nsrc] [--variant thick_alpha]
Switch: MBB_B200_NO_FUSED_SAMPLER=1 (propose / evaluate / accept kernels)."""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

import bench  # noqa: E402
from mbb_emcee_b200 import _native  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith("--")]
K = int(args[0]) if len(args) > 0 else 50
NSRC = int(args[1]) if len(args) > 1 else None
torch.cuda.set_device(0)
W = bench.build_workload("cfg5", 0, NSRC)
ctx, n, nw, nsrc = W["ctx"], W["n"], W["nw"], W["nsrc"]
dev = torch.device("cuda:0")
Pw = W["P"].clone()
lnp = torch.empty(n, dtype=torch.float64, device=dev)
nacc = torch.zeros(n, dtype=torch.int32, device=dev)
stats = torch.zeros((nsrc, _native.FIT_NSTATS), dtype=torch.float64, device=dev)
ctx.ensemble_fit_device(nsrc, nw, 0, 2, Pw.data_ptr(), lnp.data_ptr(), False, seed=7, naccept_ptr=nacc.data_ptr())
ctx.sync()
out = {}
for tag, sp, thin in (("plain", 0, 1), ("stats_thin1", stats.data_ptr(), 1), ("stats_thin10", stats.data_ptr(), 10)):
    best = 1e30
    for rep in range(2):
        ctx.ensemble_fit_device(nsrc, nw, 0, K, Pw.data_ptr(), lnp.data_ptr(), True, seed=7, step0=2 + rep * K,
                                naccept_ptr=nacc.data_ptr(), stats_ptr=sp, thin=thin)
        ctx.sync()
        best = min(best, ctx.last_kernel_ms())
    out[tag] = {"ms_per_iteration": best / K, "proposals_per_s": n * K / (best * 1e-3)}
out["acceptance"] = float(nacc.double().mean().item()) / K
out["iterations"] = K
out["nsrc"] = nsrc
out["switches"] = {k: v for k, v in os.environ.items() if k.startswith("MBB_B200_")}
print(json.dumps(out))

"""Time the device-resident sampler alone on the cfg5 workload (1e5 sources x 512 walkers):
ms per iteration (two half-steps) and proposals/s.  usage: python tools/sampler_probe.py [iterations]
Switch: MBB_B200_NO_FUSED_SAMPLER=1 (propose / evaluate / accept kernels)."""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

import bench  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 10
torch.cuda.set_device(0)
W = bench.build_workload("cfg5", 0, None)
ctx, n, nw, nsrc = W["ctx"], W["n"], W["nw"], W["nsrc"]
dev = torch.device("cuda:0")
P = W["P"] if "P" in W else W["P_host"].to(dev)
Pw = P.clone()
lnp = torch.empty(n, dtype=torch.float64, device=dev)
nacc = torch.zeros(n, dtype=torch.int32, device=dev)
ctx.ensemble_run_device(nsrc, nw, 2, Pw.data_ptr(), lnp.data_ptr(), False, seed=7, naccept_ptr=nacc.data_ptr())
ctx.sync()
best = 1e30
for rep in range(3):
    ctx.ensemble_run_device(nsrc, nw, K, Pw.data_ptr(), lnp.data_ptr(), True, seed=7, step0=2 + rep * K,
                            naccept_ptr=nacc.data_ptr())
    ctx.sync()
    best = min(best, ctx.last_kernel_ms())
print(json.dumps({"ms_per_iteration": best / K, "proposals_per_s": n * K / (best * 1e-3),
                  "acceptance": float(nacc.double().mean().item()) / (2 + 3 * K),
                  "switches": {k: v for k, v in os.environ.items() if k.startswith("MBB_B200_")}}))

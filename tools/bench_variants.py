"""Device-resident log-likelihood throughput of the four model variants (fnu.pyx a1-a4:
thin/thick x alpha/no alpha) on the 6 delta bands of BASELINE cfg1/cfg5 and on cfg2's
tabulated band set, FAST and FAST_GAUSS.  Supporting evidence for DESIGN.md section 7 (bench.py's
headline line covers thin/no-alpha delta and thick/alpha tabulated).  One JSON object."""
import json
import os
import sys

import numpy as np

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, ROOT)


def main():
    import torch
    from mbb_emcee_b200 import _native, likelihood, synthetic
    dev = torch.device("cuda:0")
    out = {}
    for tag, bands, response, nsrc in (("delta6", synthetic.CONFIGS["cfg1"]["bands"], False, 16384),
                                       ("tabulated6", synthetic.CONFIGS["cfg2"]["bands"], True, 1024)):
        nw = 512
        n = nsrc * nw
        for opthin in (True, False):
            for noalpha in (True, False):
                like = likelihood(wavenorm=500.0, opthin=opthin, noalpha=noalpha, response=response, device=0)
                nb = len(bands)
                like.set_phot(bands, np.full(nb, 30.0), np.full(nb, 3.0))
                like._stage()
                ctx = like.context
                rng = np.random.RandomState(1)
                flux = rng.uniform(10, 60, (nsrc, nb))
                ctx.set_data(flux, ivar=1.0 / np.maximum(0.1 * flux, 1.0)**2)
                truth = torch.tensor([14.0, 1.8, 400.0, 3.0, 30.0], dtype=torch.float64, device=dev)
                sig = torch.tensor(synthetic.P0_SIGMA, dtype=torch.float64, device=dev)
                g = torch.Generator(device=dev)
                g.manual_seed(5)
                P = truth + sig * torch.randn((n, 5), dtype=torch.float64, device=dev, generator=g)
                P = torch.maximum(P, torch.tensor([2.0, 0.2, 10.0, 0.2, 0.5], dtype=torch.float64, device=dev)).contiguous()
                res = torch.empty(n, dtype=torch.float64, device=dev)
                st = torch.empty(n, dtype=torch.int32, device=dev)
                row = {}
                for mode, name in ((_native.MATH_FAST, "fast"), (_native.MATH_FAST_GAUSS, "gauss")):
                    if name == "gauss" and not response:
                        continue
                    ctx.set_math_mode(mode)
                    for _ in range(3):
                        ctx.loglike_device(n, P.data_ptr(), res.data_ptr(), st.data_ptr(), walkers_per_source=nw)
                    ctx.sync()
                    ms = []
                    for _ in range(5):
                        ctx.loglike_device(n, P.data_ptr(), res.data_ptr(), st.data_ptr(), walkers_per_source=nw)
                        ctx.sync()
                        ms.append(ctx.last_kernel_ms())
                    row[name] = {"ms": float(np.median(ms)), "evals_per_s": n / (float(np.median(ms)) * 1e-3),
                                 "status_errors": int((st > 1).sum().item())}
                out["%s/%s/%s" % (tag, "thin" if opthin else "thick", "noalpha" if noalpha else "alpha")] = row
    print(json.dumps({"evals": "delta6: 16384 x 512, tabulated6: 1024 x 512 (1688 nodes)", "rows": out}))


if __name__ == "__main__":
    main()

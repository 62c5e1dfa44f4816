"""Timing of the chain post-processing kernels on BASELINE configs[3]
(10^7-sample chain; L_IR, dust mass, peak wavelength).  Device-resident chain,
CUDA-event time of the kernels (mbb_last_kernel_ms).  Prints one JSON object;
kept under profiles/ as supporting evidence (not the bench.py headline line)."""
import ctypes
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def main():
    import torch
    from mbb_emcee_b200 import _native, synthetic
    cfg = synthetic.CONFIGS["cfg4"]
    rng = np.random.RandomState(cfg["seed"])
    nw, ns = 500, 20000
    chain = synthetic.random_walk_chain(cfg["truth"], nw, ns, rng)
    uniq = 1.0 - float(np.all(chain[:, 1:] == chain[:, :-1], axis=2).mean())
    dev = torch.device("cuda:0")
    ch = torch.as_tensor(chain, device=dev)
    out = torch.empty((nw, ns), dtype=torch.float64, device=dev)
    st = torch.empty((nw, ns), dtype=torch.int32, device=dev)
    ctx = _native.Context(0)
    lib = ctx._lib
    res = {"workload": "BASELINE configs[3]: chain 500 walkers x 20000 steps = 1e7 samples, "
                       "%.0f%% of steps new; thick + alpha; z=2, dl=1.6e4 Mpc" % (100 * uniq),
           "samples": nw * ns}
    vp = ctypes.c_void_p
    for name, which, wavenorm in (("peak_lambda", 1, 500.0), ("L_IR_quadpack_replay", 2, 500.0),
                                  ("L_IR_gauss", 2, 500.0), ("dust_mass", 4, cfg["wavenorm"])):
        ctx.set_model(wavenorm, False, False)
        ctx.set_lir_method("gauss" if name.endswith("gauss") else "quadpack")
        args = [ctx._h, nw, ns, vp(ch.data_ptr()), which, cfg["z"], cfg["lumdist"], 8.0, 1000.0,
                cfg["kappa"], cfg["kappa_wave"],
                vp(out.data_ptr()) if which == 1 else None, vp(out.data_ptr()) if which == 2 else None,
                vp(out.data_ptr()) if which == 4 else None, vp(st.data_ptr()), 1]
        ms = []
        for it in range(4):
            torch.cuda.synchronize()
            rc = lib.mbb_chain_post(*args)
            assert rc == 0, lib.mbb_last_error()
            ctx.sync()
            ms.append(ctx.last_kernel_ms())
        t = float(np.mean(ms[1:]))
        res[name] = {"ms": t, "samples_per_s": nw * ns / (t * 1e-3),
                     "finite": bool(torch.isfinite(out).all().item())}
    # CPU oracle on a small slice, for scale (the reference's _map_chain is serial Python)
    try:
        import mbb_oracle as oracle
        sub = chain[:1, :400]
        t0 = time.perf_counter()
        oracle.map_chain(sub, lambda s: oracle.lir_step(s, cfg["z"], 8.0, 1000.0, False, False))
        res["cpu_oracle_L_IR_samples_per_s_per_core"] = sub.shape[1] / (time.perf_counter() - t0)
    except Exception as exc:
        res["cpu_oracle"] = repr(exc)
    print(json.dumps(res))


if __name__ == "__main__":
    main()

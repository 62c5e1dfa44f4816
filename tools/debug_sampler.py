"""Debug aid: resident sampler vs propose/evaluate/accept kernels vs the numpy twin, first mismatch."""
import os
import sys

import numpy as np

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import mbb_oracle as oracle  # noqa: E402
import philox_np  # noqa: E402
from mbb_emcee_b200 import batch_fitter  # noqa: E402


def first_diff(a, b):
    d = np.argwhere(a != b)
    return None if d.size == 0 else tuple(int(x) for x in d[0])


for nsrc, nw, nsteps in ((5, 16, 25), (11, 64, 12), (5, 250, 6), (3, 512, 4), (2, 1024, 3)):
    rng = np.random.RandomState(11)
    bands = [70.0, 100.0, 160.0, 250.0, 350.0, 500.0]
    bf = batch_fitter(nwalkers=nw, opthin=True, noalpha=True, device=0)
    bf.fix_param('alpha')
    flux = rng.uniform(10, 80, (nsrc, 6))
    unc = np.maximum(0.1 * flux, 1.0)
    bf.set_data(bands, flux, unc)
    p0 = bf.generate_initial_values((12.0, 1.8, 1300.0, 4.0, 30.0), [2, 0.2, 100, 0.3, 5.0], seed=11)
    ctx = bf._stage()
    specs = []
    for s in range(nsrc):
        sp = oracle.LikeSpec(500.0, True, True)
        sp.set_phot(bands, flux[s], unc[s])
        sp.has_uplim = list(bf.like.has_uplims)
        sp.uplim = np.array(bf.like.uplims)
        specs.append(sp)
    os.environ.pop("MBB_B200_NO_FUSED_SAMPLER", None)
    a = ctx.ensemble_fit(p0, 0, nsteps, seed=77, stats=False, chain=True)
    os.environ["MBB_B200_NO_FUSED_SAMPLER"] = "1"
    b = ctx.ensemble_fit(p0, 0, nsteps, seed=77, stats=False, chain=True)
    os.environ.pop("MBB_B200_NO_FUSED_SAMPLER", None)
    r = philox_np.replay(lambda s, Q: oracle.loglike_batch(specs[s], Q), p0, nsteps, 77, chain=True)
    print(nsrc, nw, nsteps, "resident vs split:", first_diff(a["chain"], b["chain"]),
          "resident vs twin:", first_diff(a["chain"], r[3]), "split vs twin:", first_diff(b["chain"], r[3]),
          "nacc", int(a["naccept"].sum()), int(b["naccept"].sum()), int(r[2].sum()), flush=True)
    fd = first_diff(b["chain"], r[3])
    if fd is not None:
        it, s, w, _ = fd
        print("  split", b["chain"][it, s, w], b["chain_lnprob"][it, s, w], "twin", r[3][it, s, w], r[4][it, s, w])
        if it > 0:
            print("  prev ", b["chain"][it - 1, s, w], b["chain_lnprob"][it - 1, s, w], r[4][it - 1, s, w])

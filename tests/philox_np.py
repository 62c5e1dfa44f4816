"""TEST INFRASTRUCTURE: numpy Philox4x32-10 and the stretch-move draws of
csrc/mbb_ensemble.cuh, for replaying the device sampler on the host."""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised over uint64 arrays holding 32-bit values."""
    c0, c1, c2, c3 = (np.asarray(x, dtype=np.uint64) & MASK for x in (c0, c1, c2, c3))
    k0 = np.uint64(k0 & 0xFFFFFFFF)
    k1 = np.uint64(k1 & 0xFFFFFFFF)
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        n0 = (p1 >> np.uint64(32)) ^ c1 ^ k0
        n1 = p1 & MASK
        n2 = (p0 >> np.uint64(32)) ^ c3 ^ k1
        n3 = p0 & MASK
        c0, c1, c2, c3 = n0, n1, n2, n3
        k0 = np.uint64((int(k0) + W0) & 0xFFFFFFFF)
        k1 = np.uint64((int(k1) + W1) & 0xFFFFFFFF)
    return c0, c1, c2, c3


def u53(hi, lo):
    m = ((hi >> np.uint64(5)) << np.uint64(26)) | (lo >> np.uint64(6))
    return (m.astype(np.float64) + 0.5) * 1.1102230246251565e-16


def u52w(w0, w1_top20):
    m = (w0 << np.uint64(20)) | w1_top20
    return (m.astype(np.float64) + 0.5) * 2.0**-52


def u43(w3, w1_low11):
    m = (w3 << np.uint64(11)) | w1_low11
    return (m.astype(np.float64) + 0.5) * 2.0**-43


def stretch_draw(seed, widx, hstep, a, ncomp):
    """(z, u, partner) for global walker indices ``widx`` at half-step ``hstep``: one Philox
    block per proposal (csrc/mbb_ensemble.cuh stretch_draw)."""
    widx = np.asarray(widx, dtype=np.uint64)
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    lo, hi = widx & MASK, widx >> np.uint64(32)
    h_lo = np.uint64(hstep & 0xFFFFFFFF)
    h_hi = np.uint64((hstep >> 32) & 0xFFFFFFFF)
    r = philox4x32_10(lo, hi, np.full_like(lo, h_lo), np.full_like(lo, h_hi), k0, k1)
    t = (a - 1.0) * u52w(r[0], r[1] >> np.uint64(12)) + 1.0
    z = (t * t) / a
    partner = ((r[2] * np.uint64(ncomp)) >> np.uint64(32)).astype(np.int64)
    return z, u43(r[3], r[1] & np.uint64(0x7FF)), partner


def replay(lnprob_rows, p0, nsteps, seed, a=2.0, step0=0, lnp0=None, src0=0, log_form=False, chain=False):
    """Host replay of mbb_ensemble_run: p0[nsrc][nw][5]; lnprob_rows(src, Q[m,5]) -> [m].
    log_form: emcee 2.2's own acceptance test, lnpdiff > log(u), instead of the device's
    log-free form.  chain: also return the per-iteration ensembles and log-probabilities."""
    pos = np.array(p0, dtype=np.float64)
    nsrc, nw = pos.shape[:2]
    h = nw // 2
    lnp = np.array([lnprob_rows(s, pos[s]) for s in range(nsrc)]) if lnp0 is None else np.array(lnp0)
    nacc = np.zeros((nsrc, nw), dtype=np.int64)
    ch, chl = [], []
    for it in range(nsteps):
        for half in (0, 1):
            hstep = 2 * (step0 + it) + half
            for s in range(nsrc):
                widx = (src0 + s) * h + np.arange(h)
                z, u, partner = stretch_draw(seed, widx, hstep, a, h)
                own = np.arange(h) + (0 if half == 0 else h)
                oth = partner + (h if half == 0 else 0)
                c = pos[s, oth]
                q = c - z[:, None] * (c - pos[s, own])
                newlnp = np.asarray(lnprob_rows(s, q))
                with np.errstate(invalid="ignore", over="ignore", divide="ignore"):
                    if log_form:
                        # emcee 2.2 _propose_stretch: lnpdiff = (dim - 1) ln z + newlnp - lnp; accept = lnpdiff > ln u
                        acc = (4.0 * np.log(z) + newlnp - lnp[s, own]) > np.log(u)
                    else:
                        # the device's log-free form of the same test (csrc/mbb_ensemble.cuh stretch_accept)
                        dl = np.fmin(np.fmax(newlnp - lnp[s, own], -745.0), 700.0)
                        acc = u < (z * z) * (z * z) * np.exp(dl)
                pos[s, own[acc]] = q[acc]
                lnp[s, own[acc]] = newlnp[acc]
                nacc[s, own[acc]] += 1
        if chain:
            ch.append(pos.copy())
            chl.append(lnp.copy())
    if chain:
        return pos, lnp, nacc, np.array(ch), np.array(chl)
    return pos, lnp, nacc

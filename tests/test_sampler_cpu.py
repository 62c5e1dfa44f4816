"""CPU: host-side pieces of the device sampler -- the numpy twin of its draws and the two
equivalent forms of the stretch-move acceptance test."""
import numpy as np

import philox_np


def _spec(oracle):
    sp = oracle.LikeSpec(500.0, True, True)
    sp.set_phot([70.0, 100.0, 160.0, 250.0, 350.0, 500.0], [12.0, 35.0, 60.0, 48.0, 30.0, 14.0],
                [2.0, 4.0, 6.0, 5.0, 3.0, 2.0])
    sp.auto_lambda0_uplim(500.0)
    return sp


def test_log_free_acceptance_is_emcees_test(oracle):
    """emcee 2.2 accepts where (dim-1) ln z + lnp(q) - lnp(s) > ln u; the device evaluates
    u < z^4 exp(lnp(q) - lnp(s)).  Same decisions: identical chains over ~1e4 proposals,
    including proposals below a lower limit (lnp = -inf)."""
    sp = _spec(oracle)
    rng = np.random.RandomState(7)
    nw = 48
    p0 = np.array([12.0, 1.8, 1300.0, 4.0, 30.0]) + np.array([2, 0.2, 100, 0.0, 5.0]) * rng.standard_normal((nw, 5))
    p0[:, 0] = np.abs(p0[:, 0]) + 1.5      # some walkers start close to the T >= 1 limit
    p0[:4, 0] = 1.2
    fn = lambda s, Q: oracle.loglike_batch(sp, Q)
    a = philox_np.replay(fn, p0[None], 110, 2024, log_form=False, chain=True)
    b = philox_np.replay(fn, p0[None], 110, 2024, log_form=True, chain=True)
    assert np.array_equal(a[3], b[3]) and np.array_equal(a[2], b[2])
    acc = a[2].sum() / float(nw * 110)
    assert 0.1 < acc < 0.9


def test_draw_fields():
    """One Philox block per proposal: 52 + 32 + 43 bits, uniforms strictly inside (0, 1)."""
    z, u, partner = philox_np.stretch_draw(0xDEADBEEF12345678, np.arange(200000), 77, 2.0, 256)
    assert ((z >= 0.5) & (z <= 2.0)).all() and ((u > 0) & (u < 1)).all()
    assert partner.min() == 0 and partner.max() == 255
    assert abs(u.mean() - 0.5) < 5e-3 and abs(np.bincount(partner, minlength=256).std() / (200000 / 256.) - 0.0) < 0.1
    # g(z) ~ 1/sqrt(z) on [1/a, a]: E[z] = (a^2 + a + 1) / (3a) for a = 2
    assert abs(z.mean() - 7.0 / 6.0) < 5e-3
    assert abs(np.corrcoef(z, u)[0, 1]) < 0.01
    m = np.uint64(0xFFFFFFFF)
    assert philox_np.u52w(m, np.uint64(0xFFFFF)) < 1.0 and philox_np.u43(m, np.uint64(0x7FF)) < 1.0
    assert philox_np.u52w(np.uint64(0), np.uint64(0)) > 0.0 and philox_np.u43(np.uint64(0), np.uint64(0)) > 0.0
    # a different global source offset gives different numbers
    z2, _, _ = philox_np.stretch_draw(0xDEADBEEF12345678, 256 * 1000 + np.arange(16), 77, 2.0, 256)
    assert not np.array_equal(z[:16], z2)

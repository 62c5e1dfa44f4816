"""Generate the polynomial coefficients used by csrc/mbb_fastmath.cuh.

Near-minimax polynomials by interpolation at Chebyshev nodes, solved in 60-digit
arithmetic (mpmath) and rounded to double:
    exp(r)          ~ sum_{i=0..11} E[i] r^i          on |r| <= ln2/2
    (expm1(r)-r)/r^2 ~ sum_{i=0..10} M[i] r^i         on |r| <= ln2/2
Prints C initialisers and the measured maximum relative error of each fit.
"""
import mpmath as mp

mp.mp.dps = 60
A = mp.log(2) / 2 * mp.mpf("1.0001")


def cheb_fit(f, deg):
    n = deg + 1
    xs = [A * mp.cos(mp.pi * (2 * k + 1) / (2 * n)) for k in range(n)]
    V = mp.matrix(n, n)
    b = mp.matrix(n, 1)
    for i, x in enumerate(xs):
        for j in range(n):
            V[i, j] = x**j
        b[i] = f(x)
    c = mp.lu_solve(V, b)
    return [float(c[i]) for i in range(n)]


def max_relerr(f, coef, npts=4001):
    worst = mp.mpf(0)
    for k in range(npts):
        x = -A + 2 * A * k / (npts - 1)
        p = mp.mpf(0)
        for c in reversed(coef):
            p = p * x + mp.mpf(c)
        fx = f(x)
        worst = max(worst, abs((p - fx) / fx))
    return worst


def g(x):
    if abs(x) < mp.mpf(10)**-12:
        return mp.mpf(1) / 2 + x / 6 + x**2 / 24
    return (mp.expm1(x) - x) / x**2


E = cheb_fit(mp.exp, 11)
M = cheb_fit(g, 10)
print("// exp(r), degree 11, max rel err %.2e" % float(max_relerr(mp.exp, E)))
print("{" + ", ".join("%.17g" % c for c in E) + "}")
print("// (expm1(r)-r)/r^2, degree 10, max rel err %.2e" % float(max_relerr(g, M)))
print("{" + ", ".join("%.17g" % c for c in M) + "}")

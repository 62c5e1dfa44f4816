/* TEST INFRASTRUCTURE -- not product code.
 *
 * Plain-C restatement of the four node loops of the reference's only native
 * file, mbb_emcee/fnu.pyx (Cython), used by oracle/mbb_oracle.py when the
 * compiled reference module (oracle/_ref/fnu*.so) is not available, and as a
 * cross-check of it when it is.  Arithmetic is glibc pow/expm1 in the same
 * evaluation order as the C that Cython generates, so the two agree bit for
 * bit (tests/test_oracle.py::test_c_port_matches_ref_fnu).
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off, no -ffast-math).
 */
#include <math.h>
#include <stddef.h>

static const double H_PLANCK = 6.6260693e-34; /* fnu.pyx:14 */
static const double K_BOLTZ = 1.3806505e-23;  /* fnu.pyx:15 */

/* fnu.pyx:9-27  -- retval[i] = normfac * cx**bp3 / expm1(cx) */
void oracle_fnu_thin_noalpha(const double *freq, size_t n, double T,
                             double beta, double normfac, double *out) {
  const double hokt9 = 1e9 * H_PLANCK / (K_BOLTZ * T);
  const double bp3 = beta + 3.0;
  for (size_t i = 0; i < n; ++i) {
    double cx = hokt9 * freq[i];
    double num = normfac * pow(cx, bp3);
    out[i] = num / expm1(cx);
  }
}

/* fnu.pyx:30-53 -- power law above xmerge, grey body below; normfac applied
 * to the whole array afterwards (fnu.pyx:53) */
void oracle_fnu_thin_walpha(const double *freq, size_t n, double T, double beta,
                            double alpha, double normfac, double xmerge,
                            double kappa, double *out) {
  const double hokt9 = 1e9 * H_PLANCK / (K_BOLTZ * T);
  const double bp3 = beta + 3.0;
  for (size_t i = 0; i < n; ++i) {
    double cx = hokt9 * freq[i];
    double v;
    if (cx > xmerge)
      v = kappa * pow(cx, -alpha);
    else
      v = pow(cx, bp3) / expm1(cx);
    out[i] = normfac * v;
  }
}

/* fnu.pyx:56-78 -- ((-normfac * expm1(-x0b)) * cx**3) / expm1(cx) */
void oracle_fnu_thick_noalpha(const double *freq, size_t n, double T,
                              double beta, double x0, double normfac,
                              double *out) {
  const double hokt9 = 1e9 * H_PLANCK / (K_BOLTZ * T);
  for (size_t i = 0; i < n; ++i) {
    double cx = hokt9 * freq[i];
    double x0b = pow(cx / x0, beta);
    double a = -normfac * expm1(-x0b);
    double b = a * pow(cx, 3.0);
    out[i] = b / expm1(cx);
  }
}

/* fnu.pyx:81-108 */
void oracle_fnu_thick_walpha(const double *freq, size_t n, double T,
                             double beta, double x0, double alpha,
                             double normfac, double xmerge, double kappa,
                             double *out) {
  const double hokt9 = 1e9 * H_PLANCK / (K_BOLTZ * T);
  for (size_t i = 0; i < n; ++i) {
    double cx = hokt9 * freq[i];
    double v;
    if (cx > xmerge) {
      v = kappa * pow(cx, -alpha);
    } else {
      double x0b = pow(cx / x0, beta);
      double a = -expm1(-x0b);
      double b = a * pow(cx, 3.0);
      v = b / expm1(cx);
    }
    out[i] = normfac * v;
  }
}

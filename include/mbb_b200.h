/* mbb_b200 -- C ABI of the B200 (sm_100a) modified-blackbody likelihood library.
 *
 * This is the drop-in boundary for the hot path of aconley/mbb_emcee:
 * everything the reference does per walker inside `likelihood.__call__`
 * (mbb_emcee/likelihood.py:790-834), and per chain sample inside
 * `mbb_results.compute_*` (mbb_emcee/results.py:570-801), runs in CUDA kernels
 * behind these entry points.  Plain C, plain pointers and sizes, no torch
 * types.  The shared library is libmbb_b200.so (mbb_emcee_b200/csrc).
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; the message
 *     of the last failure on the calling thread is `mbb_last_error()`.
 *   - parameter vectors are (T, beta, lambda0, alpha, fnorm), the order of
 *     likelihood.py:20-22.  `layout` MBB_AOS is emcee's [n][5] row-major
 *     block, MBB_SOA is [5][n].
 *   - `mem` says where `pars`/outputs live: MBB_HOST pointers are staged
 *     through the context's pinned buffers (H2D, kernel, D2H, synchronous on
 *     return); MBB_DEVICE pointers (e.g. torch.Tensor.data_ptr()) are used in
 *     place and the call returns after the launch -- call mbb_sync() before
 *     reading results on the host.
 *   - the caller owns every buffer it passes; set_* calls copy into
 *     library-owned device memory; nothing library-allocated is returned.
 *   - a context is bound to one device and one stream; calls on one context
 *     must be serialised by the caller; different contexts (one per GPU) may
 *     be driven from different threads.
 *   - there is NO CPU fallback: without a usable CUDA device
 *     mbb_ctx_create fails.
 */
#ifndef MBB_B200_H
#define MBB_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mbb_ctx mbb_ctx;

enum { MBB_AOS = 0, MBB_SOA = 1 };
enum { MBB_HOST = 0, MBB_DEVICE = 1 };
/* MBB_MATH_FAITHFUL: the reference's formulas in the reference's order (libdevice pow/expm1).
 * MBB_MATH_FAST: algebraically identical, lean exp family (default).
 * MBB_MATH_FAST_GAUSS: FAST, and a tabulated passband's node sum (response.py:544-576) is
 *   taken over the 32-point Gauss rule of the band's discrete measure {nu_i, w_i} wherever a
 *   per-walker bound shows the two agree to rounding (integrand analytic over the band, its
 *   exponential type over the half-band <= 12, merge point outside the band); every other
 *   (walker, band) pair uses the full table.  Same results to <= 1e-13 (tests), 4-9x fewer nodes. */
enum { MBB_MATH_FAITHFUL = 0, MBB_MATH_FAST = 1, MBB_MATH_FAST_GAUSS = 2 };
enum { MBB_LIR_QUADPACK = 0, MBB_LIR_GAUSS = 1 };

/* per-evaluation status codes written to `status[]`; 0/1 are normal outcomes,
 * the others correspond to exceptions the reference raises. */
enum {
  MBB_ST_OK = 0,
  MBB_ST_BELOW_LOWLIM = 1,  /* likelihood.py:806-807, lnlike = -inf            */
  MBB_ST_BAD_ALPHA = 2,     /* modified_blackbody.py:219-221, ValueError       */
  MBB_ST_BAD_BETA = 3,      /* modified_blackbody.py:222-224, ValueError       */
  MBB_ST_BRACKET_LOW = 4,   /* modified_blackbody.py:294-300, ValueError       */
  MBB_ST_BRACKET_HIGH = 5,  /* modified_blackbody.py:310-316, ValueError       */
  MBB_ST_NO_CONVERGE = 6,   /* scipy brentq RuntimeError (:321, :633)          */
  MBB_ST_OVERFLOW = 7,      /* modified_blackbody.py:326-328, OverflowError    */
  MBB_ST_PEAK_BRACKET = 8,  /* modified_blackbody.py:612-630, Exception        */
  MBB_ST_NONFINITE = 9      /* NaN/inf parameters                              */
};

int mbb_version(void);
const char *mbb_last_error(void);
/* number of CUDA devices visible; <0 if the runtime cannot initialise */
int mbb_device_count(void);

/* ---- context ------------------------------------------------------------ */
int mbb_ctx_create(int device_ordinal, mbb_ctx **out);
int mbb_ctx_destroy(mbb_ctx *ctx);
int mbb_sync(mbb_ctx *ctx);
/* total kernel launches issued by this context so far (bench "gpu_launches") */
int64_t mbb_launch_count(const mbb_ctx *ctx);
/* the cudaStream_t (as an integer) kernels are launched on, for event timing */
uint64_t mbb_stream_handle(const mbb_ctx *ctx);
/* milliseconds the kernels of the most recent compute call took on the device
 * (CUDA events on the context's stream around the launch); call after mbb_sync
 * for MBB_DEVICE calls. */
int mbb_last_kernel_ms(mbb_ctx *ctx, float *ms);

/* page-locked host memory for the MBB_HOST calls (cudaHostAlloc, portable across the
 * contexts of all devices): buffers the library can DMA from / to directly. */
int mbb_host_alloc(size_t bytes, void **out);
int mbb_host_free(void *p);
/* page-lock memory the caller already owns (cudaHostRegister) -- e.g. a shared-memory mapping that
 * several one-process-per-GPU ranks fill in disjoint slices: the final gather without a collective */
int mbb_host_register(void *p, size_t bytes);
int mbb_host_unregister(void *p);

/* ---- model: replaces the constructor arguments the reference threads through
 *      likelihood._set_sed (likelihood.py:765-768) ------------------------- */
int mbb_set_model(mbb_ctx *ctx, double wavenorm, int opthin, int noalpha);
int mbb_set_math_mode(mbb_ctx *ctx, int mode);

/* ---- how mbb_chain_post integrates f_nu for L_IR / freq_integrate
 * (modified_blackbody.py:671 -> scipy.integrate.quad):
 *   MBB_LIR_QUADPACK (default) replays QUADPACK dqagse (21-point Gauss-Kronrod,
 *     epsabs = epsrel = 1.49e-8, limit 50, epsilon extrapolation) on the
 *     reference's integrand: the reference's number to ~1e-15, including its
 *     own ~1e-8 quadrature error;
 *   MBB_LIR_GAUSS integrates the power-law part analytically and the grey
 *     body by a fixed 128-node Gauss-Legendre rule in ln(x): the true integral
 *     to ~1e-15, faster. */
int mbb_set_lir_method(mbb_ctx *ctx, int method);

/* ---- passbands: replaces response.__call__ (response.py:544-576).
 * Band b owns nodes [band_off[b], band_off[b+1]).  node_wave_um are the
 * wavelengths the SED is evaluated at (response._wave, or _normwave for a
 * delta band); node_weight = response._sedmult * response._normfac (1 for a
 * delta band), so band flux = sum_i f_nu(wave_i) * weight_i.
 * scalar_path[b] != 0 marks bands the reference evaluates through its numpy
 * twin (x = (h/kT)*1e9*nu, modified_blackbody.py:461-464: delta passbands
 * inside a response set) instead of the Cython loop (x = (1e9*h/kT)*nu,
 * fnu.pyx:16) -- only observable in MBB_MATH_FAITHFUL. */
int mbb_set_bands(mbb_ctx *ctx, int nbands, const int32_t *band_off,
                  const double *node_wave_um, const double *node_weight,
                  const uint8_t *scalar_path);

/* ---- data: replaces likelihood.set_phot / set_cov (likelihood.py:158-232,
 * 330-357).  flux[nsrc][nbands]; exactly one of ivar[nsrc][nbands]
 * (= 1/unc^2, :222) or cinv[nsrc][nbands][nbands] (= inv(C), :356) non-NULL. */
int mbb_set_data(mbb_ctx *ctx, int nsrc, int nbands, const double *flux,
                 const double *ivar, const double *cinv);

/* Same data with the covariance PRE-FACTORED: chol[nsrc][nbands][nbands] holds
 * the lower-triangular Cholesky factor L of each source's covariance matrix,
 * C = L L' (row-major; the upper triangle is ignored).  chi-square becomes
 * |L^-1 (data - model)|^2 by forward substitution in registers, in place of the
 * reference's explicit inverse and two dot products (likelihood.py:356
 * `np.linalg.inv`, :823 `np.dot(diff, np.dot(invcov, diff))`): no inverse is
 * ever formed and nb(nb+1)/2 instead of nb^2 multiply-adds are spent per
 * evaluation; the two forms agree to a small multiple of cond(C) * 2^-53.
 * Fails for a non-positive or non-finite diagonal. */
int mbb_set_data_chol(mbb_ctx *ctx, int nsrc, int nbands, const double *flux,
                      const double *chol);

/* ---- limits and priors: likelihood.py:73, 83-92, 462-483, 541-570.
 * Index 5 of the 6-slot arrays is the ghost parameter lambda_peak. */
int mbb_set_priors(mbb_ctx *ctx, const double lowlim[5],
                   const uint8_t has_uplim[6], const double uplim[6],
                   const uint8_t has_gprior[6], const double gmean[6],
                   const double givar[6]);

/* ---- fixed parameters: mbb_fitter.fix_param / generate_initial_values
 * (mbb_fit.py:183-200, 440-443: a fixed parameter holds ONE value in every walker
 * for the whole run -- e.g. alpha under --noalpha, run_mbb_emcee.py:233).  A promise
 * by the caller that column i of every `pars` handed to mbb_loglike equals values[i]
 * wherever fixed[i] != 0.  The host path uses it for MBB_SOA batches: fixed columns
 * are not copied host-to-device at all (cfg5 with lambda0 and alpha fixed: 24 instead
 * of 40 bytes per evaluation over PCIe), they are filled on the device once.  Other
 * paths ignore it.  fixed == NULL clears the promise. */
int mbb_set_fixed_params(mbb_ctx *ctx, const int32_t fixed[5], const double values[5]);

/* ---- THE hot entry: likelihood.__call__ for n parameter vectors
 * (likelihood.py:790-834).  Evaluation i uses source src_index[i] if
 * src_index != NULL, else i / walkers_per_source (pass n for a single
 * source).  out_status may be NULL. */
int mbb_loglike(mbb_ctx *ctx, int64_t n, const double *pars, int layout,
                const int32_t *src_index, int64_t walkers_per_source,
                double *out_lnlike, int32_t *out_status, int mem);

/* ---- SED evaluation: modified_blackbody.__call__ / f_nu
 * (modified_blackbody.py:441-554) for n parameter vectors on a common
 * frequency grid: out[n][nfreq] in mJy.  scalar_path selects the numpy-twin
 * rounding of x (see mbb_set_bands; MBB_MATH_FAITHFUL only).  math_mode:
 * MBB_MATH_FAITHFUL -- the reference's formulas in the reference's order;
 * MBB_MATH_FAST -- the per-walker setup and node arithmetic of the likelihood
 * kernels (each frequency a single-node band of weight 1); < 0 -- the mode set
 * by mbb_set_math_mode. */
int mbb_fnu(mbb_ctx *ctx, int64_t n, const double *pars, int layout,
            int nfreq, const double *freq_ghz, int scalar_path, int math_mode,
            double *out, int32_t *out_status, int mem);

/* ---- per-walker constants: modified_blackbody.__init__ + max_wave
 * (modified_blackbody.py:200-337, 581-637).  out_consts[n][6] =
 * (normfac, xmerge, kappa, x0, xnorm, max_wave_um); max_wave is computed only
 * if want_peak != 0 (else 0). */
int mbb_sed_consts(mbb_ctx *ctx, int64_t n, const double *pars, int layout,
                   int want_peak, double *out_consts, int32_t *out_status,
                   int mem);

/* ---- chain post-processing: mbb_results.compute_peaklambda / compute_lir /
 * compute_dustmass over chain[nwalkers][nsteps][5] (results.py:534-801,
 * 1267-1326), including the sequential allclose-dedupe of _map_chain
 * (results.py:553-566).  which: bit0 peak lambda, bit1 L_IR, bit2 dust mass;
 * outputs [nwalkers][nsteps], NULL where not requested.
 * L_IR: integral of f_nu over [lir_min, lir_max]*(1+z) um in the observer
 * frame times 3.11749657e4*dl_mpc^2*1e-17 (results.py:665-673); dl_mpc <= 0
 * selects prefactor 1, i.e. modified_blackbody.freq_integrate (:639-674).
 * The SED flags and wavenorm are those of mbb_set_model.  */
int mbb_chain_post(mbb_ctx *ctx, int64_t nwalkers, int64_t nsteps,
                   const double *chain, int which, double z, double dl_mpc,
                   double lir_min_um, double lir_max_um, double kappa,
                   double kappa_wave_um, double *out_peak, double *out_lir,
                   double *out_dustmass, int32_t *out_status, int mem);

/* ---- predicted band flux for every chain sample: mbb_results._predict_flux
 * (results.py:895-944) through band `band` of the table set by mbb_set_bands
 * (a single node of weight 1 = flux density at one wavelength).  Same dedupe
 * as mbb_chain_post; SED flags / wavenorm from mbb_set_model. */
int mbb_chain_flux(mbb_ctx *ctx, int64_t nwalkers, int64_t nsteps,
                   const double *chain, int band, double *out_flux,
                   int32_t *out_status, int mem);

/* ---- chain statistics: what mbb_results._parcen_internal / par_lowlim /
 * par_uplim / *_cen take from numpy.mean and numpy.percentile over the flattened
 * chain (results.py:314-431, 946-985).  x[n][ncols] row-major (ncols <= 8: the
 * parameter block of a chain, or one ancillary array with ncols = 1); values
 * outside [lowlim[c], uplim[c]] are dropped (NULL = no limit).  Per column:
 * count of kept values, their mean, and for each of the nq <= 4 quantiles q in
 * [0, 1] the two order statistics around numpy's default ('linear') virtual
 * index (count-1) q and its fractional part gamma, so that the caller forms
 * lo + (hi - lo) gamma exactly as numpy does.  Order statistics are exact (radix
 * selection); a column holding a NaN gives NaN like numpy; an empty column gives
 * count 0 and NaN.  Outputs are host arrays; `mem` describes x. */
int mbb_chain_stats(mbb_ctx *ctx, int64_t n, int ncols, const double *x,
                    const double *lowlim, const double *uplim, int nq,
                    const double *q, double *mean, int64_t *count,
                    double *q_lo, double *q_hi, double *q_gamma, int mem);

/* ---- batch fit: device-resident ensemble sampler for nsrc sources at once (SURVEY 8f
 * row 1, BASELINE configs[4]): what mbb_fitter.run asks emcee for per source
 * (mbb_fit.py:524-543: burn-in, sampler.reset(), main run -> emcee 2.2 stretch move, two
 * half-ensembles per iteration) and what mbb_results derives from the chain
 * (results.py:314-431), with proposal, log-probability, accept/reject and the
 * posterior summaries all on the GPU.
 *   pos[nsrc][nwalkers][5], lnprob[nsrc][nwalkers]: updated in place (lnprob is
 *     computed first unless have_lnprob != 0).
 *   nburn iterations are run and discarded, then nsteps iterations of the main run.
 *   naccept[nsrc][nwalkers] (may be NULL): accepted moves of the main run.
 *   status[nsrc][nwalkers] (may be NULL): first status > 1 a walker's proposals produced.
 *   stats[nsrc][MBB_FIT_NSTATS] (may be NULL): summary over the recorded samples = all
 *     walkers at every thin-th iteration of the main run; row layout MBB_FS_*.
 *   chain[nsteps/thin][nsrc][nwalkers][5], chain_lnprob[nsteps/thin][nsrc][nwalkers]
 *     (may be NULL): the recorded ensembles.  With MBB_HOST they are streamed to the
 *     host in segments behind the sampler (page-locked memory recommended).
 *   chain_nsrc: sources per record of the chain arrays (0 = nsrc).  A shard of a larger
 *     source list passes the full count and pointers offset to its first source, so
 *     that several GPUs fill disjoint slices of one [nrec][chain_nsrc][nwalkers] array.
 * Random numbers: Philox4x32-10 keyed by `seed`, one block per proposal, counter =
 * ((src0 + source)*nwalkers/2 + walker-in-half, 2*(step0 + t) + half); src0 is the global
 * index of this call's source 0, so that shards of one source list draw independent
 * numbers under one seed; pass step0 = iterations already done to continue a run.
 * Delta-band configurations (every band one node, FAST modes) run source-resident:
 * a CTA keeps whole ensembles in shared memory for all iterations of the call. */
#define MBB_FIT_NSTATS 28
enum {
  MBB_FS_N = 0,        /* number of samples                                            */
  MBB_FS_MEAN = 1,     /* [5] mean                                                     */
  MBB_FS_M2 = 6,       /* [5] sum of squared deviations (variance = M2/(N-1))          */
  MBB_FS_MIN = 11,     /* [5]                                                          */
  MBB_FS_MAX = 16,     /* [5]                                                          */
  MBB_FS_BESTLNP = 21, /* largest log-probability among the samples                    */
  MBB_FS_BEST = 22,    /* [5] the sample that has it                                   */
  MBB_FS_ACC = 27      /* mean acceptance fraction of the main run (emcee's             */
                       /* acceptance_fraction averaged over walkers)                   */
};
int mbb_ensemble_fit(mbb_ctx *ctx, int64_t nsrc, int nwalkers, int64_t nburn, int64_t nsteps,
                     double a, uint64_t seed, uint64_t step0, int64_t src0, double *pos,
                     double *lnprob, int have_lnprob, int32_t *naccept, int32_t *status,
                     double *stats, double *chain, double *chain_lnprob, int64_t chain_nsrc,
                     int thin, int mem);

/* mbb_ensemble_fit without burn-in, summaries or source offset (kept for callers that
 * drive burn-in and main run themselves through step0). */
int mbb_ensemble_run(mbb_ctx *ctx, int64_t nsrc, int nwalkers, int64_t nsteps, double a,
                     uint64_t seed, uint64_t step0, double *pos, double *lnprob,
                     int have_lnprob, int32_t *naccept, int32_t *status, double *chain,
                     double *chain_lnprob, int thin, int mem);

/* ---- measurement aid: sustained DFMA rate of this device in TFLOP/s
 * (2 flops per DFMA), the FP64 roofline denominator bench.py reports. */
int mbb_fp64_peak(mbb_ctx *ctx, int iters, double *tflops);

#ifdef __cplusplus
}
#endif
#endif /* MBB_B200_H */

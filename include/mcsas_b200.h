/*
 * mcsas_b200.h -- C-ABI of libmcsas_b200.so: the B200 (sm_100a) replacement for the
 * Monte-Carlo fitting core of McSAS (BAMresearch/McSAS 1.3.1).
 *
 * The reference is pure Python and has no FFI; its "operator API" for this path is
 *   McSAS.analyse / McSAS.mcFit            src/mcsas/mcsas/mcsas.py:191-439
 *   BackgroundScalingFit.calc              src/mcsas/mcsas/backgroundscalingfit.py:112-139
 *   ScatteringModel.calc / generateParameters   src/mcsas/bases/model/scatteringmodel.py:79-127
 *   SASModel.calcIntensity / weight        src/mcsas/bases/model/sasmodel.py:37-79
 *   <Model>.formfactor / volume / absVolume / surface   src/mcsas/models/*.py
 * Each entry point below names the reference call it replaces.  The Python host
 * (mcsas_b200/) binds these with ctypes; INTEGRATION.md shows the stub a McSAS
 * maintainer would add.
 *
 * Conventions: all pointers are caller-owned HOST memory (plain malloc/numpy buffers);
 * every function returns 0 on success or a negative mcsas_status and never throws;
 * mcsas_last_error(ctx) gives the message.  One context per GPU; a context is not
 * thread-safe, distinct contexts are independent.  All quantities are float64 in SI
 * units exactly as the reference hands them to the hot path (q in 1/m, lengths in m).
 */
#ifndef MCSAS_B200_H
#define MCSAS_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define MCSAS_MAX_PARAMS 16
#define MCSAS_MAX_ACTIVE 4
#define MCSAS_MAX_NQ 4096
#define MCSAS_MAX_SPEC 32        /* speculative proposals per batch, fp64 parity kernel (fp32 production kernel: 64) */

typedef struct mcsas_ctx mcsas_ctx;

enum mcsas_status {
    MCSAS_OK = 0,
    MCSAS_ERR_CUDA = -1,        /* CUDA runtime error, see mcsas_last_error */
    MCSAS_ERR_ARG = -2,         /* invalid argument */
    MCSAS_ERR_STATE = -3,       /* call order (no datasets / no model / nothing to fetch) */
    MCSAS_ERR_NO_DEVICE = -4,   /* no usable CUDA device: there is NO CPU fallback */
    MCSAS_ERR_STREAM = -5       /* uploaded proposal stream too short */
};

/* Model plug-ins of the reference (src/mcsas/models/<file>.py); parameter vectors
 * follow each class's `parameters` tuple order. */
enum mcsas_model_id {
    MCSAS_SPHERE = 0,                 /* sphere.py:              radius, sld */
    MCSAS_LMA_DENSE_SPHERE = 1,       /* lmadensesphere.py:      radius, volFrac, mf, sld */
    MCSAS_SPHERICAL_CORE_SHELL = 2,   /* sphericalcoreshell.py:  radius, t, eta_c, eta_s, eta_sol */
    MCSAS_ELLIPSOIDAL_CORE_SHELL = 3, /* ellipsoidalcoreshell.py: a, b, t, eta_c, eta_s, eta_sol, intDiv */
    MCSAS_ELLIPSOIDS_ISOTROPIC = 4,   /* ellipsoidsisotropic.py: a, useAspect, c, aspect, intDiv, sld */
    MCSAS_CYLINDERS_ISOTROPIC = 5,    /* cylindersisotropic.py:  radius, useAspect, length, aspect, intDiv, sld */
    MCSAS_GAUSSIAN_CHAIN = 6,         /* gaussianchain.py:       rg, bp, etas, k */
    MCSAS_KHOLODENKO = 7,             /* kholodenko.py:          radius, lenKuhn, lenContour */
    MCSAS_N_MODELS = 8
};

/* Number generators, src/mcsas/bases/algorithm/numbergenerator.py:28-31,168-189 */
enum mcsas_generator {
    MCSAS_GEN_UNIFORM = 0,   /* RandomUniform */
    MCSAS_GEN_EXP1 = 1,      /* RandomExponential / RandomExponential1 (upper = 1) */
    MCSAS_GEN_EXP2 = 2,      /* RandomExponential2 */
    MCSAS_GEN_EXP3 = 3       /* RandomExponential3 */
};

enum mcsas_quad_rule {
    MCSAS_QUAD_REFERENCE = 0,      /* the reference's own rule: linspace + mean/trapz (intDiv nodes) */
    MCSAS_QUAD_GAUSS_LEGENDRE = 1  /* Gauss-Legendre with quad_nodes nodes */
};

enum mcsas_precision {
    MCSAS_FP64 = 0,   /* parity mode: all float64, reference operation order */
    MCSAS_FP32 = 1    /* production: fp32 form factors (MUFU) and delta sums, fp64 base sums and fit */
};

/* Settings of McSAS (src/mcsas/mcsas/mcsasparameters.json) that reach the hot path. */
typedef struct mcsas_run_config {
    int n_contribs;            /* numContribs */
    int n_reps;                /* numReps: chains per dataset */
    double conv_crit;          /* convergenceCriterion */
    long long max_iter;        /* maxIterations */
    int max_retries;           /* maxRetries: up to max_retries+1 attempts (mcsas.py:220-246); < 0: one attempt */
    double comp_exp;           /* compensationExponent */
    int find_background;       /* findBackground */
    int positive_background;   /* positiveBackground */
    int start_from_minimum;    /* startFromMinimum */
    int precision;             /* mcsas_precision */
    unsigned long long seed;   /* Philox key (production proposals) */
    int spec_max;              /* max speculative proposals per batch: 0 = default (fp64 kernel 32, fp32 kernel 64, 128 on >= 4 CTAs),
                                  1 = none (sequential).  The fp64 kernel's results do not depend on it; in the
                                  fp32 kernel it decides which batches are Gram batches (different rounding) */
    int record_trace;          /* 1: keep per-step accept flag + chi-squared (parity tests) */
    int first_chain;           /* global index of this context's first chain (multi-GPU partition) */
    int cluster_size;          /* CTAs per chain (q split over a thread-block cluster): 0 = auto, 1/2/4/8 */
    long long slice_steps;     /* MC steps per chain and launch before the live chains are re-packed:
                                  0 = auto (fp32 kernel: a time budget per launch), -1 = one launch */
} mcsas_run_config;

/* Scheduler / kernel knobs for experiments (tools/, tests): no counterpart in the reference and no
 * influence on a chain's result (the fp32 kernel's results are independent of cluster size and time
 * slicing).  All 0 = defaults. */
typedef struct mcsas_tuning {
    int max_cluster;   /* largest cluster the scheduler may give a chain: 0 = default (8), 1, 2, 4, 8 */
    int slice_us;      /* time budget of one launch per chain in microseconds, 0 = default */
    int backups;       /* speculative next attempts on idle SMs: 0 = default (on), -1 = off */
    int backup_div;    /* start them when live chains * backup_div <= SMs, 0 = default (2) */
    int gram;          /* Gram batches: 0 = default (on), -1 = off */
    int qpt;           /* force q points per thread (1, 2, 4), 0 = auto */
    int profile;       /* 1: in-kernel phase cycle counters, printed to stderr by mcsas_wait */
    int verbose;       /* 1: one stderr line per time slice */
    int half;          /* 1: 4-warp CTAs, two per SM (default 0: 8-warp CTAs, one per SM) */
    int backup_pct;    /* a chain gets a backup attempt from this percentage of maxIterations on, 0 = default (40) */
    int hedge_pct;     /* with <= #SMs/4 chains alive every chain gets its backup from this percentage on:
                          0 = off (default) */
    int early_pct;     /* ... and, while the machine is still full, the early_top chains with the largest chi-squared */
    int early_top;     /* from early_pct % (0 = 30) on; early_top = 0: off (default) */
    int late_steps;    /* chains with at least this many steps take the largest clusters first, by chi-squared:
                          0 = default (never) */
    int big_batch;     /* fp32 kernel: speculative batches of up to 128 proposals (instead of 64) for chains on at
                          least this many CTAs: 0 = default (4: the chains that run on otherwise idle SMs),
                          1 / 2 / 4 / 8, -1 = never.  (A run_config.spec_max > 0 overrides it.) */
    int gram_in;       /* fp32 kernel: a move at most this many steps behind the previous one switches to Gram
                          batches, 0 = default (8).  NOT result-neutral (a Gram batch rounds a proposal's sums
                          differently): part of the chain's definition, the same on every schedule */
    int gram_out;      /* ... and a Gram batch with fewer than this many moves switches back, 0 = default (1) */
    int gpc_pack;      /* scheduler: bound a mix of cluster sizes by the clusters' shares of the machine (a cluster
                          lives inside one GPC): 0 = default (on), -1 = off (count CTAs only) */
    int l2_keep_pct;   /* fp32 kernel, contribution-row cache: percentage of the L2 the evict_last part of the resident
                          chains' rows may take (the rest of every chain's rows is streamed evict_first): 0 = default
                          (50, applied only when every SM holds a chain -- measured optimum on B200, whose L2 keeps
                          about half of its 126 MB for lines of this kind), > 0 = that percentage, in every slice,
                          -1 = off (every row evict_last, plain LRU) */
    int backup_top;    /* at most this many backup attempts alive at a time, the chains with the largest chi-squared
                          first: 0 = default (no bound) */
    int retry_pct;     /* ... and the backup threshold (percentage of maxIterations) for a repetition whose current attempt
                          is already a retry: 0 = default (at once), -1 = as for first attempts (backup_pct) */
    int repack_pct;    /* a time slice also ends once chains holding this percentage of the SMs have left their launches
                          (the freed SMs go to the remaining chains at once): 0 = default (off) */
} mcsas_tuning;

/* -- context ------------------------------------------------------------- */
int mcsas_ctx_create(int device, mcsas_ctx **out);
void mcsas_ctx_destroy(mcsas_ctx *ctx);
const char *mcsas_last_error(const mcsas_ctx *ctx);
/* Device facts for roofline reporting: sm_count, clock kHz. */
int mcsas_device_info(mcsas_ctx *ctx, int *sm_count, int *clock_khz, char *name, int name_len);

/* -- data: replaces the reads of data.q / data.f.binnedData / data.f.binnedDataU in
 *    mcFit (sasdata.py:51-54, datavector.py:96-117).  Concatenated over datasets. */
int mcsas_upload_datasets(mcsas_ctx *ctx, int n_sets, const int *nq,
                          const double *q, const double *intensity, const double *sigma);

/* -- model: replaces the ScatteringModel instance state (parameter values, which
 *    FitParameters are active, their activeRange ∩ valueRange and generator). */
int mcsas_set_model(mcsas_ctx *ctx, int model_id, const double *params, int n_params,
                    const int *active_idx, int n_active,
                    const double *active_lo, const double *active_hi, const int *generator_kind,
                    int quad_rule, int quad_nodes);

/* -- smearing: replaces data.config.smearing.prepared (qOffset, weights) and data.locs of the
 *    smeared branch of SASModel.calcIntensity (sasmodel.py:56-73; sasconfig.py:122-149, 209-233,
 *    308-339).  Applies to the reference's canSmear models (Sphere, LMADenseSphere) in
 *    mcsas_eval / mcsas_run / mcsas_fractions; two_d_coll = SmearingConfig.twoDColl (0: slit,
 *    locations sqrt(q^2 + qOffset^2); 1: 2-D averaged, q + qOffset).  n_points = 0 switches it off. */
int mcsas_set_smearing(mcsas_ctx *ctx, int n_points, const double *q_offset, const double *weights,
                       int two_d_coll);

/* -- ScatteringModel.calc(data, pset, compExp) (scatteringmodel.py:79-105) for M
 *    single contributions of dataset `set`: intensity[M][nq] (=F^2*w), v, w, s [M]. */
int mcsas_eval(mcsas_ctx *ctx, int set, const double *pset, int m, double comp_exp, int precision,
               double *intensity, double *v, double *w, double *s);

/* -- McSAS.analyse (mcsas.py:191-285): run n_sets * n_reps independent mcFit chains
 *    (mcsas.py:287-439), one CTA each.  proposal_stream: NULL (in-kernel Philox) or
 *    [n_chains][stream_stride] number-generator outputs in [0,1) in the reference's
 *    consumption order (deterministic parity mode, single attempt per chain).
 *    stop_flag: NULL or a host int the caller may set non-zero to stop (McSAS.stop);
 *    it is polled between time slices. */
int mcsas_run(mcsas_ctx *ctx, const mcsas_run_config *cfg,
              const double *proposal_stream, long long stream_stride,
              volatile int *stop_flag);
/* Split form for timing: launch asynchronously / wait; mcsas_run == launch + wait. */
int mcsas_run_async(mcsas_ctx *ctx, const mcsas_run_config *cfg,
                    const double *proposal_stream, long long stream_stride,
                    volatile int *stop_flag);
/* kernel_ms: device time (CUDA events) from the start of the first time-slice launch to the end of
 * the last one, i.e. including the scheduler's gaps between slices. */
int mcsas_wait(mcsas_ctx *ctx, float *kernel_ms);
/* Ask a running fit to stop at its next time-slice boundary (McSAS.stop). */
int mcsas_request_stop(mcsas_ctx *ctx);
/* Experiment knobs (NULL: back to the defaults). */
int mcsas_set_tuning(mcsas_ctx *ctx, const mcsas_tuning *tuning);
/* Facts about the last run: kernel launches, last cluster size, summed kernel ms. */
int mcsas_run_info(mcsas_ctx *ctx, int *n_launches, int *last_cluster, float *kernel_ms);
/* Scheduler facts of the last run: backup attempts (the next retry of a late chain, mcsas.py:220-246,
 * started ahead of time on idle SMs) started and promoted to the chain's current attempt. */
int mcsas_scheduler_info(mcsas_ctx *ctx, int *backups_started, int *backups_used);

/* -- results of the last run (any pointer may be NULL).  Layouts:
 *    contribs [set][N][P][R] (numpy zeros((N,P,R)) per set, mcsas.py:203-205),
 *    fit_meas [set][nq][R] (mcsas.py:210), everything else [set][R].  num_iter/num_moves are
 *    those of the final attempt (mcsas.py:249); steps_all counts the MC steps of all attempts,
 *    evals all contribution evaluations incl. discarded speculative ones. */
int mcsas_fetch(mcsas_ctx *ctx, double *contribs, double *fit_meas,
                double *scaling, double *background, double *conval, double *conval0,
                long long *num_iter, long long *num_moves, int *attempts, double *elapsed_s,
                long long *evals, long long *steps_all);
/* Per-step trace of chain `chain` (record_trace=1): accept[num_iter], chi[num_iter]. */
int mcsas_fetch_trace(mcsas_ctx *ctx, int chain, unsigned char *accept, double *chi,
                      long long capacity);

/* -- McSAS.histogram per-repetition part (mcsas.py:548-604): volume/number/intensity/
 *    surface fractions and minimum-observability limits of the fetched contributions.
 *    All outputs [set][N][R]; totals [set][4][R]; sc [set][2][R]. */
int mcsas_fractions(mcsas_ctx *ctx, const mcsas_run_config *cfg, const double *contribs,
                    double *vol_frac, double *num_frac, double *volsqr_frac, double *surf_frac,
                    double *min_vol, double *min_num, double *min_volsqr, double *min_surf,
                    double *totals, double *sc);

/* -- data preparation of a BATCH of raw 1-D tables: replaces, per dataset,
 *    DataObj._prepareUncertainty (dataobj/dataobj.py:204-226: sigma <- max(sigma, fu_min * I), invalid -> inf),
 *    _initMask/_applyFMasks/_applyLimits (:239-269: finite I, fMaskZero / fMaskNeg, x0Low <= q <= x0High) and
 *    DataObj._reBin (:288-345: log-spaced bins; one row: kept; several: means and max(SEM, propagated sigma);
 *    empty bins dropped) -- bit-identical to the reference's numpy arithmetic (pairwise sums, row order).
 *    q / intensity / sigma: the n_raw[s] rows of all tables concatenated, q in 1/m; sigma NULL = no uncertainty
 *    column.  n_bin <= 0: no rebinning (DataConfig.nBin = 0).  edges [n_sets][n_bin + 1]: the bin edges in 1/m as
 *    the reference computes them (numpy.logspace(log10(min q), log10(max q + last step / 100), n_bin + 1) over
 *    the masked rows, :308-311) -- computed by the caller so that they are numpy's bits.
 *    Outputs: nq_out[n_sets], and q_out / i_out / sigma_out [n_sets][cap] (cap >= n_bin, or >= the longest
 *    table when n_bin <= 0); feed them to mcsas_upload_datasets. */
int mcsas_prepare_datasets(mcsas_ctx *ctx, int n_sets, const int *n_raw, const double *q,
                           const double *intensity, const double *sigma, double fu_min, double x0_low,
                           double x0_high, int mask_zero, int mask_neg, int n_bin, const double *edges,
                           int cap, int *nq_out, double *q_out, double *i_out, double *sigma_out);

/* -- measured pipe peaks of this GPU for the FP32/SFU roofline (bench.py):
 *    MUFU (sin+cos+rcp mix) Gop/s and FP32 FMA TFLOP/s, whole chip. */
int mcsas_measure_peaks(mcsas_ctx *ctx, double *mufu_gops, double *ffma_tflops);

#ifdef __cplusplus
}
#endif
#endif

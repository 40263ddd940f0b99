/*
 * exauq_b200.h -- C ABI of libexauq_b200.so, the B200 (sm_100a) Gaussian-process engine that
 * replaces the mogp-emulator/numpy arithmetic beneath EXAUQ-Toolbox's
 * AbstractGaussianProcess interface (exauq/core/modelling.py:904-1138).
 *
 * The reference has no FFI: its "plugin API" is the Python ABC, implemented by
 * MogpEmulator (exauq/core/emulators.py:69-667), which calls mogp-emulator.  Each entry
 * point below names the reference call it replaces.  The Python host class
 * (exauq-toolbox_b200/gp.py: B200GaussianProcess) binds these with ctypes; see INTEGRATION.md.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, float64 throughout, row-major host arrays.
 *   - Return value: 0 = ok; < 0 = invalid argument / CUDA error (see exq_last_error());
 *     > 0 = numerical condition (EXQ_NOT_PD: covariance not positive definite).
 *   - Host pointers are owned by the caller; device memory is owned by the library behind
 *     opaque handles.  One device per process (exq_init), one internal stream; calls are
 *     synchronous with respect to the host unless stated.  Not thread-safe per handle.
 *   - There is no CPU fallback: every compute entry point fails with EXQ_ERR_CUDA when no
 *     B200 is present.
 */
#ifndef EXAUQ_B200_H
#define EXAUQ_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EXQ_OK 0
#define EXQ_ERR_ARG (-1)
#define EXQ_ERR_CUDA (-2)
#define EXQ_ERR_STATE (-3) /* e.g. predict before fit */
#define EXQ_NOT_PD 1       /* Cholesky met a non-positive pivot (mogp: LinAlgError) */

/* kernel ids: table at exauq/core/emulators.py:122-126 */
#define EXQ_MATERN52 0
#define EXQ_SQEXP 1
#define EXQ_PRODMAT52 2

/* nugget modes: mogp GaussianProcess(nugget=...) as passed through exauq/core/emulators.py:129-141 */
#define EXQ_NUGGET_FIXED 0    /* value supplied (float nugget, or 'fit' after estimation) */
#define EXQ_NUGGET_ADAPTIVE 1 /* 0 unless Cholesky fails, then mean(diag K)*1e-6*10^k, k < 5 */
#define EXQ_NUGGET_FIT 2      /* exq_gp_logpost_batch only: the nugget is the last raw parameter (ln nugget) */

typedef struct exq_gp_s* exq_gp_t;
typedef struct exq_cand_s* exq_cand_t;

/* ---- process / device ------------------------------------------------------------- */
int exq_init(int device);              /* select device, create stream; idempotent per device */
int exq_shutdown(void);
const char* exq_last_error(void);
int exq_device_sm_count(void);
/* Digits per operand of the int8 tensor-core variance product (csrc/ozaki_i8.cuh): 5..7 base-256
 * digits (48 bits at the default 6), or 0 when the FP64 DMMA GEMM is used instead
 * (EXQ_OZAKI_SLICES=0).  Either way the result is the FP64 quantity k^T K^-1 k of
 * mogp GaussianProcess.predict (exauq/core/emulators.py:649). */
int exq_variance_slices(void);
/* Same for the large GEMMs of the potrf + trtri recursion and LAUUM inside fits and batched
 * log-posterior evaluations (csrc/ozaki_gemm.cuh; default 7 digits = 56 bits, EXQ_OZAKI_FACTOR_SLICES). */
int exq_factor_slices(void);
/* Run-time switch of both (0 = FP64 DMMA).  Affects GP handles created afterwards; used by the
 * parity tests to compare the two arithmetic paths on identical inputs. */
int exq_set_int8_digits(int variance_digits, int factor_digits);
/* number of kernels this library launched since process start (bench.py "gpu_launches") */
int64_t exq_launch_count(void);

/* Conditioning guard of the int8 paths.  kappa = (max L_ii / min L_ii)^2 -- a lower bound of cond_2(K) -- is read
 * back with every factorisation.  kappa >= widen: products configured with 6 digits (variance, LAUUM) run with 7;
 * kappa >= fallback: the factorisation is redone on the FP64 DMMA path and the system stays there.  Defaults
 * 2^16 and 1e12 (EXQ_GUARD_WIDEN / EXQ_GUARD_FALLBACK).  The reference has no counterpart: LAPACK's FP64 dpotrf
 * (mogp GaussianProcess.fit via exauq/core/emulators.py:447-448) is what the guard falls back towards. */
int exq_set_guard(double widen, double fallback);
int exq_guard_counts(int64_t* widened, int64_t* fallbacks); /* how often each rule fired since exq_init */
/* Steps of iterative refinement of alpha = K^-1 y after a factorisation that used the int8 kernels (default 1,
 * EXQ_REFINE_STEPS; 0 switches it off): r = y - K alpha with K re-evaluated from the points in FP64,
 * alpha += Linv^T Linv r.  Keeps the predictive mean k^T alpha (mogp GaussianProcess.predict,
 * exauq/core/emulators.py:649) at the level of a backward-stable FP64 solve for ill-conditioned K. */
int exq_set_refine_steps(int steps);
/* Fused sub-tree kernel (csrc/fused_node.cuh): diagonal blocks of up to max_rows rows (256..1024, default 512;
 * 0 = off: one launch per leaf and per node GEMM as in round 1) are factored and inverted by one launch, a
 * thread-block cluster of cluster_ctas CTAs per system (0 = automatic: 16 for up to 8 systems, else 8).
 * EXQ_FUSE_N / EXQ_FUSE_G.  Same arithmetic either way (FP64 DMMA): used by tests to compare the two paths. */
int exq_set_fused(int max_rows, int cluster_ctas);
/* Prediction / PEI pipeline over candidate chunks (csrc/pei_fused.cuh).  on (default): the cross-covariance tile
 * kernel writes the int8 digit planes of the variance product directly and accumulates the mean and the repulsion
 * product per tile -- the FP64 covariance chunk is never materialised.  off: round 1's three passes (FP64 chunk,
 * slicer, per-candidate epilogue).  EXQ_PEI_FUSED.  Same quantities either way (mogp GaussianProcess.predict,
 * exauq/core/emulators.py:649; PEICalculator.compute, exauq/core/designers.py:522-706). */
int exq_set_pei_fused(int on);
/* Gradient of the log-posterior (mogp logpost_deriv, exauq/utilities/mogp_fitting.py:318-324).  on (default): when
 * K^-1 = Linv^T Linv runs on the int8 kernel (stationary kernels, d <= 8) its epilogue contracts every tile with
 * dK/dtheta while the tensor pipe works on the next one -- K^-1 is never stored and there is no gradient kernel.
 * off: LAUUM stores K^-1 and grad_fast_kernel contracts it (round 2's first half).  EXQ_GRAD_FUSED. */
int exq_set_grad_fused(int on);
/* current max_rows and how many clusters of 8 / 16 CTAs of that kernel the device co-schedules */
int exq_fused_info(int* max_rows, int* clusters8, int* clusters16);
/* kappa of the last exq_gp_fit of this handle and whether the guard sent that fit to FP64 DMMA */
int exq_gp_kappa(exq_gp_t gp, double* kappa, int* on_fp64);
/* int8 operations (2 per multiply-accumulate) oz::gemm_kernel issues for B systems of an M x N x K product with
 * `digits` base-256 digits: the kernel's own tile walk (128 x 64 tiles, 64-byte k blocks, triangular k ranges
 * krange = 0 full, 1 k <= j, 2 k >= j, 3 k <= i, 4 k >= i; lower_only skips tiles above the diagonal).  Host
 * arithmetic only -- the roofline's operation count, exposed so that tests can re-derive it.  < 0: bad shape. */
double exq_int8_gemm_ops(int M, int N, int K, int B, int krange, int lower_only, int digits);
/* The balanced static work list oz::gemm_kernel walks for this product on `grid` CTAs (csrc/ozaki_gemm.cuh, make_schedule):
 * entry i is the item of CTA i % grid in round i / grid, packed (system << 20 | row tile << 10 | column tile),
 * 0xffffffff = idle.  Writes at most `cap` entries to `out` (may be NULL) and returns the length of the list, < 0 for a
 * bad shape.  Host arithmetic only -- exposed so that tests can check coverage and balance without a GPU. */
int64_t exq_int8_gemm_schedule(int M, int N, int K, int B, int krange, int lower_only, int grid, uint32_t* out, int64_t cap);

/* CUDA-event timing on the library stream (bench.py / roofline).  exq_timer_* brackets a
 * region; exq_prof_* accumulates per-class kernel time between reset and get. */
int exq_timer_start(void);
int exq_timer_stop(double* ms);
#define EXQ_PROF_GEMM_PREDICT 0 /* variance GEMM  Z = Linv Kc^T + column sum of squares; flops = FP64-equivalent
                                   Np^2 Mc, bytes = int8 operations issued when the int8 path is active */
#define EXQ_PROF_GEMM_FACTOR 1  /* TRSM/SYRK/TRTRI/LAUUM tile GEMMs */
#define EXQ_PROF_LEAF 2         /* 128 x 128 potrf+trtri leaves */
#define EXQ_PROF_ASSEMBLE 3     /* covariance assembly */
#define EXQ_PROF_STREAM 4       /* per-candidate epilogue, PEI update/argmax, vector kernels */
#define EXQ_PROF_GEMM_FACTOR_I8 5 /* large nodes of the recursion + LAUUM on the int8 tensor cores (flops: FP64-equivalent) */
#define EXQ_PROF_SLICE 6          /* exponent + base-256 digit passes feeding the int8 kernels (bytes moved) */
#define EXQ_PROF_FUSED 7          /* fused sub-tree kernel: potrf + trtri of diagonal blocks <= 512 rows in one launch */
#define EXQ_PROF_LAUUM_GRAD 8     /* int8 LAUUM launches whose epilogue carries the gradient contraction (int8 ops issued in
                                   * `bytes`, FP64-equivalent LAUUM flops in `flops`; the FP64 pair arithmetic rides on top) */
#define EXQ_PROF_NCLASS 9
int exq_prof_enable(int class_mask); /* bit i set: time launches of class i with event pairs */
int exq_prof_reset(void);
int exq_prof_get(int cls, double* total_ms, int64_t* launches, double* flops, double* bytes);

/* ---- correlation (stateless) ---------------------------------------------------------
 * out[i*n2 + j] = kernel(A[i], B[j]; corr_raw).
 * Replaces mogp Kernel.<name>().kernel_f(x1, x2, corr_raw) as called by
 * MogpEmulator.correlation, exauq/core/emulators.py:495-567 (line 547). */
int exq_corr(int kernel_id, int d, const double* corr_raw, const double* A, int64_t n1,
             const double* B, int64_t n2, double* out);

/* ---- a Gaussian process bound to training data -------------------------------------
 * Replaces mogp GaussianProcess(inputs, targets, kernel=..., nugget=...) as constructed at
 * exauq/core/emulators.py:414,447.  X is [N][d], y is [N]. */
int exq_gp_create(int64_t N, int d, const double* X, const double* y, int kernel_id,
                  exq_gp_t* out);
int exq_gp_destroy(exq_gp_t gp);

/* Fit with given hyperparameters: K = cov*C (+ nugget I), L = chol(K), Linv, alpha = K^-1 y.
 * Replaces mogp GaussianProcess.fit(theta) as called by
 * MogpEmulator._fit_gp_with_hyperparameters, exauq/core/emulators.py:418-450.
 * corr_raw[d] = -2 ln(length scale), cov = process variance (linear scale).
 * Outputs: nugget_used (the jitter in adaptive mode), logdet = log|K|, quad = y^T K^-1 y;
 * the negative log-likelihood is 0.5*(logdet + quad + N ln 2 pi). */
int exq_gp_fit(exq_gp_t gp, const double* corr_raw, double cov, double nugget, int nugget_mode,
               double* nugget_used, double* logdet, double* quad);

/* nugget = "pivot" (mogp GaussianProcess.fit -> linalg.pivot_cholesky -> LAPACK dpstrf; the option is passed through
 * at exauq/core/emulators.py:129-137 and exercised by tests/unit/test_emulators.py:35, 683, 706).
 * exq_gp_pivot_order: pivoted Cholesky (LAPACK's unblocked lower variant dpstf2: first arg max of the running
 * Schur diagonal, stop at n * 2^-53 * max diag) of the nugget-free covariance of the handle's data; returns the
 * 0-based pivot order piv[N] and the numerical rank.  The handle is left unfitted.
 * exq_gp_fit_pivoted: the handle's data are ALREADY in pivot order (the host re-creates the handle with X[piv],
 * y[piv], so that every downstream quantity keeps its triangular structure); factors the leading `rank` columns
 * without interchanges, replaces the diagonal of the remaining rows by mogp's decreasing sequence
 * L[rank-1][rank-1] / ((rank+1)...(i+1)), forms Linv and alpha as exq_gp_fit does.  The variance of predictions carries
 * no nugget in this mode (mogp predict).  rank == N and a well-conditioned matrix: exq_gp_fit with nugget 0 on the
 * permuted handle gives the same factor through the fast recursion. */
int exq_gp_pivot_order(exq_gp_t gp, const double* corr_raw, double cov, int* piv, int* rank);
int exq_gp_fit_pivoted(exq_gp_t gp, const double* corr_raw, double cov, int rank, double* logdet, double* quad);

/* Predictive mean and variance at M points Xc[M][d] (host):
 * mean = k^T alpha, var = max(cov + nugget - k^T K^-1 k, 0).
 * Replaces mogp GaussianProcess.predict as called by MogpEmulator.predict,
 * exauq/core/emulators.py:612-667 (line 649), for a whole batch of inputs. */
int exq_gp_predict(exq_gp_t gp, const double* Xc, int64_t M, double* mean, double* var);

/* Covariance of n query points with the training inputs: out[n][N] = cov * C(Xq, train) (no
 * nugget), using the resident training data and the fitted hyperparameters.
 * Replaces AbstractGaussianProcess.covariance_matrix, exauq/core/modelling.py:1042-1080
 * (MogpEmulator.covariance_matrix, exauq/core/emulators.py:569-610). */
int exq_gp_cross_cov(exq_gp_t gp, const double* Xq, int64_t n, double* out);

/* Leave-one-out sweep with the fitted hyperparameters: for every i, the prediction at x_i
 * of the GP refitted on data \ {i}, and its NES error against y_i.
 * Replaces the loop compute_loo_errors_gp runs, exauq/core/designers.py:263-272
 * (compute_loo_gp :295-370, predict, nes_error exauq/core/modelling.py:754-825).
 * Uses the closed form mu_i = y_i - [K^-1 y]_i / [K^-1]_ii, var_i = 1/[K^-1]_ii. */
int exq_gp_loo(exq_gp_t gp, double* mean, double* var, double* nes);

/* The same quantities by EXPLICIT refits (the reference's algorithm): for each of the cnt
 * indices idx[], assemble the (N-1)-point system, factor it and predict at the left-out
 * point.  Batched over the indices.  Used to validate exq_gp_loo and as a benchmark. */
int exq_gp_loo_refit(exq_gp_t gp, const int* idx, int cnt, double* mean, double* var, double* nes);

/* Inverse of the nugget-free covariance cov*C(train, train), out[N][N].
 * Replaces AbstractGaussianProcess._compute_kinv, exauq/core/modelling.py:1095-1113
 * (np.linalg.inv).  Returns EXQ_NOT_PD if that matrix is numerically singular. */
int exq_gp_kinv(exq_gp_t gp, double* out);

/* Negative log-likelihood and its gradient for R hyperparameter vectors at once.
 * theta_raw[R][d+1] = (corr_raw[d], ln cov); nugget handling as in exq_gp_fit (one setting
 * for all R).  With EXQ_NUGGET_FIT the rows have d+2 entries (..., ln nugget) and so do the
 * gradient rows.  val[r] = 0.5*(log|K| + y^T K^-1 y + N ln 2 pi); grad[r][.] = d val / d raw;
 * info[r] = 0 or EXQ_NOT_PD; nugget_used[r] as in exq_gp_fit.  Priors are added by the host.
 * Replaces mogp GaussianProcess.logposterior / logpost_deriv as driven by
 * _fit_single_GP_MAP, exauq/utilities/mogp_fitting.py:289-346 (lines 318-324). */
int exq_gp_logpost_batch(exq_gp_t gp, int R, const double* theta_raw, double nugget,
                         int nugget_mode, double* val, double* grad, int* info,
                         double* nugget_used);

/* ---- candidate sets resident in HBM (PEI maximisation over fixed candidates) -------- */
int exq_cand_create(int64_t M, int d, const double* Xc, exq_cand_t* out);
int exq_cand_destroy(exq_cand_t c);
/* overwrite the coordinates of an existing set (same M, d): one H2D copy, no reallocation */
int exq_cand_set_points(exq_cand_t c, const double* Xc);

/* Evaluate mean, variance and pseudo-expected improvement of the fitted GP at every
 * candidate.  rep[n_rep][d] are the repulsion points other than the training inputs
 * (pseudopoints, additional points); tmax is the largest training output.
 * Replaces PEICalculator.compute = expected_improvement * repulsion,
 * exauq/core/designers.py:522-553, 610-658, 660-706, evaluated per candidate. */
int exq_pei_eval(exq_gp_t gp, exq_cand_t c, const double* rep, int n_rep, double tmax);
/* arg max of the current PEI values (lowest index wins ties; NaN ignored) */
int exq_pei_argmax(exq_cand_t c, int64_t* idx, double* val);
/* A new design point x[d] joins the repulsion set: pei *= 1 - corr(., x)
 * (PEICalculator.add_repulsion_points, exauq/core/designers.py:555-608, used at :835). */
int exq_pei_add_point(exq_gp_t gp, exq_cand_t c, const double* x);
/* batch_size rounds of (argmax, add chosen candidate) without host round trips;
 * replaces the loop at exauq/core/designers.py:830-835 with a fixed candidate set. */
int exq_pei_batch(exq_gp_t gp, exq_cand_t c, int batch_size, int64_t* idx, double* val);
/* copy results back: which = 0 mean, 1 variance, 2 pei */
int exq_cand_download(exq_cand_t c, int which, double* out);
/* device address of the {double val; int64 idx} argmax pair, for NCCL allgather */
void* exq_cand_best_devptr(exq_cand_t c);

#ifdef __cplusplus
}
#endif
#endif

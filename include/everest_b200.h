/*
 * everest_b200 -- C ABI of the B200-native acquisition-evaluation path.
 *
 * The reference (BoFire, /root/reference) is pure Python over BoTorch and has NO native
 * interface for this path; each entry point below therefore cites the Python call it
 * replaces.  Plain pointers and sizes only; row-major float64 everywhere; no torch types.
 *
 * Pointer convention: `*_dev` arguments are CUDA device pointers (the host side hands in
 * `tensor.data_ptr()`), everything else is host memory.  `stream` is a cudaStream_t passed
 * as void* (NULL = default stream).  All functions return BO_OK (0) or a negative error
 * code and never throw; `bo_last_error()` returns a message for the calling thread.
 * Numerical failure is reported LAPACK-style through `info` outputs (0 = fine, k > 0 =
 * first non-positive pivot k), mirroring BoTorch's NotPSDError-driven fallbacks
 * ("Failure detection").
 */
#ifndef EVEREST_B200_H
#define EVEREST_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BO_OK 0
#define BO_ERR_INVALID -1   /* bad argument / unsupported configuration (Python: ValueError) */
#define BO_ERR_CUDA -2      /* CUDA runtime failure */
#define BO_ERR_NOT_PSD -3   /* matrix not positive definite after jitter escalation (NotPSDError) */
#define BO_ERR_STATE -4     /* call order violated (e.g. forward before prepare) */

#define BO_MAX_LEAVES 8
#define BO_MAX_TERMS 8
#define BO_MAX_FACTORS 4
#define BO_MAX_Q 16
#define BO_MAX_OBJECTIVES 8
#define BO_MAX_CONSTRAINTS 8

/* Leaf kernels: bofire/kernels/mapper.py:31-69 (RBF, Matern), :206-253 + kernels/categorical.py:43-70
 * (Hamming on one-hot groups), :191-203 + fingerprint_kernels/base_fingerprint_kernel.py:36-53 (Tanimoto). */
enum bo_leaf_kind {
  BO_LEAF_RBF = 0,
  BO_LEAF_MATERN12 = 1,
  BO_LEAF_MATERN32 = 2,
  BO_LEAF_MATERN52 = 3,
  BO_LEAF_HAMMING = 4,
  BO_LEAF_TANIMOTO = 5
};

typedef struct {
  int32_t kind;           /* bo_leaf_kind */
  int32_t n_dims;         /* RBF/Matern: active columns; Hamming: one-hot groups; Tanimoto: bit columns */
  const int32_t* dims;    /* [n_dims] column index (Hamming: START column of each group) */
  const int32_t* cardinality; /* Hamming only: [n_dims] group sizes */
  const double* lengthscale;  /* [n_ls]; NULL for Tanimoto */
  int32_t n_ls;           /* 1 = isotropic, n_dims = ARD (Hamming: first n_dims entries are used) */
} bo_kernel_leaf;

/* A Scale/Additive/Multiplicative tree (kernels/mapper.py:126-188, surrogates/mixed_tanimoto_gp.py:101-215)
 * flattened by the host into  K = sum_t coef_t * prod_{l in factors_t} leaf_l. */
typedef struct {
  double coef;
  int32_t n_factors;
  int32_t factors[BO_MAX_FACTORS];
} bo_kernel_term;

/* One single-output exact GP (SingleTaskGPSurrogate._fit, surrogates/single_task_gp.py:39-71):
 * Normalize input transform on selected columns, Standardize outcome transform, constant mean,
 * homoskedastic noise.  M of these form the ModelListGP (botorch_surrogates.py:124-128). */
typedef struct {
  int32_t n_leaves;
  const bo_kernel_leaf* leaves;
  int32_t n_terms;
  const bo_kernel_term* terms;
  const double* in_offset; /* [d] x' = (x - in_offset) / in_scale ; 0 / 1 on untransformed columns */
  const double* in_scale;  /* [d] */
  double mean_const;       /* constant mean in standardised outcome space */
  double noise;            /* noise variance in standardised outcome space */
  double y_mean, y_std;    /* Standardize(m=1) */
  const double* y;         /* [N] raw training targets */
} bo_output_model;

typedef struct {
  int32_t N, d, M;
  const double* X_train;   /* [N, d] host, BoFire-transformed input space (one-hot / 0-1 fingerprint columns as doubles) */
  const bo_output_model* outputs; /* [M] */
} bo_state_config;

/* Objective callables as an op table (utils/torch_tools.py:384-450). p0..p2 by kind:
 *   MAX/MIN: lower_bound, upper_bound        CLOSE_TO_TARGET: target_value, exponent
 *   MIN_SIGMOID/MAX_SIGMOID: steepness, tp   TARGET: target_value, tolerance, steepness */
enum bo_objective_kind {
  BO_OBJ_MAX = 0, BO_OBJ_MIN = 1, BO_OBJ_CLOSE_TO_TARGET = 2,
  BO_OBJ_MIN_SIGMOID = 3, BO_OBJ_MAX_SIGMOID = 4, BO_OBJ_TARGET = 5
};
typedef struct { int32_t kind; int32_t out_idx; double p0, p1, p2; double w; } bo_objective_op;

/* Output constraints (constrained_objective2botorch, utils/torch_tools.py:258-337):
 * c(y) = sign * (y[out_idx] - tp), feasible iff c <= 0, smoothed weight sigmoid(-c / eta). */
typedef struct { int32_t out_idx; double sign; double tp; double eta; } bo_constraint_op;

enum bo_combine { BO_COMBINE_SINGLE = 0, BO_COMBINE_ADDITIVE = 1, BO_COMBINE_MULTIPLICATIVE = 2 };

typedef struct bo_state bo_state;

int bo_version(void);
const char* bo_last_error(void);

/* Replaces the model construction the acquisition path consumes: BotorchSurrogates.compatibilize ->
 * ModelListGP (botorch_surrogates.py:79-128) + the GPyTorch prediction-strategy caches.  Copies
 * everything; the config may be freed after the call. */
int bo_state_create(const bo_state_config* cfg, bo_state** out);

/* New hyper-parameter VALUES (lengthscales, term coefficients = products of the outputscales, noise, constant mean) for
 * output m of an existing state; kernel tree, columns, transforms and targets stay.  What one step of fit_gpytorch_mll
 * changes (surrogates/single_task_gp.py:70-71): the large buffers of the handle are reused, bo_state_factorize must follow. */
int bo_state_set_hyperparameters(bo_state* st, int32_t m, const bo_output_model* om, void* stream);
void bo_state_destroy(bo_state* st);

/* Training Gram K + sigma^2 I, psd-safe Cholesky (jitter 1e-8 .. 1e-3), mean cache alpha and the
 * inverse root L^-1 -- what gpytorch's DefaultPredictionStrategy caches on the first
 * model.posterior() call (reached from botorch.py:180,223).  info[M]: 0 or first bad pivot;
 * jitter[M]: diagonal jitter that was needed. */
int bo_state_factorize(bo_state* st, int32_t* info, double* jitter, void* stream);

/* model.posterior(X, observation_noise) marginals: BotorchStrategy._predict (botorch.py:174-194).
 * X_dev [n, d]; mean_dev, var_dev [n, M] in the original outcome space. */
int bo_posterior_marginal(bo_state* st, const double* X_dev, int32_t n, int32_t observation_noise,
                          double* mean_dev, double* var_dev, void* stream);

/* Joint posterior of one point set (used by tests and by acquisition set-up): mean_dev [n, M],
 * cov_dev [M, n, n]. */
int bo_posterior_joint(bo_state* st, const double* X_dev, int32_t n, double* mean_dev, double* cov_dev,
                       void* stream);

/* prune_inferior_points_multi_objective as called by qNoisyExpectedHypervolumeImprovement(
 * prune_baseline=True) (qnehvi.py:39-51): joint posterior samples at X [n, d] with base samples
 * z_dev [S, n, M]; counts_dev[n] receives, per point, the number of samples in which it is
 * non-dominated and strictly better than ref_point (host keeps counts > 0). */
int bo_prune_counts(bo_state* st, const double* X_dev, int32_t n, const double* z_dev, int32_t S,
                    const bo_objective_op* obj, int32_t n_obj, const bo_constraint_op* cons, int32_t n_cons,
                    const double* ref_point, int32_t* counts_dev, int32_t* info, void* stream);

/* qNoisyExpectedHypervolumeImprovement.__init__ after pruning (NoisyExpectedHypervolumeMixin.
 * _set_cell_bounds): posterior at X_baseline, cached root baseline_L, S baseline samples from
 * zb_dev [S, n_b, M], objective transform, per-sample non-dominated front + box decomposition.
 * Also serves MoboStrategy (mobo.py:72-90).  info[M]; returns max cells per sample in *max_cells. */
int bo_nehvi_prepare(bo_state* st, const double* Xb_dev, int32_t n_b, const double* zb_dev, int32_t S,
                     const bo_objective_op* obj, int32_t n_obj, const bo_constraint_op* cons, int32_t n_cons,
                     const double* ref_point, int32_t* info, int32_t* max_cells, void* stream);

/* qExpectedHypervolumeImprovement with a fixed partitioning of the observed front (qehvi.py:37-77):
 * Yobj_dev [n, n_obj] objective values (maximisation frame) of the observations. */
int bo_ehvi_prepare(bo_state* st, const double* Yobj_dev, int32_t n, int32_t S,
                    const bo_objective_op* obj, int32_t n_obj, const double* ref_point, int32_t* max_cells,
                    void* stream);

/* qLogExpectedImprovement as built by SoboStrategy._get_acqfs (sobo.py:51-90). */
int bo_logei_prepare(bo_state* st, int32_t S, int32_t combine, const bo_objective_op* obj, int32_t n_obj,
                     double best_f, void* stream);

/* The single-objective MC acquisition functions SoboStrategy can select (data_models/acquisition_functions/
 * acquisition_function.py:21-60, built through get_acquisition_function at sobo.py:64-89). */
enum bo_scalar_acqf { BO_ACQF_QLOGEI = 0, BO_ACQF_QEI = 1, BO_ACQF_QSR = 2, BO_ACQF_QUCB = 3, BO_ACQF_QPI = 4 };

/* General form of bo_logei_prepare: `variant` from bo_scalar_acqf, `param` = beta (qUCB) / tau (qPI); output
 * constraints multiply the utility by prod sigmoid(-c/eta) (qEI, qPI) or add sum log fatmoid(-c/eta) (qLogEI).
 * With n_b > 0 the NOISY variants are built (qNEI for BO_ACQF_QEI, qLogNEI for BO_ACQF_QLOGEI -- SoboStrategy's default,
 * data_models/strategies/predictives/sobo.py:15-17): posterior root at the baseline Xb_dev [n_b, d] is cached, S
 * baseline samples are drawn from zb_dev [S, n_b, M] and the incumbent becomes the best baseline objective of each MC
 * sample; best_f is then ignored. */
int bo_scalar_prepare(bo_state* st, int32_t variant, double param, int32_t S, int32_t combine, const bo_objective_op* obj,
                      int32_t n_obj, const bo_constraint_op* cons, int32_t n_cons, double best_f, const double* Xb_dev,
                      int32_t n_b, const double* zb_dev, int32_t* info, void* stream);

/* prune_inferior_points for the noisy single-objective variants: counts_dev[n] = number of joint posterior samples
 * (base samples z_dev [S, n, M]) in which point i has the best scalarised objective; samples that violate an output
 * constraint (c(y) > 0) count as -inf, as in BoTorch (SoboStrategy hands its sigmoid / target outputs over as
 * constraints, sobo.py:120-150). */
int bo_prune_counts_scalar(bo_state* st, const double* X_dev, int32_t n, const double* z_dev, int32_t S, int32_t combine,
                           const bo_objective_op* obj, int32_t n_obj, const bo_constraint_op* cons, int32_t n_cons,
                           int32_t* counts_dev, int32_t* info, void* stream);

/* [UPSTREAM] compute_best_feasible_objective for the noisy variants: recomputes the per-sample incumbent of the prepared
 * qNEI / qLogNEI with `infeasible_value` in place of the objective of baseline samples that violate an output constraint
 * (bo_scalar_prepare itself uses -inf).  *n_all_infeasible = number of MC samples without any feasible baseline point:
 * when it is > 0 BoTorch replaces -inf by a pessimistic lower bound of the objective, which the host computes and hands
 * in through a second call.  Must follow bo_scalar_prepare directly (it reads that call's baseline samples). */
int bo_scalar_baseline_best(bo_state* st, double infeasible_value, int32_t* n_all_infeasible, void* stream);

/* Named options of the prepared acquisition function: "ozaki" (run the posterior GEMM as an error-free INT8 digit-plane
 * product on the tcgen05 tensor cores instead of FP64 DMMA: 0 = never, 1 = automatically for large problems (default, or
 * EVEREST_OZAKI), 2 = whenever the shape allows, unguarded (tests); in mode 1 every INT8 call is followed by a per-row guard
 * -- digit-plane error estimate 2 sqrt(G_ii) eps against tol x the posterior variance of the row -- and the q-batches it
 * flags are recomputed by the FP64 kernel; a call that flags more than a quarter of its q-batches sends the state to the
 * FP64 kernel until the next prepare), "ozaki_guard_tol" (default 1e-10), "ozaki_guard_kappa" (standard deviations of the
 * error estimate, default 8), "ozaki_tile" (kernel variant of that product: 0 = default, 64 = one-pass 128x64 tiles,
 * 128 = two-pass 128x128 tiles, 256 = two-pass on CTA pairs), "log_hvi" (0 / 1: qLogEHVI / qLogNEHVI value instead of
 * qEHVI / qNEHVI -- MoboStrategy's default, mobo.py:72-90), "tau_relu" (default 1e-6), "tau_max" (default 1e-2).  An empty
 * t-batch (b = 0) is a no-op. */
int bo_acqf_set_option(bo_state* st, const char* name, double value);

/* AcquisitionFunction.forward(X[b, q, d]) -> [b]  (called from calc_acquisition botorch.py:223,
 * optimize_acqf's raw-sample screen, optimize_acqf_discrete botorch.py:461).  zq_dev [S, q, M] are the
 * base samples of the q new points.  info_dev[b] (may be NULL): 0, or 1 if the conditional
 * q x q root needed more jitter than 1e-3 (BoTorch would fall back to joint sampling). */
int bo_acqf_forward(bo_state* st, const double* X_dev, int32_t b, int32_t q, const double* zq_dev,
                    double* out_dev, int32_t* info_dev, void* stream);

/* forward + analytic gradient: what BoTorch obtains from autograd inside gen_candidates_scipy when
 * BotorchStrategy._optimize_acqf_continuous refines the restarts (botorch.py:384-405; "differentiable w.r.t. X",
 * SURVEY.md 8b L1).  dX_dev [b, q, d] receives d out[i] / d X[i] (each value depends on its own q-batch only, so this
 * is also the gradient of sum(out) that gen_candidates_scipy minimises).  Columns read only by Hamming / Tanimoto
 * leaves (one-hot / fingerprint columns, fixed during the optimisation) get a zero gradient. */
int bo_acqf_forward_backward(bo_state* st, const double* X_dev, int32_t b, int32_t q, const double* zq_dev,
                             double* out_dev, double* dX_dev, int32_t* info_dev, void* stream);

/* [UPSTREAM] sample_cached_cholesky's fallback: BoTorch catches the NotPSDError / NanError of the q x q conditional root and
 * samples the joint posterior over (X_baseline, X) instead.  Call after bo_acqf_forward with the same arguments: q-batches
 * whose info is non-zero are re-scored from the Cholesky root of the joint (n_b + q) x (n_b + q) covariance (own jitter
 * ladder on the whole diagonal), same base samples, same cached cells.  info_dev[i] stays 1 for a re-scored q-batch (BoTorch
 * warns there) and becomes 2 when the joint factorisation fails too (BoTorch: NotPSDError; the value stays NaN).  One
 * synchronisation (info is read back); n_resampled (HOST, may be NULL) = number of re-scored q-batches. */
int bo_acqf_resample_flagged(bo_state* st, const double* X_dev, int32_t b, int32_t q, const double* zq_dev, double* out_dev,
                             int32_t* info_dev, int32_t* n_resampled, void* stream);
/* bo_acqf_forward runs this fallback by itself (option "joint_fallback", default 1): a 4-byte counter of exhausted jitter
 * ladders leaves the device behind the conditional-root kernels and is read while the MC kernels still run, so an
 * unflagged call pays nothing; q-batches re-scored by the last forward call: */
int32_t bo_acqf_last_resampled(bo_state* st);

/* On-device multi-start refinement: replaces the host loop of [UPSTREAM] botorch.generation.gen_candidates_scipy that
 * BotorchStrategy._optimize_acqf_continuous drives through optimize_acqf (botorch.py:384-405) for the box-constrained case
 * (bounds + fixed features; linear / nonlinear constraints stay on the host SLSQP path).  X_dev [r, q_tot, d] holds the r
 * restarts on entry and the refined restarts on return; the leading q_free points of every q-batch are optimised, the
 * trailing q_tot - q_free rows (pending points of the concatenating acquisition functions) are constants.  lb / ub [d] are
 * HOST arrays (lb[j] == ub[j] fixes column j).  Each restart runs its own projected L-BFGS (`history` curvature pairs,
 * Armijo line search on the projected arc) on f = -acquisition value with the analytic gradient of
 * bo_acqf_forward_backward; stopping rules are L-BFGS-B's (projected-gradient inf-norm <= pgtol, relative decrease
 * <= ftol, maxiter iterations).  out_dev [r] receives the acquisition values at the returned points.  stats (may be
 * NULL, HOST, 4 ints): evaluations of the acquisition function launched, max iterations over the restarts, restarts
 * converged by pgtol / ftol, restarts stopped by maxiter or the evaluation budget.  Nothing crosses PCIe inside the loop
 * but one 4-byte progress counter every few evaluations. */
int bo_acqf_optimize(bo_state* st, double* X_dev, int32_t r, int32_t q_tot, int32_t q_free, const double* lb,
                     const double* ub, const double* zq_dev, int32_t maxiter, int32_t history, double pgtol, double ftol,
                     double* out_dev, int32_t* stats, void* stream);

/* Same call with HOST buffers: pinned staging, H2D of X, the launches, D2H of the values.  Columns read only by Tanimoto
 * leaves (0/1 fingerprints: molfeatures.py:31-48 hands them to BoTorch as float64) cross PCIe as BITS: the staging threads
 * pack them while they copy (a config-5 candidate shrinks from 16.5 KB to 352 bytes on the wire) and a kernel restores the
 * float64 row on the device; a fingerprint column holding anything but 0 / 1 is refused (BO_ERR_INVALID).
 * Large float64 inputs (> 8 MiB) of single-leaf models are copied in pieces by a team of staging threads while the point
 * preparation and K(X*,X) kernels already run on the pieces that have landed; every later kernel runs once over the whole
 * batch.  The values equal those of bo_acqf_forward on the same rows bit for bit.  X_host may be pageable memory. */
int bo_acqf_forward_host(bo_state* st, const double* X_host, int32_t b, int32_t q, const double* zq_dev,
                         double* out_host, void* stream);

/* The packed wire format for callers that keep their candidate set packed (the discrete choice set of
 * optimize_acqf_discrete, botorch.py:425-467, is scored again after every tell):
 *   bo_pack_layout     which columns travel as doubles (dense_cols [n_dense]) and which as bits (bit k of a row = column
 *                      bit_cols[k]; 64 bits per uint64 word, bit k in word k / 64 at position k % 64); arrays may be NULL to
 *                      query the sizes
 *   bo_pack_rows_host  X_host [rows, d] float64 -> dense_out [rows, n_dense], bits_out [rows, ceil(n_bits / 64)]
 *   bo_acqf_forward_host_packed   forward(X[b, q, d]) from the packed buffers ([b * q] rows), values to out_host [b] */
int bo_pack_layout(bo_state* st, int32_t* n_dense, int32_t* n_bits, int32_t* dense_cols, int32_t* bit_cols);
int bo_pack_rows_host(bo_state* st, const double* X_host, int64_t rows, double* dense_out, uint64_t* bits_out);
int bo_acqf_forward_host_packed(bo_state* st, const double* dense_host, const uint64_t* bits_host, int32_t b, int32_t q,
                                const double* zq_dev, double* out_host, void* stream);

/* Stand-alone multi-objective utilities (maximisation frame), the result metrics of the path:
 * botorch is_non_dominated as used by get_pareto_front (utils/multiobjective.py:58-84) and Hypervolume.compute as
 * used by compute_hypervolume (:87-130).  Y_dev [n, m]; mask_dev [n] int32 (1 = non-dominated; with deduplicate
 * only the first of identical rows is kept).  The hypervolume is exact: box(ref, ideal) minus the volume of the
 * non-dominated box decomposition clipped to the ideal point. */
int bo_pareto_mask(const double* Y_dev, int32_t n, int32_t m, int32_t deduplicate, int32_t* mask_dev, void* stream);
int bo_hypervolume(const double* Y_dev, int32_t n, int32_t m, const double* ref_point, double* hv_out, void* stream);

/* Exact marginal log likelihood of output m for the CURRENT hyper-parameters of the (factorised) state, and its
 * gradient: the objective of fit_gpytorch_mll in SingleTaskGPSurrogate._fit (surrogates/single_task_gp.py:39-71) without
 * the prior terms and the 1/N scaling (both are host-side one-liners).  Derivatives are with respect to the natural
 * parameters: noise variance, constant mean, every lengthscale (ARD dims of the continuous leaves and groups of the
 * Hamming leaves, in leaf order) and every term coefficient of the flattened kernel (products of outputscales).  All
 * outputs are HOST pointers; any gradient pointer may be NULL. */
int bo_mll_forward_backward(bo_state* st, int32_t m, double* mll_out, double* d_noise, double* d_mean_const,
                            double* d_lengthscale, int32_t n_lengthscale, double* d_coef, int32_t n_coef, void* stream);

/* Base samples on the device ([UPSTREAM] SobolQMCNormalSampler = torch SobolEngine(scramble=True, seed) + inverse normal
 * CDF, reached from every _get_acqfs and from prune_inferior_points).  bo_sobol_scramble applies the per-dimension
 * unit-lower-triangular GF(2) matrices (ltm_rows_dev [dim, 30]: bit 29-k of row p = L[p][k]) to the direction numbers
 * sobolstate_dev [dim, 30] in place; bo_sobol_normal writes z[S, n_points, M] for Sobol dimension m * n_points + i.
 * Both reproduce torch's integer pipeline bit for bit (the random bits come from torch's CPU generator). */
int bo_sobol_scramble(int64_t* sobolstate_dev, const int64_t* ltm_rows_dev, int32_t dim, void* stream);
int bo_sobol_normal(const int64_t* sobolstate_dev, const int64_t* shift_dev, int32_t n_points, int32_t M, int32_t S,
                    double* out_dev, void* stream);
/* Uniform points out[S, dim] = SobolEngine(dim, scramble=True, seed).draw(S, dtype=float64): the raw samples of
 * optimize_acqf ([UPSTREAM] draw_sobol_samples, reached from botorch.py:384-405), bit-identical to torch's engine. */
int bo_sobol_uniform(const int64_t* sobolstate_dev, const int64_t* shift_dev, int32_t dim, int32_t S, double* out_dev,
                     void* stream);

/* Introspection for tests: copies internal device buffers to the given device pointers. */
int bo_debug_get(bo_state* st, const char* name, int32_t m, double* out_dev, int64_t capacity, int64_t* n_written,
                 void* stream);
/* Kernels launched by this library since the handle was created (bench.py's gpu_launches). */
int64_t bo_launch_count(const bo_state* st);
/* Average device time (ms) of the named kernel family in the last forward (CUDA events on the launch stream). */
int bo_last_timing(const bo_state* st, const char* name, double* ms);
int bo_set_timing(bo_state* st, int32_t enabled);

#ifdef __cplusplus
}
#endif
#endif

// Argument blocks and launchers of acqf.cu.
#pragma once
#include "common.cuh"

struct CondRootArgs {
  ModelD md;
  PrepD prep_q, prep_b;
  int b, q, nb, M, m;
  const double* Gqq;     // [b, q, q]
  const double* W;       // [b*q, ldw]
  int ldw;
  const double* mu_raw;  // [b*q]
  const double* Lb;      // [nb, ldlb] cached baseline root of output m
  const double* LbInv;   // [nb, ldlb] its inverse (lower triangular): bl = Sqb LbInv^T
  const double* LbInvT;  // [nb, ldlb] transpose of the inverse (coalesced access when it does not fit in smem)
  int ldlb;
  int linv_in_smem;      // set by the launcher
  int stage_cols;        // set by the launcher: dpad of the single continuous leaf (0 = generic kernel tree)
  int wpb;               // set by the launcher: warps per q-batch (1, or 4 = the whole CTA for large baselines)
  double* root;          // [b, M, q, nb+q]
  double* BL;            // [b*q, ldbl] copy of bl for output m, zero padded (operand of the sample GEMM) or NULL
  int ldbl;
  double* mu;            // [b*q, M]
  int* info;             // [b, M] or NULL
  double* jitter;        // [b, M] or NULL
};

struct McArgs {
  int b, q, nb, M, S;
  ObjD od;
  const double* root;
  const double* mu;
  const double* zbT;      // [(e*M+m)*S + s]
  const double* zqT;      // [(k*M+m)*S + s]
  const double* Fp;       // [M][b*q][S] baseline part bl z_b of every sample (DMMA GEMM) or NULL
  size_t fp_stride;       // elements between outputs in Fp
  double* partial;        // [ceil(S/64)][b] per-sample-group sums (tiled kernel)
  const double* cell_lo;  // [(c*Mo+o)*S + s]  (or [(c*Mo+o)] when cells_shared)
  const double* cell_up;
  const int* ncells;      // [S] (or [1])
  int cells_shared;
  double best_f;
  int variant;            // scalar family: bo_scalar_acqf; HVI family: 0 = plain, 1 = log-space (qLogEHVI / qLogNEHVI)
  double vparam;          // qUCB: beta, qPI: tau
  const double* best_f_s; // [S] per-MC-sample incumbent of the noisy variants (qNEI / qLogNEI) or NULL
  double tau_relu, tau_max;  // smoothing temperatures of the log-space variants
  double* out;            // [b]
  const int* info_in;     // [b, M] flags from cond_root
  int* info_out;          // [b] or NULL
};

int launch_transpose_base_samples(const double* z, int S, int n, int M, double* zT, double* zM, int ldn, cudaStream_t st,
                                  LaunchCounter* lc);
int launch_scale_matrix(double* A, int ld, int rows, int cols, double s, cudaStream_t st, LaunchCounter* lc);
int launch_copy_scale(const double* src, int lds, double* dst, int ldd, int rows, int cols, double scale, cudaStream_t st,
                      LaunchCounter* lc);
int launch_add_diag(double* A, int ld, int n, double v, cudaStream_t st, LaunchCounter* lc);
int launch_finish_mean(const double* raw, int n, double mean_const, double y_std, double y_mean, double* out, int ldo,
                       int m, cudaStream_t st, LaunchCounter* lc);
int launch_finish_var(const ModelD& md, PrepD prep, const double* g, int n, int add_noise, double* out, int ldo, int m,
                      cudaStream_t st, LaunchCounter* lc);
int launch_cond_root(const CondRootArgs& a, cudaStream_t st, LaunchCounter* lc);
int launch_baseline_objective(const double* F, int ldf, int S, int n, int M, const double* mean, const ObjD& od,
                              double* obj, unsigned char* feas, double* samples, cudaStream_t st, LaunchCounter* lc);
int launch_front(const double* obj, const unsigned char* feas, int S, int n, int Mo, const double* ref_dev, int dedup,
                 unsigned char* front, int* counts, cudaStream_t st, LaunchCounter* lc);
int launch_partition2d(const double* obj, const unsigned char* front, int n, int S, int cap, const double* ref_dev,
                       double* lo, double* up, int* ncells, int* front_idx, cudaStream_t st, LaunchCounter* lc);
int launch_partition_nd(const double* obj, const unsigned char* front, int n, int S, int Mo, int cap,
                        const double* ref_dev, double* work, double* lo, double* up, int* ncells, int* overflow,
                        cudaStream_t st, LaunchCounter* lc);
int launch_add_inplace(double* dst, const double* src, size_t n, cudaStream_t st, LaunchCounter* lc);
int launch_count_nonzero(const int* v, int n, int* count, cudaStream_t st, LaunchCounter* lc);
int launch_joint_rows_to_root(const double* rootj, int ldn, const double* meanj, int nb, int q, int M, int m, int slot,
                              double* root, double* mu, cudaStream_t st, LaunchCounter* lc);
size_t partition_binary_work_stride(int n, int Mo);
int launch_partition_binary(const double* obj, const unsigned char* front, int n, int S, int Mo, int cap, const double* ref_dev,
                            double alpha, double* work, double* lo, double* up, int* ncells, int* overflow, cudaStream_t st,
                            LaunchCounter* lc);
int launch_mc_hvi(const McArgs& a, int max_cells, double* obj_ws, cudaStream_t st, LaunchCounter* lc);
size_t mc_hvi_obj_ws_bytes(const McArgs& a, int max_cells);
int mc_hvi_partial_groups(int S);   // sample groups of the MC/HVI partial sums (rows of McArgs::partial)
int launch_mc_scalar(const McArgs& a, cudaStream_t st, LaunchCounter* lc);
// scalarised objective of baseline samples F[m][s][ldf] + mean: per-sample best value and / or arg-max counts per point
// (infeasible samples, c(y) > 0, take `infeasible_value`; n_all_infeasible counts the MC samples without a feasible point)
int launch_baseline_best(const double* F, int ldf, int S, int n, int M, const double* mean, const ObjD& od,
                         double infeasible_value, double* best_f_s, int* counts, int* n_all_infeasible, cudaStream_t st,
                         LaunchCounter* lc);
int launch_mc_loghvi(const McArgs& a, cudaStream_t st, LaunchCounter* lc);
int launch_front_to_mask(const unsigned char* front, int n, int* mask, cudaStream_t st, LaunchCounter* lc);
int launch_hypervolume_from_cells(const double* obj, const unsigned char* front, int n, int Mo, const double* ref_dev,
                                  const double* lo, const double* up, const int* ncells, double* hv_dev, cudaStream_t st,
                                  LaunchCounter* lc);

// ---- grad.cu: analytic adjoint d acqf / d X (the backward of forward(X[b, q, d])) ----------
// MC value + d value / d f for every MC sample: dF[m * df_stride + (batch * q + j) * S + s]
int launch_mc_reduce_partials(const double* partial, int groups, int b, int S, double* out, const int* info_in, int M,
                              int* info_out, cudaStream_t st, LaunchCounter* lc);
int launch_mc_hvi_grad(const McArgs& a, int max_cells, double* dF, size_t df_stride, cudaStream_t st, LaunchCounter* lc);
int launch_mc_scalar_grad(const McArgs& a, double* dF, size_t df_stride, cudaStream_t st, LaunchCounter* lc);
// vals_ws: [b][S] doubles (per-sample values when the MC samples of a q-batch are split over several CTAs) or NULL
int launch_mc_loghvi_grad(const McArgs& a, double* dF, size_t df_stride, double* vals_ws, cudaStream_t st, LaunchCounter* lc);
// dF -> d root [b, M, q, nb+q] (= [d bl | d br]) and d mu [b*q, M]
int launch_grad_reduce(const double* dF, size_t df_stride, const double* zbT, const double* zqT, int S, int nb, int q,
                       int M, int rows, double* droot, double* dmu, cudaStream_t st, LaunchCounter* lc);

struct CondRootBwdArgs {
  int b, q, nb, M, m;
  const double* root;   // [b, M, q, nb+q] forward roots
  const double* droot;  // [b, M, q, nb+q]
  const double* dmu;    // [b*q, M]
  const double* LbInv;  // [nb, ldlb]
  int ldlb;
  double y_std;
  double* EG;           // [b, q, q]   coefficient of U_i in dKx_j  (= -2 s^2 Cbar)
  double* EW;           // [b*q, ldw]  d value / d W
  int ldw;
  double* Emu;          // [b*q]       d value / d mu_raw
};
int launch_cond_root_bwd(const CondRootBwdArgs& a, cudaStream_t st, LaunchCounter* lc);

struct KernelGradArgs {
  ModelD md;
  PrepD prep_q, prep_b;
  int rows, q, nb, N, ldk, d;
  const double* alpha;  // [ldk]
  const double* Aext;   // [nb, ldk] rows K_bX (K + s2 I)^-1
  const double* U;      // [rows, ldk] K*X (K + s2 I)^-1
  const double* EG;     // [b, q, q]
  const double* EW;     // [rows, ldw]
  int ldw;
  const double* Emu;    // [rows]
  const double* Kx;     // [rows, ldk] forward K(X*, X) (fast path of single-leaf RBF models: dK/dstat = -K/2)
  double* dX;           // [rows, d]
  int accumulate;       // 0: overwrite dX, 1: add (outputs after the first)
};
int launch_kernel_grad(const KernelGradArgs& a, cudaStream_t st, LaunchCounter* lc);

// ---- mll.cu: exact marginal log likelihood and its hyper-parameter gradient ----------------------
int launch_mll_grad(const ModelD& md, const PrepD& train, int N, int ldk, const double* alpha, const double* Kinv,
                    const int* ls_offset, int n_ls_total, double* part, double* out_params, cudaStream_t st, LaunchCounter* lc);
int launch_mll_scalars(const double* resid, const double* alpha, const double* Lmat, const double* Kinv, int N, int ldk,
                       double* out5, cudaStream_t st, LaunchCounter* lc);

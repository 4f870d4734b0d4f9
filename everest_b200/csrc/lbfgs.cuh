// Batched box-constrained L-BFGS on the device (lbfgs.cu), driven by bo_acqf_optimize (capi.cu).
#pragma once
#include <cuda_runtime.h>

#define LB_MAX_HIST 16

struct LaunchCounter;

struct LbScalars {
  double f, f_best, t, gtd, pg;
  int hist_n, head, n_iter, n_eval, n_ls, steepest, status, pad;
};

struct LbArgs {
  int q_tot, q_free, d, hist, maxiter;
  double pgtol, ftol;
  const double *lb, *ub;    // [d] device
  double* X;                // [r, q_tot, d] evaluation buffer: in = starts, out = refined restarts
  const double* dX;         // [r, q_tot, d] gradient of the acquisition value at X
  const double* vals;       // [r] acquisition values at X
  double *x, *g, *dir, *S, *Y, *rho;
  LbScalars* sc;
  int* n_running;
};

size_t lbfgs_ws_bytes(int r, int n, int hist);
void lbfgs_carve(LbArgs& a, void* ws, int r, int n, int hist);
int launch_lbfgs_step(const LbArgs& a, int r, bool first, cudaStream_t s, LaunchCounter* lc);
int launch_lbfgs_finish(const LbArgs& a, int r, cudaStream_t s, LaunchCounter* lc);

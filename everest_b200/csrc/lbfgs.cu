// On-device multi-start refinement of the restarts: a batched, box-constrained limited-memory BFGS that
// replaces the host loop of [UPSTREAM] botorch.generation.gen_candidates_scipy (scipy L-BFGS-B driven through
// strategies/predictives/botorch.py:384-405) for the case BoFire hands it most often -- bounds and fixed features only.
//
// One CTA per restart; the restarts are independent problems (the joint objective BoTorch gives scipy is their SUM, which
// is separable), so each keeps its own curvature memory, step length and convergence state.  Per evaluation of the
// acquisition function and its analytic gradient (bo_acqf_forward_backward's chain, launched by the host WITHOUT waiting
// for it) one `lbfgs_step_kernel` advances every restart's state machine:
//
//   trial accepted (Armijo on the projected arc)  ->  curvature pair (s, y) pushed, convergence tests of L-BFGS-B
//        (projected-gradient inf-norm <= pgtol, relative decrease <= ftol, maxiter), next direction from the two-loop
//        recursion on the free variables, next trial = P(x + d)
//   trial rejected                                 ->  step shortened by safeguarded quadratic interpolation, next trial
//
// Nothing crosses PCIe inside the loop; the host reads one counter (restarts still running) every few evaluations, one
// batch of launches behind the GPU.  The curvature memory ([hist][n] doubles per restart) lives in HBM: at 8 restarts x
// 120 variables x 2 x 10 vectors it is 150 KB, L2-resident.
#include <math.h>

#include "acqf.cuh"
#include "common.cuh"
#include "lbfgs.cuh"

#define LB_THREADS 128

__device__ __forceinline__ double lb_block_sum(double v, double* red) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  double t = 0.0;
#pragma unroll
  for (int i = 0; i < LB_THREADS / 32; ++i) t += red[i];
  return t;
}

__device__ __forceinline__ double lb_block_max(double v, double* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  double t = 0.0;
#pragma unroll
  for (int i = 0; i < LB_THREADS / 32; ++i) t = fmax(t, red[i]);
  return t;
}

__device__ __forceinline__ double lb_clip(double v, double lo, double hi) { return fmin(fmax(v, lo), hi); }

// variable k of a restart = point k / d, column k % d of the leading q_free points of its q-batch
struct LbView {
  double *x, *g, *dir, *S, *Y, *rho;
  double* Xt;          // the restart's rows of the evaluation buffer [q_tot, d]
  const double* dXt;   // gradient of the acquisition value at the trial
};

// d = -H g on the free variables (two-loop recursion), d = 0 on the active ones; returns g.d
__device__ double lb_direction(const LbArgs& a, const LbView& v, int n, int d, int hist_n, int head, double* alpha, double* red) {
  const int tid = threadIdx.x;
  // free set: not pinned at a bound with the gradient pushing outwards
  for (int k = tid; k < n; k += LB_THREADS) {
    const double lo = a.lb[k % d], hi = a.ub[k % d], xk = v.x[k], gk = v.g[k];
    const bool active = (hi <= lo) || (xk <= lo && gk > 0.0) || (xk >= hi && gk < 0.0);
    v.dir[k] = active ? 0.0 : gk;      // q <- g_F
  }
  __syncthreads();
  for (int j = 0; j < hist_n; ++j) {   // newest to oldest
    const int slot = (head - 1 - j + a.hist) % a.hist;
    const double* Sj = v.S + (size_t)slot * n;
    const double* Yj = v.Y + (size_t)slot * n;
    double part = 0.0;
    for (int k = tid; k < n; k += LB_THREADS) part += Sj[k] * v.dir[k];
    const double al = v.rho[slot] * lb_block_sum(part, red);
    if (tid == 0) alpha[j] = al;
    for (int k = tid; k < n; k += LB_THREADS) v.dir[k] -= al * Yj[k];
    __syncthreads();
  }
  if (hist_n > 0) {                    // initial Hessian scaling gamma = s.y / y.y of the newest pair
    const int slot = (head - 1 + a.hist) % a.hist;
    const double* Yj = v.Y + (size_t)slot * n;
    double part = 0.0;
    for (int k = tid; k < n; k += LB_THREADS) part += Yj[k] * Yj[k];
    const double yy = lb_block_sum(part, red);
    const double gamma = (yy > 0.0) ? 1.0 / (v.rho[slot] * yy) : 1.0;
    for (int k = tid; k < n; k += LB_THREADS) v.dir[k] *= gamma;
    __syncthreads();
  }
  for (int j = hist_n - 1; j >= 0; --j) {   // oldest to newest
    const int slot = (head - 1 - j + a.hist) % a.hist;
    const double* Sj = v.S + (size_t)slot * n;
    const double* Yj = v.Y + (size_t)slot * n;
    double part = 0.0;
    for (int k = tid; k < n; k += LB_THREADS) part += Yj[k] * v.dir[k];
    const double beta = v.rho[slot] * lb_block_sum(part, red);
    const double al = alpha[j];
    for (int k = tid; k < n; k += LB_THREADS) v.dir[k] += (al - beta) * Sj[k];
    __syncthreads();
  }
  double part = 0.0;
  for (int k = tid; k < n; k += LB_THREADS) {
    const double lo = a.lb[k % d], hi = a.ub[k % d], xk = v.x[k], gk = v.g[k];
    const bool active = (hi <= lo) || (xk <= lo && gk > 0.0) || (xk >= hi && gk < 0.0);
    const double dk = active ? 0.0 : -v.dir[k];
    v.dir[k] = dk;
    part += dk * gk;
  }
  return lb_block_sum(part, red);
}

// x_t = P(x + t d) into the evaluation buffer; returns max |x_t - x|
__device__ double lb_write_trial(const LbArgs& a, const LbView& v, int n, int d, double t, double* red) {
  double mv = 0.0;
  for (int k = threadIdx.x; k < n; k += LB_THREADS) {
    const double xt = lb_clip(v.x[k] + t * v.dir[k], a.lb[k % d], a.ub[k % d]);
    v.Xt[k] = xt;
    mv = fmax(mv, fabs(xt - v.x[k]));
  }
  return lb_block_max(mv, red);
}

__global__ void __launch_bounds__(LB_THREADS) lbfgs_step_kernel(LbArgs a, int first) {
  __shared__ double red[LB_THREADS / 32];
  __shared__ double alpha[LB_MAX_HIST];
  const int i = blockIdx.x, tid = threadIdx.x;
  const int d = a.d, n = a.q_free * a.d;
  LbView v;
  v.x = a.x + (size_t)i * n; v.g = a.g + (size_t)i * n; v.dir = a.dir + (size_t)i * n;
  v.S = a.S + (size_t)i * a.hist * n; v.Y = a.Y + (size_t)i * a.hist * n; v.rho = a.rho + (size_t)i * a.hist;
  v.Xt = a.X + (size_t)i * a.q_tot * d; v.dXt = a.dX + (size_t)i * a.q_tot * d;
  LbScalars& sc = a.sc[i];
  if (!first && sc.status != 0) return;                    // finished: the buffer already holds its x
  const double ft = -a.vals[i];                            // minimise f = -acquisition value
  const bool finite_t = isfinite(ft);

  if (first) {
    double part = 0.0;
    for (int k = tid; k < n; k += LB_THREADS) {
      const double xk = lb_clip(v.Xt[k], a.lb[k % d], a.ub[k % d]);
      const double gk = finite_t ? -v.dXt[k] : 0.0;
      v.x[k] = xk; v.g[k] = gk;
      part += gk * gk;
    }
    __syncthreads();
    double pg = 0.0;
    for (int k = tid; k < n; k += LB_THREADS)
      pg = fmax(pg, fabs(lb_clip(v.x[k] - v.g[k], a.lb[k % d], a.ub[k % d]) - v.x[k]));
    pg = lb_block_max(pg, red);
    const double gtd = lb_direction(a, v, n, d, 0, 0, alpha, red);
    double dn = 0.0;
    for (int k = tid; k < n; k += LB_THREADS) dn += v.dir[k] * v.dir[k];
    dn = sqrt(lb_block_sum(dn, red));
    const double t = (dn > 0.0) ? fmin(1.0, 1.0 / dn) : 1.0;   // L-BFGS-B's first step: min(1, 1 / ||d||)
    int status = 0;
    if (!finite_t) status = 5;                             // the start itself could not be scored
    else if (pg <= a.pgtol) status = 1;
    const double moved = lb_write_trial(a, v, n, d, status ? 0.0 : t, red);
    if (status == 0 && moved == 0.0) status = 3;
    if (tid == 0) {
      sc.f = ft; sc.f_best = ft; sc.t = t; sc.gtd = gtd; sc.hist_n = 0; sc.head = 0; sc.n_iter = 0; sc.n_eval = 1; sc.n_ls = 0;
      sc.steepest = 1; sc.status = status; sc.pg = pg;
      if (status) atomicSub(a.n_running, 1);
    }
    return;
  }

  // ---- Armijo test on the projected arc: f(x_t) <= f(x) + c1 g.(x_t - x) ----
  double part = 0.0;
  for (int k = tid; k < n; k += LB_THREADS) part += v.g[k] * (v.Xt[k] - v.x[k]);
  const double gts = lb_block_sum(part, red);
  const double f = sc.f;
  const bool accept = finite_t && ft <= f + 1e-4 * gts && gts <= 0.0;
  if (tid == 0) sc.n_eval += 1;
  if (!accept) {
    // shorten: minimiser of the quadratic through f, f', f(t), kept inside [0.1 t, 0.5 t]
    const double t = sc.t;
    double tn = 0.5 * t;
    if (finite_t && gts < 0.0) {
      const double denom = 2.0 * (ft - f - gts);
      if (denom > 0.0) tn = lb_clip(-gts * t / denom, 0.1 * t, 0.5 * t);
    }
    const int n_ls = sc.n_ls + 1;
    int status = 0;
    int reset = 0;
    if (n_ls >= 20) {
      // L-BFGS-B: after 20 failed trials drop the curvature memory and restart from steepest descent; a second failure
      // in a row ends the restart ("abnormal termination in lnsrch")
      if (sc.steepest) status = 3;
      else reset = 1;
    }
    double moved;
    if (reset) {
      const double gtd = lb_direction(a, v, n, d, 0, 0, alpha, red);
      double dn = 0.0;
      for (int k = tid; k < n; k += LB_THREADS) dn += v.dir[k] * v.dir[k];
      dn = sqrt(lb_block_sum(dn, red));
      tn = (dn > 0.0) ? fmin(1.0, 1.0 / dn) : 1.0;
      moved = lb_write_trial(a, v, n, d, tn, red);
      if (tid == 0) { sc.gtd = gtd; sc.hist_n = 0; sc.head = 0; sc.steepest = 1; sc.n_ls = 0; }
    } else {
      moved = lb_write_trial(a, v, n, d, status ? 0.0 : tn, red);
      if (tid == 0) sc.n_ls = n_ls;
    }
    if (status == 0 && moved == 0.0) status = 3;           // the arc collapsed onto x: nothing left to try
    if (status) lb_write_trial(a, v, n, d, 0.0, red);
    if (tid == 0) {
      sc.t = tn;
      if (status) { sc.status = status; atomicSub(a.n_running, 1); }
    }
    return;
  }

  // ---- accepted: curvature pair, convergence tests, next direction ----
  const int head = sc.head, hist_n = sc.hist_n;
  double* Sj = v.S + (size_t)head * n;
  double* Yj = v.Y + (size_t)head * n;
  double sy = 0.0, yy = 0.0;
  for (int k = tid; k < n; k += LB_THREADS) {
    const double s = v.Xt[k] - v.x[k], gn = -v.dXt[k], y = gn - v.g[k];
    Sj[k] = s; Yj[k] = y;
    sy += s * y; yy += y * y;
  }
  sy = lb_block_sum(sy, red);
  yy = lb_block_sum(yy, red);
  const bool push = sy > 2.220446049250313e-16 * yy && yy > 0.0;   // L-BFGS-B skips the update otherwise
  for (int k = tid; k < n; k += LB_THREADS) { v.x[k] = v.Xt[k]; v.g[k] = -v.dXt[k]; }
  __syncthreads();
  int new_head = head, new_hist = hist_n;
  if (push) {
    if (tid == 0) v.rho[head] = 1.0 / sy;
    new_head = (head + 1) % a.hist;
    new_hist = min(hist_n + 1, a.hist);
  }
  __syncthreads();
  double pg = 0.0;
  for (int k = tid; k < n; k += LB_THREADS)
    pg = fmax(pg, fabs(lb_clip(v.x[k] - v.g[k], a.lb[k % d], a.ub[k % d]) - v.x[k]));
  pg = lb_block_max(pg, red);
  const int n_iter = sc.n_iter + 1;
  int status = 0;
  if (pg <= a.pgtol) status = 1;
  else if ((f - ft) <= a.ftol * fmax(fmax(fabs(f), fabs(ft)), 1.0)) status = 2;
  else if (n_iter >= a.maxiter) status = 4;
  double gtd = 0.0, t = 1.0;
  int steepest = 0;
  if (!status) {
    gtd = lb_direction(a, v, n, d, new_hist, new_head, alpha, red);
    if (!(gtd < 0.0)) {                                    // not a descent direction: forget the memory
      new_hist = 0; new_head = 0;
      gtd = lb_direction(a, v, n, d, 0, 0, alpha, red);
    }
    if (new_hist == 0) {
      double dn = 0.0;
      for (int k = tid; k < n; k += LB_THREADS) dn += v.dir[k] * v.dir[k];
      dn = sqrt(lb_block_sum(dn, red));
      t = (dn > 0.0) ? fmin(1.0, 1.0 / dn) : 1.0;
      steepest = 1;
    }
  }
  const double moved = lb_write_trial(a, v, n, d, status ? 0.0 : t, red);
  if (!status && moved == 0.0) status = 3;
  if (tid == 0) {
    sc.f = ft; sc.f_best = fmin(sc.f_best, ft); sc.t = t; sc.gtd = gtd; sc.hist_n = new_hist; sc.head = new_head;
    sc.n_iter = n_iter; sc.n_ls = 0; sc.steepest = steepest; sc.pg = pg;
    if (status) { sc.status = status; atomicSub(a.n_running, 1); }
  }
}

// restarts that ran out of evaluations keep their last ACCEPTED point (the buffer may hold a rejected trial)
__global__ void __launch_bounds__(LB_THREADS) lbfgs_finish_kernel(LbArgs a) {
  const int i = blockIdx.x;
  const int n = a.q_free * a.d;
  const double* x = a.x + (size_t)i * n;
  double* Xt = a.X + (size_t)i * a.q_tot * a.d;
  for (int k = threadIdx.x; k < n; k += LB_THREADS) Xt[k] = x[k];
  if (threadIdx.x == 0 && a.sc[i].status == 0) a.sc[i].status = 6;
}

size_t lbfgs_ws_bytes(int r, int n, int hist) {
  return ((size_t)r * n * 3 + (size_t)r * hist * n * 2 + (size_t)r * hist) * sizeof(double) + (size_t)r * sizeof(LbScalars) + 64;
}

void lbfgs_carve(LbArgs& a, void* ws, int r, int n, int hist) {
  double* p = reinterpret_cast<double*>(ws);
  a.x = p; p += (size_t)r * n;
  a.g = p; p += (size_t)r * n;
  a.dir = p; p += (size_t)r * n;
  a.S = p; p += (size_t)r * hist * n;
  a.Y = p; p += (size_t)r * hist * n;
  a.rho = p; p += (size_t)r * hist;
  a.sc = reinterpret_cast<LbScalars*>(p);
  a.n_running = reinterpret_cast<int*>(reinterpret_cast<char*>(a.sc) + (size_t)r * sizeof(LbScalars));
}

int launch_lbfgs_step(const LbArgs& a, int r, bool first, cudaStream_t s, LaunchCounter* lc) {
  lbfgs_step_kernel<<<r, LB_THREADS, 0, s>>>(a, first ? 1 : 0);
  if (lc) lc->n++;
  CUDA_CHECK_RET(cudaGetLastError());
  return BO_OK;
}

int launch_lbfgs_finish(const LbArgs& a, int r, cudaStream_t s, LaunchCounter* lc) {
  lbfgs_finish_kernel<<<r, LB_THREADS, 0, s>>>(a);
  if (lc) lc->n++;
  CUDA_CHECK_RET(cudaGetLastError());
  return BO_OK;
}

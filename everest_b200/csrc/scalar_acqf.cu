// K8: the single-objective MC acquisition functions SoboStrategy builds through get_acquisition_function
// (strategies/predictives/sobo.py:64-89; data models acquisition_functions/acquisition_function.py:21-60):
//   qLogEI / qLogNEI : logmeanexp_S( fatmax_q( log_fatplus(obj - best_f, tau_relu) [+ log feasibility], tau_max ) )
//   qEI / qNEI       : mean_S max_q relu(obj - best_f) [* feasibility]
//   qSR              : mean_S max_q obj
//   qUCB             : mean_S max_q ( mean_S(obj) + sqrt(beta pi / 2) |obj - mean_S(obj)| )
//   qPI              : mean_S max_q sigmoid((obj - best_f) / tau) [* feasibility]
// The noisy variants take best_f per MC sample from the cached baseline samples (best_f_s).  [UPSTREAM] formulas:
// botorch.acquisition.monte_carlo / logei / utils.safe_math, restated in oracle/bo_oracle.py (QScalarOracle).
// One CTA per q-batch, threads over MC samples; the gradient kernel writes d value / d f per MC sample.
#include "acqf.cuh"
#include "common.cuh"
#include "mc_math.cuh"

struct ScalarSmem {
  double *root, *mu, *vals, *omean, *aj, *red;
};

__device__ __forceinline__ ScalarSmem scalar_smem(double* base, int M, int q, int nr, int S) {
  ScalarSmem sm;
  sm.root = base;
  sm.mu = sm.root + (size_t)M * q * nr;
  sm.vals = sm.mu + q * M;
  sm.omean = sm.vals + S;
  sm.aj = sm.omean + BO_MAX_Q;
  sm.red = sm.aj + BO_MAX_Q;  // [40]
  return sm;
}
static size_t scalar_smem_bytes(const McArgs& a) {
  return ((size_t)a.M * a.q * (a.nb + a.q) + a.q * a.M + a.S + 2 * BO_MAX_Q + 40) * sizeof(double);
}

// model-output sample of point j for MC sample s
__device__ __forceinline__ void sample_y(const McArgs& a, const ScalarSmem& sm, int j, int s, double* y) {
  const int q = a.q, nb = a.nb, nr = nb + q, M = a.M, S = a.S;
  for (int m = 0; m < M; ++m) {
    const double* rr = sm.root + ((size_t)m * q + j) * nr;
    double sb = 0.0, sq = 0.0;
    for (int e = 0; e < nb; ++e) sb = fma(rr[e], a.zbT[((size_t)e * M + m) * S + s], sb);
    for (int k = 0; k < q; ++k) sq = fma(rr[nb + k], a.zqT[((size_t)k * M + m) * S + s], sq);
    y[m] = (sm.mu[j * M + m] + sb) + sq;
  }
}

// block-wide sum of q per-thread values -> dst[j] (all threads see the result after the call)
__device__ __forceinline__ void block_sum_vec(const double* local, int q, double* dst, double* red) {
  for (int j = 0; j < q; ++j) {
    double t = block_sum(local[j], red);
    if (threadIdx.x == 0) dst[j] = t;
  }
  __syncthreads();
}

// per-sample utilities u_j and their reduction over the q points.  Returns h_s; fills (optionally) jstar / mx / ps.
__device__ __forceinline__ double scalar_sample_value(const McArgs& a, const ScalarSmem& sm, int s, double* u, int* jstar_out) {
  const int q = a.q, M = a.M;
  const int v = a.variant;
  const bool is_log = (v == BO_ACQF_QLOGEI);
  const double bf = a.best_f_s ? a.best_f_s[s] : a.best_f;
  const double beta_p = sqrt(a.vparam * 3.141592653589793 / 2.0);
  double mx = -INFINITY;
  int jstar = 0;
  for (int j = 0; j < q; ++j) {
    double y[2 * BO_MAX_OBJECTIVES];
    sample_y(a, sm, j, s, y);
    const double o = scalar_objective_apply(a.od, y, M, nullptr);
    double uj;
    if (is_log) {
      uj = log_fatplus_d(o - bf, a.tau_relu);
      if (a.od.n_cons) uj += log_feas_fat(a.od, y, 0.0, nullptr);
    } else if (v == BO_ACQF_QEI) {
      uj = fmax(o - bf, 0.0);
      if (a.od.n_cons) uj *= feas_sigmoid(a.od, y, 0.0, nullptr);
    } else if (v == BO_ACQF_QSR) {
      uj = o;
    } else if (v == BO_ACQF_QUCB) {
      uj = sm.omean[j] + beta_p * fabs(o - sm.omean[j]);
    } else {  // qPI
      uj = 1.0 / (1.0 + exp(-(o - bf) / a.vparam));
      if (a.od.n_cons) uj *= feas_sigmoid(a.od, y, 0.0, nullptr);
    }
    u[j] = uj;
    if (uj > mx) { mx = uj; jstar = j; }
  }
  if (jstar_out) *jstar_out = jstar;
  if (!is_log) return mx;
  double ps = 0.0;
  for (int j = 0; j < q; ++j) ps += pareto2_d((mx - u[j]) / a.tau_max);
  return mx + a.tau_max * log(ps);
}

// qUCB needs the sample mean of every point's objective before anything else
__device__ __forceinline__ void ucb_means(const McArgs& a, const ScalarSmem& sm) {
  if (a.variant != BO_ACQF_QUCB) return;
  double loc[BO_MAX_Q];
  for (int j = 0; j < a.q; ++j) loc[j] = 0.0;
  for (int s = threadIdx.x; s < a.S; s += blockDim.x)
    for (int j = 0; j < a.q; ++j) {
      double y[2 * BO_MAX_OBJECTIVES];
      sample_y(a, sm, j, s, y);
      loc[j] += scalar_objective_apply(a.od, y, a.M, nullptr);
    }
  block_sum_vec(loc, a.q, sm.omean, sm.red);
  if (threadIdx.x < a.q) sm.omean[threadIdx.x] /= (double)a.S;
  __syncthreads();
}

// final reduction of vals[S]: logmeanexp (log variants) or mean.  Returns (all threads) bm and tsum for the log case.
__device__ __forceinline__ void scalar_finish(const McArgs& a, const ScalarSmem& sm, double lmax, double lsum, double* bm_out,
                                              double* tsum_out) {
  const int tid = threadIdx.x, nt = blockDim.x, batch = blockIdx.x, S = a.S;
  double result, bm = 0.0, tsum = 0.0;
  if (a.variant == BO_ACQF_QLOGEI) {
    for (int o = 16; o > 0; o >>= 1) lmax = fmax(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
    __syncthreads();
    if ((tid & 31) == 0) sm.red[tid >> 5] = lmax;
    __syncthreads();
    bm = -INFINITY;
    for (int w = 0; w < (nt >> 5); ++w) bm = fmax(bm, sm.red[w]);
    double se = 0.0;
    for (int s = tid; s < S; s += nt) se += exp(sm.vals[s] - bm);
    double t = block_sum(se, sm.red);
    if (tid == 0) sm.red[32] = t;
    __syncthreads();
    tsum = sm.red[32];
    result = bm + log(tsum) - log((double)S);
  } else {
    double t = block_sum(lsum, sm.red);
    if (tid == 0) sm.red[32] = t;
    __syncthreads();
    result = sm.red[32] / (double)S;
  }
  if (tid == 0) {
    a.out[batch] = result;
    if (a.info_out) {
      int v = 0;
      for (int m = 0; m < a.M; ++m) v |= a.info_in[(size_t)batch * a.M + m];
      a.info_out[batch] = v;
    }
  }
  if (bm_out) *bm_out = bm;
  if (tsum_out) *tsum_out = tsum;
}

__global__ void __launch_bounds__(256)
mc_scalar_kernel(McArgs a) {
  extern __shared__ double msm[];
  const int batch = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  const int q = a.q, nr = a.nb + q, M = a.M, S = a.S;
  ScalarSmem sm = scalar_smem(msm, M, q, nr, S);
  for (int i = tid; i < M * q * nr; i += nt) sm.root[i] = a.root[(size_t)batch * M * q * nr + i];
  for (int i = tid; i < q * M; i += nt) sm.mu[i] = a.mu[(size_t)batch * q * M + i];
  __syncthreads();
  ucb_means(a, sm);
  double lmax = -INFINITY, lsum = 0.0;
  for (int s = tid; s < S; s += nt) {
    double u[BO_MAX_Q];
    const double h = scalar_sample_value(a, sm, s, u, nullptr);
    sm.vals[s] = h;
    lmax = fmax(lmax, h);
    lsum += h;
  }
  scalar_finish(a, sm, lmax, lsum, nullptr, nullptr);
}

int launch_mc_scalar(const McArgs& a, cudaStream_t st, LaunchCounter* lc) {
  if (a.b <= 0) return BO_OK;
  const int nt = 256;
  size_t smem = scalar_smem_bytes(a);
  if (smem > 220 * 1024) { bo_set_error("mc_scalar: shared memory budget exceeded"); return BO_ERR_INVALID; }
  static PerDeviceMax attr_pd; size_t& attr = attr_pd.slot();
  if (smem > 48 * 1024 && smem > attr) {
    CUDA_CHECK_RET(cudaFuncSetAttribute(mc_scalar_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  mc_scalar_kernel<<<a.b, nt, smem, st>>>(a);
  if (lc) lc->n++;
  CUDA_CHECK_RET(cudaGetLastError());
  return BO_OK;
}

// ------------------------------------------------------------------------------------------------
// value + d value / d f  (dF[m * df_stride + (batch * q + j) * S + s])
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
mc_scalar_grad_kernel(McArgs a, double* __restrict__ dF, size_t df_stride) {
  extern __shared__ double msm[];
  const int batch = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  const int q = a.q, nr = a.nb + q, M = a.M, S = a.S;
  const int v = a.variant;
  const bool is_log = (v == BO_ACQF_QLOGEI);
  ScalarSmem sm = scalar_smem(msm, M, q, nr, S);
  for (int i = tid; i < M * q * nr; i += nt) sm.root[i] = a.root[(size_t)batch * M * q * nr + i];
  for (int i = tid; i < q * M; i += nt) sm.mu[i] = a.mu[(size_t)batch * q * M + i];
  __syncthreads();
  ucb_means(a, sm);
  const double beta_p = sqrt(a.vparam * 3.141592653589793 / 2.0);
  // pass 1: per-sample values (+ for qUCB the coefficient of the sample mean: A_j = sum_s [j*(s) = j] (1 - beta' sgn_sj))
  double lmax = -INFINITY, lsum = 0.0;
  double aloc[BO_MAX_Q];
  for (int j = 0; j < q; ++j) aloc[j] = 0.0;
  for (int s = tid; s < S; s += nt) {
    double u[BO_MAX_Q];
    int jstar;
    const double h = scalar_sample_value(a, sm, s, u, &jstar);
    sm.vals[s] = h;
    lmax = fmax(lmax, h);
    lsum += h;
    if (v == BO_ACQF_QUCB) {
      // u = mean + beta' |o - mean|  ->  o - mean = +-(u - mean) / beta'; recompute the sign from the sample
      double y[2 * BO_MAX_OBJECTIVES];
      sample_y(a, sm, jstar, s, y);
      const double o = scalar_objective_apply(a.od, y, M, nullptr);
      const double sg = (o > sm.omean[jstar]) ? 1.0 : ((o < sm.omean[jstar]) ? -1.0 : 0.0);
      aloc[jstar] += 1.0 - beta_p * sg;
    }
  }
  double bm, tsum;
  scalar_finish(a, sm, lmax, lsum, &bm, &tsum);
  if (v == BO_ACQF_QUCB) block_sum_vec(aloc, q, sm.aj, sm.red);
  // pass 2: gradients
  const double invS = 1.0 / (double)S;
  for (int s = tid; s < S; s += nt) {
    const double bf = a.best_f_s ? a.best_f_s[s] : a.best_f;
    double u[BO_MAX_Q], o[BO_MAX_Q];
    double dys[BO_MAX_Q][2 * BO_MAX_OBJECTIVES];   // d o_j / d y
    double ys[BO_MAX_Q][2 * BO_MAX_OBJECTIVES];
    double mx = -INFINITY;
    int jstar = 0;
    for (int j = 0; j < q; ++j) {
      sample_y(a, sm, j, s, ys[j]);
      o[j] = scalar_objective_apply(a.od, ys[j], M, dys[j]);
      double uj;
      if (is_log) {
        uj = log_fatplus_d(o[j] - bf, a.tau_relu);
        if (a.od.n_cons) uj += log_feas_fat(a.od, ys[j], 0.0, nullptr);
      } else if (v == BO_ACQF_QEI) {
        uj = fmax(o[j] - bf, 0.0);
        if (a.od.n_cons) uj *= feas_sigmoid(a.od, ys[j], 0.0, nullptr);
      } else if (v == BO_ACQF_QSR) {
        uj = o[j];
      } else if (v == BO_ACQF_QUCB) {
        uj = sm.omean[j] + beta_p * fabs(o[j] - sm.omean[j]);
      } else {
        uj = 1.0 / (1.0 + exp(-(o[j] - bf) / a.vparam));
        if (a.od.n_cons) uj *= feas_sigmoid(a.od, ys[j], 0.0, nullptr);
      }
      u[j] = uj;
      if (uj > mx) { mx = uj; jstar = j; }
    }
    for (int j = 0; j < q; ++j) {
      double dy[2 * BO_MAX_OBJECTIVES];
      for (int m = 0; m < M; ++m) dy[m] = 0.0;
      if (is_log) {
        // h = mx + tau log sum_j P((mx - u_j) / tau); value = logmeanexp_s h_s
        double ps = 0.0, dps = 0.0;
        for (int k = 0; k < q; ++k) {
          const double x = (mx - u[k]) / a.tau_max;
          const double P = pareto2_d(x);
          ps += P;
          dps += -P * P * (1.0 + x);
        }
        const double xj = (mx - u[j]) / a.tau_max;
        const double Pj = pareto2_d(xj);
        double dh = Pj * Pj * (1.0 + xj) / ps;        // through x_j, mx held fixed
        if (j == jstar) dh += 1.0 + dps / ps;          // through mx (amax -> arg-max element)
        const double wS = exp(sm.vals[s] - bm) / tsum;
        const double gu = wS * dh;                     // d value / d u_j
        const double go = gu * log_fatplus_grad_d(o[j] - bf, a.tau_relu);
        for (int m = 0; m < M; ++m) dy[m] = go * dys[j][m];
        if (a.od.n_cons) log_feas_fat(a.od, ys[j], gu, dy);
      } else if (v == BO_ACQF_QUCB) {
        const double sg = (o[j] > sm.omean[j]) ? 1.0 : ((o[j] < sm.omean[j]) ? -1.0 : 0.0);
        const double go = invS * ((j == jstar ? beta_p * sg : 0.0) + invS * sm.aj[j]);
        for (int m = 0; m < M; ++m) dy[m] = go * dys[j][m];
      } else if (j == jstar) {
        double go, w = 1.0, base;
        if (v == BO_ACQF_QEI) { base = fmax(o[j] - bf, 0.0); go = (o[j] - bf > 0.0) ? 1.0 : 0.0; }
        else if (v == BO_ACQF_QSR) { base = o[j]; go = 1.0; }
        else { base = 1.0 / (1.0 + exp(-(o[j] - bf) / a.vparam)); go = base * (1.0 - base) / a.vparam; }
        if (a.od.n_cons && v != BO_ACQF_QSR) w = feas_sigmoid(a.od, ys[j], invS * base, dy);
        for (int m = 0; m < M; ++m) dy[m] += invS * go * w * dys[j][m];
      }
      for (int m = 0; m < M; ++m) dF[(size_t)m * df_stride + ((size_t)batch * q + j) * S + s] = dy[m];
    }
  }
}

int launch_mc_scalar_grad(const McArgs& a, double* dF, size_t df_stride, cudaStream_t st, LaunchCounter* lc) {
  if (a.b <= 0) return BO_OK;
  const int nt = 128;
  size_t smem = scalar_smem_bytes(a);
  if (smem > 220 * 1024) { bo_set_error("mc_scalar_grad: shared memory budget exceeded"); return BO_ERR_INVALID; }
  static PerDeviceMax attr_pd; size_t& attr = attr_pd.slot();
  if (smem > 48 * 1024 && smem > attr) {
    CUDA_CHECK_RET(cudaFuncSetAttribute(mc_scalar_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  mc_scalar_grad_kernel<<<a.b, nt, smem, st>>>(a, dF, df_stride);
  if (lc) lc->n++;
  CUDA_CHECK_RET(cudaGetLastError());
  return BO_OK;
}

// ------------------------------------------------------------------------------------------------
// Baseline samples of the noisy variants: best FEASIBLE scalarised objective per MC sample (qNEI / qLogNEI incumbent,
// [UPSTREAM] compute_best_feasible_objective: samples that violate an output constraint, c(y) > 0, take `infeasible_value`
// -- minus infinity unless some MC sample has no feasible baseline point at all, in which case the host hands in BoTorch's
// pessimistic lower bound) and the arg-max counts of prune_inferior_points (infeasible samples count as -inf there).
// One warp per MC sample, lanes over the baseline points.  n_all_infeasible counts the MC samples without a feasible point.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
baseline_best_kernel(const double* __restrict__ F, int ldf, int S, int n, int M, const double* __restrict__ mean, ObjD od,
                     double infeasible_value, double* __restrict__ best_f_s, int* __restrict__ counts,
                     int* __restrict__ n_all_infeasible) {
  const int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (s >= S) return;
  double best = -INFINITY;
  int arg = 0x7fffffff;
  int any_feasible = 0;
  for (int e = lane; e < n; e += 32) {
    double y[2 * BO_MAX_OBJECTIVES];
    for (int m = 0; m < M; ++m) y[m] = mean[(size_t)e * M + m] + F[((size_t)m * S + s) * ldf + e];
    double o = scalar_objective_apply(od, y, M, nullptr);
    bool ok = true;
    for (int c = 0; c < od.n_cons; ++c) ok = ok && (od.con[c].sign * (y[od.con[c].out_idx] - od.con[c].tp) <= 0.0);
    if (!ok) o = infeasible_value;
    any_feasible |= ok ? 1 : 0;
    // first arg-max like torch.argmax: a sample whose points are all -inf keeps the smallest index
    if (o > best || (o == best && e < arg)) { best = o; arg = e; }
  }
  for (int off = 16; off > 0; off >>= 1) {
    const double ob = __shfl_xor_sync(0xffffffffu, best, off);
    const int oa = __shfl_xor_sync(0xffffffffu, arg, off);
    any_feasible |= __shfl_xor_sync(0xffffffffu, any_feasible, off);
    if (ob > best || (ob == best && oa < arg)) { best = ob; arg = oa; }  // first arg-max, like torch.argmax
  }
  if (lane == 0) {
    if (best_f_s) best_f_s[s] = best;
    if (counts && arg < n) atomicAdd(counts + arg, 1);
    if (n_all_infeasible && !any_feasible) atomicAdd(n_all_infeasible, 1);
  }
}

int launch_baseline_best(const double* F, int ldf, int S, int n, int M, const double* mean, const ObjD& od,
                         double infeasible_value, double* best_f_s, int* counts, int* n_all_infeasible, cudaStream_t st,
                         LaunchCounter* lc) {
  if (S <= 0 || n <= 0) return BO_OK;
  baseline_best_kernel<<<(S * 32 + 255) / 256, 256, 0, st>>>(F, ldf, S, n, M, mean, od, infeasible_value, best_f_s, counts,
                                                             n_all_infeasible);
  if (lc) lc->n++;
  CUDA_CHECK_RET(cudaGetLastError());
  return BO_OK;
}

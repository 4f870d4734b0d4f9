// Log-space hypervolume improvement: qLogEHVI / qLogNEHVI (MoboStrategy's default acquisition function,
// strategies/predictives/mobo.py:72-90 -> get_acquisition_function("qLogNEHVI", ...)).
// [UPSTREAM] botorch.acquisition.multi_objective.logei._compute_log_qehvi, restated in oracle/bo_oracle.py
// (log_hvi_inclusion_exclusion).  Per MC sample s and cell c, for every non-empty subset J of the q points:
//   v_o   = fatmin_{j in J} obj_jo                 (tau_max)
//   w_o   = fatmin(v_o, upper_co)                  (exact v_o for an unbounded cell side)
//   la_J  = sum_o log_fatplus(w_o - lower_co, tau_relu)  [+ sum_{j in J} log feasibility_j]
//   cell  = log( sum_{|J| odd} e^la_J - sum_{|J| even} e^la_J ),   sample = logsumexp_c cell,   value = logmeanexp_s sample
// Unlike the plain HVI nothing can be skipped: the fat tails make every (cell, subset) pair contribute.
// One CTA per q-batch, threads over MC samples.  The gradient kernel uses running-max rescaling over the cells (the
// softmax weights of the cells are only known at the end) and writes d value / d f per MC sample.
#include "acqf.cuh"
#include "common.cuh"
#include "mc_math.cuh"

#include <algorithm>

#define LH_MAXO BO_MAX_OBJECTIVES

// side length of the overlap of subset `sub` with the cell along objective o; optionally d len / d obj_j for the members
__device__ __forceinline__ double subset_axis(unsigned sub, const double* __restrict__ objs_o /* [q] stride nt */, int nt,
                                              double lo, double up, double tau_max, double* dlen /* [q] or null */) {
  // fatmin over the members
  double mn = INFINITY;
  int kstar = -1;
  for (unsigned rest = sub; rest; rest &= rest - 1) {
    const int j = __ffs(rest) - 1;
    const double x = objs_o[(size_t)j * nt];
    if (x < mn) { mn = x; kstar = j; }
  }
  double v, ps = 1.0, dps = -1.0;  // the minimum itself: P(0) = 1, P'(0) = -1
  if ((sub & (sub - 1)) == 0) {
    v = mn;
  } else {
    ps = 0.0; dps = 0.0;
    for (unsigned rest = sub; rest; rest &= rest - 1) {
      const int j = __ffs(rest) - 1;
      const double y = (objs_o[(size_t)j * nt] - mn) / tau_max;
      const double P = pareto2_d(y);
      ps += P;
      dps += -P * P * (1.0 + y);
    }
    v = mn - tau_max * log(ps);
  }
  // smooth minimum with the cell's upper bound
  double w, dw_dv;
  if (isinf(up)) {
    w = v; dw_dv = 1.0;
  } else if (v <= up) {
    const double y = (up - v) / tau_max, P = pareto2_d(y);
    w = v - tau_max * log(1.0 + P);
    dw_dv = 1.0 + (-P * P * (1.0 + y)) / (1.0 + P);
  } else {
    const double y = (v - up) / tau_max, P = pareto2_d(y);
    w = up - tau_max * log(1.0 + P);
    dw_dv = P * P * (1.0 + y) / (1.0 + P);
  }
  if (dlen) {
    for (unsigned rest = sub; rest; rest &= rest - 1) {
      const int j = __ffs(rest) - 1;
      double dv;
      if ((sub & (sub - 1)) == 0) dv = 1.0;
      else {
        const double y = (objs_o[(size_t)j * nt] - mn) / tau_max;
        const double P = pareto2_d(y);
        dv = P * P * (1.0 + y) / ps;               // -P'(y_j) / sum P
        if (j == kstar) dv += 1.0 + dps / ps;       // through the minimum
      }
      dlen[j] = dw_dv * dv;
    }
  }
  return w - lo;
}

// The two halves of subset_axis for the value-only kernel: the fatmin over the members does not depend on the cell, so the
// forward kernel forms it ONCE per (MC sample, subset, objective) and re-uses it for every cell (same arithmetic, same bits).
__device__ __forceinline__ double subset_fatmin(unsigned sub, const double* __restrict__ objs_o, int nt, double tau_max) {
  double mn = INFINITY;
  for (unsigned rest = sub; rest; rest &= rest - 1) {
    const double x = objs_o[(size_t)(__ffs(rest) - 1) * nt];
    if (x < mn) mn = x;
  }
  if ((sub & (sub - 1)) == 0) return mn;
  double ps = 0.0;
  for (unsigned rest = sub; rest; rest &= rest - 1) {
    const double y = (objs_o[(size_t)(__ffs(rest) - 1) * nt] - mn) / tau_max;
    ps += pareto2_d(y);
  }
  return mn - tau_max * log(ps);
}
__device__ __forceinline__ double axis_from_fatmin(double v, double lo, double up, double tau_max) {
  double w;
  if (isinf(up)) {
    w = v;
  } else if (v <= up) {
    const double y = (up - v) / tau_max, P = pareto2_d(y);
    w = v - tau_max * log(1.0 + P);
  } else {
    const double y = (v - up) / tau_max, P = pareto2_d(y);
    w = up - tau_max * log(1.0 + P);
  }
  return w - lo;
}

// ---- value kernel only: table-driven logarithm and reciprocal temperatures ---------------------------------------------------
// The value kernel spends its time in FP64 `log` and divisions (two of each per (cell, subset, objective) after the fatmin
// cache).  Every logarithm here has a positive, finite, normal argument and enters a sum of O(1) terms, so 2e-16 ABSOLUTE
// accuracy is what matters: x = 2^e m, m in [1, 2); the top 6 mantissa bits pick c_i = 1 + (i + 1/2) / 64; r = m / c_i - 1
// (|r| < 2^-7, one FMA with the tabulated reciprocal); log x = e ln 2 + log c_i + log1p(r), log1p(r) by its series to r^8
// (|r|^9 / 9 < 2^-66).  log c_i is tabulated for the ROUNDED reciprocal, so the identity is exact up to the final roundings.
// The adjoint kernels keep the library functions; their values agree with this kernel's to ~1e-15.
#define LH_LOGTAB 64
__device__ __forceinline__ void lh_logtab_fill(double2* tab) {   // call with all threads of the CTA, then __syncthreads()
  for (int i = threadIdx.x; i < LH_LOGTAB; i += blockDim.x) {
    const double invc = 1.0 / (1.0 + ((double)i + 0.5) / (double)LH_LOGTAB);
    tab[i] = make_double2(invc, -log(invc));
  }
}
__device__ __forceinline__ double lh_log(double x, const double2* __restrict__ tab) {
  const int hi = __double2hiint(x);
  const int e = (hi >> 20) - 1023;
  const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(x));
  const double2 t = tab[(hi >> 14) & (LH_LOGTAB - 1)];
  const double r = fma(m, t.x, -1.0);
  double p = fma(r, -0.125, 1.0 / 7.0);
  p = fma(p, r, -1.0 / 6.0);
  p = fma(p, r, 0.2);
  p = fma(p, r, -0.25);
  p = fma(p, r, 1.0 / 3.0);
  p = fma(p, r, -0.5);
  p = fma(p * r, r, r);                       // r - r^2/2 + ... - r^8/8
  return fma((double)e, 0.6931471805599453, t.y) + p;
}
// axis_from_fatmin / log_fatplus_lt with the table logarithm and reciprocal temperatures (same branches, same formulas)
__device__ __forceinline__ double axis_from_fatmin_fast(double v, double lo, double up, double tau_max, double inv_tau_max,
                                                        const double2* __restrict__ tab) {
  double w;
  if (isinf(up)) {
    w = v;
  } else {
    const double y = fabs(up - v) * inv_tau_max, P = pareto2_d(y);
    w = fmin(v, up) - tau_max * lh_log(1.0 + P, tab);
  }
  return w - lo;
}
__device__ __forceinline__ double log_fatplus_fast(double x, double tau, double inv_tau, double log_tau, const double2* __restrict__ tab) {
  const double z = x * inv_tau;
  if (z > 32.0) {
    const double r = 0.1 / (z * (1.0 + z * z));
    return log_tau + lh_log(z, tab) + ((r < 1e-9) ? r : log1p(r));
  }
  const double B = -2.302585092994046 - ((z * z < 1e300) ? lh_log(1.0 + z * z, tab) : log1p(z * z));
  if (z < -750.0) return log_tau + B;
  return log_tau + logaddexp_d(log_softplus_d(z), B);
}

// running log-sum-exp as a (max, sum) pair: value = m + log s
__device__ __forceinline__ void lse_push(double& m, double& s, double v) {
  if (v > m) { s = s * exp(m - v) + 1.0; m = v; }   // exp(-inf) = 0 on the first push
  else if (!(isinf(v) && v < 0)) s += exp(v - m);
}
__device__ __forceinline__ double lse_value(double m, double s) { return (s > 0.0) ? m + log(s) : -INFINITY; }

struct LhSmem { double *root, *mu, *objs, *lfw, *gob, *glf, *ys, *vc, *vals, *red; };

__device__ __forceinline__ LhSmem lh_smem(double* base, int M, int q, int nr, int Mo, int nt, int S, bool grad, int vcn = 0) {
  LhSmem sm;
  sm.root = base;
  sm.mu = sm.root + (size_t)M * q * nr;
  sm.objs = sm.mu + q * M;
  sm.lfw = sm.objs + (size_t)q * Mo * nt;
  double* p = sm.lfw + (size_t)q * nt;
  if (grad) {
    sm.gob = p; p += (size_t)q * Mo * nt;
    sm.glf = p; p += (size_t)q * nt;
    sm.ys = p; p += (size_t)q * M * nt;
  } else {
    sm.gob = sm.glf = sm.ys = nullptr;
  }
  sm.vc = vcn ? p : nullptr;   // [vcn][nt] cell-independent fatmins of the subsets (value kernel)
  p += (size_t)vcn * nt;
  sm.vals = p;
  sm.red = sm.vals + S;
  return sm;
}

__device__ __forceinline__ void lh_load_sample(const McArgs& a, const LhSmem& sm, int batch, int s, int tid, int nt) {
  const int q = a.q, nb = a.nb, nr = nb + q, M = a.M, S = a.S, Mo = a.od.n_obj;
  for (int j = 0; j < q; ++j) {
    double y[2 * BO_MAX_OBJECTIVES];
    for (int m = 0; m < M; ++m) {
      const double* rr = sm.root + ((size_t)m * q + j) * nr;
      double sb = 0.0, sq = 0.0;
      if (a.Fp) sb = a.Fp[(size_t)m * a.fp_stride + ((size_t)batch * q + j) * S + s];
      else for (int e = 0; e < nb; ++e) sb = fma(rr[e], a.zbT[((size_t)e * M + m) * S + s], sb);
      for (int k = 0; k < q; ++k) sq = fma(rr[nb + k], a.zqT[((size_t)k * M + m) * S + s], sq);
      y[m] = (sm.mu[j * M + m] + sb) + sq;
      if (sm.ys) sm.ys[((size_t)j * M + m) * nt + tid] = y[m];
    }
    for (int o = 0; o < Mo; ++o) {
      sm.objs[((size_t)j * Mo + o) * nt + tid] = objective_apply(a.od.op[o], y);
      if (sm.gob) sm.gob[((size_t)j * Mo + o) * nt + tid] = 0.0;
    }
    sm.lfw[(size_t)j * nt + tid] = a.od.n_cons ? log_feas_fat(a.od, y, 0.0, nullptr) : 0.0;
    if (sm.glf) sm.glf[(size_t)j * nt + tid] = 0.0;
  }
}

// log area of the overlap of subset `sub` with the cell
__device__ __forceinline__ double subset_logarea(const McArgs& a, const LhSmem& sm, unsigned sub, const double* lo,
                                                 const double* up, int tid, int nt, double log_tau_relu) {
  const int Mo = a.od.n_obj;
  double la = 0.0;
  for (int o = 0; o < Mo; ++o) {
    const double len = subset_axis(sub, sm.objs + (size_t)o * nt + tid, Mo * nt, lo[o], up[o], a.tau_max, nullptr);
    la += log_fatplus_lt(len, a.tau_relu, log_tau_relu);
  }
  if (a.od.n_cons)
    for (unsigned rest = sub; rest; rest &= rest - 1) la += sm.lfw[(size_t)(__ffs(rest) - 1) * nt + tid];
  return la;
}

__device__ __forceinline__ void lh_finish(const McArgs& a, const LhSmem& sm, double lmax, double* bm_out, double* tsum_out) {
  const int tid = threadIdx.x, nt = blockDim.x, batch = blockIdx.x, S = a.S;
  for (int o = 16; o > 0; o >>= 1) lmax = fmax(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
  __syncthreads();
  if ((tid & 31) == 0) sm.red[tid >> 5] = lmax;
  __syncthreads();
  double bm = -INFINITY;
  for (int w = 0; w < (nt >> 5); ++w) bm = fmax(bm, sm.red[w]);
  double se = 0.0;
  for (int s = tid; s < S; s += nt) se += exp(sm.vals[s] - bm);
  double t = block_sum(se, sm.red);
  if (tid == 0) sm.red[32] = t;
  __syncthreads();
  const double tsum = sm.red[32];
  if (tid == 0) {
    a.out[batch] = bm + log(tsum) - log((double)S);
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
mc_loghvi_kernel(McArgs a, int vcn) {
  extern __shared__ double lsm[];
  const int batch = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  const int q = a.q, nr = a.nb + q, M = a.M, S = a.S, Mo = a.od.n_obj;
  LhSmem sm = lh_smem(lsm, M, q, nr, Mo, nt, S, false, vcn);
  __shared__ double2 logtab[LH_LOGTAB];
  lh_logtab_fill(logtab);
  for (int i = tid; i < M * q * nr; i += nt) sm.root[i] = a.root[(size_t)batch * M * q * nr + i];
  for (int i = tid; i < q * M; i += nt) sm.mu[i] = a.mu[(size_t)batch * q * M + i];
  __syncthreads();
  const unsigned full = (q >= 32) ? 0xffffffffu : ((1u << q) - 1u);
  const double log_tau_relu = log(a.tau_relu);
  const double inv_tau_relu = 1.0 / a.tau_relu, inv_tau_max = 1.0 / a.tau_max;
  double lmax = -INFINITY;
  for (int s = tid; s < S; s += nt) {
    lh_load_sample(a, sm, batch, s, tid, nt);
    if (sm.vc)
      for (unsigned sub = 1; sub <= full; ++sub)
        for (int o = 0; o < Mo; ++o)
          sm.vc[((size_t)(sub - 1) * Mo + o) * nt + tid] = subset_fatmin(sub, sm.objs + (size_t)o * nt + tid, Mo * nt, a.tau_max);
    const int nc = a.cells_shared ? a.ncells[0] : a.ncells[s];
    const int sc = a.cells_shared ? 0 : s;
    const int Sc = a.cells_shared ? 1 : S;
    double R = -INFINITY, ssum = 0.0;  // running logsumexp over the cells
    for (int c = 0; c < nc; ++c) {
      double lo[LH_MAXO], up[LH_MAXO];
      for (int o = 0; o < Mo; ++o) {
        lo[o] = a.cell_lo[((size_t)c * Mo + o) * Sc + sc];
        up[o] = a.cell_up[((size_t)c * Mo + o) * Sc + sc];
      }
      // running (max, sum) pairs: one exp per subset instead of a full logaddexp
      double mo = -INFINITY, so = 0.0, me = -INFINITY, se = 0.0;
      for (unsigned sub = 1; sub <= full; ++sub) {
        double la;
        if (sm.vc) {
          la = 0.0;
          for (int o = 0; o < Mo; ++o) {
            const double len = axis_from_fatmin_fast(sm.vc[((size_t)(sub - 1) * Mo + o) * nt + tid], lo[o], up[o], a.tau_max,
                                                     inv_tau_max, logtab);
            la += log_fatplus_fast(len, a.tau_relu, inv_tau_relu, log_tau_relu, logtab);
          }
          if (a.od.n_cons)
            for (unsigned rest = sub; rest; rest &= rest - 1) la += sm.lfw[(size_t)(__ffs(rest) - 1) * nt + tid];
        } else {
          la = subset_logarea(a, sm, sub, lo, up, tid, nt, log_tau_relu);
        }
        if (__popc(sub) & 1) lse_push(mo, so, la);
        else lse_push(me, se, la);
      }
      const double odd = lse_value(mo, so), even = lse_value(me, se);
      const double cellv = (isinf(even) && even < 0) ? odd : odd + log1mexp_d(even - odd);
      if (cellv > R) { ssum = ssum * exp(R - cellv) + 1.0; R = cellv; }
      else if (!(isinf(cellv) && cellv < 0)) ssum += exp(cellv - R);
    }
    const double sv = (nc > 0) ? R + log(ssum) : -INFINITY;
    sm.vals[s] = sv;
    lmax = fmax(lmax, sv);
  }
  lh_finish(a, sm, lmax, nullptr, nullptr);
}

static int lh_pick_threads(const McArgs& a, bool grad, size_t* smem_out, int vcn = 0) {
  const size_t fixed = (size_t)a.M * a.q * (a.nb + a.q) + (size_t)a.q * a.M + a.S + 40;
  const size_t per = (size_t)a.q * a.od.n_obj + a.q + (grad ? (size_t)a.q * a.od.n_obj + a.q + (size_t)a.q * a.M : 0) + vcn;
  for (int nt = 256; nt >= 32; nt >>= 1) {
    size_t smem = (fixed + per * nt) * sizeof(double);
    if (smem <= 200 * 1024) { *smem_out = smem; return nt; }
  }
  return 0;
}

int launch_mc_loghvi(const McArgs& a, cudaStream_t st, LaunchCounter* lc) {
  if (a.b <= 0) return BO_OK;
  size_t smem = 0;
  // cell-independent subset fatmins cached per thread when they are few: (2^q - 1) n_obj <= 64 doubles (q <= 4, or q = 5 with
  // two objectives); and only while that keeps at least 128 threads per CTA
  int vcn = (a.q <= 5) ? ((1 << a.q) - 1) * a.od.n_obj : 0;
  if (vcn > 64) vcn = 0;
  int nt = lh_pick_threads(a, false, &smem, vcn);
  if (vcn && nt < 128) { vcn = 0; nt = lh_pick_threads(a, false, &smem, 0); }
  if (!nt) { bo_set_error("mc_loghvi: shared memory budget exceeded (n_b=%d q=%d)", a.nb, a.q); return BO_ERR_INVALID; }
  static PerDeviceMax attr_pd; size_t& attr = attr_pd.slot();
  // (the kernel also holds 1 KB of static shared memory, the logarithm table: static + dynamic above 48 KB needs the opt-in)
  if (smem + 2048 > 48 * 1024 && smem > attr) {
    CUDA_CHECK_RET(cudaFuncSetAttribute(mc_loghvi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  mc_loghvi_kernel<<<a.b, nt, smem, st>>>(a, vcn);
  if (lc) lc->n++;
  CUDA_CHECK_RET(cudaGetLastError());
  return BO_OK;
}

// Few q-batches (refinement: 8 restarts): the MC samples of a q-batch are split over gridDim.y CTAs, which leave the
// per-sample values in `vals_g` [b][S] and the un-normalised d value / d f in dF; loghvi_grad_finish_kernel then forms the
// log-mean-exp over the samples and applies the samples' softmax weights.
__global__ void __launch_bounds__(256)
loghvi_grad_finish_kernel(McArgs a, const double* __restrict__ vals_g, double* __restrict__ dF, size_t df_stride) {
  __shared__ double red[40];
  const int batch = blockIdx.x, tid = threadIdx.x, nt = blockDim.x, S = a.S, q = a.q, M = a.M;
  const double* v = vals_g + (size_t)batch * S;
  double lmax = -INFINITY;
  for (int s = tid; s < S; s += nt) lmax = fmax(lmax, v[s]);
  for (int o = 16; o > 0; o >>= 1) lmax = fmax(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
  if ((tid & 31) == 0) red[tid >> 5] = lmax;
  __syncthreads();
  double bm = -INFINITY;
  for (int w = 0; w < (nt >> 5); ++w) bm = fmax(bm, red[w]);
  __syncthreads();
  double se = 0.0;
  for (int s = tid; s < S; s += nt) se += exp(v[s] - bm);
  const double t = block_sum(se, red);
  if (tid == 0) red[32] = t;
  __syncthreads();
  const double tsum = red[32];
  if (tid == 0) {
    a.out[batch] = bm + log(tsum) - log((double)S);
    if (a.info_out) {
      int f = 0;
      for (int m = 0; m < M; ++m) f |= a.info_in[(size_t)batch * M + m];
      a.info_out[batch] = f;
    }
  }
  for (int s = tid; s < S; s += nt) {
    const double wS = exp(v[s] - bm) / tsum;
    for (int j = 0; j < q; ++j)
      for (int m = 0; m < M; ++m) dF[(size_t)m * df_stride + ((size_t)batch * q + j) * S + s] *= wS;
  }
}

__global__ void __launch_bounds__(256)
mc_loghvi_grad_kernel(McArgs a, double* __restrict__ dF, size_t df_stride, double* __restrict__ vals_g) {
  extern __shared__ double lsm[];
  const int batch = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  const int q = a.q, nr = a.nb + q, M = a.M, S = a.S, Mo = a.od.n_obj;
  LhSmem sm = lh_smem(lsm, M, q, nr, Mo, nt, S, true);
  for (int i = tid; i < M * q * nr; i += nt) sm.root[i] = a.root[(size_t)batch * M * q * nr + i];
  for (int i = tid; i < q * M; i += nt) sm.mu[i] = a.mu[(size_t)batch * q * M + i];
  __syncthreads();
  const unsigned full = (q >= 32) ? 0xffffffffu : ((1u << q) - 1u);
  const bool has_cons = a.od.n_cons > 0;
  const double log_tau_relu = log(a.tau_relu);
  double lmax = -INFINITY;
  const int per_split = (S + gridDim.y - 1) / gridDim.y;
  const int s_begin = blockIdx.y * per_split, s_end = min(S, s_begin + per_split);
  for (int s = s_begin + tid; s < s_end; s += nt) {
    lh_load_sample(a, sm, batch, s, tid, nt);
    const int nc = a.cells_shared ? a.ncells[0] : a.ncells[s];
    const int sc = a.cells_shared ? 0 : s;
    const int Sc = a.cells_shared ? 1 : S;
    double R = -INFINITY, ssum = 0.0;
    for (int c = 0; c < nc; ++c) {
      double lo[LH_MAXO], up[LH_MAXO];
      for (int o = 0; o < Mo; ++o) {
        lo[o] = a.cell_lo[((size_t)c * Mo + o) * Sc + sc];
        up[o] = a.cell_up[((size_t)c * Mo + o) * Sc + sc];
      }
      double mo = -INFINITY, so = 0.0, me = -INFINITY, se = 0.0;
      for (unsigned sub = 1; sub <= full; ++sub) {
        const double la = subset_logarea(a, sm, sub, lo, up, tid, nt, log_tau_relu);
        if (__popc(sub) & 1) lse_push(mo, so, la);
        else lse_push(me, se, la);
      }
      const double odd = lse_value(mo, so), even = lse_value(me, se);
      const bool no_even = isinf(even) && even < 0;
      const double cellv = no_even ? odd : odd + log1mexp_d(even - odd);
      if (isinf(cellv) && cellv < 0) continue;
      // running-max rescaling of everything accumulated so far
      double wc;
      if (cellv > R) {
        const double sc_old = exp(R - cellv);  // 0 when R = -inf
        for (int i = 0; i < q * Mo; ++i) sm.gob[(size_t)i * nt + tid] *= sc_old;
        if (has_cons) for (int j = 0; j < q; ++j) sm.glf[(size_t)j * nt + tid] *= sc_old;
        ssum = ssum * sc_old + 1.0;
        R = cellv;
        wc = 1.0;
      } else {
        wc = exp(cellv - R);
        ssum += wc;
      }
      const double r = no_even ? 0.0 : exp(even - odd);
      const double a_odd = 1.0 / (1.0 - r), a_even = -r / (1.0 - r);
      for (unsigned sub = 1; sub <= full; ++sub) {
        // recompute the subset: side lengths and their derivatives
        double la = 0.0;
        double dlen[LH_MAXO][BO_MAX_Q];
        double dlf[LH_MAXO];
        for (int o = 0; o < Mo; ++o) {
          const double len = subset_axis(sub, sm.objs + (size_t)o * nt + tid, Mo * nt, lo[o], up[o], a.tau_max, dlen[o]);
          la += log_fatplus_lt(len, a.tau_relu, log_tau_relu);
          dlf[o] = log_fatplus_grad_d(len, a.tau_relu);
        }
        if (has_cons)
          for (unsigned rest = sub; rest; rest &= rest - 1) la += sm.lfw[(size_t)(__ffs(rest) - 1) * nt + tid];
        const bool is_odd = __popc(sub) & 1;
        const double kappa = wc * (is_odd ? a_odd * exp(la - odd) : a_even * exp(la - even));
        if (kappa == 0.0) continue;
        for (unsigned rest = sub; rest; rest &= rest - 1) {
          const int j = __ffs(rest) - 1;
          for (int o = 0; o < Mo; ++o) sm.gob[((size_t)j * Mo + o) * nt + tid] += kappa * dlf[o] * dlen[o][j];
          if (has_cons) sm.glf[(size_t)j * nt + tid] += kappa;
        }
      }
    }
    const double sv = (ssum > 0.0) ? R + log(ssum) : -INFINITY;
    sm.vals[s] = sv;
    if (vals_g) vals_g[(size_t)batch * S + s] = sv;
    lmax = fmax(lmax, sv);
    // un-normalised d sample value / d f (scaled by the sample's softmax weight after the block reduction)
    const double inv = (ssum > 0.0) ? 1.0 / ssum : 0.0;
    for (int j = 0; j < q; ++j) {
      double y[2 * BO_MAX_OBJECTIVES], dy[2 * BO_MAX_OBJECTIVES];
      for (int m = 0; m < M; ++m) { y[m] = sm.ys[((size_t)j * M + m) * nt + tid]; dy[m] = 0.0; }
      for (int o = 0; o < Mo; ++o) {
        const double gv = sm.gob[((size_t)j * Mo + o) * nt + tid] * inv;
        if (gv != 0.0) dy[a.od.op[o].out_idx] += gv * objective_grad(a.od.op[o], y);
      }
      if (has_cons) log_feas_fat(a.od, y, sm.glf[(size_t)j * nt + tid] * inv, dy);
      for (int m = 0; m < M; ++m) dF[(size_t)m * df_stride + ((size_t)batch * q + j) * S + s] = dy[m];
    }
  }
  if (vals_g) return;   // split over the samples: loghvi_grad_finish_kernel completes the q-batch
  double bm, tsum;
  lh_finish(a, sm, lmax, &bm, &tsum);
  for (int s = tid; s < S; s += nt) {
    const double wS = exp(sm.vals[s] - bm) / tsum;
    for (int j = 0; j < q; ++j)
      for (int m = 0; m < M; ++m) dF[(size_t)m * df_stride + ((size_t)batch * q + j) * S + s] *= wS;
  }
}

// Few q-batches (refinement): besides the split over the MC samples (gridDim.y), the CELLS of a sample are dealt to CL "cell
// lanes": thread (sl, cl) of a CTA holding nst samples walks the cells c = cl, cl + CL, ... with its own running maximum,
// sum and adjoint block; the lanes of a sample are merged afterwards in lane order (deterministic) by rescaling to the
// common maximum.  Leaves the per-sample values in vals_g and the un-normalised d value / d f in dF
// (loghvi_grad_finish_kernel completes the q-batch).
__global__ void __launch_bounds__(256)
mc_loghvi_grad_cl_kernel(McArgs a, double* __restrict__ dF, size_t df_stride, double* __restrict__ vals_g, int CL) {
  extern __shared__ double lsm[];
  const int batch = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  const int nst = nt / CL, sl = tid % nst, cl = tid / nst;
  const int q = a.q, nb = a.nb, nr = nb + q, M = a.M, S = a.S, Mo = a.od.n_obj;
  double* root = lsm;                            // [M][q][nr]
  double* mu = root + (size_t)M * q * nr;        // [q][M]
  double* objs = mu + q * M;                     // [q*Mo][nst]
  double* lfw = objs + (size_t)q * Mo * nst;     // [q][nst]
  double* ys = lfw + (size_t)q * nst;            // [q*M][nst]
  double* gob = ys + (size_t)q * M * nst;        // [q*Mo][nt]  per thread
  double* glf = gob + (size_t)q * Mo * nt;       // [q][nt]     per thread
  double* Rs = glf + (size_t)q * nt;             // [nt] running maximum of the thread's cells
  double* Ss = Rs + nt;                          // [nt] running sum
  for (int i = tid; i < M * q * nr; i += nt) root[i] = a.root[(size_t)batch * M * q * nr + i];
  for (int i = tid; i < q * M; i += nt) mu[i] = a.mu[(size_t)batch * q * M + i];
  const unsigned full = (q >= 32) ? 0xffffffffu : ((1u << q) - 1u);
  const bool has_cons = a.od.n_cons > 0;
  const double log_tau_relu = log(a.tau_relu);
  const int per_split = (S + gridDim.y - 1) / gridDim.y;
  const int s_begin = blockIdx.y * per_split, s_end = min(S, s_begin + per_split);
  for (int s0 = s_begin; s0 < s_end; s0 += nst) {
    const int s = s0 + sl;
    const bool valid = s < s_end;
    __syncthreads();
    if (valid)
      for (int j = cl; j < q; j += CL) {
        double y[2 * BO_MAX_OBJECTIVES];
        for (int m = 0; m < M; ++m) {
          const double* rr = root + ((size_t)m * q + j) * nr;
          double sb = 0.0, sq = 0.0;
          if (a.Fp) sb = a.Fp[(size_t)m * a.fp_stride + ((size_t)batch * q + j) * S + s];
          else for (int e = 0; e < nb; ++e) sb = fma(rr[e], a.zbT[((size_t)e * M + m) * S + s], sb);
          for (int k = 0; k < q; ++k) sq = fma(rr[nb + k], a.zqT[((size_t)k * M + m) * S + s], sq);
          y[m] = (mu[j * M + m] + sb) + sq;
          ys[((size_t)j * M + m) * nst + sl] = y[m];
        }
        for (int o = 0; o < Mo; ++o) objs[((size_t)j * Mo + o) * nst + sl] = objective_apply(a.od.op[o], y);
        lfw[(size_t)j * nst + sl] = a.od.n_cons ? log_feas_fat(a.od, y, 0.0, nullptr) : 0.0;
      }
    for (int i = 0; i < q * Mo; ++i) gob[(size_t)i * nt + tid] = 0.0;
    for (int j = 0; j < q; ++j) glf[(size_t)j * nt + tid] = 0.0;
    __syncthreads();
    double R = -INFINITY, ssum = 0.0;
    if (valid) {
      const int nc = a.cells_shared ? a.ncells[0] : a.ncells[s];
      const int sc = a.cells_shared ? 0 : s;
      const int Sc = a.cells_shared ? 1 : S;
      for (int c = cl; c < nc; c += CL) {
        double lo[LH_MAXO], up[LH_MAXO];
        for (int o = 0; o < Mo; ++o) {
          lo[o] = a.cell_lo[((size_t)c * Mo + o) * Sc + sc];
          up[o] = a.cell_up[((size_t)c * Mo + o) * Sc + sc];
        }
        double mo = -INFINITY, so = 0.0, me = -INFINITY, se = 0.0;
        for (unsigned sub = 1; sub <= full; ++sub) {
          double la = 0.0;
          for (int o = 0; o < Mo; ++o) {
            const double len = subset_axis(sub, objs + (size_t)o * nst + sl, Mo * nst, lo[o], up[o], a.tau_max, nullptr);
            la += log_fatplus_lt(len, a.tau_relu, log_tau_relu);
          }
          if (has_cons)
            for (unsigned rest = sub; rest; rest &= rest - 1) la += lfw[(size_t)(__ffs(rest) - 1) * nst + sl];
          if (__popc(sub) & 1) lse_push(mo, so, la);
          else lse_push(me, se, la);
        }
        const double odd = lse_value(mo, so), even = lse_value(me, se);
        const bool no_even = isinf(even) && even < 0;
        const double cellv = no_even ? odd : odd + log1mexp_d(even - odd);
        if (isinf(cellv) && cellv < 0) continue;
        double wc;
        if (cellv > R) {
          const double sc_old = exp(R - cellv);  // 0 when R = -inf
          for (int i = 0; i < q * Mo; ++i) gob[(size_t)i * nt + tid] *= sc_old;
          if (has_cons) for (int j = 0; j < q; ++j) glf[(size_t)j * nt + tid] *= sc_old;
          ssum = ssum * sc_old + 1.0;
          R = cellv;
          wc = 1.0;
        } else {
          wc = exp(cellv - R);
          ssum += wc;
        }
        const double r = no_even ? 0.0 : exp(even - odd);
        const double a_odd = 1.0 / (1.0 - r), a_even = -r / (1.0 - r);
        for (unsigned sub = 1; sub <= full; ++sub) {
          double la = 0.0;
          double dlen[LH_MAXO][BO_MAX_Q];
          double dlf[LH_MAXO];
          for (int o = 0; o < Mo; ++o) {
            const double len = subset_axis(sub, objs + (size_t)o * nst + sl, Mo * nst, lo[o], up[o], a.tau_max, dlen[o]);
            la += log_fatplus_lt(len, a.tau_relu, log_tau_relu);
            dlf[o] = log_fatplus_grad_d(len, a.tau_relu);
          }
          if (has_cons)
            for (unsigned rest = sub; rest; rest &= rest - 1) la += lfw[(size_t)(__ffs(rest) - 1) * nst + sl];
          const bool is_odd = __popc(sub) & 1;
          const double kappa = wc * (is_odd ? a_odd * exp(la - odd) : a_even * exp(la - even));
          if (kappa == 0.0) continue;
          for (unsigned rest = sub; rest; rest &= rest - 1) {
            const int j = __ffs(rest) - 1;
            for (int o = 0; o < Mo; ++o) gob[((size_t)j * Mo + o) * nt + tid] += kappa * dlf[o] * dlen[o][j];
            if (has_cons) glf[(size_t)j * nt + tid] += kappa;
          }
        }
      }
    }
    Rs[tid] = R;
    Ss[tid] = ssum;
    __syncthreads();
    if (valid) {
      // merge the cell lanes of this sample: common maximum, then sums in lane order
      double Rm = -INFINITY;
      for (int k = 0; k < CL; ++k) Rm = fmax(Rm, Rs[k * nst + sl]);
      double stot = 0.0, wk[8];                  // CL <= 8 (launcher); lanes without a finite cell weigh nothing
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const double Rk = (k < CL) ? Rs[k * nst + sl] : -INFINITY;
        wk[k] = (isinf(Rk) && Rk < 0) ? 0.0 : exp(Rk - Rm);
        if (k < CL) stot += Ss[k * nst + sl] * wk[k];
      }
      if (cl == 0) vals_g[(size_t)batch * S + s] = (stot > 0.0) ? Rm + log(stot) : -INFINITY;
      const double inv = (stot > 0.0) ? 1.0 / stot : 0.0;
      for (int j = cl; j < q; j += CL) {
        double y[2 * BO_MAX_OBJECTIVES], dy[2 * BO_MAX_OBJECTIVES];
        for (int m = 0; m < M; ++m) { y[m] = ys[((size_t)j * M + m) * nst + sl]; dy[m] = 0.0; }
        double gl = 0.0;
        for (int o = 0; o < Mo; ++o) {
          double gv = 0.0;
#pragma unroll
          for (int k = 0; k < 8; ++k)
            if (k < CL) gv += gob[((size_t)j * Mo + o) * nt + k * nst + sl] * wk[k];
          gv *= inv;
          if (gv != 0.0) dy[a.od.op[o].out_idx] += gv * objective_grad(a.od.op[o], y);
        }
        if (has_cons) {
#pragma unroll
          for (int k = 0; k < 8; ++k)
            if (k < CL) gl += glf[(size_t)j * nt + k * nst + sl] * wk[k];
          log_feas_fat(a.od, y, gl * inv, dy);
        }
        for (int m = 0; m < M; ++m) dF[(size_t)m * df_stride + ((size_t)batch * q + j) * S + s] = dy[m];
      }
    }
  }
}

int launch_mc_loghvi_grad(const McArgs& a, double* dF, size_t df_stride, double* vals_ws, cudaStream_t st, LaunchCounter* lc) {
  if (a.b <= 0) return BO_OK;
  size_t smem = 0;
  const int nt = lh_pick_threads(a, true, &smem);
  if (!nt) { bo_set_error("mc_loghvi_grad: shared memory budget exceeded (n_b=%d q=%d)", a.nb, a.q); return BO_ERR_INVALID; }
  // few q-batches: MC samples split over CTAs and cells over cell lanes (needs the [b][S] value workspace)
  if (vals_ws && a.b < 148) {
    const int nsplit = std::max(1, std::min((296 + a.b - 1) / a.b, (a.S + 31) / 32));
    const int per_split = (a.S + nsplit - 1) / nsplit;
    int nst = std::min(256, ((per_split + 31) / 32) * 32);
    int CL = std::max(1, std::min(8, 256 / nst));
    const size_t fixed = (size_t)a.M * a.q * (a.nb + a.q) + (size_t)a.q * a.M;
    const size_t per_sample = (size_t)a.q * a.od.n_obj + a.q + (size_t)a.q * a.M;
    const size_t per_thread = (size_t)a.q * a.od.n_obj + a.q + 2;
    size_t sm2 = 0;
    for (;;) {
      sm2 = (fixed + per_sample * nst + per_thread * (size_t)nst * CL) * sizeof(double);
      if (sm2 <= 200 * 1024) break;
      if (CL > 1) CL >>= 1;
      else if (nst > 32) nst >>= 1;
      else { bo_set_error("mc_loghvi_grad: shared memory budget exceeded (n_b=%d q=%d)", a.nb, a.q); return BO_ERR_INVALID; }
    }
    static PerDeviceMax attr2_pd; size_t& attr2 = attr2_pd.slot();
    if (sm2 > 48 * 1024 && sm2 > attr2) {
      CUDA_CHECK_RET(cudaFuncSetAttribute(mc_loghvi_grad_cl_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2));
      attr2 = sm2;
    }
    dim3 grid(a.b, nsplit);
    mc_loghvi_grad_cl_kernel<<<grid, nst * CL, sm2, st>>>(a, dF, df_stride, vals_ws, CL);
    if (lc) lc->n++;
    CUDA_CHECK_RET(cudaGetLastError());
    loghvi_grad_finish_kernel<<<a.b, 256, 0, st>>>(a, vals_ws, dF, df_stride);
    if (lc) lc->n++;
    CUDA_CHECK_RET(cudaGetLastError());
    return BO_OK;
  }
  const int nsplit = 1;
  static PerDeviceMax attr_pd; size_t& attr = attr_pd.slot();
  if (smem > 48 * 1024 && smem > attr) {
    CUDA_CHECK_RET(cudaFuncSetAttribute(mc_loghvi_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  dim3 grid(a.b, nsplit);
  mc_loghvi_grad_kernel<<<grid, nt, smem, st>>>(a, dF, df_stride, nsplit > 1 ? vals_ws : nullptr);
  if (lc) lc->n++;
  CUDA_CHECK_RET(cudaGetLastError());
  if (nsplit > 1) {
    loghvi_grad_finish_kernel<<<a.b, 256, 0, st>>>(a, vals_ws, dF, df_stride);
    if (lc) lc->n++;
    CUDA_CHECK_RET(cudaGetLastError());
  }
  return BO_OK;
}

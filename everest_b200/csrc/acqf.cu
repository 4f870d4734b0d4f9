// K3 (small batched Cholesky / triangular solve), K4/K5 (MC reparameterisation + objective and
// constraint transforms), K6 (qNEHVI / qEHVI inclusion-exclusion), K7 (non-dominated front + box
// decomposition), K8 (qLogEI).
//
// Reference call sites: qNoisyExpectedHypervolumeImprovement built at
// strategies/predictives/qnehvi.py:39-52 and mobo.py:72-90, qExpectedHypervolumeImprovement at
// qehvi.py:67-76, qLogEI via get_acquisition_function at sobo.py:64-89; the arithmetic is BoTorch's
// (sample_cached_cholesky, _compute_qehvi, FastNondominatedPartitioning, log_fatplus/fatmax), restated
// in oracle/bo_oracle.py.  Objective / constraint formulas: utils/torch_tools.py:258-337,384-450.
#include "common.cuh"
#include <algorithm>

#include "acqf.cuh"
#include "mc_math.cuh"

#include <math.h>
#include <stdlib.h>


// ------------------------------------------------------------------------------------------------
// layout helpers
// ------------------------------------------------------------------------------------------------
// z[S, n, M] -> zT[(e*M + m)*S + s]   and   zM[m][s][ldn] (row-major per output, zero padded)
__global__ void transpose_base_samples_kernel(const double* __restrict__ z, int S, int n, int M, double* __restrict__ zT,
                                              double* __restrict__ zM, int ldn) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)S * n * M;
  if (idx >= total) return;
  int m = (int)(idx % M);
  long long r = idx / M;
  int e = (int)(r % n);
  int s = (int)(r / n);
  double v = z[idx];
  if (zT) zT[((size_t)e * M + m) * S + s] = v;
  if (zM) zM[((size_t)m * S + s) * ldn + e] = v;
}

int launch_transpose_base_samples(const double* z, int S, int n, int M, double* zT, double* zM, int ldn, cudaStream_t st,
                                  LaunchCounter* lc) {
  long long total = (long long)S * n * M;
  if (total <= 0) return BO_OK;
  transpose_base_samples_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(z, S, n, M, zT, zM, ldn);
  if (lc) lc->n++;
  CUDA_CHECK_RET(cudaGetLastError());
  return BO_OK;
}

__global__ void scale_matrix_kernel(double* __restrict__ A, int ld, int rows, int cols, double s) {
  int c = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
  if (r < rows && c < cols) A[(size_t)r * ld + c] *= s;
}
int launch_scale_matrix(double* A, int ld, int rows, int cols, double s, cudaStream_t st, LaunchCounter* lc) {
  if (rows <= 0 || cols <= 0) return BO_OK;
  dim3 grid((cols + 255) / 256, rows);
  scale_matrix_kernel<<<grid, 256, 0, st>>>(A, ld, rows, cols, s);
  if (lc) lc->n++;
  CUDA_CHECK_RET(cudaGetLastError());
  return BO_OK;
}

__global__ void add_diag_kernel(double* __restrict__ A, int ld, int n, double v) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) A[(size_t)i * ld + i] += v;
}
int launch_add_diag(double* A, int ld, int n, double v, cudaStream_t st, LaunchCounter* lc) {
  if (n <= 0) return BO_OK;
  add_diag_kernel<<<(n + 255) / 256, 256, 0, st>>>(A, ld, n, v);
  if (lc) lc->n++;
  CUDA_CHECK_RET(cudaGetLastError());
  return BO_OK;
}

// dst[r*ldd + c] = scale * src[r*lds + c]
__global__ void copy_scale_kernel(const double* __restrict__ src, int lds, double* __restrict__ dst, int ldd, int rows,
                                  int cols, double scale) {
  int c = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
  if (r < rows && c < cols) dst[(size_t)r * ldd + c] = scale * src[(size_t)r * lds + c];
}
int launch_copy_scale(const double* src, int lds, double* dst, int ldd, int rows, int cols, double scale, cudaStream_t st,
                      LaunchCounter* lc) {
  if (rows <= 0 || cols <= 0) return BO_OK;
  dim3 grid((cols + 255) / 256, rows);
  copy_scale_kernel<<<grid, 256, 0, st>>>(src, lds, dst, ldd, rows, cols, scale);
  if (lc) lc->n++;
  CUDA_CHECK_RET(cudaGetLastError());
  return BO_OK;
}

// out[i*ldo + m] = (mean_const + raw[i]) * y_std + y_mean
__global__ void finish_mean_kernel(const double* __restrict__ raw, int n, double mean_const, double y_std, double y_mean,
                                   double* __restrict__ out, int ldo, int m) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[(size_t)i * ldo + m] = (mean_const + raw[i]) * y_std + y_mean;
}
int launch_finish_mean(const double* raw, int n, double mean_const, double y_std, double y_mean, double* out, int ldo,
                       int m, cudaStream_t st, LaunchCounter* lc) {
  if (n <= 0) return BO_OK;
  finish_mean_kernel<<<(n + 255) / 256, 256, 0, st>>>(raw, n, mean_const, y_std, y_mean, out, ldo, m);
  if (lc) lc->n++;
  CUDA_CHECK_RET(cudaGetLastError());
  return BO_OK;
}

// var[i*ldo + m] = (k(x_i,x_i) - g[i] (+ noise)) * y_std^2
__global__ void finish_var_kernel(ModelD md, PrepD prep, const double* __restrict__ g, int n, int add_noise,
                                  double* __restrict__ out, int ldo, int m) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double kii = model_eval_pair(md, prep, i, prep, i, true);
  double v = kii - g[i];
  if (add_noise) v += md.noise;
  out[(size_t)i * ldo + m] = v * md.y_std * md.y_std;
}
int launch_finish_var(const ModelD& md, PrepD prep, const double* g, int n, int add_noise, double* out, int ldo, int m,
                      cudaStream_t st, LaunchCounter* lc) {
  if (n <= 0) return BO_OK;
  finish_var_kernel<<<(n + 127) / 128, 128, 0, st>>>(md, prep, g, n, add_noise, out, ldo, m);
  if (lc) lc->n++;
  CUDA_CHECK_RET(cudaGetLastError());
  return BO_OK;
}

// ------------------------------------------------------------------------------------------------
// cond_root: one warp per (q-batch, output) -- or, for large baselines (launcher: wpb = 4), the four warps of a CTA on one
// q-batch.  Sqq, Sqb -> bl = Sqb L_b^-T (product with the cached inverse root), br = psd_safe_chol(Sqq - bl bl^T) with the
// 1e-8..1e-3 jitter ladder, un-standardised mean.  root[(batch*M + m)][q][nb + q] = [bl | br].
// The work split does not change the order in which a sum is accumulated (lanes own columns e; a pair (i, j) of the Gram
// update is reduced by one warp; Cholesky entries are ordered dot products).  Baselines that do not fit shared memory
// form bl on the FP64 tensor pipe (k-blocks of 4 per DMMA) instead of one FMA chain per entry.
// ------------------------------------------------------------------------------------------------
template <bool wide>   // the q-batch's team: the whole CTA (launcher: wpb = 4) or one warp
__global__ void __launch_bounds__(128)
cond_root_kernel(CondRootArgs a) {
  extern __shared__ double csm[];
  const int warp_in_cta = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int batch = wide ? blockIdx.x : blockIdx.x * 4 + warp_in_cta;
  const int tl = wide ? (int)threadIdx.x : lane;                        // thread index / count within the q-batch's team
  constexpr int tn = wide ? 128 : 32;
  const int wsub = wide ? warp_in_cta : 0;                              // warp index / count within the team
  constexpr int nw = wide ? 4 : 1;
  const int q = a.q, nb = a.nb, nr = nb + q;
  const ModelD& md = a.md;
  // per-team scratch, then (optionally) the CTA-shared copy of L_b^-1
  const int blw = max(nb, a.stage_cols);                               // columns of the BLs scratch (bl, or staged rows)
  const size_t per_team = (size_t)q * nb + 2 * q * q + (size_t)q * blw;
  double* Sqb = csm + (wide ? 0 : (size_t)warp_in_cta * per_team);      // [q][nb]
  double* Sc = Sqb + q * nb;                                            // [q][q]
  double* Lq = Sc + q * q;                                              // [q][q]
  double* BLs = Lq + q * q;                                             // [q][nb] bl
  const double* Li = a.LbInv;
  int ldli = a.ldlb;
  if (a.linv_in_smem) {
    double* Ls = csm + (size_t)(wide ? 1 : 4) * per_team;
    ldli = nb | 1;  // odd stride: lanes (rows e) hit distinct banks
    for (int idx = threadIdx.x; idx < nb * nb; idx += blockDim.x)
      Ls[(idx / nb) * ldli + (idx % nb)] = a.LbInv[(size_t)(idx / nb) * a.ldlb + (idx % nb)];
    Li = Ls;
  }
  __syncthreads();
  if (batch >= a.b) return;   // (wide: the whole CTA leaves together)
#define TEAM_SYNC() do { if (wide) __syncthreads(); else __syncwarp(); } while (0)
  const double s2 = md.y_std * md.y_std;
  const int row0 = batch * q;

  const bool one_cont_leaf = md.n_terms == 1 && md.nfac[0] == 1 && md.leaf[md.fac[0][0]].kind <= BO_LEAF_MATERN52 &&
                             q <= 8 && md.leaf[md.fac[0][0]].dpad <= 64;
  if (one_cont_leaf) {
    // SingleTaskGP default (one RBF / Matern leaf): the q candidate rows of this batch sit in shared memory and every
    // baseline row is read from global memory ONCE for all q kernel values (model_eval_pair re-reads both rows per value).
    // Same arithmetic in the same order as leaf_eval_pair -> bit-identical results.
    const int l = md.fac[0][0];
    const LeafD& L = md.leaf[l];
    double* aq = BLs;                       // [q][dpad] staging, overwritten by bl afterwards (q * dpad <= q * nb is not
                                            // guaranteed: the launcher sizes BLs for max(nb, dpad) columns)
    for (int idx = tl; idx < q * L.dpad; idx += tn) aq[idx] = a.prep_q.Xs[l][(size_t)row0 * L.dpad + idx];
    TEAM_SYNC();
    for (int e = tl; e < nb; e += tn) {
      const double* bx = a.prep_b.Xs[l] + (size_t)e * L.dpad;
      double dot[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) dot[j] = 0.0;
      for (int k = 0; k < L.nd; ++k) {
        const double bk = bx[k];
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (j < q) dot[j] = fma(aq[j * L.dpad + k], bk, dot[j]);
      }
      const double n2b = a.prep_b.n2[l][e];
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (j < q) {
          const double stat = fmax(a.prep_q.n2[l][row0 + j] + n2b - 2.0 * dot[j], 0.0);
          double kqb = 0.0;
          kqb += md.coef[0] * leaf_value_from_stat(L.kind, stat);
          Sqb[j * nb + e] = (kqb - a.W[(size_t)(row0 + j) * a.ldw + e]) * s2;
        }
    }
    TEAM_SYNC();
  } else {
    for (int j = 0; j < q; ++j)
      for (int e = tl; e < nb; e += tn) {
        double kqb = model_eval_pair(md, a.prep_q, row0 + j, a.prep_b, e, false);
        Sqb[j * nb + e] = (kqb - a.W[(size_t)(row0 + j) * a.ldw + e]) * s2;
      }
  }
  for (int p = tl; p < q * q; p += tn) {
    int i = p / q, j = p % q;
    if (j >= i) {
      double kqq = model_eval_pair(md, a.prep_q, row0 + i, a.prep_q, row0 + j, i == j);
      double v = (kqq - a.Gqq[((size_t)batch * q + i) * q + j]) * s2;
      Sc[i * q + j] = v;
      Sc[j * q + i] = v;
    }
  }
  TEAM_SYNC();
  // bl = Sqb L_b^-T with the cached inverse root: bl[j][e] = sum_{l<=e} Sqb[j][l] LbInv[e][l]  (lanes over e)
  if (a.linv_in_smem) {
    for (int e = tl; e < nb; e += tn) {
      const double* lrow = Li + (size_t)e * ldli;
      if (q <= 8) {
        // one pass over row e of L_b^-1 for all q points (same summation order per point)
        double acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = 0.0;
        for (int l = 0; l <= e; ++l) {
          const double lv = lrow[l];
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (j < q) acc[j] = fma(Sqb[j * nb + l], lv, acc[j]);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (j < q) BLs[j * nb + e] = acc[j];
      } else {
        for (int j = 0; j < q; ++j) {
          double s = 0.0;
          for (int l = 0; l <= e; ++l) s = fma(Sqb[j * nb + l], lrow[l], s);
          BLs[j * nb + e] = s;
        }
      }
    }
  } else {
    // large baselines: bl = Sqb L_b^-T on the FP64 tensor pipe (DMMA m8n8k4), straight from the TRANSPOSED inverse in
    // L2: a warp owns blocks of 32 columns e (four n-tiles), A = up to 8 points of the q-batch x 4 rows l from shared
    // memory, B[k = l][n = e] = LbInvT[l][e].  Entries with l > e are exact zeros of the triangular factor, so a block
    // stops at its last column.  (The scalar loop this replaces issued 8 LDS + 8 DFMA + 1 LDG per l and was bound by
    // instruction issue: 0.31 ms per output for 1024 q-batches at n_b = 286.)
    const int g = lane >> 2, t = lane & 3;
    for (int e0 = wsub * 32; e0 < nb; e0 += 32 * nw) {
      const int lmax = min(nb, e0 + 32);
      for (int j0 = 0; j0 < q; j0 += 8) {
        double acc[4][2];
#pragma unroll
        for (int n = 0; n < 4; ++n) acc[n][0] = acc[n][1] = 0.0;
        const bool arow = (j0 + g) < q;
        const double* sq = Sqb + (size_t)(arow ? j0 + g : 0) * nb;
#pragma unroll 4
        for (int l = 0; l < lmax; l += 4) {
          const bool kin = (l + t) < lmax;
          const double av = (arow && kin) ? sq[l + t] : 0.0;
          const double* lrow = a.LbInvT + (size_t)(kin ? l + t : 0) * a.ldlb;
          double bv[4];
#pragma unroll
          for (int n = 0; n < 4; ++n) {
            const int e = e0 + n * 8 + g;
            bv[n] = (kin && e < nb) ? lrow[e] : 0.0;
          }
#pragma unroll
          for (int n = 0; n < 4; ++n) mma_884(acc[n][0], acc[n][1], av, bv[n]);
        }
        if (arow) {
#pragma unroll
          for (int n = 0; n < 4; ++n)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int e = e0 + n * 8 + 2 * t + h;
              if (e < nb) BLs[(size_t)(j0 + g) * nb + e] = acc[n][h];
            }
        }
      }
    }
  }
  TEAM_SYNC();
  // Sc -= bl bl^T  (one warp per pair)
  for (int p = wsub; p < q * q; p += nw) {
    int i = p / q, j = p % q;
    if (j < i) continue;
    double s = 0.0;
    for (int e = lane; e < nb; e += 32) s = fma(BLs[i * nb + e], BLs[j * nb + e], s);
    s = warp_sum(s);
    if (lane == 0) {
      Sc[i * q + j] -= s;
      if (i != j) Sc[j * q + i] -= s;
    }
  }
  TEAM_SYNC();
  // psd_safe_cholesky on q x q, jitter ladder 1e-8 * 10^i, i = 0..5, added incrementally.  One warp, column by column:
  // every lane forms the pivot (same value, uniform control flow), the rows below it are spread over the lanes; each
  // entry is the ordered sum a single thread would form.
  int fail = 0;
  double jitter_total = 0.0;
  if (wsub == 0) {
    double prev = 0.0;
    for (int attempt = 0; attempt <= 6; ++attempt) {
      if (attempt > 0) {
        const double ladder[6] = {1.0, 10.0, 100.0, 1000.0, 10000.0, 100000.0};
        double nw2 = 1e-8 * ladder[attempt - 1];
        for (int i = lane; i < q; i += 32) Sc[i * q + i] += (nw2 - prev);
        jitter_total += (nw2 - prev);
        prev = nw2;
        __syncwarp();
      }
      fail = 0;
      for (int j = 0; j < q; ++j) {
        double d = Sc[j * q + j];
        for (int l = 0; l < j; ++l) d -= Lq[j * q + l] * Lq[j * q + l];
        if (!(d > 0.0)) { fail = 1; break; }
        d = sqrt(d);
        for (int i = j + 1 + lane; i < q; i += 32) {
          double s = Sc[i * q + j];
          for (int l = 0; l < j; ++l) s -= Lq[i * q + l] * Lq[j * q + l];
          Lq[i * q + j] = s / d;
        }
        if (lane == 0) Lq[j * q + j] = d;
        for (int i = lane; i < j; i += 32) Lq[i * q + j] = 0.0;
        __syncwarp();
      }
      if (!fail) break;
    }
    if (fail) {
      const double nanv = __longlong_as_double(0x7ff8000000000000LL);
      __syncwarp();
      for (int p = lane; p < q * q; p += 32) Lq[p] = nanv;
    }
    if (lane == 0) {
      if (a.info) a.info[(size_t)batch * a.M + a.m] = fail;
      if (a.jitter) a.jitter[(size_t)batch * a.M + a.m] = jitter_total;
    }
  }
  TEAM_SYNC();
  double* root = a.root + ((size_t)batch * a.M + a.m) * q * nr;
  for (int idx = tl; idx < q * nr; idx += tn) {
    int j = idx / nr, c = idx % nr;
    root[idx] = (c < nb) ? BLs[j * nb + c] : Lq[j * q + (c - nb)];
  }
  if (a.BL) {
    for (int idx = tl; idx < q * a.ldbl; idx += tn) {
      int j = idx / a.ldbl, c = idx % a.ldbl;
      a.BL[(size_t)(row0 + j) * a.ldbl + c] = (c < nb) ? BLs[j * nb + c] : 0.0;
    }
  }
  for (int j = tl; j < q; j += tn)
    a.mu[((size_t)(row0 + j)) * a.M + a.m] = (md.mean_const + a.mu_raw[row0 + j]) * md.y_std + md.y_mean;
#undef TEAM_SYNC
}

int launch_cond_root(const CondRootArgs& a0, cudaStream_t st, LaunchCounter* lc) {
  if (a0.b <= 0) return BO_OK;
  CondRootArgs a = a0;
  // the staging of the q candidate rows (single continuous leaf) shares the bl scratch: size it for max(nb, dpad) columns
  a.stage_cols = 0;
  if (a.md.n_terms == 1 && a.md.nfac[0] == 1 && a.md.leaf[a.md.fac[0][0]].kind <= BO_LEAF_MATERN52)
    a.stage_cols = a.md.leaf[a.md.fac[0][0]].dpad;
  const int blw = std::max(a.nb, a.stage_cols);
  size_t per_team = ((size_t)a.q * a.nb + 2 * (size_t)a.q * a.q + (size_t)a.q * blw) * sizeof(double);
  // large baselines: one q-batch per CTA (its four warps share the columns), so that the scratch of four q-batches does
  // not hold the SM at one resident CTA (n_b = 286, q = 8: 150 KB per CTA before, 37 KB now)
  a.wpb = (a.nb > 64) ? 4 : 1;
  size_t smem = (a.wpb == 4 ? 1 : 4) * per_team;
  size_t linv = (size_t)a.nb * (a.nb | 1) * sizeof(double);
  a.linv_in_smem = (a.nb > 0 && a.wpb == 1 && smem + linv <= 96 * 1024) ? 1 : 0;
  if (a.linv_in_smem) smem += linv;
  if (smem > 200 * 1024) { bo_set_error("cond_root: baseline too large for shared memory (n_b=%d)", a.nb); return BO_ERR_INVALID; }
  static PerDeviceMax attr_pd[2]; size_t& attr = attr_pd[a.wpb == 4].slot();
  if (smem > 48 * 1024 && smem > attr) {
    if (a.wpb == 4) CUDA_CHECK_RET(cudaFuncSetAttribute(cond_root_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    else CUDA_CHECK_RET(cudaFuncSetAttribute(cond_root_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  if (a.wpb == 4) cond_root_kernel<true><<<a.b, 128, smem, st>>>(a);
  else cond_root_kernel<false><<<(a.b + 3) / 4, 128, smem, st>>>(a);
  if (lc) lc->n++;
  CUDA_CHECK_RET(cudaGetLastError());
  return BO_OK;
}

// ------------------------------------------------------------------------------------------------
// objective transform of a sample matrix F[m][s][ldf] + mean -> obj[(s*n + e)*n_obj + o], feasibility
// ------------------------------------------------------------------------------------------------
__global__ void baseline_objective_kernel(const double* __restrict__ F, int ldf, int S, int n, int M,
                                          const double* __restrict__ mean /*[n, M]*/, ObjD od,
                                          double* __restrict__ obj, unsigned char* __restrict__ feas,
                                          double* __restrict__ samples /* optional [S, n, M] */) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)S * n) return;
  int e = (int)(idx % n), s = (int)(idx / n);
  double y[BO_MAX_OBJECTIVES * 2];
  for (int m = 0; m < M; ++m) {
    y[m] = mean[(size_t)e * M + m] + F[((size_t)m * S + s) * ldf + e];
    if (samples) samples[((size_t)s * n + e) * M + m] = y[m];
  }
  for (int o = 0; o < od.n_obj; ++o) obj[((size_t)s * n + e) * od.n_obj + o] = objective_apply(od.op[o], y);
  unsigned char ok = 1;
  for (int c = 0; c < od.n_cons; ++c) {
    double cv = od.con[c].sign * (y[od.con[c].out_idx] - od.con[c].tp);
    if (!(cv <= 0.0)) ok = 0;
  }
  feas[(size_t)s * n + e] = ok;
}
int launch_baseline_objective(const double* F, int ldf, int S, int n, int M, const double* mean, const ObjD& od,
                              double* obj, unsigned char* feas, double* samples, cudaStream_t st, LaunchCounter* lc) {
  long long total = (long long)S * n;
  if (total <= 0) return BO_OK;
  baseline_objective_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(F, ldf, S, n, M, mean, od, obj, feas, samples);
  if (lc) lc->n++;
  CUDA_CHECK_RET(cudaGetLastError());
  return BO_OK;
}

// ------------------------------------------------------------------------------------------------
// K7: per-sample non-dominated front.  One CTA per sample.  front[s][e] = 1 iff point e is feasible,
// strictly better than ref in every objective, not dominated, and (dedup) the first of its duplicates.
// With count_only (pruning, BoTorch deduplicate=False) duplicates are all kept and counts[e] += 1.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
front_kernel(const double* __restrict__ obj, const unsigned char* __restrict__ feas, int n, int Mo,
             const double* __restrict__ ref, int dedup, unsigned char* __restrict__ front, int* __restrict__ counts,
             int infeasible_to_ref) {
  const int s = blockIdx.x;
  const double* Y = obj + (size_t)s * n * Mo;
  const unsigned char* fz = feas + (size_t)s * n;
  double r[BO_MAX_OBJECTIVES];
  for (int o = 0; o < Mo; ++o) r[o] = ref[o];
  for (int e = threadIdx.x; e < n; e += blockDim.x) {
    double y[BO_MAX_OBJECTIVES];
    bool valid = fz[e] != 0;
    for (int o = 0; o < Mo; ++o) {
      y[o] = Y[(size_t)e * Mo + o];
      if (!(y[o] > r[o])) valid = false;
    }
    bool keep = valid;
    if (valid) {
      for (int k = 0; k < n && keep; ++k) {
        if (k == e) continue;
        bool kvalid = fz[k] != 0;
        if (!kvalid && !infeasible_to_ref) continue;
        // pruning semantics: infeasible points sit AT the ref point -> they can never dominate a valid point
        if (!kvalid) continue;
        bool ge = true, gt = false, eq = true;
        for (int o = 0; o < Mo; ++o) {
          double v = Y[(size_t)k * Mo + o];
          ge = ge && (v >= y[o]);
          gt = gt || (v > y[o]);
          eq = eq && (v == y[o]);
        }
        if (ge && gt) keep = false;
        if (dedup && eq && k < e) keep = false;
      }
    }
    if (front) front[(size_t)s * n + e] = keep ? 1 : 0;
    if (counts && keep) atomicAdd(&counts[e], 1);
  }
}
int launch_front(const double* obj, const unsigned char* feas, int S, int n, int Mo, const double* ref_dev, int dedup,
                 unsigned char* front, int* counts, cudaStream_t st, LaunchCounter* lc) {
  if (S <= 0 || n <= 0) return BO_OK;
  front_kernel<<<S, 256, 0, st>>>(obj, feas, n, Mo, ref_dev, dedup, front, counts, 1);
  if (lc) lc->n++;
  CUDA_CHECK_RET(cudaGetLastError());
  return BO_OK;
}

// 2-objective box decomposition of the non-dominated space.  Front sorted by objective 0 ascending
// (rank by counting, ties by row); cell c: lower = (y0[c-1] | ref0, y1[c] | ref1), upper = (y0[c] | inf, inf).
// Cells are stored sample-minor for coalesced reads: lo/up[(c*Mo + o)*S + s].
__global__ void __launch_bounds__(256)
partition2d_kernel(const double* __restrict__ obj, const unsigned char* __restrict__ front, int n, int S, int cap,
                   const double* __restrict__ ref, double* __restrict__ lo, double* __restrict__ up,
                   int* __restrict__ ncells, int* __restrict__ front_idx /* [S, cap] sorted row indices */) {
  const int s = blockIdx.x;
  const double* Y = obj + (size_t)s * n * 2;
  const unsigned char* fr = front + (size_t)s * n;
  __shared__ int count;
  if (threadIdx.x == 0) count = 0;
  __syncthreads();
  const double inf = __longlong_as_double(0x7ff0000000000000LL);
  int local = 0;
  for (int e = threadIdx.x; e < n; e += blockDim.x) {
    if (!fr[e]) continue;
    ++local;
    double y0 = Y[e * 2], y1 = Y[e * 2 + 1];
    int rank = 0;
    for (int k = 0; k < n; ++k) {
      if (!fr[k]) continue;
      double v = Y[k * 2];
      if (v < y0 || (v == y0 && k < e)) ++rank;
    }
    if (rank + 1 < cap) {
      lo[((size_t)(rank + 1) * 2 + 0) * S + s] = y0;
      up[((size_t)(rank + 1) * 2 + 1) * S + s] = inf;
    }
    if (rank < cap) {
      lo[((size_t)rank * 2 + 1) * S + s] = y1;
      up[((size_t)rank * 2 + 0) * S + s] = y0;
      if (front_idx) front_idx[(size_t)s * cap + rank] = e;
    }
  }
  if (local) atomicAdd(&count, local);
  __syncthreads();
  if (threadIdx.x == 0) {
    int p = count;
    ncells[s] = p + 1;
    lo[((size_t)0 * 2 + 0) * S + s] = ref[0];
    up[((size_t)0 * 2 + 1) * S + s] = inf;
    if (p < cap) {
      lo[((size_t)p * 2 + 1) * S + s] = ref[1];
      up[((size_t)p * 2 + 0) * S + s] = inf;
    }
  }
}
int launch_partition2d(const double* obj, const unsigned char* front, int n, int S, int cap, const double* ref_dev,
                       double* lo, double* up, int* ncells, int* front_idx, cudaStream_t st, LaunchCounter* lc) {
  partition2d_kernel<<<S, 256, 0, st>>>(obj, front, n, S, cap, ref_dev, lo, up, ncells, front_idx);
  if (lc) lc->n++;
  CUDA_CHECK_RET(cudaGetLastError());
  return BO_OK;
}

// Block-wide exclusive scan of one int per thread (256 threads); returns the exclusive prefix and the
// block total through *total.  Keeps thread order, so compactions built on it are stable.
__device__ __forceinline__ int block_excl_scan(int v, int* total, int* wsum /* [9] shared */) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  __syncthreads();
  if (lane == 31) wsum[w] = inc;
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0;
    for (int i = 0; i < 8; ++i) { int t = wsum[i]; wsum[i] = run; run += t; }
    wsum[8] = run;
  }
  __syncthreads();
  *total = wsum[8];
  return wsum[w] + inc - v;
}

// General (m >= 2) box decomposition of the non-dominated space via local upper bounds
// (Lacour et al. 2017, Alg. 3 + Eq. 2), one CTA per MC sample.  Front points are inserted in row
// order; kept bounds stay in order and new bounds are appended in (dominated-u, k) order, exactly
// like the list operations of the CPU restatement, so the cell list is bit-identical.
// Work buffers per sample: two ping-pong copies of U[cap][m] and Z[cap][m][m] (minimisation frame).
__global__ void __launch_bounds__(256)
partition_nd_kernel(const double* __restrict__ obj, const unsigned char* __restrict__ front, int n, int S, int Mo,
                    int cap, const double* __restrict__ ref, double* __restrict__ work, double* __restrict__ lo,
                    double* __restrict__ up, int* __restrict__ ncells, int* __restrict__ overflow) {
  const int s = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  const int mm = Mo * Mo;
  double* base = work + (size_t)s * 2 * cap * (Mo + mm);
  double* Ub[2] = {base, base + (size_t)cap * (Mo + mm)};
  double* Zb[2] = {Ub[0] + (size_t)cap * Mo, Ub[1] + (size_t)cap * Mo};
  __shared__ int wsum[9];
  __shared__ int sh_count, sh_over;
  const double ninf = __longlong_as_double(0xfff0000000000000LL);
  const double pinf = __longlong_as_double(0x7ff0000000000000LL);
  int cur = 0;
  if (tid == 0) {
    for (int j = 0; j < Mo; ++j) {
      Ub[0][j] = -ref[j];
      for (int i = 0; i < Mo; ++i) Zb[0][i * Mo + j] = (i == j) ? -ref[j] : ninf;
    }
    sh_count = 1;
    sh_over = 0;
  }
  __syncthreads();
  for (int e = 0; e < n; ++e) {
    if (!front[(size_t)s * n + e]) continue;
    double z[BO_MAX_OBJECTIVES];
    for (int j = 0; j < Mo; ++j) z[j] = -obj[((size_t)s * n + e) * Mo + j];
    const int count = sh_count;
    const double* U = Ub[cur];
    const double* Z = Zb[cur];
    double* U2 = Ub[cur ^ 1];
    double* Z2 = Zb[cur ^ 1];
    // pass A: how many bounds survive
    int kept_total = 0;
    for (int b0 = 0; b0 < count; b0 += nt) {
      int u = b0 + tid;
      int keep = 0;
      if (u < count) {
        bool dom = true;
        for (int j = 0; j < Mo; ++j) dom = dom && (U[(size_t)u * Mo + j] > z[j]);
        keep = dom ? 0 : 1;
      }
      int tot;
      block_excl_scan(keep, &tot, wsum);
      kept_total += tot;
    }
    // pass B: stable compaction of survivors, then the children of every dominated bound
    int kept_run = 0, new_run = 0;
    for (int b0 = 0; b0 < count; b0 += nt) {
      int u = b0 + tid;
      int keep = 0, newc = 0;
      unsigned newmask = 0;
      if (u < count) {
        bool dom = true;
        for (int j = 0; j < Mo; ++j) dom = dom && (U[(size_t)u * Mo + j] > z[j]);
        keep = dom ? 0 : 1;
        if (dom) {
          for (int k = 0; k < Mo; ++k) {
            double lower = ninf;
            for (int i = 0; i < Mo; ++i)
              if (i != k) lower = fmax(lower, Z[(size_t)u * mm + i * Mo + k]);
            if (z[k] >= lower) { newmask |= (1u << k); ++newc; }
          }
        }
      }
      int tk, tn;
      int koff = block_excl_scan(keep, &tk, wsum);
      int noff = block_excl_scan(newc, &tn, wsum);
      if (keep) {
        int dst = kept_run + koff;
        if (dst < cap) {
          for (int j = 0; j < Mo; ++j) U2[(size_t)dst * Mo + j] = U[(size_t)u * Mo + j];
          for (int j = 0; j < mm; ++j) Z2[(size_t)dst * mm + j] = Z[(size_t)u * mm + j];
        }
      }
      if (newc) {
        int dst = kept_total + new_run + noff;
        for (int k = 0; k < Mo; ++k) {
          if (!(newmask & (1u << k))) continue;
          if (dst < cap) {
            for (int j = 0; j < Mo; ++j) U2[(size_t)dst * Mo + j] = (j == k) ? z[k] : U[(size_t)u * Mo + j];
            for (int i = 0; i < Mo; ++i)
              for (int j = 0; j < Mo; ++j)
                Z2[(size_t)dst * mm + i * Mo + j] = (i == k) ? z[j] : Z[(size_t)u * mm + i * Mo + j];
          }
          ++dst;
        }
      }
      kept_run += tk;
      new_run += tn;
    }
    __syncthreads();
    if (tid == 0) {
      int nc = kept_total + new_run;
      if (nc > cap) { sh_over = 1; nc = cap; }
      sh_count = nc;
    }
    __syncthreads();
    cur ^= 1;
    if (sh_over) break;
  }
  if (sh_over) {
    if (tid == 0) { atomicExch(overflow, 1); ncells[s] = 0; }
    return;
  }
  // cell bounds (Eq. 2), dropping empty cells, written sample-minor in the maximisation frame
  const int count = sh_count;
  const double* U = Ub[cur];
  const double* Z = Zb[cur];
  int run = 0;
  for (int b0 = 0; b0 < count; b0 += nt) {
    int u = b0 + tid;
    int ok = 0;
    double lmin[BO_MAX_OBJECTIVES], umin[BO_MAX_OBJECTIVES];
    if (u < count) {
      ok = 1;
      for (int j = 0; j < Mo; ++j) {
        umin[j] = U[(size_t)u * Mo + j];
        double l = ninf;
        for (int k = 0; k < j; ++k) l = fmax(l, Z[(size_t)u * mm + k * Mo + j]);
        lmin[j] = l;
        if (umin[j] <= lmin[j]) ok = 0;
      }
    }
    int tot;
    int off = block_excl_scan(ok, &tot, wsum);
    if (ok) {
      int c = run + off;
      for (int j = 0; j < Mo; ++j) {
        lo[((size_t)c * Mo + j) * S + s] = -umin[j];
        up[((size_t)c * Mo + j) * S + s] = (lmin[j] == ninf) ? pinf : -lmin[j];
      }
    }
    run += tot;
  }
  if (tid == 0) ncells[s] = run;
}
int launch_partition_nd(const double* obj, const unsigned char* front, int n, int S, int Mo, int cap,
                        const double* ref_dev, double* work, double* lo, double* up, int* ncells, int* overflow,
                        cudaStream_t st, LaunchCounter* lc) {
  partition_nd_kernel<<<S, 256, 0, st>>>(obj, front, n, S, Mo, cap, ref_dev, work, lo, up, ncells, overflow);
  if (lc) lc->n++;
  CUDA_CHECK_RET(cudaGetLastError());
  return BO_OK;
}

// Approximate / exact box decomposition by BINARY PARTITIONING: [UPSTREAM] botorch NondominatedPartitioning._partition_space
// (Couckuyt et al. 2012), what qNEHVI builds per MC sample when BoFire passes alpha > 0 and there are more than two
// objectives (data_models/strategies/predictives/qnehvi.py:19, strategies/predictives/qnehvi.py:50).  Cells are pairs of
// index vectors into the per-objective sorted (augmented) front; a cell that lies entirely in the non-dominated region is
// kept, a cell that straddles the front is halved along its longest index edge -- unless its share of the volume between
// the ideal and the anti-ideal point is <= alpha, in which case it is DROPPED (the approximation), and cells inside the
// dominated region are dropped.  One warp per MC sample: lane 0 owns the explicit LIFO stack, the 32 lanes share the
// any / all tests against the front points; cells are emitted in the order the reference's list-based loop appends them.
// Work buffer per sample: negY [n][Mo] doubles, aug [n + 2][Mo] ints, stack [PB_STACK][2][Mo] ints.
#define PB_STACK 256
__global__ void __launch_bounds__(32)
partition_binary_kernel(const double* __restrict__ obj, const unsigned char* __restrict__ front, int n, int S, int Mo, int cap,
                        const double* __restrict__ ref, double alpha, double* __restrict__ work, size_t work_stride,
                        double* __restrict__ lo, double* __restrict__ up, int* __restrict__ ncells, int* __restrict__ overflow) {
  const int s = blockIdx.x, lane = threadIdx.x;
  char* base = reinterpret_cast<char*>(work) + (size_t)s * work_stride;
  double* negY = reinterpret_cast<double*>(base);                       // [p][Mo]
  int* aug = reinterpret_cast<int*>(negY + (size_t)n * Mo);            // [p + 2][Mo]: row of the augmented front per rank
  int* stack = aug + (size_t)(n + 2) * Mo;                              // [PB_STACK][2][Mo]
  const double pinf = __longlong_as_double(0x7ff0000000000000LL);
  // front points of this sample in row order (minimisation frame)
  int p = 0;
  for (int e0 = 0; e0 < n; e0 += 32) {
    const int e = e0 + lane;
    const bool f = e < n && front[(size_t)s * n + e];
    const unsigned m = __ballot_sync(0xffffffffu, f);
    if (f) {
      const int dst = p + __popc(m & ((1u << lane) - 1));
      for (int j = 0; j < Mo; ++j) negY[(size_t)dst * Mo + j] = -obj[((size_t)s * n + e) * Mo + j];
    }
    p += __popc(m);
  }
  __syncwarp();
  if (p == 0) {   // empty front: one cell over the whole non-dominated space
    if (lane == 0) {
      for (int j = 0; j < Mo; ++j) { lo[((size_t)0 * Mo + j) * S + s] = ref[j]; up[((size_t)0 * Mo + j) * S + s] = pinf; }
      ncells[s] = 1;
    }
    return;
  }
  // argsort of every objective (stable rank): aug[1 + rank][j] = row + 1; aug[0] = 0 (ideal), aug[p + 1] = p + 1 (anti-ideal)
  for (int idx = lane; idx < p * Mo; idx += 32) {
    const int i = idx / Mo, j = idx - i * Mo;
    const double v = negY[(size_t)i * Mo + j];
    int rank = 0;
    for (int k = 0; k < p; ++k) {
      const double w = negY[(size_t)k * Mo + j];
      rank += (w < v) || (w == v && k < i);
    }
    aug[(size_t)(1 + rank) * Mo + j] = i + 1;
  }
  for (int j = lane; j < Mo; j += 32) { aug[j] = 0; aug[(size_t)(p + 1) * Mo + j] = p + 1; }
  __syncwarp();
  double ideal[BO_MAX_OBJECTIVES], anti[BO_MAX_OBJECTIVES];
  double total = 1.0;
  for (int j = 0; j < Mo; ++j) {
    ideal[j] = negY[(size_t)(aug[(size_t)1 * Mo + j] - 1) * Mo + j] - 1.0;
    anti[j] = negY[(size_t)(aug[(size_t)p * Mo + j] - 1) * Mo + j] + 1.0;
    total *= (anti[j] - ideal[j]);
  }
  int top = 0, count = 0;
  bool over = false;
  if (lane == 0) {
    for (int j = 0; j < Mo; ++j) { stack[j] = 0; stack[Mo + j] = p + 1; }
  }
  top = 1;
  __syncwarp();
  while (top > 0) {
    --top;
    int c0[BO_MAX_OBJECTIVES], c1[BO_MAX_OBJECTIVES], r0[BO_MAX_OBJECTIVES], r1[BO_MAX_OBJECTIVES];
    double v0[BO_MAX_OBJECTIVES], v1[BO_MAX_OBJECTIVES];
    for (int j = 0; j < Mo; ++j) {
      c0[j] = stack[(size_t)top * 2 * Mo + j];
      c1[j] = stack[(size_t)top * 2 * Mo + Mo + j];
      r0[j] = aug[(size_t)c0[j] * Mo + j];
      r1[j] = aug[(size_t)c1[j] * Mo + j];
      v0[j] = (r0[j] == 0) ? ideal[j] : ((r0[j] == p + 1) ? anti[j] : negY[(size_t)(r0[j] - 1) * Mo + j]);
      v1[j] = (r1[j] == 0) ? ideal[j] : ((r1[j] == p + 1) ? anti[j] : negY[(size_t)(r1[j] - 1) * Mo + j]);
    }
    __syncwarp();
    // upper corner better than or equal to every front point in at least one objective: entirely non-dominated
    bool all_u = true, all_l = true;
    for (int k = lane; k < p; k += 32) {
      bool any_u = false, any_l = false;
      for (int j = 0; j < Mo; ++j) {
        const double y = negY[(size_t)k * Mo + j];
        any_u = any_u || (v1[j] <= y);
        any_l = any_l || (v0[j] <= y);
      }
      all_u = all_u && any_u;
      all_l = all_l && any_l;
    }
    all_u = __all_sync(0xffffffffu, all_u);
    all_l = __all_sync(0xffffffffu, all_l);
    if (all_u) {
      if (count < cap) {
        for (int j = lane; j < Mo; j += 32) {
          // minimisation bounds [aug'(r0), aug'(r1)] with aug' = (-inf, front, -ref); maximisation = negated and swapped
          lo[((size_t)count * Mo + j) * S + s] = (r1[j] == p + 1) ? ref[j] : -negY[(size_t)(r1[j] - 1) * Mo + j];
          up[((size_t)count * Mo + j) * S + s] = (r0[j] == 0) ? pinf : -negY[(size_t)(r0[j] - 1) * Mo + j];
        }
      } else over = true;
      ++count;
    } else if (all_l) {
      int length = 0, longest = 0;
      double vol = 1.0;
      for (int j = 0; j < Mo; ++j) {
        const int dj = c1[j] - c0[j];
        if (dj > length) { length = dj; longest = j; }      // first maximum, like torch.max
        vol *= (v1[j] - v0[j]);
      }
      if (length > 1 && (vol / total) > alpha) {
        // Python's round(length / 2.0): half to even
        int n1 = length / 2;
        if (length & 1) n1 += ((n1 & 1) ? 1 : 0);
        const int n2 = length - n1;
        if (top + 2 > PB_STACK) { over = true; break; }
        if (lane == 0) {
          int* a = stack + (size_t)top * 2 * Mo;        // cell 1: upper bound lowered by n1 (pushed first, popped second)
          int* b = stack + (size_t)(top + 1) * 2 * Mo;  // cell 2: lower bound raised by n2
          for (int j = 0; j < Mo; ++j) {
            a[j] = c0[j]; a[Mo + j] = c1[j] - (j == longest ? n1 : 0);
            b[j] = c0[j] + (j == longest ? n2 : 0); b[Mo + j] = c1[j];
          }
        }
        top += 2;
        __syncwarp();
      }
    }
  }
  if (lane == 0) {
    if (over) { atomicExch(overflow, 1); ncells[s] = 0; }
    else ncells[s] = count;
  }
}
size_t partition_binary_work_stride(int n, int Mo) {
  size_t b = (size_t)std::max(n, 1) * Mo * sizeof(double) + ((size_t)(n + 2) * Mo + (size_t)PB_STACK * 2 * Mo) * sizeof(int);
  return (b + 15) / 16 * 16;
}
int launch_partition_binary(const double* obj, const unsigned char* front, int n, int S, int Mo, int cap, const double* ref_dev,
                            double alpha, double* work, double* lo, double* up, int* ncells, int* overflow, cudaStream_t st,
                            LaunchCounter* lc) {
  partition_binary_kernel<<<S, 32, 0, st>>>(obj, front, n, S, Mo, cap, ref_dev, alpha, work, partition_binary_work_stride(n, Mo),
                                             lo, up, ncells, overflow);
  if (lc) lc->n++;
  CUDA_CHECK_RET(cudaGetLastError());
  return BO_OK;
}

__global__ void add_inplace_kernel(double* __restrict__ dst, const double* __restrict__ src, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] += src[i];
}
int launch_add_inplace(double* dst, const double* src, size_t n, cudaStream_t st, LaunchCounter* lc) {
  if (n == 0) return BO_OK;
  add_inplace_kernel<<<(unsigned)std::min<size_t>((n + 255) / 256, 1184), 256, 0, st>>>(dst, src, n);
  if (lc) lc->n++;
  CUDA_CHECK_RET(cudaGetLastError());
  return BO_OK;
}

__global__ void count_nonzero_kernel(const int* __restrict__ v, int n, int* __restrict__ count) {
  int c = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) c += v[i] != 0;
  c = __reduce_add_sync(0xffffffffu, c);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(count, c);
}
int launch_count_nonzero(const int* v, int n, int* count, cudaStream_t st, LaunchCounter* lc) {
  if (n <= 0) return BO_OK;
  count_nonzero_kernel<<<std::min((n + 255) / 256, 148), 256, 0, st>>>(v, n, count);
  if (lc) lc->n++;
  CUDA_CHECK_RET(cudaGetLastError());
  return BO_OK;
}

// Joint re-sampling fallback (capi.cu bo_acqf_resample_flagged): rows n_b .. n_b + q - 1 of the lower Cholesky root of the
// JOINT posterior covariance over (baseline, q-batch) are exactly [bl | br] of sample_cached_cholesky -> conditional-root
// slot `slot` of output m ([b, M, q, n_b + q]); the q-batch's posterior means go to mu [b * q, M].
__global__ void joint_rows_to_root_kernel(const double* __restrict__ rootj, int ldn, const double* __restrict__ meanj, int nb,
                                          int q, int M, int m, int slot, double* __restrict__ root, double* __restrict__ mu) {
  const int nr = nb + q;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < q * nr; idx += gridDim.x * blockDim.x) {
    const int j = idx / nr, k = idx - j * nr;
    root[(((size_t)slot * M + m) * q + j) * nr + k] = (k <= nb + j) ? rootj[(size_t)(nb + j) * ldn + k] : 0.0;
  }
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < q; j += gridDim.x * blockDim.x)
    mu[((size_t)slot * q + j) * M + m] = meanj[(size_t)(nb + j) * M + m];
}
int launch_joint_rows_to_root(const double* rootj, int ldn, const double* meanj, int nb, int q, int M, int m, int slot,
                              double* root, double* mu, cudaStream_t st, LaunchCounter* lc) {
  joint_rows_to_root_kernel<<<1, 256, 0, st>>>(rootj, ldn, meanj, nb, q, M, m, slot, root, mu);
  if (lc) lc->n++;
  CUDA_CHECK_RET(cudaGetLastError());
  return BO_OK;
}

// mask[e] = front flag as int32; ideal[o] = max over the front (or ref if the front is empty)
__global__ void front_to_mask_kernel(const unsigned char* __restrict__ front, int n, int* __restrict__ mask) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) mask[i] = front[i] ? 1 : 0;
}
int launch_front_to_mask(const unsigned char* front, int n, int* mask, cudaStream_t st, LaunchCounter* lc) {
  if (n <= 0) return BO_OK;
  front_to_mask_kernel<<<(n + 255) / 256, 256, 0, st>>>(front, n, mask);
  if (lc) lc->n++;
  CUDA_CHECK_RET(cudaGetLastError());
  return BO_OK;
}

// Exact hypervolume from the non-dominated decomposition (single sample, S = 1 layout):
// hv = prod(ideal - ref) - sum_cells prod_o max(min(up, ideal) - lo, 0).  One CTA, fixed reduction order.
__global__ void __launch_bounds__(256)
hypervolume_from_cells_kernel(const double* __restrict__ obj, const unsigned char* __restrict__ front, int n, int Mo,
                              const double* __restrict__ ref, const double* __restrict__ lo, const double* __restrict__ up,
                              const int* __restrict__ ncells, double* __restrict__ hv) {
  __shared__ double ideal[BO_MAX_OBJECTIVES];
  __shared__ double red[32];
  if (threadIdx.x < Mo) {
    double mx = ref[threadIdx.x];
    for (int e = 0; e < n; ++e)
      if (front[e]) mx = fmax(mx, obj[(size_t)e * Mo + threadIdx.x]);
    ideal[threadIdx.x] = mx;
  }
  __syncthreads();
  double acc = 0.0;
  const int nc = ncells[0];
  for (int c = threadIdx.x; c < nc; c += blockDim.x) {
    double v = 1.0;
    for (int o = 0; o < Mo; ++o) v *= fmax(fmin(up[c * Mo + o], ideal[o]) - lo[c * Mo + o], 0.0);
    acc += v;
  }
  double t = block_sum(acc, red);
  if (threadIdx.x == 0) {
    double box = 1.0;
    for (int o = 0; o < Mo; ++o) box *= (ideal[o] - ref[o]);
    hv[0] = box - t;
  }
}
int launch_hypervolume_from_cells(const double* obj, const unsigned char* front, int n, int Mo, const double* ref_dev,
                                  const double* lo, const double* up, const int* ncells, double* hv_dev, cudaStream_t st,
                                  LaunchCounter* lc) {
  hypervolume_from_cells_kernel<<<1, 256, 0, st>>>(obj, front, n, Mo, ref_dev, lo, up, ncells, hv_dev);
  if (lc) lc->n++;
  CUDA_CHECK_RET(cudaGetLastError());
  return BO_OK;
}

// ------------------------------------------------------------------------------------------------
// K4/K5/K6: one CTA per q-batch, threads over MC samples.
// f = mu + bl z_b + br z_q  ->  objective / feasibility  ->  inclusion-exclusion HVI over the sample's
// own cells, restricted to the points that overlap the cell (subsets containing a zero-length point
// contribute exactly 0, so skipping them leaves the sum unchanged).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
mc_hvi_kernel(McArgs a) {
  extern __shared__ double msm[];
  const int batch = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  const int q = a.q, nb = a.nb, nr = nb + q, M = a.M, S = a.S, Mo = a.od.n_obj;
  double* root = msm;                   // [M][q][nr]
  double* mu = root + M * q * nr;       // [q][M]
  double* objs = mu + q * M;            // [q*Mo][nt]
  double* fw = objs + (size_t)q * Mo * nt;  // [q][nt] feasibility weights
  double* red = fw + (size_t)q * nt;    // [32]
  for (int i = tid; i < M * q * nr; i += nt) root[i] = a.root[(size_t)batch * M * q * nr + i];
  for (int i = tid; i < q * M; i += nt) mu[i] = a.mu[(size_t)batch * q * M + i];
  __syncthreads();

  double total = 0.0;
  for (int s = tid; s < S; s += nt) {
    for (int j = 0; j < q; ++j) {
      double y[2 * BO_MAX_OBJECTIVES];
      for (int m = 0; m < M; ++m) {
        const double* rr = root + ((size_t)m * q + j) * nr;
        double sb = 0.0, sq = 0.0;
        if (a.Fp) sb = a.Fp[(size_t)m * a.fp_stride + ((size_t)batch * q + j) * S + s];
        else for (int e = 0; e < nb; ++e) sb = fma(rr[e], a.zbT[((size_t)e * M + m) * S + s], sb);
        for (int k = 0; k < q; ++k) sq = fma(rr[nb + k], a.zqT[((size_t)k * M + m) * S + s], sq);
        y[m] = (mu[j * M + m] + sb) + sq;
      }
      for (int o = 0; o < Mo; ++o) objs[((size_t)j * Mo + o) * nt + tid] = objective_apply(a.od.op[o], y);
      double w = 1.0;
      for (int c = 0; c < a.od.n_cons; ++c) {
        double cv = a.od.con[c].sign * (y[a.od.con[c].out_idx] - a.od.con[c].tp);
        w *= 1.0 / (1.0 + exp(cv / a.od.con[c].eta));  // sigmoid(-c / eta)
      }
      fw[(size_t)j * nt + tid] = w;
    }
    const int nc = a.cells_shared ? a.ncells[0] : a.ncells[s];
    const int sc = a.cells_shared ? 0 : s;
    const int Sc = a.cells_shared ? 1 : S;
    double acc = 0.0;
    for (int c = 0; c < nc; ++c) {
      double lo[BO_MAX_OBJECTIVES], up[BO_MAX_OBJECTIVES];
      for (int o = 0; o < Mo; ++o) {
        lo[o] = a.cell_lo[((size_t)c * Mo + o) * Sc + sc];
        up[o] = a.cell_up[((size_t)c * Mo + o) * Sc + sc];
      }
      unsigned active = 0;
      for (int j = 0; j < q; ++j) {
        bool pos = true;
        for (int o = 0; o < Mo; ++o) {
          double len = fmin(objs[((size_t)j * Mo + o) * nt + tid], up[o]) - lo[o];
          pos = pos && (len > 0.0);
        }
        if (pos) active |= (1u << j);
      }
      if (!active) continue;
      double cell = 0.0;
      // subsets of the active set, grouped by size like the reference (sizes 1..q, alternating sign)
      for (int size = 1; size <= q; ++size) {
        double asum = 0.0;
        bool any = false;
        for (unsigned sub = active; sub; sub = (sub - 1) & active) {
          if (__popc(sub) != size) continue;
          any = true;
          double vol = 1.0, wprod = 1.0;
          for (int o = 0; o < Mo; ++o) {
            double mn = up[o];
            for (unsigned rest = sub; rest; rest &= rest - 1) {
              int j = __ffs(rest) - 1;
              mn = fmin(mn, objs[((size_t)j * Mo + o) * nt + tid]);
            }
            vol *= fmax(mn - lo[o], 0.0);
          }
          if (a.od.n_cons) {
            for (unsigned rest = sub; rest; rest &= rest - 1) wprod *= fw[(size_t)(__ffs(rest) - 1) * nt + tid];
            vol *= wprod;
          }
          asum += vol;
        }
        if (any) cell += (size & 1) ? asum : -asum;
      }
      acc += cell;
    }
    total += acc;
  }
  double t = block_sum(total, red);
  if (tid == 0) {
    a.out[batch] = t / (double)S;
    if (a.info_out) {
      int v = 0;
      for (int m = 0; m < M; ++m) v |= a.info_in[(size_t)batch * M + m];
      a.info_out[batch] = v;
    }
  }
}

// Tiled variant for small cell lists (2-objective fronts): a CTA owns 64 MC samples x 64 q-batches, keeps
// the samples' cells and new-point base samples in shared memory (read once per CTA instead of once
// per q-batch) and takes the baseline part bl z_b from the sample GEMM.  Thread = (sample, batch lane);
// per-batch sums over the CTA's 64 samples go to partial[sample_group][batch], reduced by a second
// tiny kernel in a fixed order (deterministic).
// Contribution of one cell to the inclusion-exclusion sum for one MC sample.  A point overlaps the cell iff
// obj > lower in every objective (cells are non-empty, so upper > lower); the overlap test is compare-only
// and most cells are rejected after the first objective.  Only overlapping points enter the subset sums.
// One point against one cell's lower bounds: the conjunction over the objectives is carried in the predicate operand of the
// compares (setp.gt.and), and the point's bit is OR-ed in under that predicate -- MO + 1 instructions per point.  Written in
// PTX because the compiler turns the C++ form into a compare AND a select per objective (ncu source view, config 4: 32 DSETP
// + 32 SEL per cell visit).
template <int MO>
__device__ __forceinline__ unsigned overlap_bit(const double (&v)[MO], const double (&lo)[MO], unsigned active, unsigned bit) {
  static_assert(MO >= 2 && MO <= 4, "2 to 4 objectives");
  if (MO == 2)
    asm("{ .reg .pred p;\n setp.gt.f64 p, %2, %4;\n setp.gt.and.f64 p, %3, %5, p;\n @p or.b32 %0, %0, %1;\n }"
        : "+r"(active) : "r"(bit), "d"(v[0]), "d"(v[1]), "d"(lo[0]), "d"(lo[1]));
  else if (MO == 3)
    asm("{ .reg .pred p;\n setp.gt.f64 p, %2, %5;\n setp.gt.and.f64 p, %3, %6, p;\n setp.gt.and.f64 p, %4, %7, p;\n"
        " @p or.b32 %0, %0, %1;\n }"
        : "+r"(active) : "r"(bit), "d"(v[0]), "d"(v[1]), "d"(v[MO > 2 ? 2 : 0]), "d"(lo[0]), "d"(lo[1]), "d"(lo[MO > 2 ? 2 : 0]));
  else
    asm("{ .reg .pred p;\n setp.gt.f64 p, %2, %6;\n setp.gt.and.f64 p, %3, %7, p;\n setp.gt.and.f64 p, %4, %8, p;\n"
        " setp.gt.and.f64 p, %5, %9, p;\n @p or.b32 %0, %0, %1;\n }"
        : "+r"(active) : "r"(bit), "d"(v[0]), "d"(v[1]), "d"(v[MO > 2 ? 2 : 0]), "d"(v[MO > 3 ? 3 : 0]), "d"(lo[0]), "d"(lo[1]),
          "d"(lo[MO > 2 ? 2 : 0]), "d"(lo[MO > 3 ? 3 : 0]));
  return active;
}
__device__ __forceinline__ unsigned overlap_bit_f(const float* v, const float* lo, int mo, unsigned active, unsigned bit) {
  bool in = v[0] > lo[0];
  for (int o = 1; o < mo; ++o) in = in && (v[o] > lo[o]);
  return active | (in ? bit : 0u);
}
template <int MO>
__device__ __forceinline__ unsigned overlap_bit(const float (&v)[MO], const float (&lo)[MO], unsigned active, unsigned bit) {
  return overlap_bit_f(v, lo, MO, active, bit);
}

template <int QMAX, int MO, int J, int SIZE>
struct SubsetWalk {
  static __device__ __forceinline__ void go(const double (&obj)[QMAX][MO], const double (&fwt)[QMAX], unsigned active,
                                            const double (&mn)[MO], double w, const double (&lo)[MO], double (&asum)[QMAX + 1],
                                            bool has_cons) {
    SubsetWalk<QMAX, MO, J + 1, SIZE>::go(obj, fwt, active, mn, w, lo, asum, has_cons);          // subsets without point J
    if ((active >> J) & 1u) {                                                                      // subsets with point J
      double m2[MO];
#pragma unroll
      for (int o = 0; o < MO; ++o) m2[o] = fmin(mn[o], obj[J][o]);
      SubsetWalk<QMAX, MO, J + 1, SIZE + 1>::go(obj, fwt, active, m2, has_cons ? w * fwt[J] : w, lo, asum, has_cons);
    }
  }
};
template <int QMAX, int MO, int SIZE>
struct SubsetWalk<QMAX, MO, QMAX, SIZE> {
  static __device__ __forceinline__ void go(const double (&)[QMAX][MO], const double (&)[QMAX], unsigned, const double (&mn)[MO],
                                            double w, const double (&lo)[MO], double (&asum)[QMAX + 1], bool has_cons) {
    if (SIZE > 0) {
      double vol = 1.0;
#pragma unroll
      for (int o = 0; o < MO; ++o) vol *= fmax(mn[o] - lo[o], 0.0);
      if (has_cons) vol *= w;
      asum[SIZE] += vol;
    }
  }
};

template <int QMAX, int MO, bool TREE = true>
__device__ __forceinline__ double cell_contribution(const double (&obj)[QMAX][MO], const double (&fwt)[QMAX], int q,
                                                    bool has_cons, const double* __restrict__ lo_p,
                                                    const double* __restrict__ up_p, int stride,
                                                    bool single_fast = false) {
  // which points overlap the cell: one chained compare per (point, objective) -- DSETP.AND folds the conjunction over the
  // objectives into the predicate, so a point costs MO compares + one select; the per-objective bit masks this replaces cost
  // three instructions per (point, objective) and were most of the ~107 instructions per cell visit of the q = 8 kernel
  // (ncu r02: DSETP / SEL / LOP3 1.06e9 / 1.06e9 / 0.56e9 warp instructions per config-4 screen).  Unused slots hold -inf.
  double lo[MO];
#pragma unroll
  for (int o = 0; o < MO; ++o) lo[o] = lo_p[o * stride];
  unsigned active = 0;
#pragma unroll
  for (int j = 0; j < QMAX; ++j) active = overlap_bit<MO>(obj[j], lo, active, 1u << j);
  if (!active) return 0.0;
  double up[MO];
#pragma unroll
  for (int o = 0; o < MO; ++o) up[o] = up_p[o * stride];
  if (single_fast && (active & (active - 1u)) == 0u) {
    // one overlapping point (the common case): the only subset is the point itself, same operations as the general
    // loop below would do for it (vol = 1 * side_0 * side_1 ..., then the weight, then 0 + vol)
    double vol = 1.0, w = fwt[0];
#pragma unroll
    for (int j = 1; j < QMAX; ++j)
      if ((active >> j) & 1u) w = fwt[j];
#pragma unroll
    for (int o = 0; o < MO; ++o) {
      double v = obj[0][o];
#pragma unroll
      for (int j = 1; j < QMAX; ++j)
        if ((active >> j) & 1u) v = obj[j][o];
      vol *= fmax(fmin(up[o], v) - lo[o], 0.0);
    }
    if (has_cons) vol *= w;
    return 0.0 + (0.0 + vol);
  }
  // The subsets of the active set are walked as a binary decision tree over the points (include / exclude point J), fully
  // unrolled at compile time: the running minimum per objective is extended by ONE fmin per objective when a point joins
  // (nested subsets share their prefix), the subset size is a template parameter (static index into the per-size sums),
  // and a point that does not overlap the cell prunes its whole "include" subtree.  Per subset: MO fmin + MO (sub, fmax) +
  // MO mul + 1 add, no loop or index arithmetic -- the enumeration loop it replaces spent ~100 instructions per subset
  // (QMAX * MO predicated fmin, popc, QMAX predicated per-size adds), which made the q = 8 many-objective kernel issue-bound
  // at 17 % FP64 utilisation (ncu, profiles/r02_ncu_counters.json).  Sizes are summed per size first and combined with
  // alternating signs afterwards, like the reference's size-by-size loop.
  double asum[QMAX + 1];
#pragma unroll
  for (int i = 0; i <= QMAX; ++i) asum[i] = 0.0;
  if (TREE) {
    SubsetWalk<QMAX, MO, 0, 0>::go(obj, fwt, active, up, 1.0, lo, asum, has_cons);
  } else {
    // enumeration loop, kept for q = 8: the 256-leaf tree was measured on config 4 (q = 8, 4 objectives, thousands of cells
    // per sample) and changed nothing (18.89 vs 18.94 ms per screen) -- that kernel spends its ~107 instructions per
    // (sample, q-batch, cell) on the overlap test and the cell stream, not on the subset sums -- while costing 3 minutes of
    // compile time; for q <= 4 (16 leaves) the tree took config 3's MC kernel from 0.85 to 0.78 ms per screen
    for (unsigned sub = active; sub; sub = (sub - 1) & active) {
      const int size = __popc(sub);
      double vol = 1.0;
#pragma unroll
      for (int o = 0; o < MO; ++o) {
        double mn = up[o];
#pragma unroll
        for (int j = 0; j < QMAX; ++j)
          if ((sub >> j) & 1u) mn = fmin(mn, obj[j][o]);
        vol *= fmax(mn - lo[o], 0.0);
      }
      if (has_cons) {
#pragma unroll
        for (int j = 0; j < QMAX; ++j)
          if ((sub >> j) & 1u) vol *= fwt[j];
      }
#pragma unroll
      for (int i = 1; i <= QMAX; ++i)
        if (i == size) asum[i] += vol;
    }
  }
  double cell = 0.0;
#pragma unroll
  for (int size = 1; size <= QMAX; ++size) cell += (size & 1) ? asum[size] : -asum[size];
  return cell;
}

#define MT_S 64
#define MT_B 64
// STAGE: the conditional-root rows br[j][0..q) and the means of the CTA's 64 q-batches are staged in shared memory
// once (they are the same for all 64 samples), so the only global loads left in the batch loop are the Fp
// values; `prefetch` asks for those of the lane's next batch while the cells of the current one run.
template <int QMAX, int MO, bool STAGE>
__global__ void __launch_bounds__(256, (QMAX <= 4) ? 3 : 2)
mc_hvi_tiled_kernel(McArgs a, int maxc, int prefetch) {
  extern __shared__ double tsm[];
  const int tid = threadIdx.x, sl = tid & 63, bl = tid >> 6;
  const int q = a.q, nb = a.nb, nr = nb + q, M = a.M, S = a.S;
  const int s0 = blockIdx.x * MT_S, b0 = blockIdx.y * MT_B;
  double* clo = tsm;                              // [maxc*MO][64]
  double* cup = clo + (size_t)maxc * MO * MT_S;   // [maxc*MO][64]
  double* zq = cup + (size_t)maxc * MO * MT_S;    // [q*M][64]
  double* half = zq + (size_t)q * M * MT_S;       // [8] warp sums
  int* ncs = reinterpret_cast<int*>(half + 8);    // [64]
  double* rqs = half + 8 + MT_S / 2;              // [64][M*q][q]  (STAGE)
  double* mus = rqs + (size_t)MT_B * M * q * q;   // [64][q*M]     (STAGE)
  for (int i = tid; i < MT_S; i += 256) ncs[i] = (s0 + i < S) ? a.ncells[s0 + i] : 0;
  if (STAGE) {
    const int nbat = min(MT_B, a.b - b0);
    for (int idx = tid; idx < nbat * M * q * q; idx += 256) {
      int k = idx % q, row = idx / q;  // row = ((bat*M + m)*q + j)
      rqs[idx] = a.root[((size_t)b0 * M * q + row) * nr + nb + k];
    }
    for (int idx = tid; idx < nbat * q * M; idx += 256) mus[idx] = a.mu[(size_t)b0 * q * M + idx];
  }
  for (int idx = tid; idx < maxc * MO * MT_S; idx += 256) {
    int ss = idx & 63, co = idx >> 6;
    bool ok = s0 + ss < S;
    clo[idx] = ok ? a.cell_lo[(size_t)co * S + s0 + ss] : 0.0;
    cup[idx] = ok ? a.cell_up[(size_t)co * S + s0 + ss] : 0.0;
  }
  for (int idx = tid; idx < q * M * MT_S; idx += 256) {
    int ss = idx & 63, km = idx >> 6;
    zq[idx] = (s0 + ss < S) ? a.zqT[(size_t)km * S + s0 + ss] : 0.0;
  }
  __syncthreads();
  const int s = s0 + sl;
  const bool s_ok = s < S;
  const int nc = ncs[sl];
  const bool has_cons = a.od.n_cons > 0;
  for (int bat = bl; bat < MT_B; bat += 4) {
    const int batch = b0 + bat;
    if (batch >= a.b) break;  // uniform for the two warps of a batch lane
    double acc = 0.0;
    if (s_ok) {
      if ((prefetch & 1) && nb > 0 && bat + 4 < MT_B && batch + 4 < a.b) {
        for (int m = 0; m < M; ++m)
          for (int j = 0; j < q; ++j)
            asm volatile("prefetch.global.L1 [%0];\n" ::"l"(a.Fp + (size_t)m * a.fp_stride + ((size_t)(batch + 4) * q + j) * S + s));
      }
      double obj[QMAX][MO], fwt[QMAX];
      // up to two model outputs: all Fp values of the batch are requested before the first one is used, so the
      // thread waits for global memory once per batch and not once per (point, output)
      const bool pre = (M <= 2) && nb > 0 && !(prefetch & 4);
      double sbv[QMAX][2];
      if (pre) {
#pragma unroll
        for (int j = 0; j < QMAX; ++j)
#pragma unroll
          for (int m = 0; m < 2; ++m)
            sbv[j][m] = (j < q && m < M) ? a.Fp[(size_t)m * a.fp_stride + ((size_t)batch * q + j) * S + s] : 0.0;
      }
#pragma unroll
      for (int j = 0; j < QMAX; ++j) {
        fwt[j] = 1.0;
#pragma unroll
        for (int o = 0; o < MO; ++o) obj[j][o] = -INFINITY;  // unused slots never overlap a cell
        if (j < q) {
          double y[2 * BO_MAX_OBJECTIVES];
          auto point_output = [&](int m, double sb) {
            const double* rr = STAGE ? rqs + ((bat * M + m) * q + j) * q
                                     : a.root + (((size_t)batch * M + m) * q + j) * nr + nb;
            double sq = 0.0;
#pragma unroll
            for (int k = 0; k < QMAX; ++k)
              if (k < q) sq = fma(rr[k], zq[(k * M + m) * MT_S + sl], sq);
            const double mu_jm = STAGE ? mus[(bat * q + j) * M + m] : a.mu[((size_t)batch * q + j) * M + m];
            y[m] = (mu_jm + sb) + sq;
          };
          if (pre) {
#pragma unroll
            for (int m = 0; m < 2; ++m)
              if (m < M) point_output(m, sbv[j][m]);
          } else {
            for (int m = 0; m < M; ++m)
              point_output(m, (nb > 0) ? a.Fp[(size_t)m * a.fp_stride + ((size_t)batch * q + j) * S + s] : 0.0);
          }
#pragma unroll
          for (int o = 0; o < MO; ++o) obj[j][o] = objective_apply(a.od.op[o], y);
          if (has_cons) {
            double w = 1.0;
            for (int c = 0; c < a.od.n_cons; ++c) {
              double cv = a.od.con[c].sign * (y[a.od.con[c].out_idx] - a.od.con[c].tp);
              w *= 1.0 / (1.0 + exp(cv / a.od.con[c].eta));
            }
            fwt[j] = w;
          }
        }
      }
      if (prefetch & 2) {  // EVEREST_MC_FAST=0: the session-3 cell loop
        for (int c = 0; c < nc; ++c)
          acc += cell_contribution<QMAX, MO, (QMAX <= 4)>(obj, fwt, q, has_cons, clo + (c * MO) * MT_S + sl, cup + (c * MO) * MT_S + sl, MT_S);
      } else {
        // A point can only overlap a cell whose lower corner lies below the per-objective maxima over the q points:
        // two compares reject most cells before the per-point test (rejected cells contribute exactly 0).
        double mx[MO];
#pragma unroll
        for (int o = 0; o < MO; ++o) {
          mx[o] = obj[0][o];
#pragma unroll
          for (int j = 1; j < QMAX; ++j) mx[o] = fmax(mx[o], obj[j][o]);
        }
        const double* lp = clo + sl;
        const double* upp = cup + sl;
        for (int c = 0; c < nc; ++c, lp += MO * MT_S, upp += MO * MT_S) {
          bool maybe = true;
#pragma unroll
          for (int o = 0; o < MO; ++o) maybe = maybe && (mx[o] > lp[o * MT_S]);
          if (maybe) acc += cell_contribution<QMAX, MO, (QMAX <= 4)>(obj, fwt, q, has_cons, lp, upp, MT_S, true);
        }
      }
    }
    // sum over this CTA's 64 samples: two warps share one batch lane
    double wsum = warp_sum(acc);
    if ((tid & 31) == 0) half[tid >> 5] = wsum;
    __syncwarp();
    asm volatile("bar.sync %0, 64;\n" ::"r"(1 + bl));
    if (sl == 0) a.partial[(size_t)blockIdx.x * a.b + batch] = half[bl * 2] + half[bl * 2 + 1];
    asm volatile("bar.sync %0, 64;\n" ::"r"(1 + bl));
  }
}

typedef void (*McTiledFn)(McArgs, int, int);
template <int MO, bool STAGE>
static McTiledFn pick_tiled_q(int q) {
  if (q <= 2) return mc_hvi_tiled_kernel<2, MO, STAGE>;
  if (q <= 4) return mc_hvi_tiled_kernel<4, MO, STAGE>;
  if (q <= 8) return mc_hvi_tiled_kernel<8, MO, STAGE>;
  return nullptr;
}
template <bool STAGE>
static McTiledFn pick_tiled_s(int q, int Mo) {
  if (Mo == 2) return pick_tiled_q<2, STAGE>(q);
  if (Mo == 3) return pick_tiled_q<3, STAGE>(q);
  if (Mo == 4) return pick_tiled_q<4, STAGE>(q);
  return nullptr;
}
static McTiledFn pick_tiled(int q, int Mo, bool stage = false) {
  return stage ? pick_tiled_s<true>(q, Mo) : pick_tiled_s<false>(q, Mo);
}
// EVEREST_MC_STAGE=0 keeps the root rows / means in global memory (1.09 vs 0.96 ms per config-3 screen),
// EVEREST_MC_PREFETCH=1 adds the L1 prefetch of the next Fp values (measured 2 % slower, off by default)
// (experiment switches, read once).
static int mc_env_flag(const char* name, int dflt) {
  const char* e = getenv(name);
  return (e && *e) ? atoi(e) : dflt;
}
static size_t mc_stage_bytes(const McArgs& a) {
  return ((size_t)MT_B * a.M * a.q * a.q + (size_t)MT_B * a.q * a.M) * sizeof(double);
}

// ------------------------------------------------------------------------------------------------
// Many-cell variant (>2 objectives: thousands of cells per MC sample).  Pass 1 writes the objective
// values of every (q-batch, MC sample) sample-minor; pass 2 streams each sample's cell list through
// shared memory in chunks and scores 8 q-batches per thread against every chunk, so a cell is read from
// L2 once per 64 q-batches instead of once per q-batch.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
mc_objectives_kernel(McArgs a, double* __restrict__ objw) {
  // objw[(batch * (q*Mo + q) + slot) * S + s]: slots [0, q*Mo) objectives (j-major), [q*Mo, q*Mo+q) feasibility weights
  const int batch = blockIdx.x;
  const int q = a.q, nb = a.nb, nr = nb + q, M = a.M, S = a.S, Mo = a.od.n_obj;
  const int slots = q * Mo + q;
  for (int s = threadIdx.x; s < S; s += blockDim.x) {
    for (int j = 0; j < q; ++j) {
      double y[2 * BO_MAX_OBJECTIVES];
      for (int m = 0; m < M; ++m) {
        const double* rr = a.root + (((size_t)batch * M + m) * q + j) * nr + nb;
        double sb = (nb > 0) ? a.Fp[(size_t)m * a.fp_stride + ((size_t)batch * q + j) * S + s] : 0.0;
        double sq = 0.0;
        for (int k = 0; k < q; ++k) sq = fma(rr[k], a.zqT[((size_t)k * M + m) * S + s], sq);
        y[m] = (a.mu[((size_t)batch * q + j) * M + m] + sb) + sq;
      }
      for (int o = 0; o < Mo; ++o) objw[((size_t)batch * slots + j * Mo + o) * S + s] = objective_apply(a.od.op[o], y);
      double w = 1.0;
      for (int c = 0; c < a.od.n_cons; ++c) {
        double cv = a.od.con[c].sign * (y[a.od.con[c].out_idx] - a.od.con[c].tp);
        w *= 1.0 / (1.0 + exp(cv / a.od.con[c].eta));
      }
      objw[((size_t)batch * slots + q * Mo + j) * S + s] = w;
    }
  }
}

#define MC2_S 32     // samples per CTA (one warp = the 32 samples of one batch group)
#define MC2_BG 8     // batch groups per CTA (warps)
#define MC2_BPT 8    // q-batches per thread
#define MC2_CH 24    // cells per chunk
template <int QMAX, int MO>
__global__ void __launch_bounds__(256, 2)
mc_hvi_chunked_kernel(McArgs a, const double* __restrict__ objw, int maxc) {
  __shared__ double clo[MC2_CH * MO * MC2_S];
  __shared__ double cup[MC2_CH * MO * MC2_S];
  const int tid = threadIdx.x, sl = tid & 31, bg = tid >> 5;
  const int q = a.q, S = a.S;
  const int s0 = blockIdx.x * MC2_S, b0 = blockIdx.y * (MC2_BG * MC2_BPT);
  const int s = s0 + sl;
  const bool s_ok = s < S;
  const int slots = q * MO + q;
  const bool has_cons = a.od.n_cons > 0;
  double acc[MC2_BPT];
#pragma unroll
  for (int t = 0; t < MC2_BPT; ++t) acc[t] = 0.0;
  // longest cell list among the CTA's samples
  int nc = s_ok ? a.ncells[s] : 0;
  int ncmax = nc;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ncmax = max(ncmax, __shfl_xor_sync(0xffffffffu, ncmax, o));
  for (int c0 = 0; c0 < ncmax; c0 += MC2_CH) {
    __syncthreads();
    for (int idx = tid; idx < MC2_CH * MO * MC2_S; idx += 256) {
      const int ss = idx & 31, co = idx >> 5;  // co = c_local * MO + o
      const int c = c0 + co / MO;
      const int sg = s0 + ss;
      const bool ok = (sg < S) && (c < a.ncells[min(sg, S - 1)]);
      // cells past the end of a sample's list can never overlap anything
      clo[idx] = ok ? a.cell_lo[((size_t)c * MO + co % MO) * S + sg] : INFINITY;
      cup[idx] = ok ? a.cell_up[((size_t)c * MO + co % MO) * S + sg] : -INFINITY;
    }
    __syncthreads();
    if (!s_ok) continue;
    const int cn = min(MC2_CH, ncmax - c0);
#pragma unroll 1
    for (int t = 0; t < MC2_BPT; ++t) {
      const int batch = b0 + bg + t * MC2_BG;
      if (batch >= a.b) continue;
      double obj[QMAX][MO], fwt[QMAX];
      const double* ob = objw + (size_t)batch * slots * S + s;
#pragma unroll
      for (int j = 0; j < QMAX; ++j) {
        fwt[j] = 1.0;
#pragma unroll
        for (int o = 0; o < MO; ++o) obj[j][o] = (j < q) ? ob[(size_t)(j * MO + o) * S] : -INFINITY;
        if (has_cons && j < q) fwt[j] = ob[(size_t)(q * MO + j) * S];
      }
      double sum = 0.0;
      // (the tiled kernel's maxima pre-filter was tried here too: bit-identical but no faster on config 4, where the
      // q = 8 points of a batch overlap most cells: 15.05 vs 15.3 ms, tools/exp_mc4.sh)
      for (int c = 0; c < cn; ++c)
        sum += cell_contribution<QMAX, MO, (QMAX <= 4)>(obj, fwt, q, has_cons, clo + (c * MO) * MC2_S + sl, cup + (c * MO) * MC2_S + sl, MC2_S);
      acc[t] += sum;
    }
  }
#pragma unroll
  for (int t = 0; t < MC2_BPT; ++t) {
    const int batch = b0 + bg + t * MC2_BG;
    double w = warp_sum(s_ok ? acc[t] : 0.0);
    if (sl == 0 && batch < a.b) a.partial[(size_t)blockIdx.x * a.b + batch] = w;
  }
}

// Round 2: the many-cell kernel with the roles swapped -- a WARP owns one MC sample and its 32 lanes own 32 q-batches.
// The cells of a sample are then warp-uniform: one shared-memory broadcast per bound instead of 32 different addresses, one
// trip count per warp instead of 32 cell lists of different length, and every thread keeps the objective values of ITS
// q-batch in registers for the whole kernel (the kernel above re-read them from global memory for every chunk of 24 cells
// and kept 8 running sums in local memory).  8 samples per CTA, chunks of 64 cells; ncu on the kernel above (config 4):
// 21 % of the stall samples at the chunk barriers, LDL / STL 8 %, long-scoreboard waits on the objective loads.
#define MC3_SW 8      // samples (warps) per CTA
#define MC3_CH 64     // cells per chunk
template <int QMAX, int MO>
__global__ void __launch_bounds__(256, 2)
mc_hvi_bcast_kernel(McArgs a, const double* __restrict__ objw, int maxc) {
  __shared__ double clo[MC3_CH * MO * MC3_SW];   // [(c * MO + o)][sample]
  __shared__ double cup[MC3_CH * MO * MC3_SW];
  __shared__ double red[MC3_SW][33];
  __shared__ int ncs[MC3_SW];
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int q = a.q, S = a.S;
  const int s0 = blockIdx.x * MC3_SW, s = s0 + w;
  const int batch = blockIdx.y * 32 + lane;
  const bool ok = s < S && batch < a.b;
  const int slots = q * MO + q;
  const bool has_cons = a.od.n_cons > 0;
  if (tid < MC3_SW) ncs[tid] = (s0 + tid < S) ? a.ncells[s0 + tid] : 0;
  double obj[QMAX][MO], fwt[QMAX];
  {
    const double* ob = objw + (size_t)(ok ? batch : 0) * slots * S + (ok ? s : 0);
#pragma unroll
    for (int j = 0; j < QMAX; ++j) {
      fwt[j] = 1.0;
#pragma unroll
      for (int o = 0; o < MO; ++o) obj[j][o] = (ok && j < q) ? ob[(size_t)(j * MO + o) * S] : -INFINITY;
      if (has_cons && ok && j < q) fwt[j] = ob[(size_t)(q * MO + j) * S];
    }
  }
  __syncthreads();
  int ncmax = 0;
#pragma unroll
  for (int k = 0; k < MC3_SW; ++k) ncmax = max(ncmax, ncs[k]);
  const int nc = ncs[w];
  double acc = 0.0;
  for (int c0 = 0; c0 < ncmax; c0 += MC3_CH) {
    __syncthreads();
    for (int idx = tid; idx < MC3_CH * MO * MC3_SW; idx += 256) {
      const int ss = idx % MC3_SW, co = idx / MC3_SW;      // 8 consecutive samples of one (cell, objective): 64 contiguous bytes
      const int c = c0 + co / MO, sg = s0 + ss;
      const bool in = (sg < S) && (c < ncs[ss]);
      // cells past the end of a sample's list can never overlap anything
      clo[idx] = in ? a.cell_lo[((size_t)c * MO + co % MO) * S + sg] : INFINITY;
      cup[idx] = in ? a.cell_up[((size_t)c * MO + co % MO) * S + sg] : -INFINITY;
    }
    __syncthreads();
    if (ok) {
      const int cn = min(MC3_CH, nc - c0);                   // warp-uniform
      for (int c = 0; c < cn; ++c)
        acc += cell_contribution<QMAX, MO, (QMAX <= 4)>(obj, fwt, q, has_cons, clo + (c * MO) * MC3_SW + w, cup + (c * MO) * MC3_SW + w, MC3_SW);
    }
  }
  red[w][lane] = ok ? acc : 0.0;
  __syncthreads();
  if (w == 0 && batch < a.b) {
    double t = 0.0;
#pragma unroll
    for (int k = 0; k < MC3_SW; ++k) t += red[k][lane];     // fixed order over the CTA's samples
    a.partial[(size_t)blockIdx.x * a.b + batch] = t;
  }
}

// Round 2, second step: the same warp = sample / lane = q-batch layout, but the cells are SCANNED first and the subset sums
// run from a per-lane work list afterwards.  ncu source view of the kernel above on config 4 (q = 8, 4 objectives, ~1500
// cells per sample): 10.8 of 32 lanes active on average -- for a given cell only a few of the warp's 32 q-batches overlap
// it, and the others wait while those walk their subsets, each subset costing QMAX * MO predicated FP64 minima (three
// instructions each).  Here
//   scan     every lane tests every cell of the chunk (uniform, compare-only; with FILT on FP32 copies rounded outwards,
//            which can only add candidates) and notes the overlapping points of cell c in act[c][lane] and bit c of `hit`;
//   process  each lane walks ITS hit cells and their subsets as one flattened loop (one trip per subset, cell switches
//            inside the loop), so a lane waits for the longest work list of the warp instead of for the busiest lane of
//            every single cell; the members of a subset are visited by set bit from a shared-memory copy of the lane's
//            objective values (dynamic point index without local memory), not by QMAX predicated steps.
// Cells are accumulated in ascending order per lane as before (deterministic); inside a cell the subset volumes are summed
// with their signs in enumeration order instead of size by size.
#define MC4_SW 4      // samples (warps) per CTA
#define MC4_CH 64     // cells per chunk (one bit each in `hit`)
template <int QMAX, int MO>
struct Mc4Layout {
  static constexpr int CS = 2 * MO + 1;              // lower, upper, pad: odd stride, lanes read different cells
  static constexpr int CB = MC4_CH * CS;             // doubles
  static constexpr int OB = QMAX * MO * 32;          // [(j * MO + o)][lane]
  static constexpr int FW = QMAX * 32;               // [j][lane]
  static constexpr int AC = MC4_CH * 32 / 8;         // act bytes, in doubles
  static constexpr int LF = MC4_CH * MO / 2 + 1;     // FP32 lower bounds, in doubles
  static constexpr int PL = MC4_CH * 32 / 4;         // balanced variant: (owner lane, cell) pairs as u16, in doubles
  static constexpr int PV = 256;                     // balanced variant: pair values of one batch
  static constexpr int PER_WARP = CB + OB + FW + AC + LF;
  static constexpr int PER_WARP_BAL = PER_WARP + PL + PV;
};
template <bool FILT> struct Mc4Scan { typedef double T; };
template <> struct Mc4Scan<true> { typedef float T; };

template <int QMAX, int MO, bool FILT, bool BAL>
__global__ void __launch_bounds__(MC4_SW * 32)
mc_hvi_queue_kernel(McArgs a, const double* __restrict__ objw, int maxc) {
  extern __shared__ double qsm[];
  typedef Mc4Layout<QMAX, MO> L;
  typedef typename Mc4Scan<FILT>::T T;
  __shared__ double red[MC4_SW][33];
  __shared__ int ncs[MC4_SW];
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  constexpr int PW = BAL ? L::PER_WARP_BAL : L::PER_WARP;
  double* cb = qsm + (size_t)w * PW;
  double* objs = cb + L::CB;
  double* fwts = objs + L::OB;
  unsigned char* act = reinterpret_cast<unsigned char*>(fwts + L::FW);
  const float* lof = reinterpret_cast<const float*>(fwts + L::FW + L::AC);
  const int q = a.q, S = a.S;
  const int s0 = blockIdx.x * MC4_SW, s = s0 + w;
  const int batch = blockIdx.y * 32 + lane;
  const bool ok = s < S && batch < a.b;
  const int slots = q * MO + q;
  const bool has_cons = a.od.n_cons > 0;
  if (tid < MC4_SW) ncs[tid] = (s0 + tid < S) ? a.ncells[s0 + tid] : 0;
  T objr[QMAX][MO];
  {
    const double* ob = objw + (size_t)(ok ? batch : 0) * slots * S + (ok ? s : 0);
#pragma unroll
    for (int j = 0; j < QMAX; ++j) {
#pragma unroll
      for (int o = 0; o < MO; ++o) {
        const double v = (ok && j < q) ? ob[(size_t)(j * MO + o) * S] : -INFINITY;
        objs[(j * MO + o) * 32 + lane] = v;
        objr[j][o] = FILT ? (T)__double2float_ru(v) : (T)v;
      }
      fwts[j * 32 + lane] = (has_cons && ok && j < q) ? ob[(size_t)(q * MO + j) * S] : 1.0;
    }
  }
  __syncthreads();
  int ncmax = 0;
#pragma unroll
  for (int k = 0; k < MC4_SW; ++k) ncmax = max(ncmax, ncs[k]);
  const int nc = ncs[w];
  double acc = 0.0;
  for (int c0 = 0; c0 < ncmax; c0 += MC4_CH) {
    __syncthreads();
    for (int idx = tid; idx < MC4_CH * MO * MC4_SW; idx += MC4_SW * 32) {
      const int ss = idx % MC4_SW, co = idx / MC4_SW;      // consecutive samples of one (cell, objective): one 32-byte sector
      const int cc = co / MO, o = co % MO, c = c0 + cc, sg = s0 + ss;
      const bool in = (sg < S) && (c < ncs[ss]);
      const double lo = in ? a.cell_lo[((size_t)c * MO + o) * S + sg] : INFINITY;
      const double up = in ? a.cell_up[((size_t)c * MO + o) * S + sg] : -INFINITY;
      double* cbs = qsm + (size_t)ss * PW;
      cbs[cc * L::CS + o] = lo;
      cbs[cc * L::CS + MO + o] = up;
      if (FILT) reinterpret_cast<float*>(cbs + L::CB + L::OB + L::FW + L::AC)[cc * MO + o] = __double2float_rd(lo);
    }
    __syncthreads();
    const int cn = min(MC4_CH, nc - c0);                     // warp-uniform, <= 0 past the end of this sample's list
    // ---- scan
    unsigned long long hit = 0ull;
    for (int c = 0; c < cn; ++c) {
      T lo[MO];
#pragma unroll
      for (int o = 0; o < MO; ++o) lo[o] = FILT ? (T)lof[c * MO + o] : (T)cb[c * L::CS + o];
      unsigned active = 0;
#pragma unroll
      for (int j = 0; j < QMAX; ++j) active = overlap_bit<MO>(objr[j], lo, active, 1u << j);
      act[c * 32 + lane] = (unsigned char)active;
      hit |= (unsigned long long)(active != 0u) << c;
    }
    if (BAL) {
      // ---- process, balanced: in the loop below a lane walks ITS hit cells, and a few q-batches of a warp own most of them
      // (ncu: 3.5 of 32 lanes active).  Here the (owner lane, hit cell) pairs of the warp are listed in shared memory (owner
      // by owner, cells ascending) and dealt to the lanes 32 at a time; a pair's value is formed exactly as below from the
      // owner's objective values, and every owner then adds ITS pair values in list order -- the same sums in the same order.
      unsigned short* pairs = reinterpret_cast<unsigned short*>(fwts + L::FW + L::AC + L::LF);
      double* pvals = fwts + L::FW + L::AC + L::LF + L::PL;
      const int nh = __popcll(hit);
      int incl = nh;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
      }
      const int off = incl - nh;
      const int total = __shfl_sync(0xffffffffu, incl, 31);
      {
        int k = off;
        for (unsigned long long h = hit; h; h &= h - 1ull) pairs[k++] = (unsigned short)((lane << 8) | (__ffsll((long long)h) - 1));
      }
      __syncwarp();
      for (int base = 0; base < total; base += L::PV) {
        const int lim = min(total, base + L::PV);
        for (int p = base + lane; p < lim; p += 32) {
          const int owner = pairs[p] >> 8, c = pairs[p] & 255;
          const double* l = cb + c * L::CS;
          double lo[MO], up[MO];
#pragma unroll
          for (int o = 0; o < MO; ++o) { lo[o] = l[o]; up[o] = l[MO + o]; }
          const unsigned cand = act[c * 32 + owner];
          unsigned active = cand;
          if (FILT) {                                            // exact re-test of the FP32 candidates
            active = 0;
            for (unsigned m = cand; m; m &= m - 1u) {
              const int j = __ffs((int)m) - 1;
              bool in = objs[(j * MO) * 32 + owner] > lo[0];
#pragma unroll
              for (int o = 1; o < MO; ++o) in = in && (objs[(j * MO + o) * 32 + owner] > lo[o]);
              active |= in ? (1u << j) : 0u;
            }
          }
          double cell = 0.0;
          for (unsigned sub = active; sub; sub = (sub - 1u) & active) {
            unsigned m = sub;
            int j = __ffs((int)m) - 1;
            m &= m - 1u;
            double mn[MO];
#pragma unroll
            for (int o = 0; o < MO; ++o) mn[o] = fmin(up[o], objs[(j * MO + o) * 32 + owner]);
            for (; m; m &= m - 1u) {
              j = __ffs((int)m) - 1;
#pragma unroll
              for (int o = 0; o < MO; ++o) mn[o] = fmin(mn[o], objs[(j * MO + o) * 32 + owner]);
            }
            double vol = mn[0] - lo[0];
#pragma unroll
            for (int o = 1; o < MO; ++o) vol *= mn[o] - lo[o];
            if (has_cons)
              for (unsigned m2 = sub; m2; m2 &= m2 - 1u) vol *= fwts[(__ffs((int)m2) - 1) * 32 + owner];
            cell += (__popc(sub) & 1) ? vol : -vol;
          }
          pvals[p - base] = cell;
        }
        __syncwarp();
        const int k0 = max(off, base), k1 = min(off + nh, lim);
        for (int k = k0; k < k1; ++k) acc += pvals[k - base];
        __syncwarp();
      }
      continue;
    }
    // ---- process: one trip per (hit cell, subset of its overlapping points)
    unsigned sub = 0, active = 0;
    double cell = 0.0, lo[MO], up[MO];
    for (;;) {
      if (sub == 0) {
        acc += cell;
        cell = 0.0;
        if (!hit) break;
        const int c = __ffsll((long long)hit) - 1;
        hit &= hit - 1ull;
        const double* l = cb + c * L::CS;
#pragma unroll
        for (int o = 0; o < MO; ++o) { lo[o] = l[o]; up[o] = l[MO + o]; }
        const unsigned cand = act[c * 32 + lane];
        if (FILT) {                                            // exact re-test of the FP32 candidates
          active = 0;
          for (unsigned m = cand; m; m &= m - 1u) {
            const int j = __ffs((int)m) - 1;
            bool in = objs[(j * MO) * 32 + lane] > lo[0];
#pragma unroll
            for (int o = 1; o < MO; ++o) in = in && (objs[(j * MO + o) * 32 + lane] > lo[o]);
            active |= in ? (1u << j) : 0u;
          }
        } else {
          active = cand;
        }
        sub = active;
        if (!sub) continue;
      }
      // volume of the intersection of the subset's boxes with the cell; every member exceeds `lo` in every objective,
      // so the sides are positive without a clamp
      unsigned m = sub;
      int j = __ffs((int)m) - 1;
      m &= m - 1u;
      double mn[MO];
#pragma unroll
      for (int o = 0; o < MO; ++o) mn[o] = fmin(up[o], objs[(j * MO + o) * 32 + lane]);
      for (; m; m &= m - 1u) {
        j = __ffs((int)m) - 1;
#pragma unroll
        for (int o = 0; o < MO; ++o) mn[o] = fmin(mn[o], objs[(j * MO + o) * 32 + lane]);
      }
      double vol = mn[0] - lo[0];
#pragma unroll
      for (int o = 1; o < MO; ++o) vol *= mn[o] - lo[o];
      if (has_cons)
        for (unsigned m2 = sub; m2; m2 &= m2 - 1u) vol *= fwts[(__ffs((int)m2) - 1) * 32 + lane];
      cell += (__popc(sub) & 1) ? vol : -vol;
      sub = (sub - 1u) & active;
    }
  }
  red[w][lane] = ok ? acc : 0.0;
  __syncthreads();
  if (w == 0 && batch < a.b) {
    double t = 0.0;
#pragma unroll
    for (int k = 0; k < MC4_SW; ++k) t += red[k][lane];     // fixed order over the CTA's samples
    a.partial[(size_t)blockIdx.x * a.b + batch] = t;
  }
}

typedef void (*McChunkedFn)(McArgs, const double*, int);
template <int MO>
static McChunkedFn pick_queue_q(int q, bool filt, size_t* bytes) {
  // EVEREST_MC_BALANCE=1: the (q-batch, hit cell) pairs of a warp dealt evenly over its lanes (BAL).  Bit-identical values;
  // measured on config 4: 3.77 ms against 3.63 ms for the per-lane work lists -- the pair list and the pair values cost 6 KB of
  // shared memory per warp (2 instead of 3 CTAs per SM for the FP64-bound scan) and a prefix sum, a list and two warp
  // barriers per chunk, which is more than the idle lanes of the per-lane form cost.  Kept as an experiment switch.
  static const bool bal = mc_env_flag("EVEREST_MC_BALANCE", 0) != 0;
  if (q <= 4) {
    *bytes = (size_t)MC4_SW * (bal ? Mc4Layout<4, MO>::PER_WARP_BAL : Mc4Layout<4, MO>::PER_WARP) * sizeof(double);
    if (bal) return filt ? mc_hvi_queue_kernel<4, MO, true, true> : mc_hvi_queue_kernel<4, MO, false, true>;
    return filt ? mc_hvi_queue_kernel<4, MO, true, false> : mc_hvi_queue_kernel<4, MO, false, false>;
  }
  if (q <= 8) {
    *bytes = (size_t)MC4_SW * (bal ? Mc4Layout<8, MO>::PER_WARP_BAL : Mc4Layout<8, MO>::PER_WARP) * sizeof(double);
    if (bal) return filt ? mc_hvi_queue_kernel<8, MO, true, true> : mc_hvi_queue_kernel<8, MO, false, true>;
    return filt ? mc_hvi_queue_kernel<8, MO, true, false> : mc_hvi_queue_kernel<8, MO, false, false>;
  }
  return nullptr;
}
static McChunkedFn pick_queue(int q, int Mo, bool filt, size_t* bytes) {
  if (Mo == 2) return pick_queue_q<2>(q, filt, bytes);
  if (Mo == 3) return pick_queue_q<3>(q, filt, bytes);
  if (Mo == 4) return pick_queue_q<4>(q, filt, bytes);
  return nullptr;
}
template <int MO>
static McChunkedFn pick_bcast_q(int q) {
  if (q <= 2) return mc_hvi_bcast_kernel<2, MO>;
  if (q <= 4) return mc_hvi_bcast_kernel<4, MO>;
  if (q <= 8) return mc_hvi_bcast_kernel<8, MO>;
  return nullptr;
}
static McChunkedFn pick_bcast(int q, int Mo) {
  if (Mo == 2) return pick_bcast_q<2>(q);
  if (Mo == 3) return pick_bcast_q<3>(q);
  if (Mo == 4) return pick_bcast_q<4>(q);
  return nullptr;
}
int mc_hvi_partial_groups(int S) { return (S + MC4_SW - 1) / MC4_SW; }   // the finest grouping of the three kernels

template <int MO>
static McChunkedFn pick_chunked_q(int q) {
  if (q <= 2) return mc_hvi_chunked_kernel<2, MO>;
  if (q <= 4) return mc_hvi_chunked_kernel<4, MO>;
  if (q <= 8) return mc_hvi_chunked_kernel<8, MO>;
  return nullptr;
}
static McChunkedFn pick_chunked(int q, int Mo) {
  if (Mo == 2) return pick_chunked_q<2>(q);
  if (Mo == 3) return pick_chunked_q<3>(q);
  if (Mo == 4) return pick_chunked_q<4>(q);
  return nullptr;
}

__global__ void mc_reduce_partials_kernel(const double* __restrict__ partial, int groups, int b, int S,
                                          double* __restrict__ out, const int* __restrict__ info_in, int M,
                                          int* __restrict__ info_out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= b) return;
  double t = 0.0;
  for (int g = 0; g < groups; ++g) t += partial[(size_t)g * b + i];
  out[i] = t / (double)S;
  if (info_out) {
    int v = 0;
    for (int m = 0; m < M; ++m) v |= info_in[(size_t)i * M + m];
    info_out[i] = v;
  }
}

int launch_mc_reduce_partials(const double* partial, int groups, int b, int S, double* out, const int* info_in, int M,
                              int* info_out, cudaStream_t st, LaunchCounter* lc) {
  mc_reduce_partials_kernel<<<(b + 255) / 256, 256, 0, st>>>(partial, groups, b, S, out, info_in, M, info_out);
  if (lc) lc->n++;
  CUDA_CHECK_RET(cudaGetLastError());
  return BO_OK;
}

// EVEREST_MC_PATH = tiled | chunked | generic forces one of the three HVI kernels (tests); default: automatic
static int mc_forced_path() {
  const char* e = getenv("EVEREST_MC_PATH");
  if (!e) return 0;
  if (e[0] == 't') return 1;
  if (e[0] == 'c' || e[0] == 'o') return 2;
  if (e[0] == 'g') return 3;
  return 0;
}
static bool mc_use_tiled(const McArgs& a, int max_cells, size_t* bytes) {
  const int Mo = a.od.n_obj;
  const int forced = mc_forced_path();
  if (forced == 2 || forced == 3) { if (bytes) *bytes = 0; return false; }
  size_t tiled = ((size_t)2 * max_cells * Mo * MT_S + (size_t)a.q * a.M * MT_S + 8) * sizeof(double) + MT_S * sizeof(int);
  if (bytes) *bytes = tiled;
  return pick_tiled(a.q, Mo) && !a.cells_shared && a.partial && (a.nb == 0 || a.Fp) && tiled <= 100 * 1024;
}
static bool mc_use_chunked(const McArgs& a, int max_cells) {
  if (mc_forced_path() == 3) return false;
  return !mc_use_tiled(a, max_cells, nullptr) && pick_chunked(a.q, a.od.n_obj) && !a.cells_shared && a.partial &&
         (a.nb == 0 || a.Fp);
}
size_t mc_hvi_obj_ws_bytes(const McArgs& a, int max_cells) {
  if (!mc_use_chunked(a, max_cells)) return 0;
  return (size_t)a.b * (a.q * a.od.n_obj + a.q) * a.S * sizeof(double);
}

int launch_mc_hvi(const McArgs& a, int max_cells, double* obj_ws, cudaStream_t st, LaunchCounter* lc) {
  if (a.b <= 0) return BO_OK;
  const int Mo = a.od.n_obj;
  if (obj_ws && mc_use_chunked(a, max_cells)) {
    mc_objectives_kernel<<<a.b, 256, 0, st>>>(a, obj_ws);
    if (lc) lc->n++;
    // EVEREST_MC_PATH=old keeps the round-1 kernel (thread = sample, 8 q-batches per thread) for comparison
    static const bool old_kernel = []() { const char* e = getenv("EVEREST_MC_PATH"); return e && e[0] == 'o'; }();
    // EVEREST_MC_QUEUE: 0 = broadcast kernel, 1 = scan / work-list kernel with FP64 scan (default), 2 = with the FP32
    // pre-filter (measured slower on config 4: 10.2 vs 9.15 ms per screen; the exact re-test costs more than the scan saves)
    static const int queue_mode = mc_env_flag("EVEREST_MC_QUEUE", 1);
    dim3 grid;
    size_t qbytes = 0;
    McChunkedFn qf = (!old_kernel && queue_mode > 0 && a.q > 4) ? pick_queue(a.q, Mo, queue_mode > 1, &qbytes) : nullptr;
    if (qf) {
      CUDA_CHECK_RET(cudaFuncSetAttribute(qf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)qbytes));
      grid = dim3((a.S + MC4_SW - 1) / MC4_SW, (a.b + 31) / 32);
      qf<<<grid, MC4_SW * 32, qbytes, st>>>(a, obj_ws, max_cells);
    } else if (old_kernel) {
      McChunkedFn cf = pick_chunked(a.q, Mo);
      grid = dim3((a.S + MC2_S - 1) / MC2_S, (a.b + MC2_BG * MC2_BPT - 1) / (MC2_BG * MC2_BPT));
      cf<<<grid, 256, 0, st>>>(a, obj_ws, max_cells);
    } else {
      McChunkedFn cf = pick_bcast(a.q, Mo);
      grid = dim3((a.S + MC3_SW - 1) / MC3_SW, (a.b + 31) / 32);
      cf<<<grid, 256, 0, st>>>(a, obj_ws, max_cells);
    }
    if (lc) lc->n++;
    mc_reduce_partials_kernel<<<(a.b + 255) / 256, 256, 0, st>>>(a.partial, grid.x, a.b, a.S, a.out, a.info_in, a.M, a.info_out);
    if (lc) lc->n++;
    CUDA_CHECK_RET(cudaGetLastError());
    return BO_OK;
  }
  // tiled path: needs per-sample cells, the sample GEMM output (or no baseline) and a cell list that fits
  size_t tiled = 0;
  if (mc_use_tiled(a, max_cells, &tiled)) {
    static const int want_stage = mc_env_flag("EVEREST_MC_STAGE", 1),
                     prefetch = (mc_env_flag("EVEREST_MC_PREFETCH", 0) ? 1 : 0) | (mc_env_flag("EVEREST_MC_FAST", 1) ? 0 : 2) |
                                (mc_env_flag("EVEREST_MC_FPBATCH", 1) ? 0 : 4);
    // staging must leave room for two CTAs per SM
    const bool stage = want_stage && tiled + mc_stage_bytes(a) <= 100 * 1024;
    if (stage) tiled += mc_stage_bytes(a);
    McTiledFn fn = pick_tiled(a.q, Mo, stage);
    if (tiled > 48 * 1024) CUDA_CHECK_RET(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tiled));
    CUDA_CHECK_RET(cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    dim3 grid((a.S + MT_S - 1) / MT_S, (a.b + MT_B - 1) / MT_B);
    fn<<<grid, 256, tiled, st>>>(a, max_cells, prefetch);
    if (lc) lc->n++;
    mc_reduce_partials_kernel<<<(a.b + 255) / 256, 256, 0, st>>>(a.partial, grid.x, a.b, a.S, a.out, a.info_in, a.M, a.info_out);
    if (lc) lc->n++;
    CUDA_CHECK_RET(cudaGetLastError());
    return BO_OK;
  }
  const int nt = 256;
  size_t smem = ((size_t)a.M * a.q * (a.nb + a.q) + a.q * a.M + (size_t)a.q * a.od.n_obj * nt + (size_t)a.q * nt + 32) *
                sizeof(double);
  if (smem > 220 * 1024) { bo_set_error("mc_hvi: shared memory budget exceeded (n_b=%d q=%d)", a.nb, a.q); return BO_ERR_INVALID; }
  static PerDeviceMax attr_pd; size_t& attr = attr_pd.slot();
  if (smem > 48 * 1024 && smem > attr) {
    CUDA_CHECK_RET(cudaFuncSetAttribute(mc_hvi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  mc_hvi_kernel<<<a.b, nt, smem, st>>>(a);
  if (lc) lc->n++;
  CUDA_CHECK_RET(cudaGetLastError());
  return BO_OK;
}


// Exact marginal log likelihood of one output and its gradient with respect to the kernel hyper-parameters: the
// objective fit_gpytorch_mll minimises inside SingleTaskGPSurrogate._fit (surrogates/single_task_gp.py:39-71,
// mixed_single_task_gp.py:111-112) -- SURVEY.md 8f-2, "the step before" the acquisition path.  Reuses the factorisation
// of the state (L, L^-1, alpha) and the explicit inverse built for the adjoint kernels:
//   mll        = -1/2 r^T alpha - sum_i log L_ii - N/2 log(2 pi)                       (r = standardised y - mean constant)
//   d/d theta  = 1/2 sum_ij W_ij dK_ij/d theta,   W = alpha alpha^T - (K + s2 I)^-1
//   d/d noise  = 1/2 tr W,   d/d mean = sum_i alpha_i
// One CTA per training row i accumulates the row's share of every lengthscale / outputscale derivative; a second
// kernel adds the rows in a fixed order.  FP64 throughout.
#include "acqf.cuh"
#include "common.cuh"
#include "mc_math.cuh"

#define ML_CHUNK 2048

struct MllGradArgs {
  ModelD md;
  PrepD train;            // prepared training points (same buffers the leaves point to)
  int N, ldk;
  const double* alpha;    // [ldk]
  const double* Kinv;     // [N, ldk]
  int ls_offset[BO_MAX_LEAVES];  // position of each leaf's lengthscale block in the parameter vector (-1: none)
  int n_ls_total;         // lengthscale slots (ARD dims of all continuous leaves, groups of Hamming leaves)
  int n_params;           // n_ls_total + n_terms
  double* part;           // [N, n_params] per-row contributions
};

__device__ __forceinline__ double ml_leaf_dk_dstat(int kind, double stat) {
  switch (kind) {
    case BO_LEAF_RBF:
      return -0.5 * exp(-0.5 * stat);
    case BO_LEAF_MATERN12: {
      if (!(stat > 1e-30)) return 0.0;
      double r = sqrt(stat);
      return -exp(-r) / (2.0 * r);
    }
    case BO_LEAF_MATERN32: {
      if (!(stat > 1e-30)) return 0.0;
      return -1.5 * exp(-1.7320508075688772 * sqrt(stat));
    }
    case BO_LEAF_MATERN52: {
      if (!(stat > 1e-30)) return 0.0;
      double r = sqrt(stat);
      return -(5.0 / 6.0) * (1.0 + 2.23606797749979 * r) * exp(-2.23606797749979 * r);
    }
  }
  return 0.0;
}

__global__ void __launch_bounds__(256)
mll_grad_kernel(const __grid_constant__ MllGradArgs a, int dpad_max) {
  extern __shared__ double msm[];
  const int i = blockIdx.x, tid = threadIdx.x, N = a.N;
  const ModelD& md = a.md;
  double* w = msm;                  // [ML_CHUNK] W_ij
  double* g = w + ML_CHUNK;         // [ML_CHUNK]
  double* part = g + ML_CHUNK;      // [3][256]
  double* xsi = part + 768;         // [dpad_max]
  double* acc = xsi + dpad_max;     // [n_params] this row's contributions
  double* red = acc + a.n_params;   // [32]
  for (int c = tid; c < a.n_params; c += 256) acc[c] = 0.0;
  const double ai = a.alpha[i];
  for (int p0 = 0; p0 < N; p0 += ML_CHUNK) {
    const int cn = min(ML_CHUNK, N - p0);
    __syncthreads();
    for (int pl = tid; pl < cn; pl += 256) {
      const int j = p0 + pl;
      w[pl] = ai * a.alpha[j] - a.Kinv[(size_t)i * a.ldk + j];
    }
    __syncthreads();
    // outputscales: d K / d coef_t = product of the term's leaves
    for (int t = 0; t < md.n_terms; ++t) {
      double s = 0.0;
      for (int pl = tid; pl < cn; pl += 256) {
        const int j = p0 + pl;
        double prod = 1.0;
        for (int f = 0; f < md.nfac[t]; ++f) {
          const int lf = md.fac[t][f];
          prod *= leaf_eval_pair(md.leaf[lf], prep_side(a.train, lf), i, prep_side(a.train, lf), j, i == j);
        }
        s = fma(w[pl], prod, s);
      }
      s = block_sum(s, red);
      if (tid == 0) acc[a.n_ls_total + t] += 0.5 * s;
    }
    for (int l = 0; l < md.n_leaves; ++l) {
      const LeafD& L = md.leaf[l];
      if (a.ls_offset[l] < 0) continue;
      __syncthreads();
      // coefficient of leaf l in the sum-of-products tree (depends on the pair through the other leaves)
      auto tree_coef = [&](int j) {
        double cl = 0.0;
        for (int t = 0; t < md.n_terms; ++t) {
          bool has = false;
          for (int f = 0; f < md.nfac[t]; ++f) has = has || (md.fac[t][f] == l);
          if (!has) continue;
          double prod = md.coef[t];
          bool skipped = false;
          for (int f = 0; f < md.nfac[t]; ++f) {
            const int lf = md.fac[t][f];
            if (lf == l && !skipped) { skipped = true; continue; }
            prod *= leaf_eval_pair(md.leaf[lf], prep_side(a.train, lf), i, prep_side(a.train, lf), j, i == j);
          }
          cl += prod;
        }
        return cl;
      };
      if (L.kind <= BO_LEAF_MATERN52) {
        for (int k = tid; k < L.dpad; k += 256) xsi[k] = a.train.Xs[l][(size_t)i * L.dpad + k];
        __syncthreads();
        const double n2i = a.train.n2[l][i];
        for (int pl = tid; pl < cn; pl += 256) {
          const int j = p0 + pl;
          double gv = 0.0;
          if (j != i) {   // the diagonal is a forced zero distance: no dependence on the lengthscales
            const double* xb = a.train.Xs[l] + (size_t)j * L.dpad;
            double dot = 0.0;
            for (int k = 0; k < L.nd; ++k) dot = fma(xsi[k], xb[k], dot);
            const double stat = fmax(n2i + a.train.n2[l][j] - 2.0 * dot, 0.0);
            gv = w[pl] * tree_coef(j) * ml_leaf_dk_dstat(L.kind, stat);
          }
          g[pl] = gv;
        }
        __syncthreads();
        // sum_j g_ij (xs_i[a] - xs_j[a])^2 = xs_i^2 Sg - 2 xs_i Sgx + Sgxx ;  d stat / d ls_a = -2 (.)^2 / ls_a
        for (int a0 = 0; a0 < L.dpad; a0 += 256) {
          const int na = min(256, L.dpad - a0);
          const int nsl = max(1, 256 / na);
          const int ai_ = tid % na, sl = tid / na;
          double sg = 0.0, sgx = 0.0, sgxx = 0.0;
          if (sl < nsl) {
            const double* xb = a.train.Xs[l] + (size_t)p0 * L.dpad + a0 + ai_;
#pragma unroll 8
            for (int pl = sl; pl < cn; pl += nsl) {
              const double gv = g[pl], x = xb[(size_t)pl * L.dpad];
              sg += gv;
              sgx = fma(gv, x, sgx);
              sgxx = fma(gv * x, x, sgxx);
            }
          }
          part[tid] = sg; part[256 + tid] = sgx; part[512 + tid] = sgxx;
          __syncthreads();
          if (tid < na) {
            double tg = 0.0, tx = 0.0, txx = 0.0;
            for (int s = 0; s < nsl; ++s) { tg += part[s * na + tid]; tx += part[256 + s * na + tid]; txx += part[512 + s * na + tid]; }
            const int ak = a0 + tid;
            if (ak < L.nd) {
              const double xi = xsi[ak];
              const double q2 = xi * xi * tg - 2.0 * xi * tx + txx;
              acc[a.ls_offset[l] + ak] += 0.5 * (-2.0 / L.ls[ak]) * q2;
            }
          }
          __syncthreads();
        }
      } else if (L.kind == BO_LEAF_HAMMING) {
        // k = exp(-mean_f(delta_f / ls_f)):  d k / d ls_f = k delta_f / (F ls_f^2),  wls = 1 / ls
        for (int f = 0; f < L.nd; ++f) {
          double s = 0.0;
          for (int pl = tid; pl < cn; pl += 256) {
            const int j = p0 + pl;
            if (a.train.codes[l][(size_t)i * L.nd + f] != a.train.codes[l][(size_t)j * L.nd + f]) {
              const double kv = leaf_eval_pair(L, prep_side(a.train, l), i, prep_side(a.train, l), j, false);
              s = fma(w[pl] * tree_coef(j), kv * L.wls[f] * L.wls[f] / (double)L.nd, s);
            }
          }
          s = block_sum(s, red);
          if (tid == 0) acc[a.ls_offset[l] + f] += 0.5 * s;
        }
      }
    }
  }
  __syncthreads();
  for (int c = tid; c < a.n_params; c += 256) a.part[(size_t)i * a.n_params + c] = acc[c];
}

__global__ void mll_reduce_kernel(const double* __restrict__ part, int N, int n_params, double* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_params) return;
  double s = 0.0;
  for (int i = 0; i < N; ++i) s += part[(size_t)i * n_params + c];
  out[c] = s;
}

// scalars: out[0] = r^T alpha, out[1] = sum log L_ii, out[2] = sum alpha_i, out[3] = sum alpha_i^2, out[4] = tr Kinv
__global__ void __launch_bounds__(256)
mll_scalars_kernel(const double* __restrict__ resid, const double* __restrict__ alpha, const double* __restrict__ Lmat,
                   const double* __restrict__ Kinv, int N, int ldk, double* __restrict__ out) {
  __shared__ double red[32];
  double s0 = 0, s1 = 0, s2 = 0, s3 = 0, s4 = 0;
  for (int i = threadIdx.x; i < N; i += 256) {
    const double al = alpha[i];
    s0 = fma(resid[i], al, s0);
    s1 += log(Lmat[(size_t)i * ldk + i]);
    s2 += al;
    s3 = fma(al, al, s3);
    s4 += Kinv[(size_t)i * ldk + i];
  }
  double v[5] = {s0, s1, s2, s3, s4};
  for (int k = 0; k < 5; ++k) {
    double t = block_sum(v[k], red);
    if (threadIdx.x == 0) out[k] = t;
    __syncthreads();
  }
}

int launch_mll_grad(const ModelD& md, const PrepD& train, int N, int ldk, const double* alpha, const double* Kinv,
                    const int* ls_offset, int n_ls_total, double* part, double* out_params, cudaStream_t st, LaunchCounter* lc) {
  MllGradArgs a;
  a.md = md; a.train = train; a.N = N; a.ldk = ldk; a.alpha = alpha; a.Kinv = Kinv;
  for (int l = 0; l < BO_MAX_LEAVES; ++l) a.ls_offset[l] = (l < md.n_leaves) ? ls_offset[l] : -1;
  a.n_ls_total = n_ls_total; a.n_params = n_ls_total + md.n_terms; a.part = part;
  int dpad_max = 4;
  for (int l = 0; l < md.n_leaves; ++l)
    if (md.leaf[l].kind <= BO_LEAF_MATERN52) dpad_max = std::max(dpad_max, md.leaf[l].dpad);
  size_t smem = ((size_t)2 * ML_CHUNK + 768 + dpad_max + a.n_params + 32) * sizeof(double);
  if (smem > 200 * 1024) { bo_set_error("mll_grad: too many hyper-parameters for shared memory"); return BO_ERR_INVALID; }
  static PerDeviceMax attr_pd; size_t& attr = attr_pd.slot();
  if (smem > 48 * 1024 && smem > attr) {
    CUDA_CHECK_RET(cudaFuncSetAttribute(mll_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  mll_grad_kernel<<<N, 256, smem, st>>>(a, dpad_max);
  if (lc) lc->n++;
  mll_reduce_kernel<<<(a.n_params + 127) / 128, 128, 0, st>>>(part, N, a.n_params, out_params);
  if (lc) lc->n++;
  CUDA_CHECK_RET(cudaGetLastError());
  return BO_OK;
}

int launch_mll_scalars(const double* resid, const double* alpha, const double* Lmat, const double* Kinv, int N, int ldk,
                       double* out5, cudaStream_t st, LaunchCounter* lc) {
  mll_scalars_kernel<<<1, 256, 0, st>>>(resid, alpha, Lmat, Kinv, N, ldk, out5);
  if (lc) lc->n++;
  CUDA_CHECK_RET(cudaGetLastError());
  return BO_OK;
}

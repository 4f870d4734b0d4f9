// Analytic adjoint of the acquisition value with respect to the candidate points: the backward pass of
// AcquisitionFunction.forward(X[b, q, d]) that BoTorch gets from autograd inside gen_candidates_scipy
// (reached from BotorchStrategy._optimize_acqf_continuous, strategies/predictives/botorch.py:384-405).
//
// Chain (per q-batch, per output m), mirroring the forward kernels in acqf.cu / gemm.cu / kernels_eval.cu:
//   value  <- MC samples f = mu + bl z_b + br z_q            mc_hvi_grad_kernel / mc_logei_grad_kernel  (d value / d f)
//   f      <- mu, bl, br                                     grad_reduce_kernel      (sums over the MC samples)
//   br     <- chol(Sqq - bl bl^T), bl <- Sqb L_b^-T          cond_root_bwd_kernel    (Cholesky adjoint, Murray 2016)
//   Sqq, Sqb, mu <- K*X (via G = K*X K^-1 K*X^T, W = K*X K^-1 K_Xb, K*X alpha) and the prior blocks K**, K*b
//   K(x, .) <- x                                             kernel_grad_kernel      (RBF / Matern leaves; Hamming and
//                                                                                     Tanimoto columns are discrete: zero)
// Everything is float64; reductions run in a fixed order (deterministic).
#include "acqf.cuh"
#include "common.cuh"
#include "mc_math.cuh"

#include <algorithm>

// ------------------------------------------------------------------------------------------------
// MC value and d value / d f for qNEHVI / qEHVI: one CTA per (q-batch, slice of the MC samples).  A CTA holds `nst` samples
// at a time (lanes = consecutive samples: coalesced cell reads) and `CL` "cell lanes" per sample: thread (sl, cl) walks the
// cells c = cl, cl + CL, ... of its sample and accumulates its own d value / d objective block; the blocks of a sample are
// added in cell-lane order afterwards (fixed order: deterministic).  Refinement calls have 8 q-batches x 512 samples; with
// thousands of cells per sample (4 objectives) the cells are the only parallelism left.
// ------------------------------------------------------------------------------------------------
template <int MO>   // objectives at compile time (2..4: side lengths, bounds and arg-mins stay in registers), 0 = run time
__global__ void __launch_bounds__(256)
mc_hvi_grad_kernel(McArgs a, double* __restrict__ dF, size_t df_stride, int CL) {
  extern __shared__ double gsm[];
  const int batch = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  const int nst = nt / CL, sl = tid % nst, cl = tid / nst;
  const int q = a.q, nb = a.nb, nr = nb + q, M = a.M, S = a.S, Mo = MO > 0 ? MO : a.od.n_obj;
  double* root = gsm;                          // [M][q][nr]
  double* mu = root + (size_t)M * q * nr;      // [q][M]
  double* objs = mu + q * M;                   // [q*Mo][nst]
  double* fw = objs + (size_t)q * Mo * nst;    // [q][nst]
  double* ys = fw + (size_t)q * nst;           // [q*M][nst]   model-output samples
  double* gob = ys + (size_t)q * M * nst;      // [q*Mo][nt]  d value / d objective (per thread)
  double* gfw = gob + (size_t)q * Mo * nt;     // [q][nt]     d value / d feasibility weight (per thread)
  double* red = gfw + (size_t)q * nt;          // [32]
  for (int i = tid; i < M * q * nr; i += nt) root[i] = a.root[(size_t)batch * M * q * nr + i];
  for (int i = tid; i < q * M; i += nt) mu[i] = a.mu[(size_t)batch * q * M + i];
  const double invS = 1.0 / (double)S;
  const bool has_cons = a.od.n_cons > 0;

  double total = 0.0;
  // the MC samples are split over gridDim.y CTAs (few q-batches: more CTAs than q-batches); partial sums per split
  const int per_split = (S + gridDim.y - 1) / gridDim.y;
  const int s_begin = blockIdx.y * per_split, s_end = min(S, s_begin + per_split);
  for (int s0 = s_begin; s0 < s_end; s0 += nst) {
    const int s = s0 + sl;
    const bool valid = s < s_end;
    __syncthreads();   // staging done / the previous pass has been read
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
        double w = 1.0;
        for (int c = 0; c < a.od.n_cons; ++c) {
          double cv = a.od.con[c].sign * (y[a.od.con[c].out_idx] - a.od.con[c].tp);
          w *= 1.0 / (1.0 + exp(cv / a.od.con[c].eta));
        }
        fw[(size_t)j * nst + sl] = w;
      }
    for (int i = 0; i < q * Mo; ++i) gob[(size_t)i * nt + tid] = 0.0;
    for (int j = 0; j < q; ++j) gfw[(size_t)j * nt + tid] = 0.0;
    __syncthreads();
    if (valid) {
      const int nc = a.cells_shared ? a.ncells[0] : a.ncells[s];
      const int sc = a.cells_shared ? 0 : s;
      const int Sc = a.cells_shared ? 1 : S;
      double acc = 0.0;
      for (int c = cl; c < nc; c += CL) {
        double lo[BO_MAX_OBJECTIVES], up[BO_MAX_OBJECTIVES];
#pragma unroll
        for (int o = 0; o < Mo; ++o) {
          lo[o] = a.cell_lo[((size_t)c * Mo + o) * Sc + sc];
          up[o] = a.cell_up[((size_t)c * Mo + o) * Sc + sc];
        }
        unsigned active = 0;
        for (int j = 0; j < q; ++j) {
          bool pos = true;
#pragma unroll
          for (int o = 0; o < Mo; ++o) {
            double len = fmin(objs[((size_t)j * Mo + o) * nst + sl], up[o]) - lo[o];
            pos = pos && (len > 0.0);
          }
          if (pos) active |= (1u << j);
        }
        if (!active) continue;
        double cell = 0.0;
        for (int size = 1; size <= q; ++size) {
          double asum = 0.0;
          bool any = false;
          const double sgn = (size & 1) ? 1.0 : -1.0;
          for (unsigned sub = active; sub; sub = (sub - 1) & active) {
            if (__popc(sub) != size) continue;
            any = true;
            double vol = 1.0, wprod = 1.0;
            double len[BO_MAX_OBJECTIVES];
            int arg[BO_MAX_OBJECTIVES];
#pragma unroll
            for (int o = 0; o < Mo; ++o) {
              double mn = up[o];
              int aj = -1;
              for (unsigned rest = sub; rest; rest &= rest - 1) {
                int j = __ffs(rest) - 1;
                double v = objs[((size_t)j * Mo + o) * nst + sl];
                if (v < mn) { mn = v; aj = j; }
              }
              len[o] = fmax(mn - lo[o], 0.0);  // > 0: every point of an active subset overlaps the cell
              arg[o] = aj;
              vol *= len[o];
            }
            if (has_cons)
              for (unsigned rest = sub; rest; rest &= rest - 1) wprod *= fw[(size_t)(__ffs(rest) - 1) * nst + sl];
            // adjoint: the side length along o moves with the subset's minimum iff that minimum is below the cell's upper bound
#pragma unroll
            for (int o = 0; o < Mo; ++o)
              if (arg[o] >= 0) {
                double other = 1.0;
#pragma unroll
                for (int o2 = 0; o2 < Mo; ++o2)
                  if (o2 != o) other *= len[o2];
                gob[((size_t)arg[o] * Mo + o) * nt + tid] += sgn * wprod * other;
              }
            if (has_cons) {
              for (unsigned rest = sub; rest; rest &= rest - 1) {
                const int j = __ffs(rest) - 1;
                double others = 1.0;
                for (unsigned r2 = sub; r2; r2 &= r2 - 1) {
                  const int j2 = __ffs(r2) - 1;
                  if (j2 != j) others *= fw[(size_t)j2 * nst + sl];
                }
                gfw[(size_t)j * nt + tid] += sgn * vol * others;
              }
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
    __syncthreads();
    // objectives / feasibility -> model outputs (point j = cl, cl + CL, ...; the cell lanes' blocks added in lane order)
    if (valid)
      for (int j = cl; j < q; j += CL) {
        double y[2 * BO_MAX_OBJECTIVES];
        for (int m = 0; m < M; ++m) y[m] = ys[((size_t)j * M + m) * nst + sl];
        double dy[2 * BO_MAX_OBJECTIVES];
        for (int m = 0; m < M; ++m) dy[m] = 0.0;
        for (int o = 0; o < Mo; ++o) {
          double gv = gob[((size_t)j * Mo + o) * nt + sl];
          for (int c2 = 1; c2 < CL; ++c2) gv += gob[((size_t)j * Mo + o) * nt + c2 * nst + sl];
          if (gv != 0.0) dy[a.od.op[o].out_idx] += gv * objective_grad(a.od.op[o], y);
        }
        if (has_cons) {
          double gw = gfw[(size_t)j * nt + sl];
          for (int c2 = 1; c2 < CL; ++c2) gw += gfw[(size_t)j * nt + c2 * nst + sl];
          if (gw != 0.0) {
            double sg[BO_MAX_CONSTRAINTS];
            for (int c = 0; c < a.od.n_cons; ++c) {
              double cv = a.od.con[c].sign * (y[a.od.con[c].out_idx] - a.od.con[c].tp);
              sg[c] = 1.0 / (1.0 + exp(cv / a.od.con[c].eta));
            }
            for (int c = 0; c < a.od.n_cons; ++c) {
              double rest = 1.0;
              for (int c2 = 0; c2 < a.od.n_cons; ++c2)
                if (c2 != c) rest *= sg[c2];
              // d sigmoid(-cv / eta) / dy = -sign / eta * sg (1 - sg)
              dy[a.od.con[c].out_idx] += gw * rest * sg[c] * (1.0 - sg[c]) * (-a.od.con[c].sign / a.od.con[c].eta);
            }
          }
        }
        for (int m = 0; m < M; ++m) dF[(size_t)m * df_stride + ((size_t)batch * q + j) * S + s] = dy[m] * invS;
      }
  }
  __syncthreads();
  double t = block_sum(total, red);
  if (tid == 0) a.partial[(size_t)blockIdx.y * a.b + batch] = t;   // summed in a fixed order by mc_reduce_partials_kernel
}

static int pick_threads(size_t fixed_doubles, size_t per_thread_doubles, size_t* smem_out) {
  for (int nt = 256; nt >= 32; nt >>= 1) {
    size_t smem = (fixed_doubles + per_thread_doubles * nt + 32) * sizeof(double);
    if (smem <= 200 * 1024) { *smem_out = smem; return nt; }
  }
  return 0;
}

int launch_mc_hvi_grad(const McArgs& a, int max_cells, double* dF, size_t df_stride, cudaStream_t st, LaunchCounter* lc) {
  if (a.b <= 0) return BO_OK;
  const int Mo = a.od.n_obj;
  if (!a.partial) { bo_set_error("mc_hvi_grad: partial-sum workspace missing"); return BO_ERR_STATE; }
  // about two waves of CTAs: split the MC samples when there are few q-batches (partial has room for ceil(S / 32) splits)
  int nsplit = (296 + a.b - 1) / a.b;
  nsplit = std::max(1, std::min(nsplit, (a.S + 31) / 32));
  const int per_split = (a.S + nsplit - 1) / nsplit;
  // shared memory: roots + means, per-sample values (objectives, feasibility weights, model outputs), per-thread adjoints
  const size_t fixed = (size_t)a.M * a.q * (a.nb + a.q) + (size_t)a.q * a.M + 32;
  const size_t per_sample = (size_t)a.q * Mo + a.q + (size_t)a.q * a.M;
  const size_t per_thread = (size_t)a.q * Mo + a.q;
  int nst = std::min(256, ((per_split + 31) / 32) * 32);
  // cell lanes per sample: worth it when a sample has many cells and the samples alone do not fill the CTA
  int CL = (max_cells >= 64) ? 8 : (max_cells >= 8) ? 4 : 1;
  CL = std::max(1, std::min(CL, 256 / nst));
  size_t smem = 0;
  for (;;) {
    smem = (fixed + per_sample * nst + per_thread * (size_t)nst * CL) * sizeof(double);
    if (smem <= 200 * 1024) break;
    if (CL > 1) CL >>= 1;
    else if (nst > 32) nst >>= 1;
    else { bo_set_error("mc_hvi_grad: shared memory budget exceeded (n_b=%d q=%d)", a.nb, a.q); return BO_ERR_INVALID; }
  }
  void (*kern)(McArgs, double*, size_t, int) = (Mo == 2) ? mc_hvi_grad_kernel<2> : (Mo == 3) ? mc_hvi_grad_kernel<3>
                                              : (Mo == 4) ? mc_hvi_grad_kernel<4> : mc_hvi_grad_kernel<0>;
  static PerDeviceMax attr_pd[4];
  size_t& attr = attr_pd[(Mo >= 2 && Mo <= 4) ? Mo - 1 : 0].slot();
  if (smem > 48 * 1024 && smem > attr) {
    CUDA_CHECK_RET(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  dim3 grid(a.b, nsplit);
  kern<<<grid, nst * CL, smem, st>>>(a, dF, df_stride, CL);
  if (lc) lc->n++;
  CUDA_CHECK_RET(cudaGetLastError());
  return launch_mc_reduce_partials(a.partial, nsplit, a.b, a.S, a.out, a.info_in, a.M, a.info_out, st, lc);
}

// ------------------------------------------------------------------------------------------------
// d root = [ dF z_b | dF z_q ],  d mu = sum_s dF   (one CTA per (candidate point, output); warps over targets)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
grad_reduce_kernel(const double* __restrict__ dF, size_t df_stride, const double* __restrict__ zbT,
                   const double* __restrict__ zqT, int S, int nb, int q, int M, double* __restrict__ droot,
                   double* __restrict__ dmu) {
  extern __shared__ double rsm[];
  const int row = blockIdx.x, m = blockIdx.y, batch = row / q, j = row % q, nr = nb + q;
  const double* src = dF + (size_t)m * df_stride + (size_t)row * S;
  for (int s = threadIdx.x; s < S; s += blockDim.x) rsm[s] = src[s];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int t = warp; t <= nr; t += nw) {
    const double* z = (t < nb) ? zbT + ((size_t)t * M + m) * S : (t < nr) ? zqT + ((size_t)(t - nb) * M + m) * S : nullptr;
    double acc = 0.0;
    if (z) for (int s = lane; s < S; s += 32) acc = fma(rsm[s], z[s], acc);
    else for (int s = lane; s < S; s += 32) acc += rsm[s];
    acc = warp_sum(acc);
    if (lane == 0) {
      if (t < nr) droot[(((size_t)batch * M + m) * q + j) * nr + t] = acc;
      else dmu[(size_t)row * M + m] = acc;
    }
  }
}

int launch_grad_reduce(const double* dF, size_t df_stride, const double* zbT, const double* zqT, int S, int nb, int q,
                       int M, int rows, double* droot, double* dmu, cudaStream_t st, LaunchCounter* lc) {
  if (rows <= 0) return BO_OK;
  size_t smem = (size_t)S * sizeof(double);
  if (smem > 200 * 1024) { bo_set_error("grad_reduce: too many MC samples for shared memory (S=%d)", S); return BO_ERR_INVALID; }
  static PerDeviceMax attr_pd; size_t& attr = attr_pd.slot();
  if (smem > 48 * 1024 && smem > attr) {
    CUDA_CHECK_RET(cudaFuncSetAttribute(grad_reduce_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  dim3 grid(rows, M);
  grad_reduce_kernel<<<grid, 256, smem, st>>>(dF, df_stride, zbT, zqT, S, nb, q, M, droot, dmu);
  if (lc) lc->n++;
  CUDA_CHECK_RET(cudaGetLastError());
  return BO_OK;
}

// ------------------------------------------------------------------------------------------------
// Adjoint of cond_root_kernel: one warp per q-batch (output m).
//   br = chol(C), C = Sqq - bl bl^T, bl = Sqb L_b^-T, Sqq = s^2 (K** - G), Sqb = s^2 (K*b - W), mu = y_std (c + K*X alpha) + y_mean
// Cholesky adjoint: Cbar = sym( L^-T Phi(L^T Lbar) L^-1 ), Phi = lower triangle with halved diagonal.
// ------------------------------------------------------------------------------------------------
template <bool wide>
__global__ void __launch_bounds__(128)
cond_root_bwd_kernel(CondRootBwdArgs a, int warps_per_cta) {
  // Two layouts (uniform over the grid): one warp per q-batch, `warps_per_cta` q-batches per CTA -- or, `wide`, the four
  // warps of a CTA on one q-batch with the 32-column blocks of d Sqb dealt over the warps AND over gridDim.y CTAs (every
  // CTA of a q-batch repeats the small q x q adjoint).  Refinement calls have 8 q-batches: one warp each left the 286 x 286
  // inverse root of a 4-objective problem to 8 warps of the whole GPU.
  extern __shared__ double bsm[];
  const int wic = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int batch = wide ? blockIdx.x : blockIdx.x * warps_per_cta + wic;
  if (!wide && (wic >= warps_per_cta || batch >= a.b)) return;
  const int tl = wide ? (int)threadIdx.x : lane;
  constexpr int tn = wide ? 128 : 32;
  const int wsub = wide ? wic : 0;
  constexpr int nw = wide ? 4 : 1;
#define TEAM_SYNC() do { if (wide) __syncthreads(); else __syncwarp(); } while (0)
  const int q = a.q, nb = a.nb, nr = nb + q;
  double* T = bsm + (wide ? 0 : (size_t)wic * ((size_t)q * nb + 5 * q * q));  // [q][nb] total adjoint of bl
  double* Lq = T + (size_t)q * nb;                               // [q][q]
  double* Lbar = Lq + q * q;
  double* Pm = Lbar + q * q;
  double* Li = Pm + q * q;
  double* Cb = Li + q * q;
  const double* root = a.root + ((size_t)batch * a.M + a.m) * q * nr;
  const double* droot = a.droot + ((size_t)batch * a.M + a.m) * q * nr;
  const double s2 = a.y_std * a.y_std;
  const int row0 = batch * q;
  if (wsub == 0) {
    // the q x q Cholesky adjoint on one warp, one matrix entry (or one column of L^-1) per lane; every entry is the same
    // ordered dot product a single thread would form
    for (int p = lane; p < q * q; p += 32) {
      int i = p / q, j = p % q;
      Lq[p] = (j <= i) ? root[i * nr + nb + j] : 0.0;
      Lbar[p] = (j <= i) ? droot[i * nr + nb + j] : 0.0;
    }
    __syncwarp();
    // P = Phi(L^T Lbar)
    for (int p = lane; p < q * q; p += 32) {
      const int i = p / q, j = p % q;
      double s = 0.0;
      if (j <= i)
        for (int k = i; k < q; ++k) s = fma(Lq[k * q + i], Lbar[k * q + j], s);
      Pm[p] = (j > i) ? 0.0 : (i == j) ? 0.5 * s : s;
    }
    // Li = L^-1 (lower): forward substitution down column c
    for (int c = lane; c < q; c += 32)
      for (int i = 0; i < q; ++i) {
        if (i < c) { Li[i * q + c] = 0.0; continue; }
        double s = (i == c) ? 1.0 : 0.0;
        for (int k = c; k < i; ++k) s -= Lq[i * q + k] * Li[k * q + c];
        Li[i * q + c] = s / Lq[i * q + i];
      }
    __syncwarp();
    // Cb = Li^T P Li, then symmetrise.  tmp = P Li (lower x lower = lower) into Lbar (no longer needed)
    for (int p = lane; p < q * q; p += 32) {
      const int i = p / q, j = p % q;
      double s = 0.0;
      for (int k = j; k <= i; ++k) s = fma(Pm[i * q + k], Li[k * q + j], s);
      Lbar[p] = s;
    }
    __syncwarp();
    for (int p = lane; p < q * q; p += 32) {
      const int i = p / q, j = p % q;
      double s = 0.0;
      for (int k = i; k < q; ++k) s = fma(Li[k * q + i], Lbar[k * q + j], s);
      Cb[p] = s;
    }
    __syncwarp();
    for (int p = lane; p < q * q; p += 32) {
      const int i = p / q, j = p % q;
      if (j < i) {
        double v = 0.5 * (Cb[i * q + j] + Cb[j * q + i]);
        Cb[i * q + j] = v;
        Cb[j * q + i] = v;
      }
    }
  }
  TEAM_SYNC();
  // T = d bl - 2 Cbar bl
  for (int e = tl; e < nb; e += tn)
    for (int j = 0; j < q; ++j) {
      double s = droot[j * nr + e];
      for (int i = 0; i < q; ++i) s = fma(-2.0 * Cb[j * q + i], root[i * nr + e], s);
      T[(size_t)j * nb + e] = s;
    }
  TEAM_SYNC();
  // d Sqb[j][l] = sum_{e >= l} T[j][e] LbInv[e][l];  EW = d value / d W = -s^2 d Sqb.  Blocks of 32 columns l on the FP64
  // tensor pipe (DMMA m8n8k4): A = up to 8 points of the q-batch x 4 rows e of T (shared memory), B[k = e][n = l] =
  // LbInv[e][l] read straight from L2; the rows above a block's first column contribute nothing and are skipped.
  {
    const int g = lane >> 2, t = lane & 3;
    for (int l0 = (blockIdx.y * nw + wsub) * 32; l0 < nb; l0 += 32 * nw * gridDim.y) {
      for (int j0 = 0; j0 < q; j0 += 8) {
        double acc[4][2];
#pragma unroll
        for (int n = 0; n < 4; ++n) acc[n][0] = acc[n][1] = 0.0;
        const bool arow = (j0 + g) < q;
        const double* tq = T + (size_t)(arow ? j0 + g : 0) * nb;
#pragma unroll 4
        for (int e0 = l0; e0 < nb; e0 += 4) {
          const int e = e0 + t;
          const bool kin = e < nb;
          const double av = (arow && kin) ? tq[e] : 0.0;
          const double* lrow = a.LbInv + (size_t)(kin ? e : 0) * a.ldlb;
          double bv[4];
#pragma unroll
          for (int n = 0; n < 4; ++n) {
            const int l = l0 + n * 8 + g;
            bv[n] = (kin && l < nb && e >= l) ? lrow[l] : 0.0;
          }
#pragma unroll
          for (int n = 0; n < 4; ++n) mma_884(acc[n][0], acc[n][1], av, bv[n]);
        }
        if (arow) {
#pragma unroll
          for (int n = 0; n < 4; ++n)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int l = l0 + n * 8 + 2 * t + h;
              if (l < nb) a.EW[(size_t)(row0 + j0 + g) * a.ldw + l] = -s2 * acc[n][h];
            }
        }
      }
    }
  }
  if (blockIdx.y == 0) {
    for (int p = tl; p < q * q; p += tn) a.EG[(size_t)batch * q * q + p] = -2.0 * s2 * Cb[p];
    for (int j = tl; j < q; j += tn) a.Emu[row0 + j] = a.y_std * a.dmu[(size_t)(row0 + j) * a.M + a.m];
  }
#undef TEAM_SYNC
}

int launch_cond_root_bwd(const CondRootBwdArgs& a, cudaStream_t st, LaunchCounter* lc) {
  if (a.b <= 0) return BO_OK;
  size_t per_warp = ((size_t)a.q * a.nb + 5 * a.q * a.q) * sizeof(double);
  // few q-batches (refinement) or a large baseline: the whole CTA on one q-batch, column blocks split over up to 4 CTAs
  const int wide = (a.nb > 64 || a.b <= 148) ? 1 : 0;
  int wpc = wide ? 1 : ((4 * per_warp <= 96 * 1024) ? 4 : 1);
  size_t smem = wpc * per_warp;
  if (smem > 200 * 1024) { bo_set_error("cond_root_bwd: baseline too large for shared memory (n_b=%d)", a.nb); return BO_ERR_INVALID; }
  static PerDeviceMax attr_pd; size_t& attr = attr_pd.slot();
  if (smem > 48 * 1024 && smem > attr) {
    CUDA_CHECK_RET(cudaFuncSetAttribute(cond_root_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CUDA_CHECK_RET(cudaFuncSetAttribute(cond_root_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  int ysplit = 1;
  if (wide) {
    const int blocks = (a.nb + 31) / 32;
    ysplit = std::max(1, std::min((blocks + 3) / 4, std::max(1, 296 / a.b)));
  }
  dim3 grid(wide ? a.b : (a.b + wpc - 1) / wpc, ysplit);
  if (wide) cond_root_bwd_kernel<true><<<grid, 128, smem, st>>>(a, wpc);
  else cond_root_bwd_kernel<false><<<grid, 128, smem, st>>>(a, wpc);
  if (lc) lc->n++;
  CUDA_CHECK_RET(cudaGetLastError());
  return BO_OK;
}

// ------------------------------------------------------------------------------------------------
// d value / d x for one candidate point per CTA: partners = N training points (weights from mu, W, G),
// n_b baseline points (K*b) and the other points of the same q-batch (K**).
// ------------------------------------------------------------------------------------------------
#define KG_CHUNK 2048

__device__ __forceinline__ double leaf_dk_dstat(int kind, double stat) {
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
kernel_grad_kernel(const __grid_constant__ KernelGradArgs a, int dpad_max) {
  extern __shared__ double ksm[];
  const int row = blockIdx.x, tid = threadIdx.x;
  const int q = a.q, nb = a.nb, N = a.N, batch = row / q, jj = row % q;
  const ModelD& md = a.md;
  double* w = ksm;               // [KG_CHUNK] d value / d K(x_row, partner)
  double* g = w + KG_CHUNK;      // [KG_CHUNK] w * dK/dstat for the current leaf
  double* part = g + KG_CHUNK;   // [2][256]
  double* xsj = part + 512;      // [dpad_max]
  double* dxrow = xsj + dpad_max;  // [d]
  for (int c = tid; c < a.d; c += 256) dxrow[c] = 0.0;
  const int P = N + nb + q;
  const double emu = a.Emu[row];
  const double* ew = a.EW + (size_t)row * a.ldw;
  const double* eg = a.EG + ((size_t)batch * q + jj) * q;
  bool any_cont = false;
  for (int l = 0; l < md.n_leaves; ++l) any_cont = any_cont || (md.leaf[l].kind <= BO_LEAF_MATERN52);
  // single RBF leaf (SingleTaskGP default): dK/dstat = -K/2 with K already in the forward K(X*, X) buffer
  const bool fast_rbf = a.Kx && md.n_terms == 1 && md.nfac[0] == 1 && md.leaf[md.fac[0][0]].kind == BO_LEAF_RBF;

  for (int p0 = 0; any_cont && p0 < P; p0 += KG_CHUNK) {
    const int cn = min(KG_CHUNK, P - p0);
    __syncthreads();
    for (int pl = tid; pl < cn; pl += 256) {
      const int p = p0 + pl;
      double wv;
      if (p < N) {
        wv = emu * a.alpha[p];
        for (int e = 0; e < nb; ++e) wv = fma(ew[e], a.Aext[(size_t)e * a.ldk + p], wv);
        for (int i = 0; i < q; ++i) wv = fma(eg[i], a.U[((size_t)batch * q + i) * a.ldk + p], wv);
      } else if (p < N + nb) {
        wv = -ew[p - N];
      } else {
        const int i = p - N - nb;
        wv = (i == jj) ? 0.0 : -eg[i];
      }
      w[pl] = wv;
    }
    for (int l = 0; l < md.n_leaves; ++l) {
      const LeafD& L = md.leaf[l];
      if (L.kind > BO_LEAF_MATERN52) continue;
      __syncthreads();
      for (int k = tid; k < L.dpad; k += 256) xsj[k] = a.prep_q.Xs[l][(size_t)row * L.dpad + k];
      __syncthreads();
      const double n2j = a.prep_q.n2[l][row];
      for (int pl = tid; pl < cn; pl += 256) {
        const int p = p0 + pl;
        const double wv = w[pl];
        double gv = 0.0;
        if (wv != 0.0 && fast_rbf && p < N) {
          gv = wv * (-0.5) * a.Kx[(size_t)row * a.ldk + p];
        } else if (wv != 0.0) {
          // partner side / index
          const PrepD* pp;
          int pi;
          bool train = false;
          if (p < N) { train = true; pp = nullptr; pi = p; }
          else if (p < N + nb) { pp = &a.prep_b; pi = p - N; }
          else { pp = &a.prep_q; pi = batch * q + (p - N - nb); }
          // coefficient of this leaf in the sum-of-products tree
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
              const LeafSide B = train ? train_side(md.leaf[lf]) : prep_side(*pp, lf);
              prod *= leaf_eval_pair(md.leaf[lf], prep_side(a.prep_q, lf), row, B, pi, false);
            }
            cl += prod;
          }
          const double* xb = (train ? L.Xs : pp->Xs[l]) + (size_t)pi * L.dpad;
          const double n2b = train ? L.n2[pi] : pp->n2[l][pi];
          double dot = 0.0;
          for (int k = 0; k < L.nd; ++k) dot = fma(xsj[k], xb[k], dot);
          const double stat = fmax(n2j + n2b - 2.0 * dot, 0.0);
          gv = wv * cl * leaf_dk_dstat(L.kind, stat);
        }
        g[pl] = gv;
      }
      __syncthreads();
      // d stat / d xs_row[a] = 2 (xs_row[a] - xs_partner[a]):  accumulate sum_p g_p and sum_p g_p xs_p[a]
      for (int a0 = 0; a0 < L.dpad; a0 += 256) {
        const int na = min(256, L.dpad - a0);
        const int nsl = max(1, 256 / na);
        const int ai = tid % na, sl = tid / na;
        double accx = 0.0, accg = 0.0;
        if (sl < nsl) {
          // three partner segments with a fixed base pointer each; g = 0 entries simply add nothing, so the loops carry
          // no branch and the (L2-resident) row loads of 8 iterations are in flight together
          const int n_tr = max(0, min(cn, N - p0));                 // training partners in this chunk
          {
            const double* xb = L.Xs + (size_t)p0 * L.dpad + a0 + ai;
#pragma unroll 8
            for (int pl = sl; pl < n_tr; pl += nsl) {
              const double gv = g[pl];
              accx = fma(gv, xb[(size_t)pl * L.dpad], accx);
              accg += gv;
            }
          }
          for (int pl = n_tr + ((sl - n_tr % nsl + nsl) % nsl); pl < cn; pl += nsl) {   // baseline / same-batch partners
            const double gv = g[pl];
            if (gv == 0.0) continue;
            const int p = p0 + pl;
            const double* xb = (p < N + nb) ? a.prep_b.Xs[l] + (size_t)(p - N) * L.dpad
                                            : a.prep_q.Xs[l] + (size_t)(batch * q + (p - N - nb)) * L.dpad;
            accx = fma(gv, xb[a0 + ai], accx);
            accg += gv;
          }
        }
        part[tid] = accx;
        part[256 + tid] = accg;
        __syncthreads();
        if (tid < na) {
          double sx = 0.0, sg = 0.0;
          for (int s = 0; s < nsl; ++s) { sx += part[s * na + tid]; sg += part[256 + s * na + tid]; }
          const int ak = a0 + tid;
          if (ak < L.nd) dxrow[L.col[ak]] += 2.0 * (xsj[ak] * sg - sx) / (L.in_scl[ak] * L.ls[ak]);
        }
        __syncthreads();
      }
    }
  }
  __syncthreads();
  for (int c = tid; c < a.d; c += 256) {
    double* dst = a.dX + (size_t)row * a.d + c;
    *dst = (a.accumulate ? *dst : 0.0) + dxrow[c];
  }
}

int launch_kernel_grad(const KernelGradArgs& a, cudaStream_t st, LaunchCounter* lc) {
  if (a.rows <= 0) return BO_OK;
  int dpad_max = 4;
  for (int l = 0; l < a.md.n_leaves; ++l)
    if (a.md.leaf[l].kind <= BO_LEAF_MATERN52) dpad_max = std::max(dpad_max, a.md.leaf[l].dpad);
  size_t smem = ((size_t)2 * KG_CHUNK + 512 + dpad_max + a.d) * sizeof(double);
  if (smem > 200 * 1024) { bo_set_error("kernel_grad: input dimension too large for shared memory (d=%d)", a.d); return BO_ERR_INVALID; }
  static PerDeviceMax attr_pd; size_t& attr = attr_pd.slot();
  if (smem > 48 * 1024 && smem > attr) {
    CUDA_CHECK_RET(cudaFuncSetAttribute(kernel_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  kernel_grad_kernel<<<a.rows, 256, smem, st>>>(a, dpad_max);
  if (lc) lc->n++;
  CUDA_CHECK_RET(cudaGetLastError());
  return BO_OK;
}

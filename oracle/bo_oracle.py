"""CPU float64 oracle for BoFire's acquisition-evaluation hot path.

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.  Nothing in the product
package imports this file.

What it restates (reference = /root/reference, BoFire):

* in-tree arithmetic, followed line by line and PINNED by golden vectors generated
  from the reference's own code (``tests/golden/gen_reference_golden.py``):
  Tanimoto (``bofire/kernels/fingerprint_kernels/base_fingerprint_kernel.py:36-53,81-84``),
  one-hot Hamming (``bofire/kernels/categorical.py:43-70``), objective / constraint
  callables (``bofire/utils/torch_tools.py:258-337,384-450,662-727``), ref-point helpers
  (``bofire/utils/multiobjective.py:18-55,133-159``, ``strategies/predictives/qehvi.py:87-110``),
  benchmark functions (``bofire/benchmarks/multi.py:95-132``, ``single.py:409-446``,
  ``detergent.py:10-88``).
* third-party arithmetic the reference calls but does not vendor: ``botorch>=0.13``
  (pyproject.toml:34) and transitively gpytorch / linear_operator.  Neither is
  installed here and there is no network, so these parts follow the published
  algorithms and are **PARITY UNPINNED** against real BoTorch:
  RBF / Matern distance formulas, exact-GP posterior with cached ``L^-T`` root,
  cached-root conditional sampling (``sample_cached_cholesky``), ``psd_safe_cholesky``
  jitter escalation, qNEHVI / qEHVI inclusion-exclusion, Lacour-2017 box decomposition,
  ``prune_inferior_points_multi_objective``, qLogEI (``log_fatplus`` / ``fatmax``).
  They are anchored on the reference's call sites (``qnehvi.py:39-52``, ``mobo.py:72-90``,
  ``sobo.py:64-89``) and on the known-answer tests the reference holds
  (``tests/bofire/utils/test_multiobjective.py:75-266``, ``tests/bofire/kernels/test_categorical.py``).
"""

from __future__ import annotations

import itertools
import math
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import torch

DT = torch.float64


# ----------------------------------------------------------------------------------------
# Kernel trees (mirrors bofire/kernels/mapper.py:31-253 -> gpytorch modules)
# ----------------------------------------------------------------------------------------
@dataclass
class RBF:
    dims: Sequence[int]
    lengthscale: Sequence[float]  # 1 value (isotropic) or len(dims) (ARD)


@dataclass
class Matern:
    nu: float
    dims: Sequence[int]
    lengthscale: Sequence[float]


@dataclass
class Hamming:
    """One-hot Hamming kernel. groups = [(start_column, cardinality), ...]."""

    groups: Sequence[Tuple[int, int]]
    lengthscale: Sequence[float]  # 1 value or >= len(groups) values (first F are used)


@dataclass
class Tanimoto:
    dims: Sequence[int]


@dataclass
class Scale:
    base: object
    outputscale: float


@dataclass
class Add:
    children: Sequence[object]


@dataclass
class Mul:
    children: Sequence[object]


def _ls_tensor(ls, n):
    # a tensor is passed through untouched so that autograd can differentiate the marginal likelihood w.r.t. it
    t = ls.to(DT).reshape(-1) if isinstance(ls, torch.Tensor) else torch.as_tensor(list(ls), dtype=DT)
    if t.numel() == 1:
        t = t.expand(n)
    return t[:n]


def _sq_dist(x1, x2, same):
    # [UPSTREAM] gpytorch Kernel.covar_dist / sq_dist: quadratic expansion, clamp_min(0),
    # diagonal forced to zero when x1 is x2.
    n1 = x1.pow(2).sum(-1, keepdim=True)
    n2 = x2.pow(2).sum(-1, keepdim=True)
    res = n1 + n2.transpose(-1, -2) - 2.0 * (x1 @ x2.transpose(-1, -2))
    if same:
        res.diagonal(dim1=-2, dim2=-1).fill_(0.0)
    return res.clamp_min(0.0)


def one_hot_to_codes(X, groups):
    """[UPSTREAM] OneHotToNumeric: argmax over each one-hot group."""
    return torch.stack([X[..., s : s + c].argmax(dim=-1) for s, c in groups], dim=-1)


def eval_kernel(k, X1, X2, center, same=False):
    """K(X1, X2) on transformed inputs. ``center`` is a length-d vector subtracted from the
    continuous columns before scaling (gpytorch subtracts x1.mean(-2); the oracle and the
    device path both use the training-set mean so that the value is batch-independent)."""
    if isinstance(k, (RBF, Matern)):
        dims = list(k.dims)
        ls = _ls_tensor(k.lengthscale, len(dims))
        a = (X1[..., dims] - center[dims]) / ls
        b = (X2[..., dims] - center[dims]) / ls
        d2 = _sq_dist(a, b, same)
        if isinstance(k, RBF):
            return torch.exp(-0.5 * d2)
        r = d2.clamp_min(1e-30).sqrt()
        e = torch.exp(-math.sqrt(2.0 * k.nu) * r)
        if k.nu == 0.5:
            return e
        if k.nu == 1.5:
            return (math.sqrt(3.0) * r + 1.0) * e
        if k.nu == 2.5:
            return (math.sqrt(5.0) * r + 1.0 + (5.0 / 3.0) * r * r) * e
        raise ValueError(k.nu)
    if isinstance(k, Hamming):
        # bofire/kernels/categorical.py:43-70
        c1 = one_hot_to_codes(X1, k.groups)
        c2 = one_hot_to_codes(X2, k.groups)
        delta = (c1.unsqueeze(-2) != c2.unsqueeze(-3)).to(DT)
        ls = _ls_tensor(k.lengthscale, len(k.groups))
        return torch.exp(-(delta / ls).mean(-1))
    if isinstance(k, Tanimoto):
        # base_fingerprint_kernel.py:36-53 (+ clamp_min_ at :83)
        a = X1[..., list(k.dims)]
        b = X2[..., list(k.dims)]
        eps = 1e-6
        dot = a @ b.transpose(-1, -2)
        s1 = (a**2).sum(-1).unsqueeze(-1)
        s2 = (b**2).sum(-1).unsqueeze(-1)
        return ((dot + eps) / (eps + s1 + s2.transpose(-1, -2) - dot)).clamp_min(0.0)
    if isinstance(k, Scale):
        return k.outputscale * eval_kernel(k.base, X1, X2, center, same)
    if isinstance(k, Add):
        out = eval_kernel(k.children[0], X1, X2, center, same)
        for c in k.children[1:]:
            out = out + eval_kernel(c, X1, X2, center, same)
        return out
    if isinstance(k, Mul):
        out = eval_kernel(k.children[0], X1, X2, center, same)
        for c in k.children[1:]:
            out = out * eval_kernel(c, X1, X2, center, same)
        return out
    raise TypeError(type(k))


# ----------------------------------------------------------------------------------------
# psd_safe_cholesky  ([UPSTREAM] linear_operator.utils.cholesky._psd_safe_cholesky)
# ----------------------------------------------------------------------------------------
class NotPSDError(RuntimeError):
    pass


def psd_safe_cholesky(A, max_tries=6, jitter=1e-8, return_jitter=False):
    L, info = torch.linalg.cholesky_ex(A)
    applied = torch.zeros(A.shape[:-2], dtype=DT)
    if not torch.any(info):
        return (L, applied) if return_jitter else L
    if torch.isnan(A).any():
        raise NotPSDError("NaN in input")
    Ap = A.clone()
    prev = 0.0
    for i in range(max_tries):
        new = jitter * (10**i)
        failing = (info > 0).to(DT)
        add = (failing * (new - prev)).unsqueeze(-1).expand(*Ap.shape[:-1])
        Ap.diagonal(dim1=-1, dim2=-2).add_(add)
        applied = applied + failing * (new - prev)
        prev = new
        L, info = torch.linalg.cholesky_ex(Ap)
        if not torch.any(info):
            return (L, applied) if return_jitter else L
    raise NotPSDError(f"not p.d. after jitter {prev:.1e}")


# ----------------------------------------------------------------------------------------
# Exact GP (ModelListGP of independent single-output GPs; botorch_surrogates.py:124-128)
# ----------------------------------------------------------------------------------------
@dataclass
class GPOutput:
    kernel: object
    in_offset: torch.Tensor  # [d]  Normalize: x' = (x - offset) / scale (identity columns: 0 / 1)
    in_scale: torch.Tensor  # [d]
    mean_const: float  # constant mean, standardised space
    noise: float  # homoskedastic noise variance, standardised space
    y: torch.Tensor  # [N] raw targets
    y_mean: float
    y_std: float


def standardize_stats(y):
    """[UPSTREAM] botorch Standardize(m=1): unbiased std, floor 1e-8."""
    y = torch.as_tensor(y, dtype=DT)
    s = y.std(unbiased=True) if y.numel() > 1 else torch.tensor(1.0, dtype=DT)
    s = float(s)
    if not (s >= 1e-8):
        s = 1.0
    return float(y.mean()), s


class GPOracle:
    def __init__(self, X_train, outputs: List[GPOutput]):
        self.X = torch.as_tensor(X_train, dtype=DT)
        self.outputs = outputs
        self.N, self.d = self.X.shape
        self.M = len(outputs)
        self._fact = None

    def _t(self, X, o):
        return (X - o.in_offset) / o.in_scale

    def factorize(self):
        fact = []
        for o in self.outputs:
            Xt = self._t(self.X, o)
            center = Xt.mean(0)
            K = eval_kernel(o.kernel, Xt, Xt, center, same=True)
            G = K + o.noise * torch.eye(self.N, dtype=DT)
            L, jit = psd_safe_cholesky(G, return_jitter=True)
            Linv = torch.linalg.solve_triangular(L, torch.eye(self.N, dtype=DT), upper=False)
            r = (o.y - o.y_mean) / o.y_std - o.mean_const
            alpha = Linv.T @ (Linv @ r)
            fact.append(dict(Xt=Xt, center=center, L=L, Linv=Linv, alpha=alpha, jitter=float(jit)))
        self._fact = fact
        return self

    def cross(self, X, m):
        """K(X, X_train) for output m, standardised space. X raw [n, d]."""
        o, f = self.outputs[m], self._fact[m]
        return eval_kernel(o.kernel, self._t(X, o), f["Xt"], f["center"])

    def prior(self, X1, X2, m, same=False):
        o, f = self.outputs[m], self._fact[m]
        return eval_kernel(o.kernel, self._t(X1, o), self._t(X2, o), f["center"], same=same)

    def posterior(self, X, observation_noise=False):
        """mean [n, M], cov [M, n, n] in the original (un-standardised) outcome space."""
        X = torch.as_tensor(X, dtype=DT)
        n = X.shape[0]
        mean = torch.empty(n, self.M, dtype=DT)
        cov = torch.empty(self.M, n, n, dtype=DT)
        for m, (o, f) in enumerate(zip(self.outputs, self._fact)):
            Ks = self.cross(X, m)
            V = Ks @ f["Linv"].T
            mu = o.mean_const + Ks @ f["alpha"]
            S = self.prior(X, X, m, same=True) - V @ V.T
            if observation_noise:
                S = S + o.noise * torch.eye(n, dtype=DT)
            mean[:, m] = mu * o.y_std + o.y_mean
            cov[m] = S * (o.y_std**2)
        return mean, cov


def log_marginal_likelihood(X_train, out: "GPOutput"):
    """[UPSTREAM] gpytorch ExactMarginalLogLikelihood without the prior terms and the 1/N scaling -- the objective
    SingleTaskGPSurrogate._fit hands to fit_gpytorch_mll (surrogates/single_task_gp.py:69-71).  Hyper-parameters may be
    tensors that require grad (lengthscales, outputscales, noise, mean constant): autograd gives the reference gradient."""
    X = torch.as_tensor(X_train, dtype=DT)
    N = X.shape[0]
    Xt = (X - out.in_offset) / out.in_scale
    center = Xt.mean(0)
    Kmat = eval_kernel(out.kernel, Xt, Xt, center, same=True) + out.noise * torch.eye(N, dtype=DT)
    r = (out.y - out.y_mean) / out.y_std - out.mean_const
    L = torch.linalg.cholesky(Kmat)
    alpha = torch.cholesky_solve(r.unsqueeze(-1), L).squeeze(-1)
    return -0.5 * (r * alpha).sum() - torch.log(torch.diagonal(L)).sum() - 0.5 * N * math.log(2.0 * math.pi)


# ----------------------------------------------------------------------------------------
# Base samples ([UPSTREAM] botorch.sampling.qmc.NormalQMCEngine / draw_sobol_normal_samples)
# ----------------------------------------------------------------------------------------
def draw_sobol_normal_samples(d, n, seed):
    eng = torch.quasirandom.SobolEngine(d, scramble=True, seed=seed)
    u = eng.draw(n, dtype=DT)
    v = 0.5 + (1 - 1e-10) * (u - 0.5)
    return torch.erfinv(2 * v - 1) * math.sqrt(2)


def base_samples_points_by_outputs(n_points, M, S, seed):
    """z[S, n_points, M]; flat Sobol dimension index = m * n_points + i (non-interleaved
    MultitaskMultivariateNormal layout, [UPSTREAM] _reshape_base_samples_non_interleaved)."""
    z = draw_sobol_normal_samples(n_points * M, S, seed)
    return z.view(S, M, n_points).transpose(1, 2).contiguous()


# ----------------------------------------------------------------------------------------
# Objectives and output constraints as op tables (torch_tools.py:258-337, 384-450, 662-727)
# ----------------------------------------------------------------------------------------
def apply_objective_op(op, Y):
    kind, idx = op[0], op[1]
    y = Y[..., idx]
    if kind == "max":
        lo, hi = op[2], op[3]
        return (y - lo) / (hi - lo)
    if kind == "min":
        lo, hi = op[2], op[3]
        return -1.0 * ((y - lo) / (hi - lo))
    if kind == "close_to_target":
        t, p = op[2], op[3]
        return -1.0 * (torch.abs(y - t) ** p)
    if kind == "min_sigmoid":
        st, tp = op[2], op[3]
        return 1.0 - 1.0 / (1.0 + torch.exp(-1.0 * st * (y - tp)))
    if kind == "max_sigmoid":
        st, tp = op[2], op[3]
        return 1.0 / (1.0 + torch.exp(-1.0 * st * (y - tp)))
    if kind == "target":
        t, tol, st = op[2], op[3], op[4]
        return (
            1.0
            / (1.0 + torch.exp(-1 * st * (y - (t - tol))))
            * (1.0 - 1.0 / (1.0 + torch.exp(-1.0 * st * (y - (t + tol)))))
        )
    raise ValueError(kind)


def multi_objective(ops, Y):
    """get_multiobjective_objective (torch_tools.py:699-727): stack of per-output callables."""
    return torch.stack([apply_objective_op(op, Y) for op in ops], dim=-1)


def scalar_objective(spec, Y):
    """('single', op) | ('additive', [(op, w), ...]) | ('multiplicative', [(op, w), ...])
    (torch_tools.py:384-450, 662-697)."""
    mode = spec[0]
    if mode == "single":
        return apply_objective_op(spec[1], Y)
    if mode == "additive":
        val = torch.tensor(0.0, dtype=DT)
        for op, w in spec[1]:
            val = val + apply_objective_op(op, Y) * w
        return val
    if mode == "multiplicative":
        val = torch.tensor(1.0, dtype=DT)
        for op, w in spec[1]:
            val = val * apply_objective_op(op, Y) ** w
        return val
    raise ValueError(mode)


def constraint_values(cons, Y):
    """cons = [(idx, sign, tp, eta)]: c(y) = sign * (y[idx] - tp); feasible iff c <= 0."""
    return [sign * (Y[..., idx] - tp) for idx, sign, tp, _eta in cons]


def smoothed_feasibility(cons, Y):
    """[UPSTREAM] compute_smoothed_feasibility_indicator(fat=False): prod sigmoid(-c/eta)."""
    w = torch.ones(Y.shape[:-1], dtype=DT)
    for (idx, sign, tp, eta), c in zip(cons, constraint_values(cons, Y)):
        w = w * torch.sigmoid(-c / eta)
    return w


def feasibility_indicator(cons, Y):
    """[UPSTREAM] compute_feasibility_indicator: hard feasibility, every c(y) <= 0."""
    ok = torch.ones(Y.shape[:-1], dtype=torch.bool)
    for c in constraint_values(cons, Y):
        ok = ok & (c <= 0)
    return ok


def objective_lower_bound(gp, spec, X, generator=None):
    """[UPSTREAM] botorch.acquisition.utils._estimate_objective_lower_bound: -get_infeasible_cost at 32 random convex
    combinations of X -- the objective of (mean - 6 sd), minimum over the points, clamped at 0 from above.  The convex
    weights come from torch's global generator in BoTorch; `generator` lets both paths share them."""
    X = torch.as_tensor(X, dtype=DT)
    w = torch.rand(32, X.shape[-2], dtype=DT, generator=generator)
    w = w / w.sum(dim=0, keepdim=True)
    Xc = w @ X
    mean, cov = gp.posterior(Xc)
    var = torch.diagonal(cov, dim1=-2, dim2=-1).transpose(0, 1)           # [n, M]
    lb = scalar_objective(spec, mean - 6.0 * var.clamp_min(0.0).sqrt())
    return float(lb.min().clamp_max(0.0))


def best_feasible_objective(gp, spec, cons, samples, obj, X_baseline, generator=None):
    """[UPSTREAM] compute_best_feasible_objective: max over the last dim of `obj` with infeasible entries replaced by -inf
    when every leading index has at least one feasible point, else by the pessimistic lower bound above."""
    if not cons:
        return obj.amax(dim=-1)
    feas = feasibility_indicator(cons, samples)
    if bool(feas.any(dim=-1).all()):
        infeasible_value = -float("inf")
    else:
        infeasible_value = objective_lower_bound(gp, spec, X_baseline, generator=generator)
    return torch.where(feas, obj, torch.full_like(obj, infeasible_value)).amax(dim=-1)


# ----------------------------------------------------------------------------------------
# Pareto / partitioning / hypervolume ([UPSTREAM] botorch.utils.multi_objective)
# ----------------------------------------------------------------------------------------
def is_non_dominated(Y, deduplicate=True):
    """Maximisation. Y [..., n, m] -> bool [..., n]."""
    n = Y.shape[-2]
    if n == 0:
        return torch.zeros(Y.shape[:-1], dtype=torch.bool)
    Y1 = Y.unsqueeze(-3)
    Y2 = Y.unsqueeze(-2)
    dominates = (Y1 >= Y2).all(dim=-1) & (Y1 > Y2).any(dim=-1)
    nd = ~(dominates.any(dim=-1))
    if deduplicate:
        idx = (Y1 == Y2).all(dim=-1).long().argmax(dim=-1)
        keep = torch.zeros_like(nd)
        keep.scatter_(dim=-1, index=idx, value=True)
        return nd & keep
    return nd


def is_non_dominated_chunked(Y, deduplicate=False, max_elems=1 << 27):
    """Same result as is_non_dominated for Y [S, n, m], evaluated a few samples at a time so that the
    S x n x n comparison tensor stays bounded (BoTorch switches to a loop variant for n > 1000)."""
    S, n, m = Y.shape
    step = max(1, int(max_elems // max(1, n * n * m)))
    return torch.cat([is_non_dominated(Y[i : i + step], deduplicate=deduplicate) for i in range(0, S, step)], dim=0)


def pareto_front_above_ref(Y, ref):
    """Unique non-dominated points strictly better than ref in every objective, in the
    original row order.  Returns (points [p, m], row indices [p])."""
    better = (Y > ref).all(dim=-1)
    idx = torch.nonzero(better).view(-1)
    Yb = Y[idx]
    if Yb.shape[0] == 0:
        return Yb, idx
    nd = is_non_dominated(Yb, deduplicate=True)
    return Yb[nd], idx[nd]


def partition_2d(pareto_Y, ref):
    """Non-dominated-space cells for m=2, maximisation.  Front sorted by objective 0 ascending
    (objective 1 descending).  Cell i (i=0..p): lower=(y0[i-1] | ref0, y1[i] | ref1),
    upper=(y0[i] | +inf, +inf).  Returns (lower [p+1, 2], upper [p+1, 2], order [p])."""
    p = pareto_Y.shape[0]
    inf = float("inf")
    lower = torch.empty(p + 1, 2, dtype=DT)
    upper = torch.full((p + 1, 2), inf, dtype=DT)
    if p == 0:
        lower[0] = ref
        return lower, upper, torch.empty(0, dtype=torch.long)
    # stable sort by objective 0 ascending; ties (impossible on a strict front) by row order
    order = torch.argsort(pareto_Y[:, 0], stable=True)
    Ys = pareto_Y[order]
    lower[0, 0] = ref[0]
    lower[1:, 0] = Ys[:, 0]
    lower[:p, 1] = Ys[:, 1]
    lower[p, 1] = ref[1]
    upper[:p, 0] = Ys[:, 0]
    return lower, upper, order


def _lub_update(U, Z, z):
    """[UPSTREAM] compute_local_upper_bounds (Lacour et al. 2017, Alg. 3), minimisation."""
    m = U.shape[-1]
    dominated = (U > z).all(dim=-1)
    A, AZ = U[dominated], Z[dominated]
    P, PZ = [], []
    for j in range(A.shape[0]):
        u, Zu = A[j], AZ[j]
        for k in range(m):
            others = torch.cat([Zu[:k, k], Zu[k + 1 :, k]])
            if z[k] >= others.max():
                nu = u.clone()
                nu[k] = z[k]
                nZ = Zu.clone()
                nZ[k] = z
                P.append(nu)
                PZ.append(nZ)
    keep = ~dominated
    if P:
        U = torch.cat([U[keep], torch.stack(P)], dim=0)
        Z = torch.cat([Z[keep], torch.stack(PZ)], dim=0)
    else:
        U, Z = U[keep], Z[keep]
    return U, Z


def partition_nd(pareto_Y, ref):
    """[UPSTREAM] FastNondominatedPartitioning (alpha=0) for m>=2 via local upper bounds.
    pareto_Y [p, m] (maximisation, any order; processed in the given row order).
    Returns (lower [C, m], upper [C, m])."""
    m = ref.shape[-1]
    inf = float("inf")
    if pareto_Y.shape[0] == 0:
        return ref.view(1, m).clone(), torch.full((1, m), inf, dtype=DT)
    neg_ref = -ref
    U = neg_ref.view(1, m).clone()
    Z = torch.full((1, m, m), -inf, dtype=DT)
    for j in range(m):
        Z[0, j, j] = U[0, j]
    for z in -pareto_Y:
        U, Z = _lub_update(U, Z, z)
    # Lacour Eq. 2 in minimisation space: the search region is the disjoint union over local
    # upper bounds u of  (-inf, u_0) x prod_{j>=1} [max_{k<j} z^k_j(u), u_j).
    C = U.shape[0]
    lo_min = torch.full((C, m), -inf, dtype=DT)
    up_min = U.clone()
    for j in range(1, m):
        lo_min[:, j] = Z[:, :j, j].max(dim=1).values
    empty = (up_min <= lo_min).any(dim=-1)
    lo_min, up_min = lo_min[~empty], up_min[~empty]
    return -up_min, -lo_min


def partition_binary(pareto_Y, ref, alpha=0.0):
    """[UPSTREAM] botorch NondominatedPartitioning._partition_space + get_hypercell_bounds (binary partitioning of Couckuyt
    et al. 2012), what BoTorch's qNEHVI builds per MC sample for alpha > 0 and m > 2 (BoFire passes `alpha`:
    data_models/strategies/predictives/qnehvi.py:19, strategies/predictives/qnehvi.py:50).  Restated from the published
    algorithm as BoTorch implements it, UNPINNED like the rest of the BoTorch arithmetic:
      * minimisation frame; per objective the front is sorted, index 0 = ideal point (min - 1; -inf in the final bounds), index
        p + 1 = anti-ideal point (max + 1; the reference point in the final bounds);
      * LIFO stack of cells [lower idx, upper idx] per objective, start = [0, p + 1]^m;
      * upper corner <= every front point in some objective -> accept; else lower corner <= ... -> if an index edge is longer
        than 1 and volume / total > alpha: halve the longest index edge (first maximum; `round(length / 2)` half-to-even off the
        upper bound, the rest onto the lower bound of the second child), otherwise DROP; else (dominated) drop.
    pareto_Y [p, m] maximisation.  Returns (lower [C, m], upper [C, m]) in the maximisation frame, acceptance order."""
    m = ref.shape[-1]
    inf = float("inf")
    p = pareto_Y.shape[0]
    if p == 0:
        return ref.view(1, m).clone(), torch.full((1, m), inf, dtype=DT)
    neg = -pareto_Y
    outcome = torch.arange(m)
    aug_idcs = torch.cat([torch.zeros(1, m, dtype=torch.long), torch.argsort(neg, dim=0, stable=True) + 1,
                          torch.full((1, m), p + 1, dtype=torch.long)], dim=0)
    ideal = neg.min(dim=0, keepdim=True).values - 1
    anti = neg.max(dim=0, keepdim=True).values + 1
    aug_Y = torch.cat([ideal, neg, anti], dim=0)
    total_volume = (anti - ideal).prod()
    cell = torch.zeros(2, m, dtype=torch.long)
    cell[1] = p + 1
    stack = [cell]
    hyper = []
    while stack:
        cell = stack.pop()
        idcs = aug_idcs[cell, outcome]                 # [2, m] rows of the augmented front
        vals = aug_Y[idcs, outcome]                    # [2, m]
        if bool((vals[1] <= neg).any(dim=1).all()):
            hyper.append(idcs)
        elif bool((vals[0] <= neg).any(dim=1).all()):
            idx_dist = cell[1] - cell[0]
            volume = float((vals[1] - vals[0]).prod())
            if bool((idx_dist > 1).any()) and (volume / float(total_volume)) > alpha:
                length, longest = torch.max(idx_dist, dim=0)
                length, longest = int(length), int(longest)
                n1 = int(round(length / 2.0))
                n2 = length - n1
                for bound, delta in ((1, -n1), (0, n2)):
                    child = cell.clone()
                    child[bound, longest] += delta
                    stack.append(child)
    if not hyper:
        return torch.zeros(0, m, dtype=DT), torch.zeros(0, m, dtype=DT)
    H = torch.stack(hyper)                              # [C, 2, m]
    aug2 = torch.cat([torch.full((1, m), -inf, dtype=DT), neg, (-ref).view(1, m)], dim=0)
    lo_min = aug2[H[:, 0], outcome]
    up_min = aug2[H[:, 1], outcome]
    return -up_min, -lo_min


def hypervolume(pareto_Y, ref):
    """Exact dominated hypervolume (maximisation) by slicing on the last objective."""
    Y = pareto_Y[(pareto_Y > ref).all(dim=-1)]
    if Y.shape[0] == 0:
        return 0.0
    m = Y.shape[1]
    if m == 1:
        return float(Y[:, 0].max() - ref[0])
    order = torch.argsort(Y[:, -1], descending=True)
    Y = Y[order]
    hv = 0.0
    for i in range(Y.shape[0]):
        nxt = Y[i + 1, -1] if i + 1 < Y.shape[0] else ref[-1]
        depth = float(Y[i, -1] - nxt)
        if depth > 0:
            hv += depth * hypervolume(Y[: i + 1, :-1], ref[:-1])
    return hv


# ----------------------------------------------------------------------------------------
# qNEHVI / qEHVI ([UPSTREAM] botorch.acquisition.multi_objective; call sites qnehvi.py:39-52,
# mobo.py:72-90, qehvi.py:67-76)
# ----------------------------------------------------------------------------------------
def hvi_inclusion_exclusion(obj, lower, upper, feas=None):
    """[UPSTREAM] qExpectedHypervolumeImprovement._compute_qehvi.
    obj [S, b, q, m]; lower/upper [S|1, C, m]; feas [S, b, q] or None -> [b]."""
    S, b, q, m = obj.shape
    lo = lower.view(lower.shape[0], 1, lower.shape[1], 1, m)
    up = upper.view(upper.shape[0], 1, upper.shape[1], 1, m)
    areas = torch.zeros(S, b, lower.shape[1], dtype=DT)
    for i in range(1, q + 1):
        comb = torch.tensor(list(itertools.combinations(range(q), i)), dtype=torch.long)
        sub = obj[:, :, comb.view(-1), :].view(S, b, comb.shape[0], i, m)
        vert = sub.min(dim=-2).values  # [S, b, nC, m]
        vert = torch.min(vert.unsqueeze(-3), up)  # [S, b, C, nC, m]
        lengths = (vert - lo).clamp_min(0.0)
        a = lengths.prod(dim=-1)
        if feas is not None:
            fs = feas[:, :, comb.view(-1)].view(S, b, comb.shape[0], i).prod(dim=-1)
            a = a * fs.unsqueeze(-2)
        areas = areas + ((-1) ** (i + 1)) * a.sum(dim=-1)
    return areas.sum(dim=-1).mean(dim=0)


class QNEHVIOracle:
    """qNoisyExpectedHypervolumeImprovement(model, ref_point, X_baseline, prune_baseline,
    objective, constraints, eta, alpha, cache_root=True, X_pending); alpha > 0 switches the per-sample box decomposition of
    more than two objectives to the approximate binary partitioning (`partition_binary`)."""

    def __init__(self, gp: GPOracle, ref_point, X_baseline, objective_ops, constraints=None,
                 mc_samples=512, seed=1234, prune_baseline=True, prune_samples=2048,
                 prune_seed=4321, X_pending=None, base_samples_baseline=None, cell_bounds=None, alpha=0.0):
        self.gp = gp
        self.alpha = float(alpha)
        self.ref = torch.as_tensor(ref_point, dtype=DT)
        self.ops = objective_ops
        self.cons = constraints
        self.S = mc_samples
        self.seed = seed
        self.Mo = len(objective_ops)
        Xb = torch.as_tensor(X_baseline, dtype=DT)
        self.prune_idx = None
        if prune_baseline:
            self.prune_idx = self.prune(Xb, prune_samples, prune_seed)
            Xb = Xb[self.prune_idx]
        if X_pending is not None:
            Xb = torch.cat([Xb, torch.as_tensor(X_pending, dtype=DT)], dim=0)
        self.Xb = Xb
        self.nb = Xb.shape[0]
        self.zb = base_samples_baseline
        # cell_bounds=(lower, upper) [S, C, m] skips the (slow, pure-Python) box decomposition: used only by
        # bench.py to TIME the forward pass of the many-objective config, never by the parity tests
        self._injected_cells = cell_bounds
        self._set_cell_bounds()

    # -- [UPSTREAM] prune_inferior_points_multi_objective
    def prune(self, X, num_samples, seed, base_samples=None):
        mean, cov = self.gp.posterior(X)
        n, M = mean.shape
        z = base_samples if base_samples is not None else base_samples_points_by_outputs(n, M, num_samples, seed)
        L = psd_safe_cholesky(cov)  # [M, n, n]
        samples = mean.unsqueeze(0) + torch.einsum("mij,sjm->sim", L, z)
        obj = multi_objective(self.ops, samples)
        if self.cons:
            infeas = torch.stack([c > 0 for c in constraint_values(self.cons, samples)], 0).any(0)
            obj = obj.clone()
            obj[infeas] = self.ref
        mask = is_non_dominated_chunked(obj, deduplicate=False) & (obj > self.ref).all(dim=-1)
        probs = mask.to(DT).mean(dim=0)
        return probs.nonzero().view(-1)

    # -- [UPSTREAM] NoisyExpectedHypervolumeMixin._set_cell_bounds
    def _set_cell_bounds(self):
        gp, nb, M, S = self.gp, self.nb, self.gp.M, self.S
        if self.zb is None:
            self.zb = base_samples_points_by_outputs(nb, M, S, self.seed) if nb > 0 else torch.zeros(S, 0, M, dtype=DT)
        if nb > 0:
            mean, cov = gp.posterior(self.Xb)
            self.baseline_L = psd_safe_cholesky(cov)  # [M, nb, nb]
            self.mean_b = mean
            fb = mean.unsqueeze(0) + torch.einsum("mij,sjm->sim", self.baseline_L, self.zb)
            self.samples_b = fb
            obj = multi_objective(self.ops, fb)  # [S, nb, Mo]
            feas = None
            if self.cons:
                feas = torch.stack([c <= 0 for c in constraint_values(self.cons, fb)], 0).all(0)
        else:
            self.baseline_L = torch.zeros(M, 0, 0, dtype=DT)
            obj = torch.zeros(S, 0, self.Mo, dtype=DT)
            feas = None
        self.obj_b = obj
        if self._injected_cells is not None:
            self.cell_lower, self.cell_upper = self._injected_cells
            self.n_cells = torch.full((S,), self.cell_lower.shape[1])
            self.fronts = None
            return
        lows, ups, fronts = [], [], []
        for s in range(S):
            Ys = obj[s] if feas is None else obj[s][feas[s]]
            rows = torch.arange(nb) if feas is None else torch.nonzero(feas[s]).view(-1)
            P, idx = pareto_front_above_ref(Ys, self.ref)
            idx = rows[idx]
            if self.Mo == 2:
                lo, up, order = partition_2d(P, self.ref)
                fronts.append(idx[order])
            else:
                lo, up = partition_binary(P, self.ref, self.alpha) if self.alpha > 0 else partition_nd(P, self.ref)
                fronts.append(idx)
            lows.append(lo)
            ups.append(up)
        C = max(l.shape[0] for l in lows)
        # pad with empty cells (lower = upper = ref) -> zero volume
        self.cell_lower = self.ref.view(1, 1, -1).repeat(S, C, 1)
        self.cell_upper = self.ref.view(1, 1, -1).repeat(S, C, 1)
        for s in range(S):
            c = lows[s].shape[0]
            self.cell_lower[s, :c] = lows[s]
            self.cell_upper[s, :c] = ups[s]
        self.n_cells = torch.tensor([l.shape[0] for l in lows])
        self.fronts = fronts

    def base_samples_q(self, q):
        """New-point base samples: the last q points of a fresh (nb+q)-point draw with the
        sampler's seed; the first nb slots are overwritten by the cached baseline draw
        ([UPSTREAM] NormalMCSampler._update_base_samples)."""
        z = base_samples_points_by_outputs(self.nb + q, self.gp.M, self.S, self.seed)
        return z[:, self.nb :, :].contiguous()

    def posterior_blocks(self, X):
        """X [b, q, d] -> mu [b, q, M], Sqq [M, b, q, q], Sqb [M, b, q, nb] (un-standardised)."""
        b, q, d = X.shape
        gp = self.gp
        Xf = X.reshape(b * q, d)
        mu = torch.empty(b, q, gp.M, dtype=DT)
        Sqq = torch.empty(gp.M, b, q, q, dtype=DT)
        Sqb = torch.empty(gp.M, b, q, self.nb, dtype=DT)
        for m, (o, f) in enumerate(zip(gp.outputs, gp._fact)):
            Kx = gp.cross(Xf, m)
            V = (Kx @ f["Linv"].T).view(b, q, -1)
            mu[:, :, m] = ((o.mean_const + Kx @ f["alpha"]) * o.y_std + o.y_mean).view(b, q)
            Xt = gp._t(X, o)
            Kqq = torch.stack([eval_kernel(o.kernel, Xt[i], Xt[i], f["center"], same=True) for i in range(b)])
            Sqq[m] = (Kqq - V @ V.transpose(-1, -2)) * o.y_std**2
            if self.nb > 0:
                Vb = gp.cross(self.Xb, m) @ f["Linv"].T  # [nb, N]
                Kqb = eval_kernel(o.kernel, Xt.reshape(b * q, d), gp._t(self.Xb, o), f["center"]).view(b, q, self.nb)
                Sqb[m] = (Kqb - V @ Vb.T) * o.y_std**2
        return mu, Sqq, Sqb

    def conditional_root(self, Sqq, Sqb):
        """[UPSTREAM] sample_cached_cholesky: bl = Sqb L_b^-T ; br = psd_safe_chol(Sqq - bl bl^T)."""
        M, b, q, _ = Sqq.shape
        if self.nb > 0:
            Lb = self.baseline_L.unsqueeze(1).expand(M, b, self.nb, self.nb)
            bl = torch.linalg.solve_triangular(Lb.transpose(-1, -2), Sqb, upper=True, left=False)
        else:
            bl = Sqb
        br, jit = psd_safe_cholesky(Sqq - bl @ bl.transpose(-1, -2), return_jitter=True)
        return bl, br, jit

    def joint_root_rows(self, X):
        """[UPSTREAM] the fallback of sample_cached_cholesky (`except (NanError, NotPSDError)` in _get_f_X_samples): the joint
        posterior over (X_baseline, X) is factorised as a whole -- psd_safe_cholesky of the (n_b + q) x (n_b + q) covariance,
        jitter on the whole diagonal -- and the samples of the q new points are the last q rows of  mu + L z  with the SAME
        base samples z = [z_b; z_q].  Returns (mu [b, q, M], bl [M, b, q, n_b], br [M, b, q, q]) = those rows, so that
        f_q = mu + bl z_b + br z_q as in the cached path (identical to it whenever no jitter is needed)."""
        b, q, d = X.shape
        gp, nb = self.gp, self.nb
        mu = torch.empty(b, q, gp.M, dtype=DT)
        bl = torch.empty(gp.M, b, q, nb, dtype=DT)
        br = torch.empty(gp.M, b, q, q, dtype=DT)
        for i in range(b):
            mean, cov = gp.posterior(torch.cat([self.Xb, X[i]], dim=0))
            Lfull = psd_safe_cholesky(cov)                        # [M, nb + q, nb + q]
            mu[i] = mean[nb:]
            bl[:, i] = Lfull[:, nb:, :nb]
            br[:, i] = Lfull[:, nb:, nb:]
        return mu, bl, br

    def forward_joint(self, X, zq=None):
        """forward() with every q-batch scored through the joint re-sampling fallback."""
        X = torch.as_tensor(X, dtype=DT)
        if zq is None:
            zq = self.base_samples_q(X.shape[1])
        mu, bl, br = self.joint_root_rows(X)
        f = mu.unsqueeze(0) + torch.einsum("mbqk,skm->sbqm", bl, self.zb) + torch.einsum("mbqk,skm->sbqm", br, zq)
        obj = multi_objective(self.ops, f)
        feas = smoothed_feasibility(self.cons, f) if self.cons else None
        return hvi_inclusion_exclusion(obj, self.cell_lower, self.cell_upper, feas)

    def sample_q(self, X, zq=None):
        """f_q samples [S, b, q, M] consistent with the cached baseline samples."""
        b, q, _ = X.shape
        if zq is None:
            zq = self.base_samples_q(q)
        mu, Sqq, Sqb = self.posterior_blocks(X)
        bl, br, jit = self.conditional_root(Sqq, Sqb)
        f = mu.unsqueeze(0) + torch.einsum("mbqk,skm->sbqm", bl, self.zb) + torch.einsum("mbqk,skm->sbqm", br, zq)
        return f, dict(mu=mu, Sqq=Sqq, Sqb=Sqb, bl=bl, br=br, jitter=jit)

    def forward(self, X, zq=None, return_parts=False):
        X = torch.as_tensor(X, dtype=DT)
        f, parts = self.sample_q(X, zq)
        obj = multi_objective(self.ops, f)
        feas = smoothed_feasibility(self.cons, f) if self.cons else None
        val = hvi_inclusion_exclusion(obj, self.cell_lower, self.cell_upper, feas)
        if return_parts:
            parts.update(samples=f, obj=obj)
            return val, parts
        return val

    def forward_reference_style(self, X, zq=None, batch_limit=None):
        """The op sequence BoTorch runs per call: joint posterior over cat(X_baseline, X)
        for every q-batch (recomputing the nb baseline rows), then the cached-root
        update.  Used as the timed CPU baseline; equals ``forward`` to rounding."""
        X = torch.as_tensor(X, dtype=DT)
        b, q, d = X.shape
        if zq is None:
            zq = self.base_samples_q(q)
        out = []
        step = batch_limit or b
        gp, nb = self.gp, self.nb
        for s0 in range(0, b, step):
            Xc = X[s0 : s0 + step]
            bc = Xc.shape[0]
            Xfull = torch.cat([self.Xb.unsqueeze(0).expand(bc, nb, d), Xc], dim=1)  # [bc, nb+q, d]
            n = nb + q
            mu = torch.empty(bc, q, gp.M, dtype=DT)
            bls, brs = [], []
            for m, (o, f) in enumerate(zip(gp.outputs, gp._fact)):
                Xt = gp._t(Xfull, o)
                Kx = eval_kernel(o.kernel, Xt.reshape(bc * n, d), f["Xt"], f["center"])
                V = (Kx @ f["Linv"].T).view(bc, n, -1)
                mean = ((o.mean_const + Kx @ f["alpha"]) * o.y_std + o.y_mean).view(bc, n)
                Kss = torch.stack([eval_kernel(o.kernel, Xt[i], Xt[i], f["center"], same=True) for i in range(bc)])
                cov = (Kss - V @ V.transpose(-1, -2)) * o.y_std**2
                mu[:, :, m] = mean[:, nb:]
                bl_mat = cov[:, nb:, :nb]
                Lb = self.baseline_L[m].unsqueeze(0).expand(bc, nb, nb)
                bl = torch.linalg.solve_triangular(Lb.transpose(-1, -2), bl_mat, upper=True, left=False) if nb > 0 else bl_mat
                br = psd_safe_cholesky(cov[:, nb:, nb:] - bl @ bl.transpose(-1, -2))
                bls.append(bl)
                brs.append(br)
            bl = torch.stack(bls)
            br = torch.stack(brs)
            fq = mu.unsqueeze(0) + torch.einsum("mbqk,skm->sbqm", bl, self.zb) + torch.einsum("mbqk,skm->sbqm", br, zq)
            obj = multi_objective(self.ops, fq)
            feas = smoothed_feasibility(self.cons, fq) if self.cons else None
            out.append(hvi_inclusion_exclusion(obj, self.cell_lower, self.cell_upper, feas))
        return torch.cat(out)


class QEHVIOracle:
    """qExpectedHypervolumeImprovement with a fixed partitioning of the observed front
    (qehvi.py:37-77): no baseline conditioning, cells shared by all MC samples."""

    def __init__(self, gp, ref_point, Y_obj_observed, objective_ops, mc_samples=512, seed=1234):
        self.gp, self.ops, self.S, self.seed = gp, objective_ops, mc_samples, seed
        self.ref = torch.as_tensor(ref_point, dtype=DT)
        P, _ = pareto_front_above_ref(torch.as_tensor(Y_obj_observed, dtype=DT), self.ref)
        if self.ref.numel() == 2:
            lo, up, _ = partition_2d(P, self.ref)
        else:
            lo, up = partition_nd(P, self.ref)
        self.cell_lower, self.cell_upper = lo.unsqueeze(0), up.unsqueeze(0)

    def forward(self, X, zq=None):
        X = torch.as_tensor(X, dtype=DT)
        b, q, d = X.shape
        gp = self.gp
        if zq is None:
            zq = base_samples_points_by_outputs(q, gp.M, self.S, self.seed)
        f = torch.empty(self.S, b, q, gp.M, dtype=DT)
        for i in range(b):
            mean, cov = gp.posterior(X[i])
            L = psd_safe_cholesky(cov)
            f[:, i] = mean.unsqueeze(0) + torch.einsum("mij,sjm->sim", L, zq)
        obj = multi_objective(self.ops, f)
        return hvi_inclusion_exclusion(obj, self.cell_lower, self.cell_upper)


# ----------------------------------------------------------------------------------------
# qLogEI ([UPSTREAM] botorch.acquisition.logei + utils.safe_math; call site sobo.py:64-89)
# ----------------------------------------------------------------------------------------
TAU_RELU = 1e-6
TAU_MAX = 1e-2
FATMAX_ALPHA = 2.0
FATPLUS_ALPHA = 1e-1


def log_softplus(x, tau=1.0):
    lower, upper = -35.0, 32.0
    mask = x / tau > lower
    xs = x.masked_fill(~mask, lower)
    sp = torch.nn.functional.softplus(xs, beta=1.0 / tau, threshold=upper)
    return torch.where(mask, sp.log(), x / tau + math.log(tau))


def log_fatplus(x, tau=1.0):
    def _lfp(x):
        return torch.logaddexp(log_softplus(x), math.log(FATPLUS_ALPHA) - torch.log1p(x.square()))

    return math.log(tau) + _lfp(x / tau)


def _pareto(x, alpha):
    a = alpha / 2.0
    b1 = 2.0 * a
    b0 = a * b1
    return (b0 / (b0 + b1 * x + x.square())).pow(a)


def fatmax(x, dim, tau=1.0, alpha=FATMAX_ALPHA):
    mx = x.amax(dim=dim, keepdim=True)
    y = (mx - x) / tau
    out = mx + tau * _pareto(y, alpha).sum(dim=dim, keepdim=True).log()
    return out.squeeze(dim)


def logmeanexp(x, dim):
    return torch.logsumexp(x, dim=dim) - math.log(x.shape[dim])


class QLogEIOracle:
    def __init__(self, gp, objective_spec, X_observed, mc_samples=512, seed=1234, best_f=None):
        self.gp, self.spec, self.S, self.seed = gp, objective_spec, mc_samples, seed
        if best_f is None:
            mean, _ = gp.posterior(torch.as_tensor(X_observed, dtype=DT))
            best_f = float(scalar_objective(objective_spec, mean).max())
        self.best_f = best_f

    def forward(self, X, zq=None, return_parts=False):
        X = torch.as_tensor(X, dtype=DT)
        b, q, d = X.shape
        gp = self.gp
        if zq is None:
            zq = base_samples_points_by_outputs(q, gp.M, self.S, self.seed)
        Xf = X.reshape(b * q, d)
        f = torch.empty(self.S, b, q, gp.M, dtype=DT)
        jit = []
        for m, (o, fa) in enumerate(zip(gp.outputs, gp._fact)):
            Kx = gp.cross(Xf, m)
            V = (Kx @ fa["Linv"].T).view(b, q, -1)
            mu = ((o.mean_const + Kx @ fa["alpha"]) * o.y_std + o.y_mean).view(b, q)
            Xt = gp._t(X, o)
            Kqq = torch.stack([eval_kernel(o.kernel, Xt[i], Xt[i], fa["center"], same=True) for i in range(b)])
            Sqq = (Kqq - V @ V.transpose(-1, -2)) * o.y_std**2
            L, j = psd_safe_cholesky(Sqq, return_jitter=True)
            jit.append(j)
            f[..., m] = mu.unsqueeze(0) + torch.einsum("bqk,sk->sbq", L, zq[..., m])
        obj = scalar_objective(self.spec, f)  # [S, b, q]
        li = log_fatplus(obj - self.best_f, tau=TAU_RELU)
        val = logmeanexp(fatmax(li, dim=-1, tau=TAU_MAX), dim=0)
        if return_parts:
            return val, dict(samples=f, obj=obj, jitter=torch.stack(jit))
        return val


# ----------------------------------------------------------------------------------------
# Log-space hypervolume improvement: qLogEHVI / qLogNEHVI ([UPSTREAM] botorch.acquisition.multi_objective.logei,
# botorch.utils.safe_math; call site mobo.py:72-90 -- qLogNEHVI is MoboStrategy's default,
# data_models/strategies/predictives/mobo.py).  PARITY UNPINNED: restated from the published algorithm
# (Ament et al. 2023, "Unexpected Improvements to Expected Improvement", section 4 / appendix).
# ----------------------------------------------------------------------------------------
def log1mexp(x):
    """log(1 - exp(x)) for x < 0."""
    return torch.where(-math.log(2.0) < x, (-torch.expm1(x)).log(), torch.log1p(-torch.exp(x)))


def logdiffexp(log_a, log_b):
    """log(b - a) from log a and log b (b > a > 0)."""
    is_inf = log_b.isinf() & log_a.isinf()
    return log_b + log1mexp(log_a - log_b.masked_fill(is_inf, 0.0))


def fatmin(x, dim, tau=1.0, alpha=FATMAX_ALPHA):
    return -fatmax(-x, dim=dim, tau=tau, alpha=alpha)


def fatminimum(a, b, tau):
    """Smooth element-wise minimum; an infinite `b` (unbounded cell) returns `a` exactly."""
    a, b = torch.broadcast_tensors(a, b)
    inf = torch.isinf(b)
    bf = torch.where(inf, a.detach(), b)
    sm = fatmin(torch.stack([a, bf], dim=-1), dim=-1, tau=tau)
    return torch.where(inf, a, sm)


def fatmoid(x, tau=1.0):
    """Twice differentiable step approximation with an O(1/x^2) tail."""
    x = x / tau
    m = 1.0 / math.sqrt(3.0)
    cauchy = lambda t: 1.0 / (1.0 + t.square())  # noqa: E731
    return torch.where(x < 0, (2.0 / 3.0) * cauchy(x - m), 1.0 - (2.0 / 3.0) * cauchy(x + m))


def log_smoothed_feasibility(cons, Y):
    """[UPSTREAM] compute_smoothed_feasibility_indicator(log=True, fat=True): sum_c log fatmoid(-c / eta)."""
    w = torch.zeros(Y.shape[:-1], dtype=DT)
    for (idx, sign, tp, eta), c in zip(cons, constraint_values(cons, Y)):
        w = w + fatmoid(-c / eta).log()
    return w


def log_hvi_inclusion_exclusion(obj, lower, upper, n_cells, log_feas=None, tau_relu=TAU_RELU, tau_max=TAU_MAX):
    """[UPSTREAM] qLogExpectedHypervolumeImprovement._compute_log_qehvi.
    obj [S, b, q, m]; lower / upper [S|1, C, m]; n_cells [S|1] = valid cells per MC sample (the padding cells that make
    the per-sample lists rectangular are ignored); log_feas [S, b, q] or None -> [b]."""
    S, b, q, m = obj.shape
    Sc, C = lower.shape[0], lower.shape[1]
    lo = lower.view(Sc, 1, C, 1, m)
    up = upper.view(Sc, 1, C, 1, m)
    neg_inf = torch.full((S, b, C), -math.inf, dtype=DT)
    acc = [neg_inf, neg_inf]  # even, odd
    for i in range(1, q + 1):
        comb = torch.tensor(list(itertools.combinations(range(q), i)), dtype=torch.long)
        sub = obj[:, :, comb.view(-1), :].view(S, b, comb.shape[0], i, m)
        vert = fatmin(sub, dim=-2, tau=tau_max)                    # [S, b, nC, m]
        vert = fatminimum(vert.unsqueeze(-3), up, tau=tau_max)     # [S, b, C, nC, m]
        ll = log_fatplus(vert - lo, tau=tau_relu).sum(dim=-1)      # [S, b, C, nC]
        if log_feas is not None:
            ll = ll + log_feas[:, :, comb.view(-1)].view(S, b, comb.shape[0], i).sum(dim=-1).unsqueeze(-2)
        acc[i % 2] = torch.logaddexp(acc[i % 2], torch.logsumexp(ll, dim=-1))
    cellv = logdiffexp(log_a=acc[0], log_b=acc[1])                 # [S, b, C]
    valid = torch.arange(C).view(1, 1, C) < torch.as_tensor(n_cells).view(Sc, 1, 1)
    cellv = torch.where(valid.expand(S, b, C), cellv, neg_inf)
    return logmeanexp(torch.logsumexp(cellv, dim=-1), dim=0)


class QLogNEHVIOracle(QNEHVIOracle):
    """qLogNoisyExpectedHypervolumeImprovement: same construction as qNEHVI (pruning, cached root, per-sample
    cells), log-space smoothed value."""

    def forward(self, X, zq=None, return_parts=False):
        X = torch.as_tensor(X, dtype=DT)
        f, parts = self.sample_q(X, zq)
        obj = multi_objective(self.ops, f)
        lf = log_smoothed_feasibility(self.cons, f) if self.cons else None
        val = log_hvi_inclusion_exclusion(obj, self.cell_lower, self.cell_upper, self.n_cells, lf)
        if return_parts:
            parts.update(samples=f, obj=obj)
            return val, parts
        return val


class QLogEHVIOracle(QEHVIOracle):
    def forward(self, X, zq=None):
        X = torch.as_tensor(X, dtype=DT)
        b, q, d = X.shape
        gp = self.gp
        if zq is None:
            zq = base_samples_points_by_outputs(q, gp.M, self.S, self.seed)
        fs = []
        for i in range(b):
            mean, cov = gp.posterior(X[i])
            L = psd_safe_cholesky(cov)
            fs.append(mean.unsqueeze(0) + torch.einsum("mij,sjm->sim", L, zq))
        obj = multi_objective(self.ops, torch.stack(fs, dim=1))
        return log_hvi_inclusion_exclusion(obj, self.cell_lower, self.cell_upper, [self.cell_lower.shape[1]])


# ----------------------------------------------------------------------------------------
# Single-objective MC acquisition functions of SoboStrategy (sobo.py:51-90 via get_acquisition_function):
# qEI, qLogEI, qSR, qUCB, qPI (fixed best_f) and qNEI, qLogNEI (best value of the cached baseline samples, per MC
# sample).  [UPSTREAM] botorch.acquisition.monte_carlo / logei -- PARITY UNPINNED.
# ----------------------------------------------------------------------------------------
class QScalarOracle(QNEHVIOracle):
    def __init__(self, gp, kind, objective_spec, X_observed, mc_samples=512, seed=1234, beta=0.2, tau=1e-3,
                 prune_baseline=True, prune_samples=2048, prune_seed=4321, X_pending=None, constraints=None,
                 best_f=None, lb_generator=None):
        self.gp, self.kind, self.spec, self.S, self.seed = gp, kind, objective_spec, mc_samples, seed
        self.beta, self.tau, self.cons = beta, tau, constraints
        # [UPSTREAM] @concatenate_pending_points: pending points are scored jointly with X (also for qNEI / qLogNEI)
        self.X_pending = None if X_pending is None else torch.as_tensor(X_pending, dtype=DT).reshape(-1, gp.d)
        self.noisy = kind in ("qNEI", "qLogNEI")
        Xo = torch.as_tensor(X_observed, dtype=DT)
        self.zb = None
        if self.noisy:
            self.prune_idx = None
            if prune_baseline:
                self.prune_idx = self.prune_so(Xo, prune_samples, prune_seed)
                Xo = Xo[self.prune_idx]
            self.Xb, self.nb = Xo, Xo.shape[0]
            M = gp.M
            self.zb = base_samples_points_by_outputs(self.nb, M, self.S, seed)
            mean, cov = gp.posterior(self.Xb)
            self.baseline_L = psd_safe_cholesky(cov)
            fb = mean.unsqueeze(0) + torch.einsum("mij,sjm->sim", self.baseline_L, self.zb)
            self.samples_b = fb
            # [UPSTREAM] qNoisyExpectedImprovement.compute_best_f -> compute_best_feasible_objective: [S]
            self.best_f_s = best_feasible_objective(gp, self.spec, self.cons, fb, scalar_objective(self.spec, fb), self.Xb,
                                                    generator=lb_generator)
        else:
            self.Xb, self.nb = torch.zeros(0, gp.d, dtype=DT), 0
            self.baseline_L = torch.zeros(gp.M, 0, 0, dtype=DT)
            self.zb = torch.zeros(self.S, 0, gp.M, dtype=DT)
            if best_f is None:
                # [UPSTREAM] get_acquisition_function (qEI / qLogEI / qPI): best FEASIBLE objective of the posterior mean
                mean, _ = gp.posterior(Xo)
                best_f = float(best_feasible_objective(gp, objective_spec, self.cons, mean, scalar_objective(objective_spec, mean),
                                                       Xo, generator=lb_generator))
            self.best_f = best_f

    def prune_so(self, X, num_samples, seed):
        """[UPSTREAM] prune_inferior_points: keep the points that are the best one in at least one joint sample."""
        mean, cov = self.gp.posterior(X)
        n, M = mean.shape
        z = base_samples_points_by_outputs(n, M, num_samples, seed)
        L = psd_safe_cholesky(cov)
        samples = mean.unsqueeze(0) + torch.einsum("mij,sjm->sim", L, z)
        obj = scalar_objective(self.spec, samples)
        if self.cons:   # [UPSTREAM] infeasible samples cannot be the best point
            obj = torch.where(feasibility_indicator(self.cons, samples), obj, torch.full_like(obj, -float("inf")))
        best = obj.argmax(dim=-1)
        return torch.unique(best)

    def base_samples_q(self, q):
        z = base_samples_points_by_outputs(self.nb + q, self.gp.M, self.S, self.seed)
        return z[:, self.nb:, :].contiguous()

    def forward(self, X, zq=None):
        X = torch.as_tensor(X, dtype=DT)
        if self.X_pending is not None:
            X = torch.cat([X, self.X_pending.unsqueeze(0).expand(X.shape[0], -1, -1)], dim=1)
        f, _ = self.sample_q(X, zq)
        obj = scalar_objective(self.spec, f)  # [S, b, q]
        bf = self.best_f_s.view(-1, 1, 1) if self.noisy else self.best_f
        k = self.kind
        if k in ("qLogEI", "qLogNEI"):
            li = log_fatplus(obj - bf, tau=TAU_RELU)
            if self.cons:
                li = li + log_smoothed_feasibility(self.cons, f)
            return logmeanexp(fatmax(li, dim=-1, tau=TAU_MAX), dim=0)
        if k in ("qEI", "qNEI"):
            u = (obj - bf).clamp_min(0.0)
        elif k == "qSR":
            u = obj
        elif k == "qPI":
            u = torch.sigmoid((obj - bf) / self.tau)
        elif k == "qUCB":
            mean = obj.mean(dim=0)
            u = mean + math.sqrt(self.beta * math.pi / 2.0) * (obj - mean).abs()
        else:
            raise ValueError(k)
        if self.cons:
            if k in ("qSR", "qUCB"):
                raise ValueError("constraints need a non-negative utility")
            u = u * smoothed_feasibility(self.cons, f)
        return u.max(dim=-1).values.mean(dim=0)


# ----------------------------------------------------------------------------------------
# Test functions used as input generators for the BASELINE configs
# ----------------------------------------------------------------------------------------
def zdt1(X):
    """[UPSTREAM] botorch.test_functions.ZDT1 (reference wraps it at benchmarks/multi.py:454,467)."""
    f1 = X[..., 0]
    g = 1 + 9 * X[..., 1:].mean(dim=-1)
    f2 = g * (1 - (f1 / g).sqrt())
    return torch.stack([f1, f2], dim=-1)


def dtlz2(X, num_objectives):
    """bofire/benchmarks/multi.py:95-132."""
    d = X.shape[-1]
    k = d - num_objectives + 1
    Xm = X[..., -k:]
    g1 = 1 + ((Xm - 0.5) ** 2).sum(dim=-1)
    fs = []
    for i in range(num_objectives):
        idx = num_objectives - 1 - i
        f = g1.clone()
        f = f * torch.cos(X[..., :idx] * (math.pi / 2)).prod(dim=-1)
        if i > 0:
            f = f * torch.sin(X[..., idx] * (math.pi / 2))
        fs.append(f)
    return torch.stack(fs, dim=-1)


def himmelblau(X):
    """bofire/benchmarks/single.py:409-446."""
    x1, x2 = X[..., 0], X[..., 1]
    return (x1**2 + x2 - 11) ** 2 + (x1 + x2**2 - 7) ** 2

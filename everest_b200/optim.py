"""Candidate generation around the acquisition function: the L2 boundary of SURVEY.md 8b
(`BotorchStrategy._optimize_acqf_continuous`, strategies/predictives/botorch.py:326-406, which calls
botorch.optim.optimize_acqf / optimize_acqf_discrete).

Round-1 scope: the raw-sample screen runs as ONE device call over all raw_samples q-batches (the
reference scores them in raw_samples / batch_limit sequential CPU calls), then
`initialize_q_batch` picks the restarts, then `gen_candidates_scipy` refines them with scipy's L-BFGS-B
exactly like [UPSTREAM] botorch.generation.gen_candidates_scipy (all restarts as one vector, loss =
-sum acqf, box bounds, fixed features).  The gradient comes from the analytic adjoint kernels
(bo_acqf_forward_backward, csrc/grad.cu) -- what BoTorch obtains from autograd; options["gradient"] = "fd"
switches to central finite differences evaluated in ONE batched device call per iteration (kept as a
cross-check: the MC acquisition value is deterministic for fixed base samples).
"""
from typing import Dict, Optional, Tuple

import numpy as np
import torch
from scipy.optimize import minimize  # imported up front: the first import costs ~0.5 s, keep it out of ask()

from . import sampling


def draw_sobol_samples(bounds: torch.Tensor, n: int, q: int, seed: Optional[int] = None, device=None) -> torch.Tensor:
    """[UPSTREAM] botorch.utils.sampling.draw_sobol_samples: [n, q, d] inside bounds [2, d].  With a CUDA `device` the
    points are generated there (csrc/sobol.cu, bit-identical to torch's engine for the same seed)."""
    d = bounds.shape[-1]
    if device is not None and torch.device(device).type == "cuda" and seed is not None:
        u = sampling.sobol_uniform_device(q * d, n, int(seed), device).view(n, q, d)
        lo, hi = bounds[0].to(u), bounds[1].to(u)
        return lo + (hi - lo) * u
    eng = torch.quasirandom.SobolEngine(q * d, scramble=True, seed=seed)
    u = eng.draw(n, dtype=torch.double).view(n, q, d)
    lo, hi = bounds[0].to(torch.double), bounds[1].to(torch.double)
    return lo + (hi - lo) * u


def initialize_q_batch(X: torch.Tensor, acq_vals: torch.Tensor, n: int, eta: float = 2.0,
                       generator: Optional[torch.Generator] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """[UPSTREAM] botorch.optim.initializers.initialize_q_batch: Boltzmann-weighted multinomial pick of n
    restarts (softmax of eta * z-score), always including the arg-max.  Returns (X[idcs], idcs)."""
    n_samples = X.shape[0]
    if n > n_samples:
        raise RuntimeError(f"n ({n}) cannot be larger than the number of provided samples ({n_samples})")
    if n == n_samples:
        return X, torch.arange(n)
    Y = acq_vals.detach().to("cpu", torch.double)
    Ystd = Y.std(dim=0)
    if torch.any(Ystd == 0):
        perm = torch.randperm(n_samples, generator=generator)[:n]
        return X[perm], perm
    max_val, max_idx = torch.max(Y, dim=0)
    Z = (Y - Y.mean(dim=0)) / Ystd
    etaZ = eta * Z
    weights = torch.exp(etaZ)
    while torch.isinf(weights).any():
        etaZ = etaZ * 0.5
        weights = torch.exp(etaZ)
    idcs = torch.multinomial(weights, n, generator=generator)
    if max_idx not in idcs:
        idcs[-1] = max_idx
    return X[idcs], idcs


def apply_fixed_features(X: torch.Tensor, fixed_features: Optional[Dict[int, float]]) -> torch.Tensor:
    if fixed_features:
        X = X.clone()
        for idx, val in fixed_features.items():
            X[..., idx] = val
    return X


def dense_linear_constraints(constraints, q: int, d: int) -> Tuple[np.ndarray, np.ndarray]:
    """BoTorch-form linear constraints -> dense rows over the flattened q*d variables of ONE q-batch.

    Each constraint is (indices, coefficients, rhs) as produced by get_linear_constraints / get_interpoint_constraints
    (utils/torch_tools.py:45-144): 1-d `indices` [k] = an intra-point constraint sum_i coef_i x[idx_i] (>=|==) rhs that
    holds for every one of the q points; 2-d `indices` [k, 2] = (point, feature) pairs of an inter-point constraint
    ([UPSTREAM] botorch.optim.parameter_constraints.make_scipy_linear_constraints)."""
    rows, rhs = [], []
    for idx, coef, r in constraints or []:
        idx = torch.as_tensor(idx).long()
        coef = torch.as_tensor(coef, dtype=torch.double)
        if idx.dim() == 1:
            if idx.numel() and (int(idx.max()) >= d or int(idx.min()) < 0):
                raise RuntimeError(f"Index out of bounds for {d}-dim parameter tensor")
            for j in range(q):
                row = np.zeros(q * d)
                for i, c in zip(idx.tolist(), coef.tolist()):
                    row[j * d + i] += c
                rows.append(row)
                rhs.append(float(r))
        elif idx.dim() == 2:
            if int(idx[:, 0].max()) >= q or int(idx[:, 1].max()) >= d:
                raise RuntimeError(f"Index out of bounds for ({q}, {d})-dim parameter tensor")
            row = np.zeros(q * d)
            for (j, i), c in zip(idx.tolist(), coef.tolist()):
                row[j * d + i] += c
            rows.append(row)
            rhs.append(float(r))
        else:
            raise ValueError("`indices` must be 1- or 2-dimensional")
    if not rows:
        return np.zeros((0, q * d)), np.zeros(0)
    return np.stack(rows), np.asarray(rhs)


def _has_interpoint(constraints) -> bool:
    return any(torch.as_tensor(c[0]).dim() == 2 for c in (constraints or []))


NLC_TOL = -1e-6   # [UPSTREAM] botorch.optim.parameter_constraints.NLC_TOL: slack of the feasibility test of a start point


def nchoosek_constraints(indices, max_count: Optional[int] = None, min_count: int = 0, ell: float = 1e-3):
    """NChooseK as nonlinear inequality callables (utils/torch_tools.py:147-207): the number of zeros among the listed
    columns is counted by a sum of narrow Gaussians exp(-(x / ell)^2 / 2); at most `max_count` non-zero features means
    zeros >= n - max_count, at least `min_count` means zeros <= n - min_count.  Returns [(callable, True)] (intra-point)."""
    idx = torch.as_tensor(indices, dtype=torch.long)
    n = int(idx.numel())

    def zeros(x):
        return torch.exp(-0.5 * (x[..., idx] / ell) ** 2).sum(dim=-1)

    out = []
    if max_count is not None and max_count != n:
        out.append((lambda x, k=n - int(max_count): zeros(x) - k, True))
    if min_count > 0:
        out.append((lambda x, k=n - int(min_count): k - zeros(x), True))
    return out


def product_constraint(indices, exponents, rhs: float, sign: int = 1):
    """ProductInequalityConstraint as a nonlinear inequality callable (utils/torch_tools.py:210-236):
    feasible iff  -sign * prod_i x_i^e_i + rhs >= 0."""
    idx = torch.as_tensor(indices, dtype=torch.long)
    ex = torch.as_tensor(exponents, dtype=torch.double)
    return (lambda x: -1.0 * sign * (x[..., idx] ** ex).prod(dim=-1) + rhs, True)


def nchoosek_generator(bounds, specs, fixed_features: Optional[Dict[int, float]] = None):
    """Feasible raw samples for NChooseK-constrained spaces in the (n, q, seed) -> [n, q, d] form that
    get_initial_conditions_generator hands to gen_batch_initial_conditions (utils/torch_tools.py:809-864; the reference
    asks its RandomStrategy, strategies/random.py:180-353): per point and constraint a uniformly drawn number of active
    features in [min_count, max_count], the others exactly 0.  specs: [(indices, max_count, min_count)]."""
    bounds = torch.as_tensor(bounds, dtype=torch.double)

    def generator(n: int, q: int, seed: int) -> torch.Tensor:
        g = torch.Generator().manual_seed(int(seed))
        lo, hi = bounds[0], bounds[1]
        X = lo + (hi - lo) * torch.rand(n, q, bounds.shape[1], dtype=torch.double, generator=g)
        for indices, max_count, min_count in specs:
            idx = torch.as_tensor(indices, dtype=torch.long)
            m = int(idx.numel())
            kmax = m if max_count is None else int(max_count)
            k = torch.randint(int(min_count), kmax + 1, (n, q), generator=g)
            order = torch.rand(n, q, m, generator=g).argsort(dim=-1)          # random ranking of the features
            active = order < k.unsqueeze(-1)
            X[..., idx] = torch.where(active, X[..., idx], torch.zeros((), dtype=torch.double))
        return apply_fixed_features(X, fixed_features)

    return generator


def _split_nlc(nonlinear_inequality_constraints):
    out = []
    for c in nonlinear_inequality_constraints or []:
        if isinstance(c, (tuple, list)):
            out.append((c[0], bool(c[1])))
        else:
            out.append((c, True))     # older BoTorch convention: a bare callable is intra-point
    return out


def nonlinear_constraints_satisfied(X: torch.Tensor, nonlinear_inequality_constraints, tol: float = NLC_TOL) -> torch.Tensor:
    """[r] bool: every intra-point constraint holds at each of the q points and every inter-point one on the q-batch."""
    X = torch.as_tensor(X, dtype=torch.double)
    ok = torch.ones(X.shape[0], dtype=torch.bool)
    for fn, intra in _split_nlc(nonlinear_inequality_constraints):
        if intra:
            ok &= (fn(X) >= tol).reshape(X.shape[0], -1).all(dim=1)
        else:
            ok &= torch.stack([torch.as_tensor(fn(X[i]) >= tol).reshape(-1).all() for i in range(X.shape[0])])
    return ok


def sample_polytope(A: np.ndarray, b: np.ndarray, C: np.ndarray, c: np.ndarray, lb: np.ndarray, ub: np.ndarray, n: int,
                    seed: Optional[int] = None, n_burnin: int = 10000, n_thinning: int = 32, n_chains: int = 64) -> np.ndarray:
    """Uniform samples from {x : A x <= b, C x = c, lb <= x <= ub} by hit-and-run ([UPSTREAM]
    botorch.utils.sampling.HitAndRunPolytopeSampler as reached through sample_q_batches_from_polytope; BoFire's own
    RandomStrategy uses the same sampler, strategies/random.py:180-353).  The chains are vectorised: `n_chains` chains
    each burn in for n_burnin / n_chains * 8 steps (at least 64) and are thinned by n_thinning."""
    from scipy.optimize import linprog

    k = lb.shape[0]
    rng = np.random.default_rng(seed)
    Ab = np.concatenate([A.reshape(-1, k), np.eye(k), -np.eye(k)], axis=0)
    bb = np.concatenate([b.reshape(-1), ub, -lb])
    # equality constraints: x = x0 + Nz with N a basis of the null space of C
    if C.size:
        Cm = C.reshape(-1, k)
        x0 = np.linalg.lstsq(Cm, c, rcond=None)[0]
        _, sv, Vt = np.linalg.svd(Cm)
        rank = int((sv > 1e-10 * max(1.0, sv.max())).sum())
        Nb = Vt[rank:].T
        if Nb.shape[1] == 0:
            raise ValueError("equality constraints leave no degree of freedom")
    else:
        x0 = np.zeros(k)
        Nb = np.eye(k)
    Az = Ab @ Nb
    bz = bb - Ab @ x0
    kz = Nb.shape[1]
    # Chebyshev centre: max r  s.t.  Az z + r |Az_i| <= bz
    norms = np.linalg.norm(Az, axis=1)
    keep = norms > 1e-14
    if np.any(bz[~keep] < -1e-9):
        raise ValueError("the polytope is empty")
    Az, bz, norms = Az[keep], bz[keep], norms[keep]
    res = linprog(np.concatenate([np.zeros(kz), [-1.0]]), A_ub=np.concatenate([Az, norms[:, None]], axis=1), b_ub=bz,
                  bounds=[(None, None)] * kz + [(0, None)], method="highs")
    if res.status != 0 or res.x[-1] <= 0:
        raise ValueError("no interior point: the polytope is empty or has no volume")
    z = np.tile(res.x[:kz], (n_chains, 1))
    per_chain = -(-n // n_chains)
    burn = max(64, (n_burnin * 8) // n_chains)
    out = np.empty((per_chain, n_chains, kz))
    total = burn + per_chain * n_thinning
    for it in range(total):
        dirs = rng.standard_normal((n_chains, kz))
        dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
        slack = bz[None, :] - z @ Az.T          # >= 0 inside
        rate = dirs @ Az.T
        with np.errstate(divide="ignore", invalid="ignore"):
            t = slack / rate
        t_hi = np.where(rate > 1e-14, t, np.inf).min(axis=1)
        t_lo = np.where(rate < -1e-14, t, -np.inf).max(axis=1)
        step = t_lo + (t_hi - t_lo) * rng.random(n_chains)
        z = z + step[:, None] * dirs
        if it >= burn and (it - burn + 1) % n_thinning == 0:
            out[(it - burn + 1) // n_thinning - 1] = z
    zs = out.reshape(-1, kz)[rng.permutation(per_chain * n_chains)[:n]]
    return x0[None, :] + zs @ Nb.T


def sample_q_batches_from_polytope(n: int, q: int, bounds: torch.Tensor, inequality_constraints=None,
                                   equality_constraints=None, seed: Optional[int] = None, n_burnin: int = 10000,
                                   n_thinning: int = 32, fixed_features: Optional[Dict[int, float]] = None) -> torch.Tensor:
    """[UPSTREAM] botorch.optim.initializers.sample_q_batches_from_polytope: [n, q, d] raw samples inside the bounds
    and the linear constraints (inequalities in BoTorch's  sum coef x >= rhs  form).  Fixed features are pinned by
    collapsing their bounds."""
    bounds = torch.as_tensor(bounds, dtype=torch.double)
    d = bounds.shape[-1]
    lb, ub = bounds[0].numpy().copy(), bounds[1].numpy().copy()
    for j, v in (fixed_features or {}).items():
        lb[j] = ub[j] = float(v)
    inter = _has_interpoint(inequality_constraints) or _has_interpoint(equality_constraints)
    qq = q if inter else 1          # intra-point constraints + box = product of q identical d-polytopes
    Ai, bi = dense_linear_constraints(inequality_constraints, qq, d)
    Ce, ce = dense_linear_constraints(equality_constraints, qq, d)
    lbq, ubq = np.tile(lb, qq), np.tile(ub, qq)
    # pinned coordinates (fixed features / degenerate bounds) become equalities so that the polytope keeps a volume
    pinned = np.nonzero(ubq <= lbq)[0]
    if pinned.size:
        E = np.zeros((pinned.size, qq * d))
        E[np.arange(pinned.size), pinned] = 1.0
        Ce = np.concatenate([Ce, E], axis=0)
        ce = np.concatenate([ce, lbq[pinned]])
        lbq = lbq.copy(); ubq = ubq.copy()
        lbq[pinned] -= 1.0
        ubq[pinned] += 1.0
    x = sample_polytope(-Ai, -bi, Ce, ce, lbq, ubq, n * (q // qq), seed=seed, n_burnin=n_burnin, n_thinning=n_thinning)
    X = torch.from_numpy(x).view(n, q, d)
    return apply_fixed_features(X, fixed_features)


def gen_batch_initial_conditions(acq_function, bounds: torch.Tensor, q: int, num_restarts: int, raw_samples: int,
                                 fixed_features: Optional[Dict[int, float]] = None, options: Optional[dict] = None,
                                 seed: Optional[int] = None, inequality_constraints=None, equality_constraints=None,
                                 generator=None):
    """[UPSTREAM] gen_batch_initial_conditions: Sobol raw samples (hit-and-run polytope samples when linear
    constraints are given; `generator(n, q, seed) -> [n, q, d]` when the caller supplies one, as BoFire does for
    nonlinear constraints, botorch.py:257-265), one screened forward over ALL of them on the device, then
    initialize_q_batch."""
    options = options or {}
    if seed is None:
        seed = int(torch.randint(0, 1000000, (1,)).item())
    if generator is not None:
        X_rnd = torch.as_tensor(generator(raw_samples, q, seed), dtype=torch.double)
        if tuple(X_rnd.shape) != (raw_samples, q, bounds.shape[-1]):
            raise ValueError(f"generator must return [{raw_samples}, {q}, {bounds.shape[-1]}] samples")
        X_rnd = apply_fixed_features(X_rnd, fixed_features)
    elif inequality_constraints or equality_constraints:
        X_rnd = sample_q_batches_from_polytope(raw_samples, q, bounds, inequality_constraints, equality_constraints,
                                               seed=seed, n_burnin=options.get("n_burnin", 10000),
                                               n_thinning=options.get("thinning", 32), fixed_features=fixed_features)
    else:
        # raw samples are drawn where they are scored: no host Sobol draw, no H2D of raw_samples * q * d doubles
        dev = acq_function.model.device
        X_rnd = apply_fixed_features(draw_sobol_samples(bounds, raw_samples, q, seed=seed,
                                                        device=dev if torch.device(dev).type == "cuda" else None),
                                     fixed_features)
    with torch.no_grad():
        Y_rnd = acq_function(X_rnd.to(acq_function.model.device))
    X_ic, idcs = initialize_q_batch(X_rnd, Y_rnd, n=num_restarts, eta=options.get("eta", 2.0))
    return X_ic.cpu(), Y_rnd.cpu()[idcs], X_rnd, Y_rnd


def gen_candidates_scipy(initial_conditions: torch.Tensor, acquisition_function, lower_bounds, upper_bounds,
                         fixed_features: Optional[Dict[int, float]] = None, options: Optional[dict] = None,
                         inequality_constraints=None, equality_constraints=None, nonlinear_inequality_constraints=None):
    """[UPSTREAM] botorch.generation.gen_candidates_scipy: joint L-BFGS-B over all restarts for box bounds, SLSQP
    with the linear constraints replicated per restart when (in)equality constraints are given.  Nonlinear inequality
    constraints ([(callable, is_intrapoint)], feasible iff callable(x) >= 0; NChooseK / Product, utils/torch_tools.py:
    147-252) also go to SLSQP, one restart at a time (BoFire forces batch_limit = 1 for them, botorch.py:117-121), with
    their Jacobians from torch autograd on the host; every start point must satisfy them.
    Returns (candidates [r, q, d] CPU, acq values [r] CPU).  options: maxiter (default 2000, BoFire's
    `maxiter`), gradient ("analytic" | "fd"), fd_step (relative to the bound width, default 1e-6)."""
    options = options or {}
    maxiter = int(options.get("maxiter", 2000))
    use_fd = options.get("gradient", "analytic") == "fd"
    rel_h = float(options.get("fd_step", 1e-6))
    X0 = torch.as_tensor(initial_conditions, dtype=torch.double).cpu().clone()
    r, q, d = X0.shape
    nlcs = _split_nlc(nonlinear_inequality_constraints)
    if nlcs and r > 1:
        parts = [gen_candidates_scipy(X0[i:i + 1], acquisition_function, lower_bounds, upper_bounds, fixed_features=fixed_features,
                                      options=options, inequality_constraints=inequality_constraints,
                                      equality_constraints=equality_constraints,
                                      nonlinear_inequality_constraints=nonlinear_inequality_constraints) for i in range(r)]
        return (torch.cat([p_[0] for p_ in parts]), torch.cat([p_[1] for p_ in parts]),
                {"nit": max(p_[2]["nit"] for p_ in parts), "n_acqf_evals": sum(p_[2]["n_acqf_evals"] for p_ in parts),
                 "message": "; ".join(sorted({p_[2]["message"] for p_ in parts}))})
    lb = torch.as_tensor(lower_bounds, dtype=torch.double).cpu().expand(d).clone()
    ub = torch.as_tensor(upper_bounds, dtype=torch.double).cpu().expand(d).clone()
    free = [j for j in range(d) if not (fixed_features and j in fixed_features) and float(ub[j]) > float(lb[j])]
    X0 = apply_fixed_features(X0, fixed_features)
    h = rel_h * (ub - lb)
    nf = len(free)
    free_t = torch.tensor(free, dtype=torch.long)
    device = acquisition_function.model.device
    state = {"n_eval": 0}

    def unpack(x):
        X = X0.clone()
        X[:, :, free_t] = torch.from_numpy(x).view(r, q, nf)
        return X

    # index tables of the 2*q*nf perturbed copies: copy k = 1 + 2*(p*nf + a) moves point p along free dim a by +h,
    # copy k + 1 by -h
    pp, aa = torch.meshgrid(torch.arange(q), torch.arange(nf), indexing="ij")
    pp, aa = pp.reshape(-1), aa.reshape(-1)
    kk = 1 + 2 * (pp * nf + aa)
    jj = free_t[aa]
    hvec = h[free_t]                       # [nf]
    lbf, ubf = lb[free_t], ub[free_t]

    def f_and_grad_analytic(x):
        X = unpack(np.ascontiguousarray(x))
        vals, dX = acquisition_function.forward_backward(X.to(device))
        state["n_eval"] += r
        g = dX.cpu()[:, :, free_t]
        f = float(vals.sum())
        if not np.isfinite(f):             # a q-batch whose conditional root failed: steer L-BFGS-B away
            return 1e300, np.zeros(r * q * nf)
        return -f, (-g).reshape(-1).numpy().astype(np.float64)

    def f_and_grad_fd(x):
        X = unpack(np.ascontiguousarray(x))
        Xf = X[:, :, free_t]               # [r, q, nf]
        # stay inside the box: shrink the step on the side that would leave it (one-sided at the boundary)
        SP = torch.minimum((ubf - Xf).clamp_min(0.0), hvec.expand_as(Xf))
        SM = torch.minimum((Xf - lbf).clamp_min(0.0), hvec.expand_as(Xf))
        P = X.unsqueeze(1).repeat(1, 1 + 2 * q * nf, 1, 1)   # [r, 1 + 2*q*nf, q, d]
        P[:, kk, pp, jj] += SP.reshape(r, -1)
        P[:, kk + 1, pp, jj] -= SM.reshape(r, -1)
        with torch.no_grad():
            vals = acquisition_function(P.view(-1, q, d).to(device)).cpu().view(r, 1 + 2 * q * nf)
        state["n_eval"] += vals.numel()
        f0 = vals[:, 0]
        vp = vals[:, 1::2].reshape(r, q, nf)
        vm = vals[:, 2::2].reshape(r, q, nf)
        width = SP + SM
        g = torch.where(width > 0, (vp - vm) / width.clamp_min(1e-300), torch.zeros_like(width))
        return -float(f0.sum()), (-g).reshape(-1).numpy().astype(np.float64)

    x0 = X0[:, :, free_t].reshape(-1).numpy().astype(np.float64)
    bnds = [(float(lb[j]), float(ub[j])) for _ in range(r * q) for j in free]
    fun = f_and_grad_fd if use_fd else f_and_grad_analytic
    if inequality_constraints or equality_constraints or nlcs:
        cons = []
        if nlcs:
            if not bool(nonlinear_constraints_satisfied(X0, nlcs).all()):
                raise ValueError("`batch_initial_conditions` must satisfy the non-linear inequality constraints.")

            def nlc_pair(fn, point):
                # value and Jacobian (w.r.t. the free variables) of one constraint at one point (or on the whole q-batch)
                def evaluate(x):
                    xt = torch.from_numpy(np.ascontiguousarray(x)).requires_grad_(True)
                    X = X0.clone()
                    X[:, :, free_t] = xt.view(r, q, nf)
                    v = fn(X[0, point]) if point is not None else fn(X[0])
                    v = torch.as_tensor(v, dtype=torch.double).reshape(())
                    (gx,) = torch.autograd.grad(v, xt, allow_unused=True)
                    return float(v.detach()), (np.zeros(x.shape[0]) if gx is None else gx.numpy())
                return evaluate

            for fn, intra in nlcs:
                for point in (range(q) if intra else [None]):
                    ev = nlc_pair(fn, point)
                    cons.append({"type": "ineq", "fun": (lambda x, ev=ev: ev(x)[0]), "jac": (lambda x, ev=ev: ev(x)[1])})
        free_cols = np.asarray([j * d + a for j in range(q) for a in free])
        x_fixed = X0[0].reshape(-1).numpy()     # fixed columns hold the same value in every restart
        fixed_cols = np.setdiff1d(np.arange(q * d), free_cols)
        for cset, ctype in ((inequality_constraints, "ineq"), (equality_constraints, "eq")):
            A1, r1 = dense_linear_constraints(cset, q, d)
            if not A1.shape[0]:
                continue
            r1 = r1 - A1[:, fixed_cols] @ x_fixed[fixed_cols]
            A1 = A1[:, free_cols]
            Abig = np.kron(np.eye(r), A1)       # the same rows for every restart
            rbig = np.tile(r1, r)
            cons.append({"type": ctype, "fun": (lambda x, A=Abig, rr=rbig: A @ x - rr), "jac": (lambda x, A=Abig: A)})
        res = minimize(fun, x0, jac=True, method="SLSQP", bounds=bnds, constraints=cons, options={"maxiter": maxiter})
    else:
        res = minimize(fun, x0, jac=True, method="L-BFGS-B", bounds=bnds, options={"maxiter": maxiter})
    Xf = unpack(np.clip(res.x, [b_[0] for b_ in bnds], [b_[1] for b_ in bnds]))
    with torch.no_grad():
        vals = acquisition_function(Xf.to(device)).cpu()
    return Xf, vals, {"nit": int(res.nit), "n_acqf_evals": state["n_eval"] + r, "message": str(res.message)}


def gen_candidates_device(initial_conditions: torch.Tensor, acquisition_function, lower_bounds, upper_bounds,
                          fixed_features: Optional[Dict[int, float]] = None, options: Optional[dict] = None):
    """The box-constrained case of gen_candidates_scipy without the host in the loop: the restarts are
    refined by bo_acqf_optimize (csrc/lbfgs.cu) -- one projected L-BFGS per restart on the device, scipy's L-BFGS-B
    defaults for memory and stopping rules.  BoTorch hands scipy the SUM over the restarts as one problem; the sum is
    separable, so the restarts are optimised independently here (own curvature memory, own step length, own stopping
    test) and reach the local maxima the joint run converges to.  Fixed features are columns with lb == ub.
    Returns (candidates [r, q, d] CPU, acq values [r] CPU, info) like gen_candidates_scipy."""
    options = options or {}
    X0 = torch.as_tensor(initial_conditions, dtype=torch.double).cpu().clone()
    d = X0.shape[-1]
    lb = torch.as_tensor(lower_bounds, dtype=torch.double).cpu().expand(d).clone()
    ub = torch.as_tensor(upper_bounds, dtype=torch.double).cpu().expand(d).clone()
    X0 = apply_fixed_features(X0, fixed_features)
    for j, v in (fixed_features or {}).items():
        lb[j] = ub[j] = float(v)
    X0 = torch.minimum(torch.maximum(X0, lb), ub)
    X, vals, info = acquisition_function.optimize(X0, lb, ub, maxiter=int(options.get("maxiter", 2000)),
                                                  history=int(options.get("history", 10)), pgtol=float(options.get("pgtol", 1e-5)),
                                                  ftol=float(options.get("ftol", 2.220446049250313e-09)))
    info = dict(info, message="device L-BFGS", optimizer="on-device batched projected L-BFGS (bo_acqf_optimize)")
    return X.cpu(), vals.cpu(), info


def linear_feasibility(X: torch.Tensor, inequality_constraints, equality_constraints, rel_tol: float = 1e-6) -> np.ndarray:
    """[r] bool: the q-batches X[r, q, d] satisfy the linear constraints up to a RELATIVE slack -- rel_tol times
    max(1, |rhs|, ||row||_1 max|x|) per constraint row, so that e.g. a mixture constraint "sums to 100" is not held to
    1e-8 absolute (SLSQP's own stopping tolerance is ~1e-6 on the scaled problem)."""
    r, q, d = X.shape
    flat = X.reshape(r, -1).numpy()
    ok = np.ones(r, dtype=bool)
    xmax = max(1.0, float(np.abs(flat).max())) if flat.size else 1.0
    for cset, is_eq in ((inequality_constraints, False), (equality_constraints, True)):
        A, b = dense_linear_constraints(cset, q, d)
        if not A.shape[0]:
            continue
        tol = rel_tol * np.maximum(1.0, np.maximum(np.abs(b), np.abs(A).sum(axis=1) * xmax))
        res = flat @ A.T - b
        ok &= ((np.abs(res) <= tol) if is_eq else (res >= -tol)).all(axis=1)
    return ok


def refine_restarts(acq_function, X_ic: torch.Tensor, Y_ic: torch.Tensor, bounds: torch.Tensor, fixed_features=None,
                    options: Optional[dict] = None, inequality_constraints=None, equality_constraints=None,
                    nonlinear_inequality_constraints=None):
    """The refine / filter / keep-if-better block shared by optimize_acqf and distributed.sharded_optimize_acqf:
    gen_candidates_scipy on the restarts, refined restarts that left the feasible set are discarded (SLSQP may stop
    slightly outside), and a restart never ends worse than its screened start (piecewise-smooth MC estimate)."""
    import logging

    q = X_ic.shape[1]
    constrained = bool(inequality_constraints or equality_constraints or nonlinear_inequality_constraints)
    which = (options or {}).get("optimizer", "device")
    if which not in ("device", "scipy"):
        raise ValueError("options['optimizer'] must be 'device' or 'scipy'")
    if not constrained and which == "device" and hasattr(acq_function, "optimize") and (options or {}).get("gradient", "analytic") == "analytic":
        X_ref, Y_ref, info = gen_candidates_device(X_ic, acq_function, bounds[0], bounds[1], fixed_features=fixed_features,
                                                   options=options)
    else:
        X_ref, Y_ref, info = gen_candidates_scipy(X_ic, acq_function, bounds[0], bounds[1], fixed_features=fixed_features,
                                                  options=options, inequality_constraints=inequality_constraints,
                                                  equality_constraints=equality_constraints,
                                                  nonlinear_inequality_constraints=nonlinear_inequality_constraints)
    if nonlinear_inequality_constraints:
        ok_n = nonlinear_constraints_satisfied(X_ref, nonlinear_inequality_constraints)
        Y_ref = torch.where(ok_n, Y_ref, torch.full_like(Y_ref, -float("inf")))
    if inequality_constraints or equality_constraints:
        ok = torch.from_numpy(linear_feasibility(X_ref, inequality_constraints, equality_constraints))
        if not bool(ok.any()):
            logging.getLogger(__name__).warning("every refined restart violates the linear constraints: keeping the raw starts")
        Y_ref = torch.where(ok, Y_ref, torch.full_like(Y_ref, -float("inf")))
    better = Y_ref >= Y_ic
    X_out = torch.where(better.view(-1, 1, 1), X_ref, X_ic.cpu())
    Y_out = torch.where(better, Y_ref, Y_ic)
    return X_out, Y_out


def optimize_acqf(acq_function, bounds: torch.Tensor, q: int, num_restarts: int, raw_samples: int,
                  fixed_features: Optional[Dict[int, float]] = None, options: Optional[dict] = None,
                  return_best_only: bool = True, seed: Optional[int] = None, refine: bool = True, **unsupported):
    """Same signature / return convention as botorch.optim.optimize_acqf as BoFire calls it
    (botorch.py:384-405): (candidates [q, d] on CPU, acq_value scalar tensor)."""
    nlcs = unsupported.get("nonlinear_inequality_constraints") or None
    generator = unsupported.get("generator")          # ic_gen_kwargs of botorch.py:258-265
    if nlcs and generator is None:
        # [UPSTREAM] optimize_acqf: "`batch_initial_conditions` or `ic_generator` must be given if there are non-linear
        # inequality constraints" -- Sobol samples are not feasible for NChooseK
        raise RuntimeError("`ic_generator` (ic_gen_kwargs['generator']) must be given if there are non-linear inequality "
                           "constraints.")
    ineq = unsupported.get("inequality_constraints") or None
    eq = unsupported.get("equality_constraints") or None
    bounds = torch.as_tensor(bounds, dtype=torch.double)
    X_ic, Y_ic, _, _ = gen_batch_initial_conditions(acq_function, bounds, q, num_restarts, raw_samples,
                                                    fixed_features=fixed_features, options=options, seed=seed,
                                                    inequality_constraints=ineq, equality_constraints=eq, generator=generator)
    if refine:
        X_ic, Y_ic = refine_restarts(acq_function, X_ic, Y_ic, bounds, fixed_features=fixed_features, options=options,
                                     inequality_constraints=ineq, equality_constraints=eq,
                                     nonlinear_inequality_constraints=nlcs)
    if return_best_only:
        best = int(torch.argmax(Y_ic))
        return X_ic[best].cpu(), Y_ic[best]
    return X_ic.cpu(), Y_ic


def _pending_base(acq_function):
    Xp = getattr(acq_function, "X_pending", None)
    return None if Xp is None else torch.as_tensor(Xp, dtype=torch.double).cpu().reshape(-1, acq_function.model.d)


def optimize_acqf_mixed(acq_function, bounds: torch.Tensor, q: int, num_restarts: int, raw_samples: int,
                        fixed_features_list, options: Optional[dict] = None, seed: Optional[int] = None,
                        refine: bool = True, **unsupported):
    """[UPSTREAM] botorch.optim.optimize_acqf_mixed as BoFire calls it for categorical combinations
    (botorch.py:358-378): for each of the q points in turn, one optimize_acqf(q=1) per fixed-feature dictionary, the
    best value wins, and the chosen point becomes pending (set_X_pending) for the next one.  Returns
    (candidates [q, d], joint acquisition value of the final batch)."""
    if not fixed_features_list:
        raise ValueError("fixed_features_list must be non-empty.")
    base_pending = _pending_base(acq_function)
    chosen = []
    best_v = None
    try:
        for _ in range(q):
            best_c, best_v = None, None
            for ff in fixed_features_list:
                c, v = optimize_acqf(acq_function, bounds, 1, num_restarts, raw_samples, fixed_features=ff, options=options,
                                     seed=seed, refine=refine, **unsupported)
                if best_v is None or float(v) > float(best_v):
                    best_c, best_v = c, v
            chosen.append(best_c.reshape(1, -1))
            if q > 1:
                pend = torch.cat(([base_pending] if base_pending is not None else []) + chosen, dim=0)
                acq_function.set_X_pending(pend)
    finally:
        if q > 1:
            acq_function.set_X_pending(base_pending)
    cands = torch.cat(chosen, dim=0)
    if q > 1:
        with torch.no_grad():
            best_v = acq_function(cands.unsqueeze(0).to(acq_function.model.device)).cpu()[0]
    return cands, best_v


def optimize_acqf_list(acq_function_list, bounds: torch.Tensor, num_restarts: int, raw_samples: int,
                       fixed_features: Optional[Dict[int, float]] = None, fixed_features_list=None,
                       options: Optional[dict] = None, seed: Optional[int] = None, refine: bool = True, **unsupported):
    """[UPSTREAM] botorch.optim.optimize_acqf_list as BotorchStrategy._optimize_acqf_continuous calls it when a strategy
    hands it more than one acquisition function (botorch.py:337-356, e.g. one scalarisation per candidate): the candidates
    are generated one at a time, acquisition function i sees the candidates of 0..i-1 as pending points, each step is an
    optimize_acqf(q=1) -- or optimize_acqf_mixed(q=1) when `fixed_features_list` is given.  Returns
    (candidates [len(list), d], acquisition values [len(list)]).  All functions may share one DeviceGPState: each is
    re-activated on the handle before its turn."""
    if not acq_function_list:
        raise ValueError("acq_function_list must be non-empty.")
    if fixed_features and fixed_features_list:
        raise ValueError("Either fixed_features or fixed_features_list can be provided, not both.")
    chosen, values = [], []
    for acqf in acq_function_list:
        acqf.activate()
        base_pending = _pending_base(acqf)
        if chosen:
            acqf.set_X_pending(torch.cat(([base_pending] if base_pending is not None else []) + chosen, dim=0))
        try:
            if fixed_features_list:
                c, v = optimize_acqf_mixed(acqf, bounds, 1, num_restarts, raw_samples, fixed_features_list=fixed_features_list,
                                           options=options, seed=seed, refine=refine, **unsupported)
            else:
                c, v = optimize_acqf(acqf, bounds, 1, num_restarts, raw_samples, fixed_features=fixed_features,
                                     options=options, seed=seed, refine=refine, **unsupported)
        finally:
            if chosen:
                acqf.set_X_pending(base_pending)
        chosen.append(c.reshape(1, -1))
        values.append(torch.as_tensor(v, dtype=torch.double).reshape(()))
    return torch.cat(chosen, dim=0), torch.stack(values)


def calc_acquisition(acq_function, candidates, combined: bool = False):
    """BotorchStrategy.calc_acquisition (botorch.py:196-225) on already transformed candidates [n, d]:
    one value per row, or a single value for the whole set as one q-batch when `combined`."""
    X = torch.as_tensor(candidates, dtype=torch.double)
    if not combined:
        X = X.unsqueeze(-2)
    with torch.no_grad():
        return acq_function(X).cpu().numpy()


def optimize_acqf_discrete(acq_function, q: int, choices: torch.Tensor, max_batch_size: int = 1 << 20, unique: bool = True):
    """[UPSTREAM] optimize_acqf_discrete as used by the all-categorical branch (botorch.py:425-467): sequential greedy
    selection over a discrete choice set, forward-only; after each pick the point becomes pending and (unique=True)
    leaves the choice set.  Returns (candidates [q, d], values [q] of the sequential picks; a scalar for q == 1)."""
    choices = torch.as_tensor(choices, dtype=torch.double)
    if choices.dim() != 2 or choices.shape[0] == 0:
        raise ValueError("`choices` must be a non-empty [n, d] tensor.")
    if unique and q > choices.shape[0]:
        raise ValueError(f"Requested {q} unique candidates from a choice set of {choices.shape[0]}.")
    base_pending = _pending_base(acq_function)
    picked, vals = [], []
    remaining = choices
    try:
        for _ in range(q):
            with torch.no_grad():
                v = torch.cat([acq_function(remaining[i:i + max_batch_size].unsqueeze(-2).to(acq_function.model.device)).cpu()
                               for i in range(0, remaining.shape[0], max_batch_size)])
            best = int(torch.argmax(v))
            picked.append(remaining[best].unsqueeze(0))
            vals.append(v[best])
            if q > 1:
                pend = torch.cat(([base_pending] if base_pending is not None else []) + picked, dim=0)
                acq_function.set_X_pending(pend)
                if unique:
                    remaining = torch.cat([remaining[:best], remaining[best + 1:]], dim=0)
    finally:
        if q > 1:
            acq_function.set_X_pending(base_pending)
    if q == 1:
        return picked[0], vals[0]
    return torch.cat(picked, dim=0), torch.stack(vals)

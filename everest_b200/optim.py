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


def draw_sobol_samples(bounds: torch.Tensor, n: int, q: int, seed: Optional[int] = None) -> torch.Tensor:
    """[UPSTREAM] botorch.utils.sampling.draw_sobol_samples: [n, q, d] inside bounds [2, d]."""
    d = bounds.shape[-1]
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


def gen_batch_initial_conditions(acq_function, bounds: torch.Tensor, q: int, num_restarts: int, raw_samples: int,
                                 fixed_features: Optional[Dict[int, float]] = None, options: Optional[dict] = None,
                                 seed: Optional[int] = None):
    """[UPSTREAM] gen_batch_initial_conditions without linear constraints: Sobol raw samples, one screened
    forward over ALL of them on the device, then initialize_q_batch."""
    options = options or {}
    if seed is None:
        seed = int(torch.randint(0, 1000000, (1,)).item())
    X_rnd = apply_fixed_features(draw_sobol_samples(bounds, raw_samples, q, seed=seed), fixed_features)
    with torch.no_grad():
        Y_rnd = acq_function(X_rnd.to(acq_function.model.device))
    X_ic, idcs = initialize_q_batch(X_rnd, Y_rnd, n=num_restarts, eta=options.get("eta", 2.0))
    return X_ic, Y_rnd.cpu()[idcs], X_rnd, Y_rnd


def gen_candidates_scipy(initial_conditions: torch.Tensor, acquisition_function, lower_bounds, upper_bounds,
                         fixed_features: Optional[Dict[int, float]] = None, options: Optional[dict] = None):
    """[UPSTREAM] botorch.generation.gen_candidates_scipy for box bounds: joint L-BFGS-B over all restarts.
    Returns (candidates [r, q, d] CPU, acq values [r] CPU).  options: maxiter (default 2000, BoFire's
    `maxiter`), gradient ("analytic" | "fd"), fd_step (relative to the bound width, default 1e-6)."""
    options = options or {}
    maxiter = int(options.get("maxiter", 2000))
    use_fd = options.get("gradient", "analytic") == "fd"
    rel_h = float(options.get("fd_step", 1e-6))
    X0 = torch.as_tensor(initial_conditions, dtype=torch.double).cpu().clone()
    r, q, d = X0.shape
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
    res = minimize(f_and_grad_fd if use_fd else f_and_grad_analytic, x0, jac=True, method="L-BFGS-B", bounds=bnds, options={"maxiter": maxiter})
    Xf = unpack(np.clip(res.x, [b_[0] for b_ in bnds], [b_[1] for b_ in bnds]))
    with torch.no_grad():
        vals = acquisition_function(Xf.to(device)).cpu()
    return Xf, vals, {"nit": int(res.nit), "n_acqf_evals": state["n_eval"] + r, "message": str(res.message)}


def optimize_acqf(acq_function, bounds: torch.Tensor, q: int, num_restarts: int, raw_samples: int,
                  fixed_features: Optional[Dict[int, float]] = None, options: Optional[dict] = None,
                  return_best_only: bool = True, seed: Optional[int] = None, refine: bool = True, **unsupported):
    """Same signature / return convention as botorch.optim.optimize_acqf as BoFire calls it
    (botorch.py:384-405): (candidates [q, d] on CPU, acq_value scalar tensor)."""
    for key in ("equality_constraints", "inequality_constraints", "nonlinear_inequality_constraints"):
        if unsupported.get(key):
            raise NotImplementedError(f"{key} are not handled by the accelerated optimiser yet")
    bounds = torch.as_tensor(bounds, dtype=torch.double)
    X_ic, Y_ic, _, _ = gen_batch_initial_conditions(acq_function, bounds, q, num_restarts, raw_samples,
                                                    fixed_features=fixed_features, options=options, seed=seed)
    if refine:
        X_ref, Y_ref, _ = gen_candidates_scipy(X_ic, acq_function, bounds[0], bounds[1], fixed_features=fixed_features,
                                               options=options)
        # never return something worse than the screened start (piecewise-smooth MC estimate)
        better = Y_ref >= Y_ic
        X_ic = torch.where(better.view(-1, 1, 1), X_ref, X_ic.cpu())
        Y_ic = torch.where(better, Y_ref, Y_ic)
    if return_best_only:
        best = int(torch.argmax(Y_ic))
        return X_ic[best].cpu(), Y_ic[best]
    return X_ic.cpu(), Y_ic


def optimize_acqf_mixed(acq_function, bounds: torch.Tensor, q: int, num_restarts: int, raw_samples: int,
                        fixed_features_list, options: Optional[dict] = None, seed: Optional[int] = None,
                        refine: bool = True, **unsupported):
    """[UPSTREAM] botorch.optim.optimize_acqf_mixed as BoFire calls it for categorical combinations
    (botorch.py:358-378), q == 1: one optimize_acqf per fixed-feature dictionary, best value wins.
    (For q > 1 BoTorch adds points sequentially with X_pending updates; that needs an acqf rebuild per pick
    and is not accelerated.)"""
    if not fixed_features_list:
        raise ValueError("fixed_features_list must be non-empty.")
    if q != 1:
        raise NotImplementedError("optimize_acqf_mixed with q > 1 (sequential X_pending updates) is not accelerated")
    best_c, best_v = None, None
    for ff in fixed_features_list:
        c, v = optimize_acqf(acq_function, bounds, q, num_restarts, raw_samples, fixed_features=ff, options=options,
                             seed=seed, refine=refine, **unsupported)
        if best_v is None or float(v) > float(best_v):
            best_c, best_v = c, v
    return best_c, best_v


def calc_acquisition(acq_function, candidates, combined: bool = False):
    """BotorchStrategy.calc_acquisition (botorch.py:196-225) on already transformed candidates [n, d]:
    one value per row, or a single value for the whole set as one q-batch when `combined`."""
    X = torch.as_tensor(candidates, dtype=torch.double)
    if not combined:
        X = X.unsqueeze(-2)
    with torch.no_grad():
        return acq_function(X).cpu().numpy()


def optimize_acqf_discrete(acq_function, q: int, choices: torch.Tensor, max_batch_size: int = 1 << 20, unique: bool = True):
    """[UPSTREAM] optimize_acqf_discrete as used by the all-categorical branch (botorch.py:425-467):
    sequential greedy selection over a discrete choice set, forward-only."""
    if q != 1:
        raise NotImplementedError("sequential greedy q > 1 needs X_pending updates (set-up cost per pick); q == 1 is accelerated")
    choices = torch.as_tensor(choices, dtype=torch.double)
    with torch.no_grad():
        vals = acq_function(choices.unsqueeze(-2).to(acq_function.model.device)).cpu()
    best = int(torch.argmax(vals))
    return choices[best].unsqueeze(0), vals[best]

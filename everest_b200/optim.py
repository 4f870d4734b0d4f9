"""Candidate generation around the acquisition function: the L2 boundary of SURVEY.md 8b
(`BotorchStrategy._optimize_acqf_continuous`, strategies/predictives/botorch.py:326-406, which calls
botorch.optim.optimize_acqf / optimize_acqf_discrete).

Round-1 scope: the raw-sample screen runs as ONE device call over all raw_samples q-batches (the
reference scores them in raw_samples / batch_limit sequential CPU calls), then
`initialize_q_batch` picks the restarts.  The gradient refinement of the restarts ([UPSTREAM]
gen_candidates_scipy, L-BFGS-B) needs d acqf / d X and is listed as the next step in DESIGN.md; until
then `optimize_acqf` returns the best screened restart.
"""
from typing import Dict, Optional, Tuple

import torch

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


def optimize_acqf(acq_function, bounds: torch.Tensor, q: int, num_restarts: int, raw_samples: int,
                  fixed_features: Optional[Dict[int, float]] = None, options: Optional[dict] = None,
                  return_best_only: bool = True, seed: Optional[int] = None, **unsupported):
    """Same signature / return convention as botorch.optim.optimize_acqf as BoFire calls it
    (botorch.py:384-405): (candidates [q, d] on CPU, acq_value scalar tensor)."""
    for key in ("equality_constraints", "inequality_constraints", "nonlinear_inequality_constraints"):
        if unsupported.get(key):
            raise NotImplementedError(f"{key} are not handled by the accelerated optimiser yet")
    X_ic, Y_ic, _, _ = gen_batch_initial_conditions(acq_function, bounds, q, num_restarts, raw_samples,
                                                    fixed_features=fixed_features, options=options, seed=seed)
    if return_best_only:
        best = int(torch.argmax(Y_ic))
        return X_ic[best].cpu(), Y_ic[best]
    return X_ic.cpu(), Y_ic


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

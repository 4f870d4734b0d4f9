"""Multi-GPU sharding of the raw-sample screen: one process per GPU, q-batches split
contiguously across ranks, model state replicated, and ONE exchange at the end -- an all-gather of the
acquisition values (initialize_q_batch needs max / mean / std of all of them) or, for pure arg-max
selection (optimize_acqf_discrete), an all-reduce of the packed (value, index) pair.  torch.distributed
(NCCL over NVLink on the GPU box, gloo in the CPU tests) is only the plumbing."""
from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced split of n units: the first n % world ranks get one extra unit."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def sharded_forward(acq: Callable[[torch.Tensor], torch.Tensor], X: torch.Tensor, group=None) -> torch.Tensor:
    """Every rank scores its slice of X[n, q, d]; returns all n values on every rank."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return acq(X)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n = X.shape[0]
    lo, hi = shard_bounds(n, rank, world)
    out_device = getattr(getattr(acq, "model", None), "device", None) or X.device
    vals = acq(X[lo:hi]) if hi > lo else torch.empty(0, dtype=torch.double, device=out_device)
    width = (n + world - 1) // world
    buf = torch.zeros(width, dtype=torch.double, device=vals.device)
    buf[: hi - lo] = vals
    gathered = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(gathered, buf, group=group)
    out = []
    for r in range(world):
        rlo, rhi = shard_bounds(n, r, world)
        out.append(gathered[r][: rhi - rlo])
    return torch.cat(out)


def sharded_argmax(acq: Callable[[torch.Tensor], torch.Tensor], X: torch.Tensor, group=None) -> Tuple[float, int]:
    """Best acquisition value and its global index with a single MAX all-reduce of a packed key.
    Ties resolve to the smallest index, like torch.argmax on the gathered vector."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        v = acq(X)
        i = int(torch.argmax(v))
        return float(v[i]), i
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n = X.shape[0]
    lo, hi = shard_bounds(n, rank, world)
    if hi > lo:
        v = acq(X[lo:hi])
        i = int(torch.argmax(v))
        best = torch.tensor([float(v[i]), -float(lo + i)], dtype=torch.double, device=v.device)
    else:
        best = torch.tensor([float("-inf"), -float(n)], dtype=torch.double, device=X.device)
    # lexicographic max of (value, -index): gather the world pairs (2 doubles per rank) and reduce locally
    pairs = [torch.empty_like(best) for _ in range(world)]
    dist.all_gather(pairs, best, group=group)
    top = max(pairs, key=lambda t: (float(t[0]), float(t[1])))
    return float(top[0]), int(-float(top[1]))


def sharded_optimize_acqf(acq_function, bounds: torch.Tensor, q: int, num_restarts: int, raw_samples: int,
                          fixed_features=None, options: Optional[dict] = None, seed: int = 0, group=None,
                          inequality_constraints=None, equality_constraints=None, nonlinear_inequality_constraints=None,
                          generator=None):
    """optimize_acqf over the GPUs of one box: every rank scores its slice of the raw samples (one
    all-gather of the values, because initialize_q_batch needs all of them), every rank refines ITS share of the
    restarts with L-BFGS-B / SLSQP without any per-iteration collective, and the only other exchange is the arg-max over
    the per-rank best (value, rank) -- "only the per-restart best-acquisition argmax is reduced" -- followed by the
    broadcast of the winning [q, d] candidate.  All ranks must pass the same `seed`; returns the same (candidate, value)
    on every rank."""
    from . import optim

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return optim.optimize_acqf(acq_function, bounds, q, num_restarts, raw_samples, fixed_features=fixed_features,
                                   options=options, seed=seed, inequality_constraints=inequality_constraints,
                                   equality_constraints=equality_constraints,
                                   nonlinear_inequality_constraints=nonlinear_inequality_constraints, generator=generator)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    bounds = torch.as_tensor(bounds, dtype=torch.double)
    dev = acq_function.model.device
    if nonlinear_inequality_constraints and generator is None:
        raise RuntimeError("`ic_generator` (generator) must be given if there are non-linear inequality constraints.")
    # identical raw samples on every rank (same seed), scored in slices
    if generator is not None:
        X_rnd = optim.apply_fixed_features(torch.as_tensor(generator(raw_samples, q, seed), dtype=torch.double), fixed_features)
    elif inequality_constraints or equality_constraints:
        X_rnd = optim.sample_q_batches_from_polytope(raw_samples, q, bounds, inequality_constraints, equality_constraints,
                                                     seed=seed, fixed_features=fixed_features)
    else:
        # drawn where they are scored, like optimize_acqf: no host Sobol draw and no H2D of raw_samples * q * d doubles (the
        # device generator is bit-identical to torch's engine for the same seed, so every rank still holds the same set)
        X_rnd = optim.apply_fixed_features(optim.draw_sobol_samples(bounds, raw_samples, q, seed=seed,
                                                                    device=dev if torch.device(dev).type == "cuda" else None),
                                           fixed_features)
    with torch.no_grad():
        Y_rnd = sharded_forward(lambda x: acq_function(x.to(dev)), X_rnd, group=group).cpu()
    gen = torch.Generator().manual_seed(int(seed))          # the same multinomial draw on every rank
    X_ic, idcs = optim.initialize_q_batch(X_rnd, Y_rnd, n=num_restarts, eta=(options or {}).get("eta", 2.0), generator=gen)
    X_ic = X_ic.cpu()
    Y_ic = Y_rnd[idcs]
    lo, hi = shard_bounds(num_restarts, rank, world)
    best_val, best_x = float("-inf"), torch.zeros(q, bounds.shape[-1], dtype=torch.double)
    if hi > lo:
        Xb, Yb = optim.refine_restarts(acq_function, X_ic[lo:hi], Y_ic[lo:hi], bounds, fixed_features=fixed_features,
                                       options=options, inequality_constraints=inequality_constraints,
                                       equality_constraints=equality_constraints,
                                       nonlinear_inequality_constraints=nonlinear_inequality_constraints)
        i = int(torch.argmax(Yb))
        best_val, best_x = float(Yb[i]), Xb[i]
    pair = torch.tensor([best_val, -float(rank)], dtype=torch.double, device=dev)
    pairs = [torch.empty_like(pair) for _ in range(world)]
    dist.all_gather(pairs, pair, group=group)
    top = max(pairs, key=lambda t: (float(t[0]), float(t[1])))
    winner = int(-float(top[1]))
    cand = best_x.to(dev).contiguous()
    dist.broadcast(cand, src=dist.get_global_rank(group, winner) if group is not None else winner, group=group)
    return cand.cpu(), torch.tensor(float(top[0]), dtype=torch.double)

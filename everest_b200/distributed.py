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
    vals = acq(X[lo:hi]) if hi > lo else torch.empty(0, dtype=torch.double, device=X.device)
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

"""Kernel specifications: the host-side mirror of ``bofire.kernels.mapper`` (kernels/mapper.py:31-302).

Where the reference maps a kernel data model to a gpytorch module, this module maps it to a plain
spec tree and flattens the tree into the sum-of-products form the CUDA kernels evaluate:
``K = sum_t coef_t * prod_{l in t} leaf_l`` (include/everest_b200.h, bo_kernel_term).
Class and argument names follow bofire.data_models.kernels.
"""
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple, Union

from . import _lib as L


@dataclass
class RBFKernel:
    """data_models/kernels/continuous.py RBFKernel -> gpytorch RBFKernel (mapper.py:31-46)."""
    active_dims: Sequence[int]
    lengthscale: Sequence[float]  # 1 value or one per active dim (ard=True)


@dataclass
class MaternKernel:
    """data_models/kernels/continuous.py:19-29 MaternKernel, nu in {0.5, 1.5, 2.5} (mapper.py:49-69)."""
    active_dims: Sequence[int]
    lengthscale: Sequence[float]
    nu: float = 2.5


@dataclass
class HammingDistanceKernel:
    """One-hot Hamming kernel (mapper.py:206-253, kernels/categorical.py:43-70).
    categorical_features = {start_column: cardinality} exactly like HammingKernelWithOneHots."""
    categorical_features: Dict[int, int]
    lengthscale: Sequence[float] = (1.0,)


@dataclass
class TanimotoKernel:
    """mapper.py:191-203, fingerprint_kernels/tanimoto_kernel.py:68-77 (no lengthscale)."""
    active_dims: Sequence[int]


@dataclass
class ScaleKernel:
    """mapper.py:168-188."""
    base_kernel: object
    outputscale: float = 1.0


@dataclass
class AdditiveKernel:
    """mapper.py:126-145."""
    kernels: Sequence[object]


@dataclass
class MultiplicativeKernel:
    """mapper.py:148-165."""
    kernels: Sequence[object]


AnyKernel = Union[RBFKernel, MaternKernel, HammingDistanceKernel, TanimotoKernel, ScaleKernel, AdditiveKernel,
                  MultiplicativeKernel]

_NU_TO_KIND = {0.5: L.LEAF_MATERN12, 1.5: L.LEAF_MATERN32, 2.5: L.LEAF_MATERN52}


@dataclass
class FlatKernel:
    leaves: List[object] = field(default_factory=list)
    terms: List[Tuple[float, List[int]]] = field(default_factory=list)  # (coef, leaf indices)


def flatten(kernel: AnyKernel) -> FlatKernel:
    """Expand a Scale/Additive/Multiplicative tree into a sum of products of leaves."""
    flat = FlatKernel()

    def rec(k) -> List[Tuple[float, List[int]]]:
        if isinstance(k, (RBFKernel, MaternKernel, HammingDistanceKernel, TanimotoKernel)):
            if isinstance(k, MaternKernel) and k.nu not in _NU_TO_KIND:
                raise ValueError(f"Matern nu={k.nu} not in (0.5, 1.5, 2.5)")
            if isinstance(k, TanimotoKernel):
                # parameter-free: the Tanimoto leaf of the additive part and of the product part of MixedTanimotoGP
                # (mixed_tanimoto_gp.py:165-215) are the same function -- share the leaf so that the device evaluates the
                # 2048-bit AND + POPC once per pair (leaf-major evaluation in crosscov_kernel2).  Leaves with
                # hyper-parameters are never merged: their lengthscale slots belong to the fit.
                for i, other in enumerate(flat.leaves):
                    if isinstance(other, TanimotoKernel) and list(other.active_dims) == list(k.active_dims):
                        return [(1.0, [i])]
            flat.leaves.append(k)
            return [(1.0, [len(flat.leaves) - 1])]
        if isinstance(k, ScaleKernel):
            return [(c * float(k.outputscale), f) for c, f in rec(k.base_kernel)]
        if isinstance(k, AdditiveKernel):
            out = []
            for c in k.kernels:
                out += rec(c)
            return out
        if isinstance(k, MultiplicativeKernel):
            acc = [(1.0, [])]
            for c in k.kernels:
                terms = rec(c)
                acc = [(c1 * c2, f1 + f2) for c1, f1 in acc for c2, f2 in terms]
            return acc
        raise NotImplementedError(f"kernel {type(k).__name__} is outside the accelerated path "
                                  "(supported: RBF, Matern, Hamming, Tanimoto, Scale, Additive, Multiplicative)")

    flat.terms = rec(kernel)
    if len(flat.leaves) > 8 or len(flat.terms) > 8 or any(len(f) > L.BO_MAX_FACTORS for _, f in flat.terms):
        raise ValueError("kernel tree too large for the device representation (<= 8 leaves, 8 terms, 4 factors)")
    return flat


def leaf_to_c(leaf, keep):
    """Fill a KernelLeaf ctypes struct; `keep` collects the arrays that must stay alive."""
    import ctypes as C

    out = L.KernelLeaf()

    def iarr(v):
        a = (C.c_int32 * len(v))(*[int(x) for x in v])
        keep.append(a)
        return C.cast(a, L.c_int_p)

    def darr(v):
        a = (C.c_double * len(v))(*[float(x) for x in v])
        keep.append(a)
        return C.cast(a, L.c_double_p)

    if isinstance(leaf, (RBFKernel, MaternKernel)):
        out.kind = L.LEAF_RBF if isinstance(leaf, RBFKernel) else _NU_TO_KIND[leaf.nu]
        out.n_dims = len(leaf.active_dims)
        out.dims = iarr(leaf.active_dims)
        ls = list(leaf.lengthscale)
        if len(ls) not in (1, out.n_dims):
            raise ValueError("lengthscale must have 1 or len(active_dims) entries")
        out.lengthscale = darr(ls)
        out.n_ls = len(ls)
    elif isinstance(leaf, HammingDistanceKernel):
        groups = sorted(leaf.categorical_features.items())
        out.kind = L.LEAF_HAMMING
        out.n_dims = len(groups)
        out.dims = iarr([g[0] for g in groups])
        out.cardinality = iarr([g[1] for g in groups])
        ls = list(leaf.lengthscale)
        if len(ls) != 1 and len(ls) < len(groups):
            raise ValueError("Hamming lengthscale needs 1 or >= n_groups entries")
        out.lengthscale = darr(ls)
        out.n_ls = len(ls)
    elif isinstance(leaf, TanimotoKernel):
        out.kind = L.LEAF_TANIMOTO
        out.n_dims = len(leaf.active_dims)
        out.dims = iarr(leaf.active_dims)
        out.n_ls = 0
    else:
        raise TypeError(type(leaf))
    return out

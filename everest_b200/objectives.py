"""Objective / constraint op tables: host-side mirror of ``bofire.utils.torch_tools``
(get_objective_callable :384-450, get_multiobjective_objective :699-727,
get_additive_botorch_objective :676-697, get_multiplicative_botorch_objective :662-674,
constrained_objective2botorch :258-337).  Each spec is also callable on CPU tensors with the
reference's formula, because BoFire evaluates these callables outside the acquisition loop too
(e.g. utils/multiobjective.py:76-83)."""
from dataclasses import dataclass
from typing import List, Optional, Sequence

import torch

from . import _lib as L


@dataclass
class ObjectiveSpec:
    kind: str          # max | min | close_to_target | min_sigmoid | max_sigmoid | target
    idx: int           # output index
    p0: float = 0.0
    p1: float = 1.0
    p2: float = 0.0
    w: float = 1.0

    _KINDS = {"max": L.OBJ_MAX, "min": L.OBJ_MIN, "close_to_target": L.OBJ_CLOSE_TO_TARGET,
              "min_sigmoid": L.OBJ_MIN_SIGMOID, "max_sigmoid": L.OBJ_MAX_SIGMOID, "target": L.OBJ_TARGET}

    def to_c(self):
        if self.kind not in self._KINDS:
            raise NotImplementedError(f"Objective {self.kind} not implemented.")
        return L.ObjectiveOp(self._KINDS[self.kind], int(self.idx), float(self.p0), float(self.p1), float(self.p2),
                             float(self.w))

    def __call__(self, y: torch.Tensor, X=None) -> torch.Tensor:
        v = y[..., self.idx]
        k = self.kind
        if k == "max":
            return (v - self.p0) / (self.p1 - self.p0)
        if k == "min":
            return -1.0 * ((v - self.p0) / (self.p1 - self.p0))
        if k == "close_to_target":
            return -1.0 * (torch.abs(v - self.p0) ** self.p1)
        if k == "min_sigmoid":
            return 1.0 - 1.0 / (1.0 + torch.exp(-1.0 * self.p0 * (v - self.p1)))
        if k == "max_sigmoid":
            return 1.0 / (1.0 + torch.exp(-1.0 * self.p0 * (v - self.p1)))
        if k == "target":
            return (1.0 / (1.0 + torch.exp(-1 * self.p2 * (v - (self.p0 - self.p1))))
                    * (1.0 - 1.0 / (1.0 + torch.exp(-1.0 * self.p2 * (v - (self.p0 + self.p1))))))
        raise NotImplementedError(f"Objective {k} not implemented.")


def MaximizeObjective(idx, lower_bound=0.0, upper_bound=1.0, w=1.0):
    return ObjectiveSpec("max", idx, lower_bound, upper_bound, 0.0, w)


def MinimizeObjective(idx, lower_bound=0.0, upper_bound=1.0, w=1.0):
    return ObjectiveSpec("min", idx, lower_bound, upper_bound, 0.0, w)


def CloseToTargetObjective(idx, target_value, exponent=1.0, w=1.0):
    return ObjectiveSpec("close_to_target", idx, target_value, exponent, 0.0, w)


def MinimizeSigmoidObjective(idx, steepness, tp, w=1.0):
    return ObjectiveSpec("min_sigmoid", idx, steepness, tp, 0.0, w)


def MaximizeSigmoidObjective(idx, steepness, tp, w=1.0):
    return ObjectiveSpec("max_sigmoid", idx, steepness, tp, 0.0, w)


def TargetObjective(idx, target_value, tolerance, steepness, w=1.0):
    return ObjectiveSpec("target", idx, target_value, tolerance, steepness, w)


@dataclass
class MultiObjective:
    """get_multiobjective_objective: stack of per-output callables (maximisation frame)."""
    ops: Sequence[ObjectiveSpec]

    def __call__(self, samples: torch.Tensor, X=None) -> torch.Tensor:
        return torch.stack([op(samples, None) for op in self.ops], dim=-1)


@dataclass
class ScalarObjective:
    """single | additive (sum w*c(y)) | multiplicative (prod c(y)**w)."""
    ops: Sequence[ObjectiveSpec]
    combine: str = "single"

    def __post_init__(self):
        if self.combine == "single" and len(self.ops) != 1:
            raise ValueError("a single objective takes exactly one op")

    @property
    def combine_code(self):
        return {"single": L.COMBINE_SINGLE, "additive": L.COMBINE_ADDITIVE,
                "multiplicative": L.COMBINE_MULTIPLICATIVE}[self.combine]

    def __call__(self, samples: torch.Tensor, X=None) -> torch.Tensor:
        if self.combine == "single":
            return self.ops[0](samples)
        if self.combine == "additive":
            val = torch.tensor(0.0, dtype=samples.dtype)
            for op in self.ops:
                val = val + op(samples) * op.w
            return val
        val = torch.tensor(1.0, dtype=samples.dtype)
        for op in self.ops:
            val = val * op(samples) ** op.w
        return val


@dataclass
class OutputConstraint:
    """c(y) = sign * (y[idx] - tp); feasible iff c <= 0; eta = 1 / steepness."""
    idx: int
    sign: float
    tp: float
    eta: float

    def to_c(self):
        return L.ConstraintOp(int(self.idx), float(self.sign), float(self.tp), float(self.eta))

    def __call__(self, Z: torch.Tensor) -> torch.Tensor:
        return self.sign * (Z[..., self.idx] - self.tp)


def constraints_from_sigmoid_objectives(ops: Sequence[ObjectiveSpec]) -> List[OutputConstraint]:
    """constrained_objective2botorch for Maximize/MinimizeSigmoid and Target objectives."""
    out = []
    for op in ops:
        if op.kind == "max_sigmoid":
            out.append(OutputConstraint(op.idx, -1.0, op.p1, 1.0 / op.p0))
        elif op.kind == "min_sigmoid":
            out.append(OutputConstraint(op.idx, 1.0, op.p1, 1.0 / op.p0))
        elif op.kind == "target":
            out.append(OutputConstraint(op.idx, -1.0, op.p0 - op.p1, 1.0 / op.p2))
            out.append(OutputConstraint(op.idx, 1.0, op.p0 + op.p1, 1.0 / op.p2))
        else:
            raise ValueError(f"Objective {op.kind} is not a ConstrainedObjective.")
    return out


def c_array(items, ctype):
    if not items:
        return None, 0
    arr = (ctype * len(items))(*[it.to_c() for it in items])
    return arr, len(items)

"""Multi-objective helpers of the path: host-side mirror of ``bofire/utils/multiobjective.py``.

* ``get_ref_point_mask`` (:18-55), ``infer_ref_point`` (:133-159) and ``get_adjusted_refpoint``
  (strategies/predictives/qehvi.py:87-110) are plain host arithmetic on the objective specs.
* ``is_non_dominated`` / ``get_pareto_front`` (:58-84) and ``compute_hypervolume`` (:87-130) run on the device
  (``bo_pareto_mask``, ``bo_hypervolume``): the same front and box-decomposition kernels the acquisition
  function uses.  Everything is in the maximisation frame the objective callables produce.
"""
import ctypes as C
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _lib as L
from .model import _dev_ptr, _stream
from .objectives import MultiObjective


def get_ref_point_mask(objective: MultiObjective) -> np.ndarray:
    """+1 for Maximize, -1 for Minimize / CloseToTarget (utils/multiobjective.py:18-55)."""
    if len(objective.ops) < 2:
        raise ValueError("At least two output features have to be provided.")
    mask = []
    for op in objective.ops:
        if op.kind == "max":
            mask.append(1.0)
        elif op.kind in ("min", "close_to_target"):
            mask.append(-1.0)
        else:
            raise ValueError("Only `MaximizeObjective` and `MinimizeObjective` supported")
    return np.array(mask)


def infer_ref_point(objective: MultiObjective, Y, return_masked: bool = False) -> np.ndarray:
    """Worst observed objective value per output (utils/multiobjective.py:133-159). Y [n, M] raw outputs."""
    obj = objective(torch.as_tensor(np.asarray(Y), dtype=torch.double), None).numpy()
    ref = obj.min(axis=0)
    if not return_masked:
        ref = ref / get_ref_point_mask(objective)
    return ref


def get_adjusted_refpoint(objective: MultiObjective, Y, ref_point: Optional[Sequence[float]] = None) -> List[float]:
    """mask * ref_point, inferring the point from the data when it is not given (qehvi.py:87-110)."""
    if ref_point is None:
        ref_point = infer_ref_point(objective, Y, return_masked=False)
    return (get_ref_point_mask(objective) * np.asarray(ref_point, dtype=np.float64)).tolist()


def _device():
    if not torch.cuda.is_available():
        raise L.EverestError("everest_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def is_non_dominated(Y, deduplicate: bool = True) -> torch.Tensor:
    """[UPSTREAM] botorch.utils.multi_objective.is_non_dominated (maximisation), evaluated on the device."""
    dev = _device()
    Yd = torch.as_tensor(Y, dtype=torch.double).to(dev).contiguous()
    if Yd.dim() != 2:
        raise ValueError("Y must be [n, m]")
    n, m = Yd.shape
    mask = torch.zeros(n, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        L.check(L.load().bo_pareto_mask(_dev_ptr(Yd), n, m, int(deduplicate), _dev_ptr(mask), _stream()))
    return mask.bool().cpu()


def get_pareto_front(objective: MultiObjective, Y) -> np.ndarray:
    """Indices of the Pareto-optimal rows of Y [n, M] (raw outputs), like get_pareto_front's df.loc[pareto_mask]."""
    obj = objective(torch.as_tensor(np.asarray(Y), dtype=torch.double), None)
    return torch.nonzero(is_non_dominated(obj)).view(-1).numpy()


def compute_hypervolume(objective: MultiObjective, Y_optimal, ref_point: Sequence[float]) -> float:
    """Hypervolume(ref_point * mask).compute(objective(Y)) (utils/multiobjective.py:87-130); ref_point is given
    in the ORIGINAL output frame, like the reference's `ref_point` dict."""
    dev = _device()
    obj = objective(torch.as_tensor(np.asarray(Y_optimal), dtype=torch.double), None).to(dev).contiguous()
    ref = (get_ref_point_mask(objective) * np.asarray(ref_point, dtype=np.float64)).astype(np.float64)
    n, m = obj.shape
    ref_c = (C.c_double * m)(*ref.tolist())
    hv = C.c_double(0.0)
    with torch.cuda.device(dev):
        L.check(L.load().bo_hypervolume(_dev_ptr(obj), n, m, ref_c, C.byref(hv), _stream()))
    return float(hv.value)

"""Bridges from BoFire / BoTorch objects to the device specs.

* ``map_kernel``: BoFire kernel DATA MODEL + fitted hyper-parameters -> spec tree.  Mirrors
  ``bofire.kernels.mapper.map`` (kernels/mapper.py:289-302): same arguments (``active_dims``,
  ``features_to_idx_mapper``), same active-dims rules (:17-28), one-hot layout of the Hamming kernel (:206-253).
  Runs wherever ``bofire.data_models`` imports (it does in the build image; tests/test_host_logic.py).
* ``objectives_from_outputs``: ``Outputs`` -> op tables, mirrors ``_callables_and_weights`` /
  ``get_multiobjective_objective`` / ``get_output_constraints`` (utils/torch_tools.py:340-381, 598-727).
* ``state_from_botorch_model``: fitted ModelListGP / SingleTaskGP -> DeviceGPState.  Needs botorch + gpytorch,
  which are NOT installable in the build image, so this one function is untested here.
"""
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np

from . import kernels as K
from . import objectives as Ob


# ------------------------------------------------------------------------------------------------
# kernels
# ------------------------------------------------------------------------------------------------
def _active(data_model, active_dims, features_to_idx_mapper):
    if getattr(data_model, "features", None):
        if features_to_idx_mapper is None:
            raise RuntimeError("features_to_idx_mapper must be defined when using only a subset of features")
        return list(features_to_idx_mapper(data_model.features))
    return list(active_dims)


def map_kernel(data_model, active_dims: List[int], features_to_idx_mapper: Optional[Callable] = None,
               hyper: Optional[Dict] = None, _path: str = ""):
    """`hyper` holds the fitted values keyed by the path of the node in the tree ("" for the root,
    "base_kernel", "kernels.0", ...): {"lengthscale": [...]} for leaves, {"outputscale": s} for ScaleKernel.
    Missing entries default to lengthscale 1 / outputscale 1 (gpytorch's raw initialisation differs; a fitted
    model always provides them)."""
    hyper = hyper or {}
    h = hyper.get(_path, {})
    t = type(data_model).__name__

    def sub(child, name):
        return map_kernel(child, active_dims, features_to_idx_mapper, hyper, f"{_path}.{name}".lstrip("."))

    if t == "RBFKernel" or t == "MaternKernel":
        dims = _active(data_model, active_dims, features_to_idx_mapper)
        n_ls = len(dims) if data_model.ard else 1
        ls = list(np.atleast_1d(h.get("lengthscale", [1.0] * n_ls)).astype(float))
        if len(ls) != n_ls:
            raise ValueError(f"{t}: expected {n_ls} lengthscales, got {len(ls)}")
        return K.RBFKernel(dims, ls) if t == "RBFKernel" else K.MaternKernel(dims, ls, nu=data_model.nu)
    if t == "TanimotoKernel":
        return K.TanimotoKernel(_active(data_model, active_dims, features_to_idx_mapper))
    if t == "HammingDistanceKernel":
        if getattr(data_model, "features", None) is None:
            # CategoricalKernel on integer codes (OneHotToNumeric upstream): each active dim is one feature,
            # expressed as a one-hot group of unknown width -> needs the one-hot layout; refuse loudly.
            raise NotImplementedError("HammingDistanceKernel without `features` (integer-coded CategoricalKernel) "
                                      "needs the one-hot layout: pass categorical_features via features")
        if features_to_idx_mapper is None:
            raise RuntimeError("features_to_idx_mapper must be defined when using only a subset of features")
        cats, used = {}, []
        for k in data_model.features:
            idx = list(features_to_idx_mapper([k]))
            if any(i in used for i in idx):
                raise RuntimeError(f"indices {[i for i in idx if i in used]} are used in more than one categorical feature")
            if len(idx) == 1:
                raise RuntimeError(f"feature {k} is supposed to be one-hot encoded but is mapped to a single dimension")
            if idx != list(range(idx[0], idx[0] + len(idx))):
                raise NotImplementedError("one-hot columns of a categorical feature must be contiguous")
            cats[idx[0]] = len(idx)
            used += idx
        n_ls = len(used) if data_model.ard else 1
        ls = list(np.atleast_1d(h.get("lengthscale", [1.0] * n_ls)).astype(float))
        return K.HammingDistanceKernel(cats, ls)
    if t == "ScaleKernel":
        return K.ScaleKernel(sub(data_model.base_kernel, "base_kernel"), float(h.get("outputscale", 1.0)))
    if t == "AdditiveKernel":
        return K.AdditiveKernel([sub(c, f"kernels.{i}") for i, c in enumerate(data_model.kernels)])
    if t == "MultiplicativeKernel":
        return K.MultiplicativeKernel([sub(c, f"kernels.{i}") for i, c in enumerate(data_model.kernels)])
    raise NotImplementedError(f"kernel {t} is outside the accelerated path")


# ------------------------------------------------------------------------------------------------
# objectives / constraints
# ------------------------------------------------------------------------------------------------
def objective_spec(idx: int, objective, x_adapt=None) -> Ob.ObjectiveSpec:
    """get_objective_callable (utils/torch_tools.py:384-450) for the accelerated closed set."""
    t = type(objective).__name__
    w = float(getattr(objective, "w", 1.0))
    if t == "MaximizeObjective":
        return Ob.MaximizeObjective(idx, objective.lower_bound, objective.upper_bound, w)
    if t == "MinimizeObjective":
        return Ob.MinimizeObjective(idx, objective.lower_bound, objective.upper_bound, w)
    if t == "CloseToTargetObjective":
        return Ob.CloseToTargetObjective(idx, objective.target_value, objective.exponent, w)
    if t == "MinimizeSigmoidObjective":
        return Ob.MinimizeSigmoidObjective(idx, objective.steepness, objective.tp, w)
    if t == "MaximizeSigmoidObjective":
        return Ob.MaximizeSigmoidObjective(idx, objective.steepness, objective.tp, w)
    if t == "MovingMaximizeSigmoidObjective":
        if x_adapt is None:
            raise ValueError("x_adapt is needed for MovingMaximizeSigmoidObjective")
        return Ob.MaximizeSigmoidObjective(idx, objective.steepness, float(np.max(x_adapt)) + objective.tp, w)
    if t == "TargetObjective":
        return Ob.TargetObjective(idx, objective.target_value, objective.tolerance, objective.steepness, w)
    raise NotImplementedError(f"Objective {t} not implemented.")


def objectives_from_outputs(outputs, experiments=None):
    """(MultiObjective over Max/Min/CloseToTarget outputs, output constraints) like
    get_multiobjective_objective + get_output_constraints."""
    feats = outputs.get()
    ops, cons = [], []
    for i, feat in enumerate(feats):
        if feat.objective is None:
            continue
        t = type(feat.objective).__name__
        x_adapt = None if experiments is None else experiments[feat.key].values
        if t in ("MaximizeObjective", "MinimizeObjective", "CloseToTargetObjective"):
            ops.append(objective_spec(i, feat.objective, x_adapt))
        elif t in ("MaximizeSigmoidObjective", "MinimizeSigmoidObjective", "TargetObjective", "MovingMaximizeSigmoidObjective"):
            cons += Ob.constraints_from_sigmoid_objectives([objective_spec(i, feat.objective, x_adapt)])
    return Ob.MultiObjective(ops), cons


# ------------------------------------------------------------------------------------------------
# fitted BoTorch model -> device state   (needs botorch / gpytorch; untested in the build image)
# ------------------------------------------------------------------------------------------------
def _kernel_from_gpytorch(k, d, col_map):  # pragma: no cover
    name = type(k).__name__
    dims = [col_map[int(i)] for i in (k.active_dims.tolist() if k.active_dims is not None else range(d))]
    if name == "RBFKernel":
        return K.RBFKernel(dims, k.lengthscale.detach().cpu().double().view(-1).tolist())
    if name == "MaternKernel":
        return K.MaternKernel(dims, k.lengthscale.detach().cpu().double().view(-1).tolist(), nu=float(k.nu))
    if name == "TanimotoKernel":
        return K.TanimotoKernel(dims)
    if name == "HammingKernelWithOneHots":
        cf = {dims[s]: c for s, c in k.trx.categorical_features.items()}
        return K.HammingDistanceKernel(cf, k.lengthscale.detach().cpu().double().view(-1).tolist())
    if name == "ScaleKernel":
        return K.ScaleKernel(_kernel_from_gpytorch(k.base_kernel, d, col_map), float(k.outputscale.detach().cpu()))
    if name == "AdditiveKernel":
        return K.AdditiveKernel([_kernel_from_gpytorch(c, d, col_map) for c in k.kernels])
    if name == "ProductKernel":
        return K.MultiplicativeKernel([_kernel_from_gpytorch(c, d, col_map) for c in k.kernels])
    raise NotImplementedError(f"gpytorch kernel {name} is outside the accelerated path")


def state_from_botorch_model(model, device=None):  # pragma: no cover
    from .model import DeviceGPState, SingleTaskGPSpec

    models = list(model.models) if hasattr(model, "models") else [model]
    X_full, specs = None, []
    for sub in models:
        tf = getattr(sub, "input_transform", None)
        chain = list(tf.values()) if tf is not None and hasattr(tf, "values") else ([tf] if tf is not None else [])
        X_raw = sub.train_inputs[0].detach().cpu().double()
        col_map = list(range(X_raw.shape[-1]))
        off, scl = None, None
        for t in chain:
            tn = type(t).__name__
            if tn == "FilterFeatures":
                col_map = [col_map[int(i)] for i in t.feature_indices.tolist()]
            elif tn == "Normalize":
                idx = t.indices.tolist() if getattr(t, "indices", None) is not None else list(range(len(col_map)))
                off = {col_map[i]: float(t.offset.view(-1)[j if t.offset.numel() > 1 else 0]) for j, i in enumerate(idx)}
                scl = {col_map[i]: float(t.coefficient.view(-1)[j if t.coefficient.numel() > 1 else 0]) for j, i in enumerate(idx)}
            elif tn in ("OneHotToNumeric",):
                raise NotImplementedError("OneHotToNumeric + CategoricalKernel: rebuild with HammingKernelWithOneHots")
            else:
                raise NotImplementedError(f"input transform {tn} is outside the accelerated path")
        # BoTorch stores the TRANSFORMED inputs on the model in eval mode; the untransformed ones are what BoFire passes
        X_un = tf.untransform(X_raw) if (tf is not None and hasattr(tf, "untransform") and not sub.training) else X_raw
        if X_full is None:
            X_full = X_un.numpy()
        d = X_full.shape[1]
        octf = sub.outcome_transform
        y_mean, y_std = float(octf.means.view(-1)[0]), float(octf.stdvs.view(-1)[0])
        y = sub.train_targets.detach().cpu().double().view(-1).numpy() * y_std + y_mean
        in_off, in_scl = np.zeros(d), np.ones(d)
        if off:
            for c, v in off.items():
                in_off[c] = v
            for c, v in scl.items():
                in_scl[c] = v
        specs.append(SingleTaskGPSpec(kernel=_kernel_from_gpytorch(sub.covar_module, d, col_map), y=y, in_offset=in_off,
                                      in_scale=in_scl, mean_const=float(sub.mean_module.constant.detach().cpu()),
                                      noise=float(sub.likelihood.noise.detach().cpu().view(-1)[0]), y_mean=y_mean, y_std=y_std))
    return DeviceGPState(X_full, specs, device=device).factorize()

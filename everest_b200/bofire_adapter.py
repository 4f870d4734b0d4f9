"""Bridges from BoFire / BoTorch objects to the device specs.

* ``map_kernel``: BoFire kernel DATA MODEL + fitted hyper-parameters -> spec tree.  Mirrors
  ``bofire.kernels.mapper.map`` (kernels/mapper.py:289-302): same arguments (``active_dims``,
  ``features_to_idx_mapper``), same active-dims rules (:17-28), one-hot layout of the Hamming kernel (:206-253).
  Runs wherever ``bofire.data_models`` imports (it does in the build image; tests/test_host_logic.py).
* ``objectives_from_outputs``: ``Outputs`` -> op tables, mirrors ``_callables_and_weights`` /
  ``get_multiobjective_objective`` / ``get_output_constraints`` (utils/torch_tools.py:340-381, 598-727).
* ``specs_from_botorch_model`` / ``state_from_botorch_model``: fitted ModelListGP / SingleTaskGP / MixedSingleTaskGP ->
  specs -> DeviceGPState.  botorch / gpytorch are NOT installable in the build image: the attribute walk is tested on CPU with
  duck-typed stand-ins (tests/test_host_logic.py) and against the real classes in tests/test_botorch_parity.py where BoTorch is
  installed.
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
def _kernel_from_gpytorch(k, d, col_map):
    """gpytorch / BoTorch / BoFire kernel module -> spec tree.  `col_map[j]` is what dimension j of the space the kernel sees
    is in the BoFire-transformed layout: a column index, or ("cat", start, cardinality) for an integer-coded categorical that
    an upstream OneHotToNumeric produced from a one-hot block."""
    name = type(k).__name__
    dims = [col_map[int(i)] for i in (k.active_dims.tolist() if k.active_dims is not None else range(d))]

    def ls():
        return k.lengthscale.detach().cpu().double().view(-1).tolist()

    def columns():
        if any(isinstance(c, tuple) for c in dims):
            raise NotImplementedError(f"{name} over integer-coded categorical dimensions is outside the accelerated path")
        return dims

    if name == "RBFKernel":
        return K.RBFKernel(columns(), ls())
    if name == "MaternKernel":
        return K.MaternKernel(columns(), ls(), nu=float(k.nu))
    if name == "TanimotoKernel":
        return K.TanimotoKernel(columns())
    if name == "HammingKernelWithOneHots":
        cols = columns()
        cf = {cols[s]: c for s, c in k.trx.categorical_features.items()}
        return K.HammingDistanceKernel(cf, ls())
    if name == "CategoricalKernel":
        # [UPSTREAM] botorch CategoricalKernel on integer codes: exp(-mean_f(delta_f / l_f)) -- the one-hot Hamming kernel
        # (equality pinned by tests/bofire/kernels/test_categorical.py), one ARD lengthscale per categorical FEATURE
        if not all(isinstance(c, tuple) for c in dims):
            raise NotImplementedError("CategoricalKernel over columns that no OneHotToNumeric produced")
        vals = ls()
        if len(vals) == 1:
            vals = vals * len(dims)
        order = sorted(range(len(dims)), key=lambda i: dims[i][1])
        return K.HammingDistanceKernel({dims[i][1]: dims[i][2] for i in order}, [vals[i] for i in order])
    if name == "ScaleKernel":
        return K.ScaleKernel(_kernel_from_gpytorch(k.base_kernel, d, col_map), float(k.outputscale.detach().cpu()))
    if name == "AdditiveKernel":
        return K.AdditiveKernel([_kernel_from_gpytorch(c, d, col_map) for c in k.kernels])
    if name == "ProductKernel":
        return K.MultiplicativeKernel([_kernel_from_gpytorch(c, d, col_map) for c in k.kernels])
    raise NotImplementedError(f"gpytorch kernel {name} is outside the accelerated path")


def specs_from_botorch_model(model, X_train=None):
    """Fitted ModelListGP / SingleTaskGP / MixedSingleTaskGP as `BotorchSurrogates.compatibilize` returns it
    (surrogates/botorch_surrogates.py:79-128) -> (X_train in the BoFire-transformed layout, [SingleTaskGPSpec]).
    Input transforms understood: FilterFeatures (compatibilize), Normalize / InputStandardize (get_scaler,
    surrogates/utils.py:103-164), OneHotToNumeric (mixed_single_task_gp.py:82-88: the literal MixedSingleTaskGPSurrogate;
    its CategoricalKernel becomes the one-hot Hamming leaf on the original one-hot block).  `X_train` [N, d]: the transformed
    experiments BoFire fitted on (`inputs.transform(experiments, specs)`); needed when a FilterFeatures hides columns from a
    sub-model, otherwise recovered from the model.  Pure host code that only reads attributes: exercised on CPU with
    duck-typed stand-ins (tests/test_host_logic.py) and against real BoTorch where installed (tests/test_botorch_parity.py)."""
    from .model import SingleTaskGPSpec

    models = list(model.models) if hasattr(model, "models") else [model]
    X_full = None if X_train is None else np.ascontiguousarray(np.asarray(
        X_train.detach().cpu().numpy() if hasattr(X_train, "detach") else X_train, dtype=np.float64))
    specs = []
    for sub in models:
        tf = getattr(sub, "input_transform", None)
        chain = list(tf.values()) if tf is not None and hasattr(tf, "values") else ([tf] if tf is not None else [])
        names = [type(t).__name__ for t in chain]
        for tn in names:
            if tn not in ("FilterFeatures", "Normalize", "InputStandardize", "OneHotToNumeric"):
                raise NotImplementedError(f"input transform {tn} is outside the accelerated path")
        X_raw = sub.train_inputs[0].detach().cpu().double()
        X_raw = X_raw.reshape(-1, X_raw.shape[-1])
        o2n = next((t for t in chain if type(t).__name__ == "OneHotToNumeric"), None)
        if X_full is None:
            if "FilterFeatures" in names:
                raise ValueError("a sub-model sees a subset of the columns (FilterFeatures): pass X_train, the transformed "
                                 "experiments of the whole domain")
            if sub.training:     # train_inputs are as passed to the constructor: unscaled, but already integer-coded
                X_un = o2n.untransform(X_raw) if o2n is not None else X_raw
            else:                # eval mode: BoTorch keeps the transformed inputs on the model
                X_un = tf.untransform(X_raw) if tf is not None else X_raw
            X_full = np.ascontiguousarray(X_un.detach().cpu().double().numpy())
        d = X_full.shape[1]
        # what dimension j of the CURRENT space is in the original layout: a column, or ("cat", start column, cardinality)
        col_map = list(range(d))
        in_off, in_scl = np.zeros(d), np.ones(d)
        for t, tn in zip(chain, names):
            if tn == "FilterFeatures":
                col_map = [col_map[int(i)] for i in t.feature_indices.tolist()]
            elif tn in ("Normalize", "InputStandardize"):
                idx = [int(i) for i in t.indices.tolist()] if getattr(t, "indices", None) is not None else list(range(len(col_map)))
                offs = (t.offset if tn == "Normalize" else t.means).detach().cpu().double().reshape(-1)
                cofs = (t.coefficient if tn == "Normalize" else t.stds).detach().cpu().double().reshape(-1)

                def pick(v, j, i):   # statistics are kept over all columns of the space, over `indices` only, or as a scalar
                    return float(v[i] if v.numel() == len(col_map) else (v[j] if v.numel() == len(idx) else v[0]))

                for j, i in enumerate(idx):
                    c = col_map[i]
                    if isinstance(c, tuple):
                        raise NotImplementedError("a scaler on integer-coded categorical dimensions is outside the accelerated path")
                    in_off[c], in_scl[c] = pick(offs, j, i), pick(cofs, j, i)
            else:   # OneHotToNumeric: numeric columns first, one integer code per one-hot block after them
                cats = {int(s): int(c) for s, c in t.categorical_features.items()}
                in_group = {s + k for s, c in cats.items() for k in range(c)}
                for s in cats:
                    if isinstance(col_map[s], tuple) or col_map[s:s + cats[s]] != list(range(col_map[s], col_map[s] + cats[s])):
                        raise NotImplementedError("one-hot blocks must be contiguous columns")
                col_map = [col_map[i] for i in range(len(col_map)) if i not in in_group] + \
                          [("cat", col_map[s], c) for s, c in sorted(cats.items())]
        octf = getattr(sub, "outcome_transform", None)
        if octf is not None:
            y_mean, y_std = float(octf.means.reshape(-1)[0]), float(octf.stdvs.reshape(-1)[0])
        else:
            y_mean, y_std = 0.0, 1.0
        y = sub.train_targets.detach().cpu().double().reshape(-1).numpy() * y_std + y_mean
        specs.append(SingleTaskGPSpec(kernel=_kernel_from_gpytorch(sub.covar_module, len(col_map), col_map), y=y, in_offset=in_off,
                                      in_scale=in_scl, mean_const=float(sub.mean_module.constant.detach().cpu()),
                                      noise=float(sub.likelihood.noise.detach().cpu().reshape(-1)[0]), y_mean=y_mean, y_std=y_std))
    return X_full, specs


def state_from_botorch_model(model, device=None, X_train=None):
    """specs_from_botorch_model + factorisation on the device."""
    from .model import DeviceGPState

    X_full, specs = specs_from_botorch_model(model, X_train=X_train)
    return DeviceGPState(X_full, specs, device=device).factorize()

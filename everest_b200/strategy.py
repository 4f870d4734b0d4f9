"""Host-side mirror of the predictive strategies' tell / ask / predict orchestration:
`BotorchStrategy` (strategies/predictives/botorch.py:174-225 `_predict` / `calc_acquisition`, :226-296 `_setup_ask`,
:408-521 `_ask`, :696-724 `get_acqf_input_tensors`) and the `_get_acqfs` bodies of QnehviStrategy (qnehvi.py:23-53),
QehviStrategy (qehvi.py:37-77), MoboStrategy (mobo.py:43-91) and SoboStrategy (sobo.py:51-90).

It works on the TRANSFORMED input layout BoFire hands to BoTorch (`Inputs.transform`, data_models/domain/features.py:
423-533): continuous columns with bounds, one-hot blocks for categoricals, 0/1 fingerprint columns.  The pandas / pydantic
layer around it (Domain, data models) is out of scope; a BoFire maintainer keeps those and swaps the
strategy bodies as shown in INTEGRATION.md.  Hyper-parameters are handed in (the fit is SURVEY.md 8f-2): `surrogate_factory
(X, Y) -> [SingleTaskGPSpec]` plays the role of `_fit`.
"""
import itertools
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import acquisition as A
from . import multiobjective as MO
from . import optim
from .model import DeviceGPState, SingleTaskGPSpec
from .objectives import MultiObjective, OutputConstraint, ScalarObjective


@dataclass
class InputSpace:
    """Transformed input space: `bounds` [2, d]; `categorical_groups` {start column: cardinality} of the one-hot blocks;
    `fixed_features` {column: value}; linear constraints in BoTorch's (indices, coefficients, rhs) form as produced by
    get_linear_constraints / get_interpoint_constraints (utils/torch_tools.py:45-144); `allowed_categories` optionally
    restricts the levels of a group (CategoricalInput.allowed)."""
    bounds: np.ndarray
    categorical_groups: Dict[int, int] = field(default_factory=dict)
    fixed_features: Dict[int, float] = field(default_factory=dict)
    inequality_constraints: list = field(default_factory=list)
    equality_constraints: list = field(default_factory=list)
    allowed_categories: Dict[int, Sequence[int]] = field(default_factory=dict)
    discrete_values: Dict[int, Sequence[float]] = field(default_factory=dict)  # DiscreteInput columns
    # get_nonlinear_constraints (utils/torch_tools.py:239-252): [(callable, is_intrapoint)], feasible iff callable(x) >= 0;
    # with them BoFire passes a generator of feasible raw samples (n, q, seed) -> [n, q, d] (botorch.py:257-265)
    nonlinear_constraints: list = field(default_factory=list)
    initial_conditions_generator: Optional[Callable] = None

    def __post_init__(self):
        self.bounds = np.asarray(self.bounds, dtype=np.float64)
        if self.bounds.ndim != 2 or self.bounds.shape[0] != 2:
            raise ValueError("bounds must be [2, d]")

    @property
    def d(self) -> int:
        return self.bounds.shape[1]

    def combinatorial_columns(self) -> List[int]:
        cols = list(self.discrete_values.keys())
        for s, c in self.categorical_groups.items():
            cols += list(range(s, s + c))
        return sorted(cols)

    def is_fully_combinatorial(self) -> bool:
        """botorch.py:424-427: every input is Discrete or Categorical."""
        free = [j for j in range(self.d) if j not in self.fixed_features]
        return len(free) > 0 and set(free) <= set(self.combinatorial_columns())

    def categorical_combinations(self) -> List[Dict[int, float]]:
        """get_categorical_combinations (botorch.py:523-640) for one-hot encoded categoricals and discrete inputs:
        one fixed-feature dictionary per combination of levels."""
        axes = []
        for s, c in sorted(self.categorical_groups.items()):
            levels = list(self.allowed_categories.get(s, range(c)))
            axes.append([{s + k: (1.0 if k == lv else 0.0) for k in range(c)} for lv in levels])
        for col, vals in sorted(self.discrete_values.items()):
            axes.append([{col: float(v)} for v in vals])
        combos = []
        for parts in itertools.product(*axes):
            ff = dict(self.fixed_features)
            for part in parts:
                ff.update(part)
            combos.append(ff)
        return combos


def drop_duplicate_rows(X: np.ndarray, Y: Optional[np.ndarray] = None):
    """get_acqf_input_tensors (botorch.py:696-724): experiments.drop_duplicates(subset=inputs, keep="first")."""
    _, first = np.unique(X, axis=0, return_index=True)
    keep = np.sort(first)
    return (X[keep], None if Y is None else Y[keep])


class DeviceBotorchStrategy:
    """tell / ask / predict / calc_acquisition with the reference's semantics; subclasses provide `_get_acqfs`."""

    def __init__(self, space: InputSpace, surrogate_factory: Callable[[np.ndarray, np.ndarray], List[SingleTaskGPSpec]],
                 num_restarts: int = 8, num_raw_samples: int = 1024, maxiter: int = 2000, seed: Optional[int] = None,
                 device=None):
        self.space = space
        self.surrogate_factory = surrogate_factory
        self.num_restarts, self.num_raw_samples, self.maxiter = int(num_restarts), int(num_raw_samples), int(maxiter)
        self.seed = int(np.random.default_rng().integers(1, 1000000)) if seed is None else int(seed)
        torch.manual_seed(self.seed)            # botorch.py:86
        self.device = device
        self.X: Optional[np.ndarray] = None
        self.Y: Optional[np.ndarray] = None
        self.candidates: Optional[np.ndarray] = None   # pending candidates (Strategy.add_candidates)
        self.model: Optional[DeviceGPState] = None

    # -- tell ----------------------------------------------------------------------------------------
    def tell(self, X, Y, replace: bool = False):
        """Strategy.tell (strategies/strategy.py:79-109) + PredictiveStrategy fit: store the experiments and rebuild
        the device state from the surrogate specs."""
        X = np.atleast_2d(np.asarray(X, dtype=np.float64))
        Y = np.asarray(Y, dtype=np.float64).reshape(X.shape[0], -1)
        if X.shape[1] != self.space.d:
            raise ValueError(f"experiments must have {self.space.d} input columns")
        if replace or self.X is None:
            self.X, self.Y = X, Y
        else:
            self.X, self.Y = np.concatenate([self.X, X]), np.concatenate([self.Y, Y])
        self.candidates = None
        self._fit()

    def _fit(self):
        if self.model is not None:
            self.model.close()
        self.model = DeviceGPState(self.X, self.surrogate_factory(self.X, self.Y), device=self.device).factorize()

    @property
    def is_fitted(self) -> bool:
        return self.model is not None

    def add_candidates(self, X):
        X = np.atleast_2d(np.asarray(X, dtype=np.float64))
        self.candidates = X if self.candidates is None else np.concatenate([self.candidates, X])

    def get_acqf_input_tensors(self) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
        if self.X is None:
            raise ValueError("No experiments have been provided yet.")
        Xc, _ = drop_duplicate_rows(self.X)
        Xp = None if self.candidates is None else torch.as_tensor(self.candidates, dtype=torch.double)
        return torch.as_tensor(Xc, dtype=torch.double), Xp

    # -- predict / calc_acquisition --------------------------------------------------------------------
    def predict(self, X):
        """(preds [n, M], stds [n, M]) with observation noise (botorch.py:174-194)."""
        assert self.is_fitted, "Model not trained."
        return self.model.predict(np.atleast_2d(np.asarray(X, dtype=np.float64)))

    def calc_acquisition(self, candidates, combined: bool = False) -> np.ndarray:
        acqf = self._get_acqfs(np.atleast_2d(candidates).shape[0] if combined else 1)[0]
        return optim.calc_acquisition(acqf, np.atleast_2d(np.asarray(candidates, dtype=np.float64)), combined=combined)

    # -- ask ---------------------------------------------------------------------------------------------
    def _get_acqfs(self, n: int) -> list:
        raise NotImplementedError

    def _get_optimizer_options(self) -> dict:
        return {"maxiter": self.maxiter}

    def ask(self, candidate_count: int = 1):
        """BotorchStrategy._ask: (candidates [q, d], predictions [q, M], stds [q, M]) as numpy."""
        assert candidate_count > 0, "candidate_count has to be larger than zero."
        if self.X is None:
            raise ValueError("No experiments have been provided yet.")
        acqfs = self._get_acqfs(candidate_count)
        sp = self.space
        bounds = torch.as_tensor(sp.bounds)
        if sp.is_fully_combinatorial():
            # botorch.py:425-467: enumerate the choices, drop those already measured, arg-max over the rest
            d = sp.d
            choices = np.zeros((0, d))
            combos = sp.categorical_combinations()
            if combos:
                choices = np.zeros((len(combos), d))
                for i, ff in enumerate(combos):
                    for j, v in ff.items():
                        choices[i, j] = v
            seen = {tuple(r) for r in np.round(self.X, 12)}
            keep = np.array([tuple(r) not in seen for r in np.round(choices, 12)], dtype=bool)
            if not keep.any():
                raise ValueError("every combination of the combinatorial search space has been measured already")
            cand, _ = optim.optimize_acqf_discrete(acqfs[0], q=candidate_count, choices=torch.as_tensor(choices[keep]))
        else:
            combos = sp.categorical_combinations() if (sp.categorical_groups or sp.discrete_values) else []
            kw = dict(options=self._get_optimizer_options(), seed=int(torch.randint(0, 1000000, (1,)).item()),
                      inequality_constraints=sp.inequality_constraints or None,
                      equality_constraints=sp.equality_constraints or None)
            if sp.nonlinear_constraints:
                kw.update(nonlinear_inequality_constraints=list(sp.nonlinear_constraints),
                          generator=sp.initial_conditions_generator)
            if len(acqfs) > 1:
                # botorch.py:337-356: one acquisition function per candidate -> sequential optimize_acqf_list
                if len(acqfs) != candidate_count:
                    raise ValueError("one acquisition function per candidate is expected")
                ff = (combos[0] if len(combos) == 1 else None) or (dict(sp.fixed_features) or None)
                cand, _ = optim.optimize_acqf_list(acqfs, bounds, self.num_restarts, self.num_raw_samples,
                                                   fixed_features=None if len(combos) > 1 else ff,
                                                   fixed_features_list=combos if len(combos) > 1 else None, **kw)
            elif len(combos) > 1:
                cand, _ = optim.optimize_acqf_mixed(acqfs[0], bounds, candidate_count, self.num_restarts,
                                                    self.num_raw_samples, fixed_features_list=combos, **kw)
            else:
                ff = combos[0] if combos else (dict(sp.fixed_features) or None)
                cand, _ = optim.optimize_acqf(acqfs[0], bounds, candidate_count, self.num_restarts, self.num_raw_samples,
                                              fixed_features=ff, **kw)
        cand = cand.detach().numpy()      # botorch.py:314: candidates must be CPU tensors
        preds, stds = self.predict(cand)
        return cand, preds, stds


class _MultiObjectiveStrategy(DeviceBotorchStrategy):
    def __init__(self, space, surrogate_factory, objective: MultiObjective, ref_point: Optional[Sequence[float]] = None,
                 constraints: Optional[List[OutputConstraint]] = None, n_mc_samples: int = 512, **kw):
        super().__init__(space, surrogate_factory, **kw)
        self.objective, self.ref_point, self.constraints = objective, ref_point, constraints
        self.n_mc_samples = int(n_mc_samples)

    def get_adjusted_refpoint(self) -> List[float]:
        """qehvi.py:87-110 / mobo.py:93-116: masked reference point, inferred from the data when not given."""
        return MO.get_adjusted_refpoint(self.objective, self.Y, self.ref_point)


class MoboStrategy(_MultiObjectiveStrategy):
    """mobo.py:43-91; acquisition_function in {"qLogNEHVI" (default), "qNEHVI", "qEHVI", "qLogEHVI"}."""

    def __init__(self, *args, acquisition_function: str = "qLogNEHVI", alpha: float = 0.0, prune_baseline: bool = True, **kw):
        super().__init__(*args, **kw)
        if acquisition_function not in ("qLogNEHVI", "qNEHVI", "qEHVI", "qLogEHVI"):
            raise ValueError(f"{acquisition_function} is not a multi-objective acquisition function")
        self.acquisition_function, self.alpha, self.prune_baseline = acquisition_function, float(alpha), bool(prune_baseline)

    def _get_acqfs(self, n: int) -> list:
        assert self.is_fitted, "Model not trained."
        X_train, X_pending = self.get_acqf_input_tensors()
        cons = self.constraints if self.constraints else None
        name = self.acquisition_function
        return [A.get_acquisition_function(
            name, self.model, self.objective, X_train, X_pending=X_pending, constraints=cons,
            ref_point=self.get_adjusted_refpoint(), mc_samples=self.n_mc_samples, alpha=self.alpha,
            prune_baseline=self.prune_baseline if name in ("qLogNEHVI", "qNEHVI") else True,
            Y=self.Y if name in ("qEHVI", "qLogEHVI") else None)]


class QnehviStrategy(MoboStrategy):
    """qnehvi.py:23-53 (fixed to qNEHVI, prune_baseline=True, cache_root=True)."""

    def __init__(self, *args, **kw):
        kw.pop("acquisition_function", None)
        super().__init__(*args, acquisition_function="qNEHVI", **kw)


class QehviStrategy(MoboStrategy):
    """qehvi.py:37-77: fixed partitioning of the observed objective values that beat the reference point."""

    def __init__(self, *args, **kw):
        kw.pop("acquisition_function", None)
        super().__init__(*args, acquisition_function="qEHVI", **kw)


class SoboStrategy(DeviceBotorchStrategy):
    """sobo.py:51-90; acquisition_function in {"qLogNEI" (default), "qNEI", "qLogEI", "qEI", "qSR", "qUCB", "qPI"};
    `objective` is the scalarised objective (single / additive / multiplicative, sobo.py:92-152)."""

    def __init__(self, space, surrogate_factory, objective: ScalarObjective, acquisition_function: str = "qLogNEI",
                 constraints: Optional[List[OutputConstraint]] = None, n_mc_samples: int = 512, beta: float = 0.2,
                 tau: float = 1e-3, prune_baseline: bool = True, **kw):
        super().__init__(space, surrogate_factory, **kw)
        if acquisition_function not in ("qLogNEI", "qNEI", "qLogEI", "qEI", "qSR", "qUCB", "qPI"):
            raise ValueError(f"{acquisition_function} is not a single-objective acquisition function")
        self.objective, self.acquisition_function, self.constraints = objective, acquisition_function, constraints
        self.n_mc_samples, self.beta, self.tau, self.prune_baseline = int(n_mc_samples), float(beta), float(tau), bool(prune_baseline)

    def _get_acqfs(self, n: int) -> list:
        assert self.is_fitted, "Model not trained."
        X_train, X_pending = self.get_acqf_input_tensors()
        name = self.acquisition_function
        return [A.get_acquisition_function(
            name, self.model, self.objective, X_train, X_pending=X_pending, constraints=self.constraints or None,
            mc_samples=self.n_mc_samples, beta=self.beta, tau=self.tau,
            prune_baseline=self.prune_baseline if name in ("qNEI", "qLogNEI") else True)]

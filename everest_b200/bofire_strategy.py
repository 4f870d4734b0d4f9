"""Drop-in behind BoFire's OWN API: strategy DATA MODELS + `Domain` in, pandas DataFrames in and out.

    import everest_b200.bofire_strategy as strategies            # instead of bofire.strategies.api
    strategy = strategies.map(QnehviStrategy(domain=domain))     # the unchanged BoFire data model
    strategy.tell(experiments)                                   # DataFrame
    candidates = strategy.ask(candidate_count=1)                 # DataFrame: inputs + <key>_pred / _sd / _des

Mirrors, for the predictive strategies on the accelerated path:

  Strategy / PredictiveStrategy      strategies/strategy.py:14-262, strategies/predictives/predictive.py:23-216
  BotorchStrategy                    strategies/predictives/botorch.py:57-750 (_fit, _predict, calc_acquisition, _setup_ask,
                                     _postprocess_candidates, _optimize_acqf_continuous, _ask, get_fixed_features,
                                     get_categorical_combinations, has_sufficient_experiments, get_acqf_input_tensors)
  QehviStrategy / QnehviStrategy     qehvi.py:23-110, qnehvi.py:14-53
  MoboStrategy                       mobo.py:31-124
  SoboStrategy and its Additive / Multiplicative variants   sobo.py:42-263
  RandomStrategy                     strategies/random.py:29-353 (initial designs and feasible raw samples)
  BotorchSurrogates.fit / compatibilize   surrogates/botorch_surrogates.py:19-128 with SingleTaskGPSurrogate._fit
                                     (single_task_gp.py:39-71), MixedSingleTaskGPSurrogate._fit (mixed_single_task_gp.py:46-112),
                                     TanimotoGPSurrogate / MixedTanimotoGPSurrogate (mixed_tanimoto_gp.py:101-215), get_scaler
                                     (surrogates/utils.py:103-164)
  get_linear_constraints / get_interpoint_constraints / get_nonlinear_constraints / get_output_constraints and the
  objective builders                 utils/torch_tools.py:45-252, 258-381, 598-727

Everything numerical runs on the device through the C ABI (DeviceGPState, the acquisition classes, optim); the pandas side
(`Inputs.transform` / `inverse_transform` / `get_bounds`, `Outputs.__call__`, validation) is BoFire's own data-model code,
which this module CALLS, never re-implements -- `bofire.data_models` must be importable (it is pure pydantic / pandas).
Outside the accelerated path and refused loudly: outlier detection, hyper-parameter-optimisation runs, LSR-BO local search,
MultiTask / non-GP surrogates, categorical outputs, desirability objectives.
"""
import copy
import math
import warnings
from typing import Callable, Dict, List, Optional, Tuple

import numpy as np
import pandas as pd
import torch

from . import acquisition as A
from . import fit as F
from . import kernels as K
from . import objectives as Ob
from . import optim
from .bofire_adapter import map_kernel, objective_spec
from .model import DeviceGPState, SingleTaskGPSpec


def _name(obj) -> str:
    return type(obj).__name__


def _is(obj, *names) -> bool:
    """isinstance by class name over the MRO: the data-model classes are not imported here, so that any BoFire version
    that names them alike works (and `bofire.data_models.strategies.api`, which needs `formulaic`, is never touched)."""
    return any(c.__name__ in names for c in type(obj).__mro__)


# ------------------------------------------------------------------------------------------------------------------
# utils/torch_tools.py mirrors that take a Domain
# ------------------------------------------------------------------------------------------------------------------
def get_linear_constraints(domain, constraint: str) -> List[Tuple[torch.Tensor, torch.Tensor, float]]:
    """utils/torch_tools.py:45-100 (unit_scaled=False): `constraint` is "LinearEqualityConstraint" or
    "LinearInequalityConstraint"; BoFire's `sum c x <= rhs` becomes BoTorch's `sum -c x >= -rhs`; fixed features move to the
    right-hand side; the column index is the position of the feature among the inputs (torch_tools.py:66)."""
    out = []
    keys = domain.inputs.get_keys()
    for c in domain.constraints.constraints:
        if _name(c) != constraint:
            continue
        indices, coefficients, rhs = [], [], 0.0
        for i, featkey in enumerate(c.features):
            feat = domain.inputs.get_by_key(featkey)
            if feat.is_fixed():
                rhs -= feat.fixed_value()[0] * c.coefficients[i]
            else:
                indices.append(keys.index(featkey))
                coefficients.append(c.coefficients[i])
        out.append((torch.tensor(indices, dtype=torch.long), -torch.tensor(coefficients, dtype=torch.double), -(rhs + c.rhs)))
    return out


def get_interpoint_constraints(domain, n_candidates: int):
    """utils/torch_tools.py:103-144: InterpointEqualityConstraint -> pairwise equalities across the q points."""
    out = []
    if n_candidates == 1:
        return out
    keys = domain.inputs.get_keys()
    for c in domain.constraints.constraints:
        if _name(c) != "InterpointEqualityConstraint":
            continue
        feat = domain.inputs.get_by_key(c.feature)
        if feat.is_fixed():
            continue
        feat_idx = keys.index(c.feature)
        multiplicity = c.multiplicity or n_candidates
        for i in range(math.ceil(n_candidates / multiplicity)):
            all_idx = list(range(i * multiplicity, min((i + 1) * multiplicity, n_candidates)))
            for k in range(len(all_idx) - 1):
                out.append((torch.tensor([[all_idx[0], feat_idx], [all_idx[k + 1], feat_idx]], dtype=torch.long),
                            torch.tensor([1.0, -1.0], dtype=torch.double), 0.0))
    return out


def _continuous_keys(domain) -> List[str]:
    return [f.key for f in domain.inputs.get() if _is(f, "ContinuousInput")]


def get_nonlinear_constraints(domain):
    """utils/torch_tools.py:147-252: NChooseK (narrow-Gaussian relaxation) and ProductInequality constraints as
    [(callable, is_intrapoint)], indices = position among the CONTINUOUS inputs (torch_tools.py:180, 228)."""
    ckeys = _continuous_keys(domain)
    out = []
    for c in domain.constraints.constraints:
        if _name(c) == "NChooseKConstraint":
            out += optim.nchoosek_constraints([ckeys.index(k) for k in c.features], max_count=c.max_count, min_count=c.min_count)
    for c in domain.constraints.constraints:
        if _name(c) == "ProductInequalityConstraint":
            out.append(optim.product_constraint([ckeys.index(k) for k in c.features], list(c.exponents), c.rhs, c.sign))
    return out


def _x_adapt(outputs, feat, experiments):
    return outputs.preprocess_experiments_one_valid_output(feat.key, experiments)[feat.key].values


def get_output_constraints(outputs, experiments) -> List[Ob.OutputConstraint]:
    """utils/torch_tools.py:340-381: one or two `c(y) <= 0` callables per ConstrainedObjective output, eta = 1 / steepness
    carried inside each OutputConstraint."""
    cons = []
    for idx, feat in enumerate(outputs.get()):
        if feat.objective is not None and _is(feat.objective, "ConstrainedObjective"):
            if _is(feat.objective, "ConstrainedCategoricalObjective"):
                raise NotImplementedError("categorical outputs are outside the accelerated path")
            cons += Ob.constraints_from_sigmoid_objectives([objective_spec(idx, feat.objective, _x_adapt(outputs, feat, experiments))])
    return cons


def _callables_and_weights(outputs, experiments, exclude_constraints=False, allowed: Optional[Tuple[str, ...]] = None,
                           adapt_weights_to_1_inf=False) -> List[Ob.ObjectiveSpec]:
    """utils/torch_tools.py:598-659 -> op table entries carrying their weights."""
    ops = []
    for i, feat in enumerate(outputs.get()):
        if feat.objective is None:
            continue
        if exclude_constraints and _is(feat.objective, "ConstrainedObjective"):
            continue
        if allowed is not None and not _is(feat.objective, *allowed):
            continue
        ops.append(objective_spec(i, feat.objective, _x_adapt(outputs, feat, experiments)))
    if adapt_weights_to_1_inf and ops:
        min_w = min(op.w for op in ops)
        for op in ops:
            op.w = op.w / min_w
    return ops


def get_multiobjective_objective(outputs, experiments) -> Ob.MultiObjective:
    """utils/torch_tools.py:699-727."""
    return Ob.MultiObjective(_callables_and_weights(outputs, experiments,
                                                    allowed=("MaximizeObjective", "MinimizeObjective", "CloseToTargetObjective")))


def get_ref_point_mask(domain) -> np.ndarray:
    """utils/multiobjective.py:18-55."""
    mask = []
    for feat in domain.outputs.get():
        if feat.objective is None:
            continue
        if _is(feat.objective, "MaximizeObjective"):
            mask.append(1.0)
        elif _is(feat.objective, "MinimizeObjective", "CloseToTargetObjective"):
            mask.append(-1.0)
    if len(mask) < 2:
        raise ValueError("At least two output features have to be provided.")
    return np.array(mask)


def _mo_keys(domain) -> List[str]:
    return [f.key for f in domain.outputs.get() if f.objective is not None
            and _is(f.objective, "MaximizeObjective", "MinimizeObjective", "CloseToTargetObjective")]


def _non_constrained_objective_keys(domain) -> List[str]:
    """outputs.get_keys_by_objective(excludes=ConstrainedObjective)."""
    return [f.key for f in domain.outputs.get() if f.objective is not None and not _is(f.objective, "ConstrainedObjective")]


# ------------------------------------------------------------------------------------------------------------------
# priors / kernels: data model -> device spec + fit priors
# ------------------------------------------------------------------------------------------------------------------
def map_prior(data_model, d: int = 1):
    """priors/mapper.py:9-61."""
    if data_model is None:
        return None
    t = _name(data_model)
    if t == "GammaPrior":
        return F.GammaPrior(float(data_model.concentration), float(data_model.rate))
    if t == "NormalPrior":
        return F.NormalPrior(float(data_model.loc), float(data_model.scale))
    if t == "LogNormalPrior":
        return F.LogNormalPrior(float(data_model.loc), float(data_model.scale))
    if t == "DimensionalityScaledLogNormalPrior":
        return F.DimensionalityScaledLogNormalPrior(d, data_model.loc, data_model.loc_scaling, data_model.scale, data_model.scale_scaling)
    raise NotImplementedError(f"prior {t} is outside the accelerated path")


def collect_kernel_priors(data_model, spec) -> Tuple[Dict[int, object], Dict[int, object]]:
    """Walk the kernel DATA MODEL and its mapped spec tree side by side (leaves depth first, ScaleKernels pre-order: the
    numbering of fit._Layout) and return ({leaf index: lengthscale prior}, {scale index: outputscale prior})."""
    ls, os_ = {}, {}
    counters = {"leaf": 0, "scale": 0}

    def rec(dm, sp):
        t = _name(dm)
        if t in ("RBFKernel", "MaternKernel", "HammingDistanceKernel", "TanimotoKernel"):
            i = counters["leaf"]
            counters["leaf"] += 1
            prior = getattr(dm, "lengthscale_prior", None)
            if prior is not None and not isinstance(sp, K.TanimotoKernel):
                n = len(sp.categorical_features) if isinstance(sp, K.HammingDistanceKernel) else len(sp.active_dims)
                ls[i] = map_prior(prior, d=n)
        elif t == "ScaleKernel":
            i = counters["scale"]
            counters["scale"] += 1
            if dm.outputscale_prior is not None:
                os_[i] = map_prior(dm.outputscale_prior)
            rec(dm.base_kernel, sp.base_kernel)
        elif t in ("AdditiveKernel", "MultiplicativeKernel"):
            for c_dm, c_sp in zip(dm.kernels, sp.kernels):
                rec(c_dm, c_sp)
        else:
            raise NotImplementedError(f"kernel {t} is outside the accelerated path")

    rec(data_model, spec)
    return ls, os_


# ------------------------------------------------------------------------------------------------------------------
# surrogates: data model + experiments -> fitted SingleTaskGPSpec on the GLOBAL transformed column layout
# ------------------------------------------------------------------------------------------------------------------
def _specs_of(surrogate_dm) -> dict:
    return dict(surrogate_dm.input_preprocessing_specs)


def _continuous_feature_keys(inputs, specs) -> List[str]:
    """surrogates/utils.py:46-73."""
    non_cont = [k for k, v in specs.items() if not (str(getattr(v, "value", v)) == "DESCRIPTOR" or _name(v) == "MordredDescriptors")]
    return sorted(f.key for f in inputs.get() if f.key not in non_cont)


def _categorical_feature_keys(specs) -> List[str]:
    """surrogates/utils.py:76-100."""
    return sorted(k for k, v in specs.items()
                  if str(getattr(v, "value", v)) != "DESCRIPTOR"
                  and _name(v) not in ("Fingerprints", "Fragments", "FingerprintsFragments", "MordredDescriptors"))


def _molecular_feature_keys(specs) -> List[str]:
    """surrogates/utils.py:20-43."""
    return sorted(k for k, v in specs.items() if _name(v) in ("Fingerprints", "Fragments", "FingerprintsFragments"))


class _SurrogateFit:
    """What `BotorchSurrogates.fit` + `compatibilize` produce for ONE output: a spec on the domain-wide column layout."""

    def __init__(self, surrogate_dm, domain_inputs, all_specs: dict, features2idx: Dict[str, Tuple[int, ...]], d: int):
        self.dm = surrogate_dm
        self.inputs = surrogate_dm.inputs
        self.all_specs = all_specs
        self.features2idx = features2idx
        self.d = d
        if len(surrogate_dm.outputs) != 1:
            raise NotImplementedError("multi-output surrogates are outside the accelerated path")
        self.output_key = surrogate_dm.outputs.get_keys()[0]
        # the columns this surrogate sees (FilterFeatures of compatibilize, botorch_surrogates.py:98-118)
        self.columns = [i for key in self.inputs.get_keys() for i in features2idx[key]]

    def idx(self, feats) -> List[int]:
        return sorted(i for f in feats for i in self.features2idx[f])

    def scaler(self, X: pd.DataFrame):
        """get_scaler (surrogates/utils.py:103-164) -> (in_offset, in_scale) on the global layout, identity elsewhere."""
        off, scl = np.zeros(self.d), np.ones(self.d)
        kind = str(getattr(self.dm.scaler, "value", self.dm.scaler))
        if kind == "IDENTITY":
            return off, scl
        specs = _specs_of(self.dm)
        ord_keys = _continuous_feature_keys(self.inputs, specs)
        ord_dims = self.idx(ord_keys)
        if not ord_dims:
            return off, scl
        if kind == "NORMALIZE":
            lower, upper = self.inputs.get_bounds(specs=specs, experiments=X)
            local = [i for key in self.inputs.get_keys() for i in self.features2idx[key]]
            lo = dict(zip(local, lower))
            up = dict(zip(local, upper))
            for j in ord_dims:
                off[j] = lo[j]
                scl[j] = (up[j] - lo[j]) if up[j] > lo[j] else 1.0
        elif kind == "STANDARDIZE":
            tX = self.inputs.transform(X, specs).values
            local = [i for key in self.inputs.get_keys() for i in self.features2idx[key]]
            pos = {g: k for k, g in enumerate(local)}
            for j in ord_dims:
                col = tX[:, pos[j]]
                s = float(col.std(ddof=1)) if len(col) > 1 else 1.0
                off[j] = float(col.mean())
                scl[j] = s if s >= 1e-8 else 1.0
        else:
            raise ValueError("Scaler enum not known.")
        return off, scl

    def kernel(self):
        """(spec tree, lengthscale priors, outputscale priors) for the surrogate type."""
        t = _name(self.dm)
        mapper = lambda feats: self.idx(feats)  # noqa: E731  (features_to_idx_mapper of kernels.map)
        specs = _specs_of(self.dm)
        if t in ("SingleTaskGPSurrogate", "TanimotoGPSurrogate"):
            spec = map_kernel(self.dm.kernel, active_dims=list(self.columns), features_to_idx_mapper=mapper)
            ls, os_ = collect_kernel_priors(self.dm.kernel, spec)
            return spec, ls, os_
        if t == "MixedSingleTaskGPSurrogate":
            # [UPSTREAM] botorch MixedSingleTaskGP: Scale(K_c + Scale(K_h)) + Scale(K_c * K_h), the continuous kernel from
            # BoFire's cont_kernel_factory on the ordinal columns, CategoricalKernel (= the one-hot Hamming kernel, pinned by
            # tests/bofire/kernels/test_categorical.py) with one ARD lengthscale per categorical FEATURE; every factor has
            # its own hyper-parameters.  BoFire's `categorical_kernel` field is not handed to BoTorch
            # (mixed_single_task_gp.py:90-108), so it is not used here either.
            ord_dims = self.idx(_continuous_feature_keys(self.inputs, specs))
            cats = {self.features2idx[k][0]: len(self.features2idx[k]) for k in _categorical_feature_keys(specs)}

            def cont():
                return map_kernel(self.dm.continuous_kernel, active_dims=list(ord_dims), features_to_idx_mapper=mapper)

            def ham():
                return K.HammingDistanceKernel(dict(cats), [1.0] * len(cats))

            if not ord_dims:
                spec = K.ScaleKernel(ham(), 1.0)
                return spec, {}, {}
            spec = K.AdditiveKernel([K.ScaleKernel(K.AdditiveKernel([cont(), K.ScaleKernel(ham(), 1.0)]), 1.0),
                                     K.ScaleKernel(K.MultiplicativeKernel([cont(), ham()]), 1.0)])
            lp = map_prior(getattr(self.dm.continuous_kernel, "lengthscale_prior", None), d=len(ord_dims))
            ls = {0: lp, 2: lp} if lp is not None else {}
            return spec, ls, {}
        if t == "MixedTanimotoGPSurrogate":
            # mixed_tanimoto_gp.py:101-215: (s1 K_c + s2 K_m + s3 K_h) + (s4 K_c * s5 K_m * s6 K_h), parts dropped when the
            # space has no such columns
            ord_dims = self.idx(_continuous_feature_keys(self.inputs, specs))
            mol_dims = self.idx(_molecular_feature_keys(specs))
            cats = {self.features2idx[k][0]: len(self.features2idx[k]) for k in _categorical_feature_keys(specs)}
            parts = []
            lp = map_prior(getattr(self.dm.continuous_kernel, "lengthscale_prior", None), d=max(len(ord_dims), 1))

            def build():
                leaves = []
                if ord_dims:
                    leaves.append(("c", map_kernel(self.dm.continuous_kernel, active_dims=list(ord_dims), features_to_idx_mapper=mapper)))
                if mol_dims:
                    leaves.append(("m", K.TanimotoKernel(list(mol_dims))))
                if cats:
                    leaves.append(("h", K.HammingDistanceKernel(dict(cats), [1.0] * len(cats))))
                return leaves

            add = build()
            mul = build()
            parts = [K.AdditiveKernel([K.ScaleKernel(k, 1.0) for _, k in add]),
                     K.MultiplicativeKernel([K.ScaleKernel(k, 1.0) for _, k in mul])]
            spec = K.AdditiveKernel(parts)
            ls = {}
            if lp is not None:
                for i, (tag, _) in enumerate(add + mul):
                    if tag == "c":
                        ls[i] = lp
            return spec, ls, {}
        raise NotImplementedError(f"surrogate {t} is outside the accelerated path (supported: SingleTaskGPSurrogate, "
                                  "MixedSingleTaskGPSurrogate, TanimotoGPSurrogate, MixedTanimotoGPSurrogate)")

    def fit(self, tX_global: np.ndarray, X_df: pd.DataFrame, y: np.ndarray, device=None, options: Optional[dict] = None,
            hyperparameters: Optional[SingleTaskGPSpec] = None) -> SingleTaskGPSpec:
        off, scl = self.scaler(X_df)
        if hyperparameters is not None:   # fitted values handed in (e.g. shared with a CPU BoFire run): no device fit
            hp = hyperparameters
            return SingleTaskGPSpec(kernel=hp.kernel, y=y, in_offset=off, in_scale=scl, mean_const=hp.mean_const, noise=hp.noise,
                                    y_mean=hp.y_mean, y_std=hp.y_std)
        spec, ls_priors, os_priors = self.kernel()
        res = F.fit_gp(tX_global, y, spec, in_offset=off, in_scale=scl, noise_prior=map_prior(self.dm.noise_prior),
                       lengthscale_priors=ls_priors, outputscale_priors=os_priors, options=options, device=device)
        out = res.spec
        if str(getattr(self.dm.output_scaler, "value", self.dm.output_scaler)) != "STANDARDIZE":
            raise NotImplementedError("output_scaler other than STANDARDIZE is outside the accelerated path")
        return out


# ------------------------------------------------------------------------------------------------------------------
# strategies
# ------------------------------------------------------------------------------------------------------------------
class Strategy:
    """strategies/strategy.py:14-262."""

    def __init__(self, data_model):
        self.domain = data_model.domain
        self.seed = data_model.seed if data_model.seed is not None else int(np.random.default_rng().integers(1000))
        self.rng = np.random.default_rng(self.seed)
        self._experiments = None
        self._candidates = None

    def _get_seed(self) -> int:
        return int(self.rng.integers(1, 100000))

    @classmethod
    def from_spec(cls, data_model):
        return cls(data_model=data_model)

    @property
    def experiments(self) -> Optional[pd.DataFrame]:
        return self._experiments

    @property
    def candidates(self) -> Optional[pd.DataFrame]:
        return self._candidates

    def tell(self, experiments: pd.DataFrame, replace: bool = False) -> None:
        if len(experiments) == 0:
            return
        if replace:
            self.set_experiments(experiments)
        else:
            self.add_experiments(experiments)
        self._tell()

    def _tell(self) -> None:
        pass

    def ask(self, candidate_count: Optional[int] = None, add_pending: bool = False, raise_validation_error: bool = True) -> pd.DataFrame:
        if candidate_count is not None and candidate_count < 1:
            raise ValueError(f"Candidate_count has to be at least 1 but got {candidate_count}.")
        if not self.has_sufficient_experiments():
            raise ValueError("Not enough experiments available to execute the strategy.")
        candidates = self._ask(candidate_count=candidate_count)
        self.domain.validate_candidates(candidates=candidates, only_inputs=True, raise_validation_error=raise_validation_error)
        if candidate_count is not None and len(candidates) != candidate_count:
            warnings.warn(f"Expected {candidate_count} candidates, got {len(candidates)}", UserWarning)
        if add_pending:
            self.add_candidates(candidates)
        return candidates

    def has_sufficient_experiments(self) -> bool:
        raise NotImplementedError

    def _ask(self, candidate_count=None) -> pd.DataFrame:
        raise NotImplementedError

    def set_candidates(self, candidates: pd.DataFrame):
        keys = self.domain.inputs.get_keys()
        candidates = self.domain.inputs.validate_experiments(candidates[keys], strict=False)
        self._candidates = candidates[keys]

    def add_candidates(self, candidates: pd.DataFrame):
        keys = self.domain.inputs.get_keys()
        candidates = self.domain.inputs.validate_experiments(candidates[keys], strict=False)
        if self.candidates is None:
            self._candidates = candidates[keys]
        else:
            self._candidates = pd.concat((self.candidates, candidates[keys]), ignore_index=True)

    def reset_candidates(self):
        self._candidates = None

    @property
    def num_candidates(self) -> int:
        return 0 if self.candidates is None else len(self.candidates)

    def set_experiments(self, experiments: pd.DataFrame):
        self._experiments = self.domain.validate_experiments(experiments)

    def add_experiments(self, experiments: pd.DataFrame):
        experiments = self.domain.validate_experiments(experiments)
        if self.experiments is None:
            self._experiments = experiments
        else:
            self._experiments = pd.concat((self.experiments, experiments), ignore_index=True)

    @property
    def num_experiments(self) -> int:
        return 0 if self.experiments is None else len(self.experiments)


class RandomStrategy(Strategy):
    """strategies/random.py:29-353: uniform samples of the constrained input space -- `Inputs.sample` without constraints,
    hit-and-run over the linear polytope (optim.sample_q_batches_from_polytope) with categoricals / discretes drawn
    uniformly, one NChooseK sub-space per sample, rejection for the product constraints."""

    def __init__(self, data_model):
        super().__init__(data_model)
        self.num_base_samples = data_model.num_base_samples
        self.max_iters = data_model.max_iters
        self.fallback_sampling_method = data_model.fallback_sampling_method
        self.n_burnin = data_model.n_burnin
        self.n_thinning = data_model.n_thinning

    def has_sufficient_experiments(self) -> bool:
        return True

    def _ask(self, candidate_count) -> pd.DataFrame:
        names = {_name(c) for c in self.domain.constraints.constraints}
        if names <= {"LinearInequalityConstraint", "LinearEqualityConstraint", "NChooseKConstraint", "InterpointEqualityConstraint"}:
            return self._sample_with_nchooseks(candidate_count)
        num_base = self.num_base_samples or candidate_count
        n_iters, n_found, valid_samples = 0, 0, []
        while n_found < candidate_count:
            if n_iters > self.max_iters:
                raise ValueError("Maximum iterations exceeded in rejection sampling.")
            samples = self._sample_with_nchooseks(num_base)
            valid = self.domain.constraints.is_fulfilled(samples)
            n_found += int(np.sum(valid))
            valid_samples.append(samples[valid])
            n_iters += 1
        return pd.concat(valid_samples, ignore_index=True).iloc[:candidate_count]

    def _sample_with_nchooseks(self, candidate_count: int) -> pd.DataFrame:
        if any(_name(c) == "NChooseKConstraint" for c in self.domain.constraints.constraints):
            _, unused = self.domain.get_nchoosek_combinations()
            if candidate_count <= len(unused):
                combos = [unused[i] for i in self.rng.choice(len(unused), size=candidate_count, replace=False)]
                per_it = 1
            else:
                combos = unused
                per_it = math.ceil(candidate_count / len(unused))
            samples = []
            for u in combos:
                domain = copy.deepcopy(self.domain)
                domain.constraints = domain.constraints.get(excludes=type(next(c for c in self.domain.constraints.constraints
                                                                               if _name(c) == "NChooseKConstraint")))
                for key in u:
                    domain.inputs.get_by_key(key).bounds = [0.0, 0.0]
                samples.append(self._sample_from_polytope(domain, per_it, self._get_seed()))
            samples = pd.concat(samples, axis=0, ignore_index=True)
            return samples.sample(n=candidate_count, replace=False, ignore_index=True, random_state=self._get_seed())
        return self._sample_from_polytope(self.domain, candidate_count, self._get_seed())

    def _sample_from_polytope(self, domain, n: int, seed: int) -> pd.DataFrame:
        """strategies/random.py:180-353."""
        if len(domain.constraints) == 0:
            return domain.inputs.sample(n, self.fallback_sampling_method, seed=seed)
        ckeys = _continuous_keys(domain)
        allkeys = domain.inputs.get_keys()
        fixed = {f.key: f.fixed_value()[0] for f in domain.inputs.get() if _is(f, "ContinuousInput") and f.is_fixed()}
        eqs = []
        for idx, coef, rhs in get_linear_constraints(domain, "LinearEqualityConstraint"):
            if len(idx) == 1:   # pseudo-fixed feature (random.py:216-226)
                fixed[allkeys[int(idx[0])]] = float(rhs / coef[0])
            else:
                eqs.append((idx, coef, rhs))
        ineqs = get_linear_constraints(domain, "LinearInequalityConstraint")
        inter = get_interpoint_constraints(domain, n_candidates=n)
        free = [k for k in ckeys if k not in fixed]
        if not free:
            warnings.warn("Nothing to sample, all is fixed. Just the fixed set is returned.", UserWarning)
            samples = pd.DataFrame(index=range(n))
        else:
            col = {k: j for j, k in enumerate(free)}

            def unfix(cons, eq):
                """_generate_unfixed_lin_constraints (random.py): drop fixed columns, move them to the right-hand side."""
                out = []
                for idx, coef, rhs in cons:
                    if idx.dim() == 2:   # inter-point: [[point, column], ...]
                        key = allkeys[int(idx[0, 1])]
                        if key in fixed:
                            continue
                        out.append((torch.stack([idx[:, 0], torch.full_like(idx[:, 0], col[key])], dim=1), coef, rhs))
                        continue
                    new_i, new_c, r = [], [], float(rhs)
                    for i, c in zip(idx.tolist(), coef.tolist()):
                        key = allkeys[i]
                        if key in fixed:
                            r -= c * fixed[key]
                        else:
                            new_i.append(col[key])
                            new_c.append(c)
                    if new_i:
                        out.append((torch.tensor(new_i, dtype=torch.long), torch.tensor(new_c, dtype=torch.double), r))
                return out

            lower = [domain.inputs.get_by_key(k).lower_bound for k in free]
            upper = [domain.inputs.get_by_key(k).upper_bound for k in free]
            bounds = torch.tensor([lower, upper], dtype=torch.double)
            cand = optim.sample_q_batches_from_polytope(1, n, bounds, unfix(ineqs, False) or None,
                                                        (unfix(eqs, True) + unfix(inter, True)) or None, seed=seed,
                                                        n_burnin=self.n_burnin, n_thinning=self.n_thinning)[0]
            samples = pd.DataFrame(cand.numpy(), index=range(n), columns=free)
        others = [f for f in domain.inputs.get() if _is(f, "CategoricalInput", "DiscreteInput")]
        if others:
            sub = type(domain.inputs)(features=others)
            samples = pd.concat([samples, sub.sample(n, seed=seed)], axis=1)
        for k, v in fixed.items():
            samples[k] = v
        return domain.inputs.validate_candidates(samples)[allkeys]


class PredictiveStrategy(Strategy):
    """strategies/predictives/predictive.py:23-216."""

    def __init__(self, data_model):
        super().__init__(data_model)
        self.is_fitted = False

    @property
    def input_preprocessing_specs(self) -> dict:
        raise NotImplementedError

    def ask(self, candidate_count=None, add_pending=False, raise_validation_error=True) -> pd.DataFrame:
        candidates = super().ask(candidate_count=candidate_count, add_pending=add_pending, raise_validation_error=raise_validation_error)
        self.domain.validate_candidates(candidates=candidates, raise_validation_error=raise_validation_error)
        return candidates

    def tell(self, experiments: pd.DataFrame, replace: bool = False, retrain: bool = True):
        if len(experiments) == 0:
            return
        if replace:
            self.set_experiments(experiments)
        else:
            self.add_experiments(experiments)
        cleaned = self.domain.outputs.preprocess_experiments_all_valid_outputs(experiments=experiments)
        for feature in self.domain.inputs.get_fixed():
            if _is(feature, "TaskInput"):
                continue
            fixed_value = feature.fixed_value()
            assert fixed_value is not None
            if (cleaned[feature.key] == fixed_value[0]).all():
                raise ValueError(f"No variance in experiments for fixed feature {feature.key}")
        if retrain and self.has_sufficient_experiments():
            self.fit()
            self._tell()

    def predict(self, experiments: pd.DataFrame) -> pd.DataFrame:
        if self.is_fitted is not True:
            raise ValueError("Model not yet fitted.")
        transformed = self.domain.inputs.transform(experiments=experiments, specs=self.input_preprocessing_specs)
        preds, stds = self._predict(transformed)
        keys = [f.key for f in self.domain.outputs.get() if _is(f, "ContinuousOutput")]   # utils/naming_conventions.py:9-34
        predictions = pd.DataFrame(data=np.hstack((preds, stds)), columns=[f"{k}_pred" for k in keys] + [f"{k}_sd" for k in keys])
        objectives = self.domain.outputs(predictions, experiments_adapt=self.experiments, predictions=True)
        predictions = pd.concat((predictions, objectives), axis=1)
        predictions.index = experiments.index
        return predictions

    def fit(self):
        assert self.experiments is not None and len(self.experiments) > 0, "No fitting data available"
        self.domain.validate_experiments(self.experiments, strict=True)
        self._fit(self.experiments)
        self.is_fitted = True


class BotorchStrategy(PredictiveStrategy):
    """strategies/predictives/botorch.py:57-750 with the device state in place of the ModelListGP."""

    def __init__(self, data_model, device=None, fit_options: Optional[dict] = None, hyperparameters: Optional[Dict[str, SingleTaskGPSpec]] = None):
        super().__init__(data_model)
        self.num_restarts = data_model.num_restarts
        self.num_raw_samples = data_model.num_raw_samples
        self.descriptor_method = data_model.descriptor_method
        self.categorical_method = data_model.categorical_method
        self.discrete_method = data_model.discrete_method
        self.surrogate_specs = data_model.surrogate_specs
        if data_model.outlier_detection_specs is not None:
            raise NotImplementedError("outlier detection is outside the accelerated path")
        if data_model.frequency_hyperopt > 0:
            raise NotImplementedError("hyper-parameter optimisation runs (frequency_hyperopt > 0) are outside the accelerated path")
        if data_model.local_search_config is not None:
            raise NotImplementedError("LSR-BO local search is outside the accelerated path")
        for out in self.domain.outputs.get():
            if not _is(out, "ContinuousOutput"):
                raise NotImplementedError("categorical outputs are outside the accelerated path")
        self.maxiter = data_model.maxiter
        self.batch_limit = data_model.batch_limit
        self.device = device
        self.fit_options = fit_options
        self.hyperparameters = hyperparameters      # {output key: SingleTaskGPSpec with fitted values}: skips the device fit
        self.model: Optional[DeviceGPState] = None
        self.fitted_specs: Dict[str, SingleTaskGPSpec] = {}
        torch.manual_seed(self.seed)

    # -- layout ---------------------------------------------------------------------------------------------------
    @property
    def input_preprocessing_specs(self) -> dict:
        return dict(self.surrogate_specs.input_preprocessing_specs)

    @property
    def _features2idx(self) -> Dict[str, Tuple[int, ...]]:
        return self.domain.inputs._get_transform_info(self.input_preprocessing_specs)[0]

    @property
    def _features2names(self) -> Dict[str, Tuple[str, ...]]:
        return self.domain.inputs._get_transform_info(self.input_preprocessing_specs)[1]

    def _nonlinear_constraint_count(self) -> int:
        return sum(1 for c in self.domain.constraints.constraints if _name(c) in ("NChooseKConstraint", "ProductInequalityConstraint"))

    def _get_optimizer_options(self) -> Dict[str, int]:
        return {"batch_limit": self.batch_limit if self._nonlinear_constraint_count() == 0 else 1, "maxiter": self.maxiter}

    # -- fit / predict --------------------------------------------------------------------------------------------
    def _fit(self, experiments: pd.DataFrame):
        """BotorchSurrogates.fit + compatibilize (surrogates/botorch_surrogates.py:43-128): one exact GP per output, fitted
        on the device (fit.fit_gp: marginal likelihood + priors, L-BFGS-B), all on ONE training matrix in the domain-wide
        transformed column layout."""
        specs = self.input_preprocessing_specs
        features2idx = self._features2idx
        d = sum(len(v) for v in features2idx.values())
        out_keys = self.domain.outputs.get_keys()
        by_key = {s.outputs.get_keys()[0]: s for s in self.surrogate_specs.surrogates}
        clean = self.domain.outputs.preprocess_experiments_all_valid_outputs(experiments)
        for key in out_keys:
            own = by_key[key].outputs.preprocess_experiments_all_valid_outputs(experiments, output_feature_keys=[key])
            if len(own) != len(clean):
                raise NotImplementedError("outputs with different sets of valid experiments need one training matrix per "
                                          "output: outside the accelerated path (all outputs share X_train on the device)")
        tX = np.ascontiguousarray(self.domain.inputs.transform(clean, specs).values.astype(np.float64))
        fitted = []
        for key in out_keys:
            sf = _SurrogateFit(by_key[key], self.domain.inputs, specs, features2idx, d)
            hp = None if self.hyperparameters is None else self.hyperparameters.get(key)
            spec = sf.fit(tX, clean[sf.inputs.get_keys()], clean[key].values.astype(np.float64), device=self.device,
                          options=self.fit_options, hyperparameters=hp)
            self.fitted_specs[key] = spec
            fitted.append(spec)
        if self.model is not None:
            self.model.close()
        self.model = DeviceGPState(tX, fitted, device=self.device).factorize()

    def _predict(self, transformed: pd.DataFrame) -> Tuple[np.ndarray, np.ndarray]:
        """botorch.py:174-194: posterior with observation noise."""
        return self.model.predict(np.ascontiguousarray(transformed.values.astype(np.float64)))

    def calc_acquisition(self, candidates: pd.DataFrame, combined: bool = False) -> np.ndarray:
        acqf = self._get_acqfs(1)[0]
        transformed = self.domain.inputs.transform(candidates, self.input_preprocessing_specs)
        return optim.calc_acquisition(acqf, transformed.values.astype(np.float64), combined=combined)

    # -- ask ------------------------------------------------------------------------------------------------------
    def _setup_ask(self):
        """botorch.py:227-295."""
        n_cat = sum(1 for f in self.domain.inputs.get() if _is(f, "CategoricalInput", "DiscreteInput"))
        n_combos = len(self.domain.inputs.get_categorical_combinations())
        lower, upper = self.domain.inputs.get_bounds(specs=self.input_preprocessing_specs)
        bounds = torch.tensor([lower, upper], dtype=torch.double)
        if self._nonlinear_constraint_count() == 0:
            generator, nonlinear = None, None
        else:
            sampler = RandomStrategy(_RandomSpec(self.domain, seed=self._get_seed()))
            specs = self.input_preprocessing_specs

            def generator(n: int, q: int, seed: int) -> torch.Tensor:   # get_initial_conditions_generator, torch_tools.py:809-864
                out = []
                for _ in range(n):
                    c = sampler.ask(q)
                    out.append(torch.from_numpy(self.domain.inputs.transform(c, specs).values.astype(np.float64)))
                return torch.stack(out, dim=0)

            nonlinear = get_nonlinear_constraints(self.domain)
        free = [str(getattr(m, "value", m)) == "FREE" for m in (self.categorical_method, self.descriptor_method, self.discrete_method)]
        if n_cat == 0 or n_combos == 1 or all(free):
            fixed_features, fixed_features_list = self.get_fixed_features(), None
        else:
            fixed_features, fixed_features_list = None, self.get_categorical_combinations()
        return bounds, generator, nonlinear, fixed_features, fixed_features_list

    def _postprocess_candidates(self, candidates: torch.Tensor) -> pd.DataFrame:
        """botorch.py:297-324."""
        names = [item for key in self.domain.inputs.get_keys() for item in self._features2names[key]]
        df = pd.DataFrame(data=candidates.detach().numpy(), columns=names)
        df = self.domain.inputs.inverse_transform(df, self.input_preprocessing_specs)
        preds = self.predict(df)
        return pd.concat((df, preds), axis=1)

    def _optimize_acqf_continuous(self, candidate_count: int, acqfs: list, bounds: torch.Tensor, generator, nonlinear_constraints,
                                  fixed_features, fixed_features_list) -> Tuple[torch.Tensor, torch.Tensor]:
        """botorch.py:326-406."""
        kw = dict(options=self._get_optimizer_options(), seed=int(torch.randint(0, 1000000, (1,)).item()),
                  inequality_constraints=get_linear_constraints(self.domain, "LinearInequalityConstraint") or None)
        eqs = get_linear_constraints(self.domain, "LinearEqualityConstraint")
        if nonlinear_constraints:
            kw.update(nonlinear_inequality_constraints=nonlinear_constraints, generator=generator)
        if len(acqfs) > 1:
            return optim.optimize_acqf_list(acqfs, bounds, self.num_restarts, self.num_raw_samples, fixed_features=fixed_features,
                                            fixed_features_list=fixed_features_list, equality_constraints=eqs or None, **kw)
        if fixed_features_list:
            return optim.optimize_acqf_mixed(acqfs[0], bounds, candidate_count, self.num_restarts, self.num_raw_samples,
                                             fixed_features_list=fixed_features_list, equality_constraints=eqs or None, **kw)
        eqs = eqs + get_interpoint_constraints(self.domain, candidate_count)
        return optim.optimize_acqf(acqfs[0], bounds, candidate_count, self.num_restarts, self.num_raw_samples,
                                   fixed_features=fixed_features or None, equality_constraints=eqs or None, **kw)

    def _ask(self, candidate_count: int) -> pd.DataFrame:
        """botorch.py:408-521."""
        assert candidate_count is not None and candidate_count > 0, "candidate_count has to be larger than zero."
        if self.experiments is None:
            raise ValueError("No experiments have been provided yet.")
        acqfs = self._get_acqfs(candidate_count)
        inputs = self.domain.inputs
        if sum(1 for f in inputs.get() if _is(f, "DiscreteInput", "CategoricalInput")) == len(inputs):
            if len(acqfs) > 1:
                raise NotImplementedError("Multiple Acqfs are currently not supported for purely combinatorial search spaces.")
            choices = pd.DataFrame.from_dict([{e[0]: e[1] for e in combi} for combi in inputs.get_categorical_combinations()])
            for feat in inputs.get_fixed():
                choices[feat.key] = feat.fixed_value()[0]
            merged = choices.merge(self.experiments[inputs.get_keys()], on=list(choices.columns), how="left", indicator=True)
            filtered = merged[merged["_merge"] == "left_only"].copy()
            filtered.drop(columns=["_merge"], inplace=True)
            t_choices = torch.from_numpy(inputs.transform(filtered, specs=self.input_preprocessing_specs).values.astype(np.float64))
            candidates, _ = optim.optimize_acqf_discrete(acqfs[0], q=candidate_count, unique=True, choices=t_choices)
            return self._postprocess_candidates(candidates.reshape(candidate_count, -1))
        bounds, generator, nonlinears, fixed_features, fixed_features_list = self._setup_ask()
        candidates, _ = self._optimize_acqf_continuous(candidate_count, acqfs, bounds, generator, nonlinears, fixed_features,
                                                       fixed_features_list)
        return self._postprocess_candidates(candidates.reshape(candidate_count, -1))

    def _get_acqfs(self, n: int) -> list:
        raise NotImplementedError

    def get_fixed_features(self) -> Dict[int, float]:
        """botorch.py:523-595."""
        fixed = {}
        features2idx = self._features2idx
        specs = self.input_preprocessing_specs
        for feat in self.domain.inputs.get():
            if feat.fixed_value() is not None:
                vals = feat.fixed_value(transform_type=specs.get(feat.key))
                for j, idx in enumerate(features2idx[feat.key]):
                    fixed[idx] = vals[j]
        enc = [str(getattr(v, "value", v)) for v in specs.values()]
        if str(getattr(self.categorical_method, "value", self.categorical_method)) == "FREE" and "ONE_HOT" in enc:
            for feat in self.domain.inputs.get():
                if _is(feat, "CategoricalInput") and str(getattr(specs.get(feat.key), "value", specs.get(feat.key))) == "ONE_HOT" \
                        and feat.is_fixed() is False:
                    for cat in feat.get_forbidden_categories():
                        transformed = feat.to_onehot_encoding(pd.Series([cat]))
                        for j, idx in enumerate(features2idx[feat.key]):
                            if transformed.values[0, j] == 1.0:
                                fixed[idx] = 0
        if str(getattr(self.descriptor_method, "value", self.descriptor_method)) == "FREE" and "DESCRIPTOR" in enc:
            for feat in self.domain.inputs.get():
                if _is(feat, "CategoricalDescriptorInput") and str(getattr(specs.get(feat.key), "value", specs.get(feat.key))) == "DESCRIPTOR" \
                        and feat.is_fixed() is False:
                    lower, upper = feat.get_bounds(specs[feat.key])
                    for j, idx in enumerate(features2idx[feat.key]):
                        if lower[j] == upper[j]:
                            fixed[idx] = lower[j]
        return fixed

    def get_categorical_combinations(self) -> List[Dict[int, float]]:
        """botorch.py:597-675."""
        fixed_basis = self.get_fixed_features()
        meth = {k: str(getattr(m, "value", m)) for k, m in (("descriptor", self.descriptor_method), ("discrete", self.discrete_method),
                                                            ("categorical", self.categorical_method))}
        if all(v == "FREE" for v in meth.values()):
            return [{}]

        def wanted(f) -> bool:
            is_desc = _is(f, "CategoricalDescriptorInput")
            if _is(f, "DiscreteInput"):
                return meth["discrete"] == "EXHAUSTIVE"
            if is_desc:
                return meth["descriptor"] == "EXHAUSTIVE"
            if _is(f, "CategoricalInput"):
                return meth["categorical"] == "EXHAUSTIVE"
            return False

        cats = [f for f in self.domain.inputs.get() if _is(f, "CategoricalInput") and wanted(f) and not f.is_fixed()]
        discs = [f for f in self.domain.inputs.get() if _is(f, "DiscreteInput") and wanted(f) and not f.is_fixed()]
        import itertools
        combos = list(itertools.product(*([[(f.key, c) for c in f.get_allowed_categories()] for f in cats]
                                          + [[(f.key, v) for v in f.values] for f in discs])))
        if len(combos) == 1:
            return [fixed_basis]
        features2idx = self._features2idx
        specs = self.input_preprocessing_specs
        out = []
        for combo in combos:
            ff = copy.deepcopy(fixed_basis)
            for key, val in combo:
                feature = self.domain.inputs.get_by_key(key)
                enc = specs.get(key)
                if _is(feature, "CategoricalDescriptorInput") and str(getattr(enc, "value", enc)) == "DESCRIPTOR":
                    index = feature.categories.index(val)
                    for j, idx in enumerate(features2idx[key]):
                        ff[idx] = feature.values[index][j]
                elif _is(feature, "CategoricalMolecularInput"):
                    transformed = feature.to_descriptor_encoding(enc, pd.Series([val]))
                    for j, idx in enumerate(features2idx[key]):
                        ff[idx] = transformed.values[0, j]
                elif _is(feature, "CategoricalInput"):
                    transformed = feature.to_onehot_encoding(pd.Series([val]))
                    for j, idx in enumerate(features2idx[key]):
                        ff[idx] = transformed.values[0, j]
                elif _is(feature, "DiscreteInput"):
                    ff[features2idx[key][0]] = val
            out.append(ff)
        return out

    def has_sufficient_experiments(self) -> bool:
        if self.experiments is None:
            return False
        return len(self.domain.outputs.preprocess_experiments_all_valid_outputs(experiments=self.experiments)) > 1

    def get_acqf_input_tensors(self):
        """botorch.py:696-724."""
        assert self.experiments is not None
        experiments = self.domain.outputs.preprocess_experiments_all_valid_outputs(self.experiments)
        clean = experiments.drop_duplicates(subset=self.domain.inputs.get_keys(), keep="first", inplace=False)
        X_train = torch.from_numpy(self.domain.inputs.transform(clean, self.input_preprocessing_specs).values.astype(np.float64))
        X_pending = None
        if self.candidates is not None:
            X_pending = torch.from_numpy(self.domain.inputs.transform(self.candidates, self.input_preprocessing_specs).values.astype(np.float64))
        return X_train, X_pending

    def _constraints(self) -> Optional[List[Ob.OutputConstraint]]:
        cons = get_output_constraints(self.domain.outputs, self.experiments)
        return cons or None


class _RandomSpec:
    """The few fields of the RandomStrategy data model RandomStrategy reads (botorch.py:250-256 builds it with defaults)."""

    def __init__(self, domain, seed=None):
        self.domain, self.seed = domain, seed
        self.num_base_samples, self.max_iters = None, 1000
        self.fallback_sampling_method = _sampling_enum("UNIFORM")
        self.n_burnin, self.n_thinning = 1000, 32


def _sampling_enum(name: str):
    from bofire.data_models.enum import SamplingMethodEnum

    return SamplingMethodEnum[name]


class _MultiObjectiveBase(BotorchStrategy):
    def __init__(self, data_model, **kw):
        super().__init__(data_model, **kw)
        self.ref_point = data_model.ref_point
        self.ref_point_mask = get_ref_point_mask(self.domain)

    def _get_objective(self) -> Ob.MultiObjective:
        assert self.experiments is not None, "No experiments available."
        return get_multiobjective_objective(self.domain.outputs, self.experiments)

    def get_adjusted_refpoint(self) -> List[float]:
        """qehvi.py:87-110 / mobo.py:93-124; the inferred point is utils/multiobjective.py:133-159."""
        assert self.experiments is not None, "No experiments available."
        keys = _mo_keys(self.domain)
        if self.ref_point is None:
            df = self.domain.outputs.preprocess_experiments_all_valid_outputs(self.experiments, output_feature_keys=keys)
            obj = self._get_objective()
            Y = torch.from_numpy(df[self.domain.outputs.get_keys()].values.astype(np.float64))
            ref = obj(Y, None).numpy().min(axis=0) / self.ref_point_mask
            ref_point = dict(zip(keys, ref))
        else:
            ref_point = self.ref_point
        return (self.ref_point_mask * np.array([ref_point[k] for k in _non_constrained_objective_keys(self.domain)])).tolist()


class QehviStrategy(_MultiObjectiveBase):
    """qehvi.py:23-110."""

    def __init__(self, data_model, **kw):
        super().__init__(data_model, **kw)
        self.num_sobol_samples = data_model.num_sobol_samples

    def _get_acqfs(self, n: int) -> list:
        assert self.experiments is not None, "No experiments available."
        df = self.domain.outputs.preprocess_experiments_all_valid_outputs(self.experiments)
        keys = [f.key for f in self.domain.outputs.get() if f.objective is not None]
        train_obj = df[keys].values * self.ref_point_mask
        ref_point = self.get_adjusted_refpoint()
        better = (train_obj > ref_point).all(axis=-1)
        _, X_pending = self.get_acqf_input_tensors()
        assert self.model is not None
        return [A.qExpectedHypervolumeImprovement(self.model, ref_point, train_obj[better], self._get_objective(),
                                                  mc_samples=self.num_sobol_samples, X_pending=X_pending)]


class QnehviStrategy(QehviStrategy):
    """qnehvi.py:14-53."""

    def __init__(self, data_model, **kw):
        super().__init__(data_model, **kw)
        self.alpha = data_model.alpha

    def _get_acqfs(self, n: int) -> list:
        assert self.experiments is not None, "No experiments available."
        X_train, X_pending = self.get_acqf_input_tensors()
        assert self.model is not None
        return [A.qNoisyExpectedHypervolumeImprovement(
            self.model, self.get_adjusted_refpoint(), X_train, self._get_objective(), constraints=self._constraints(),
            prune_baseline=True, alpha=self.alpha, cache_root=True, X_pending=X_pending, mc_samples=self.num_sobol_samples)]


class MoboStrategy(_MultiObjectiveBase):
    """mobo.py:31-124."""

    def __init__(self, data_model, **kw):
        super().__init__(data_model, **kw)
        self.acquisition_function = data_model.acquisition_function

    def _get_acqfs(self, n: int) -> list:
        assert self.is_fitted is True, "Model not trained."
        assert self.experiments is not None, "No experiments available."
        X_train, X_pending = self.get_acqf_input_tensors()
        af = self.acquisition_function
        name = _name(af)
        Y = None
        if name in ("qLogEHVI", "qEHVI"):
            Y = self.domain.outputs.preprocess_experiments_all_valid_outputs(self.experiments)[self.domain.outputs.get_keys()].values
        assert self.model is not None
        return [A.get_acquisition_function(
            name, self.model, self._get_objective(), X_train, X_pending=X_pending, constraints=self._constraints(),
            ref_point=self.get_adjusted_refpoint(), mc_samples=af.n_mc_samples, alpha=af.alpha, cache_root=True,
            prune_baseline=af.prune_baseline if name in ("qLogNEHVI", "qNEHVI") else True, Y=Y)]


class SoboStrategy(BotorchStrategy):
    """sobo.py:42-152."""

    def __init__(self, data_model, **kw):
        super().__init__(data_model, **kw)
        self.acquisition_function = data_model.acquisition_function

    def _get_acqfs(self, n: int) -> list:
        assert self.is_fitted is True, "Model not trained."
        X_train, X_pending = self.get_acqf_input_tensors()
        objective, constraints = self._get_objective_and_constraints()
        af = self.acquisition_function
        name = _name(af)
        assert self.model is not None
        return [A.get_acquisition_function(
            name, self.model, objective, X_train, X_pending=X_pending, constraints=constraints, mc_samples=af.n_mc_samples,
            beta=af.beta if name == "qUCB" else 0.2, tau=af.tau if name == "qPI" else 1e-3, cache_root=True,
            prune_baseline=af.prune_baseline if name in ("qNEI", "qLogNEI") else True)]

    def _reject_constrained_sr_ucb(self, constraints):
        if constraints and _name(self.acquisition_function) in ("qSR", "qUCB"):
            raise NotImplementedError("qSR / qUCB with output constraints (ConstrainedMCObjective with an infeasible cost) are "
                                      "outside the accelerated path")

    def _get_objective_and_constraints(self):
        assert self.experiments is not None, "No experiments available."
        outs = self.domain.outputs
        with_obj = [f for f in outs.get() if f.objective is not None]
        plain = [f for f in with_obj if not _is(f.objective, "ConstrainedObjective")]
        target = plain[0] if plain else with_obj[0]
        idx = outs.get_keys().index(target.key)
        op = objective_spec(idx, target.objective, _x_adapt(outs, target, self.experiments))
        constrained = [f for f in with_obj if _is(f.objective, "ConstrainedObjective")]
        constraints = get_output_constraints(outs, self.experiments) if (constrained and len(with_obj) > 1) else None
        self._reject_constrained_sr_ucb(constraints)
        return Ob.ScalarObjective([op], "single"), constraints or None


class AdditiveSoboStrategy(SoboStrategy):
    """sobo.py:155-224."""

    def __init__(self, data_model, **kw):
        super().__init__(data_model, **kw)
        self.use_output_constraints = data_model.use_output_constraints

    def _get_objective_and_constraints(self):
        assert self.experiments is not None, "No experiments available."
        outs = self.domain.outputs
        with_obj = [f for f in outs.get() if f.objective is not None]
        constrained = [f for f in with_obj if _is(f.objective, "ConstrainedObjective")]
        constraints = None
        if constrained and len(with_obj) > 1 and self.use_output_constraints:
            constraints = get_output_constraints(outs, self.experiments)
        self._reject_constrained_sr_ucb(constraints)
        ops = _callables_and_weights(outs, self.experiments, exclude_constraints=bool(self.use_output_constraints))
        return Ob.ScalarObjective(ops, "additive"), constraints or None


class MultiplicativeSoboStrategy(SoboStrategy):
    """sobo.py:227-256."""

    def _get_objective_and_constraints(self):
        assert self.experiments is not None, "No experiments available."
        ops = _callables_and_weights(self.domain.outputs, self.experiments, adapt_weights_to_1_inf=True)
        return Ob.ScalarObjective(ops, "multiplicative"), None


STRATEGY_MAP = {
    "RandomStrategy": RandomStrategy,
    "QehviStrategy": QehviStrategy,
    "QnehviStrategy": QnehviStrategy,
    "MoboStrategy": MoboStrategy,
    "SoboStrategy": SoboStrategy,
    "AdditiveSoboStrategy": AdditiveSoboStrategy,
    "MultiplicativeSoboStrategy": MultiplicativeSoboStrategy,
}


def map(data_model, **kwargs) -> Strategy:
    """bofire.strategies.api.map (strategies/mapper.py) for the strategies on the accelerated path."""
    cls = STRATEGY_MAP.get(_name(data_model))
    if cls is None:
        raise NotImplementedError(f"strategy {_name(data_model)} is outside the accelerated path "
                                  f"(supported: {', '.join(STRATEGY_MAP)})")
    return cls(data_model, **kwargs)

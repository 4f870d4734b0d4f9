"""Hyper-parameter fit of one exact GP on the device: host-side mirror of what `SingleTaskGPSurrogate._fit` does through
`fit_gpytorch_mll(ExactMarginalLogLikelihood(...))` (surrogates/single_task_gp.py:39-71; kernels/mapper.py:31-69 for the
lengthscale priors, priors/mapper.py:9-61, data_models/priors/api.py:30-51 for the prior families and the Hvarfner
defaults) -- SURVEY.md 8f-2, the step before the acquisition path.

The objective and its gradient come from `bo_mll_forward_backward` (csrc/mll.cu: Cholesky, inverse root, explicit inverse
and the pairwise derivative reduction all on the device); the host adds the prior terms, the parameter transforms and
runs scipy's L-BFGS-B like [UPSTREAM] botorch.optim.fit.fit_gpytorch_mll_scipy:

    loss(raw) = -( mll(theta(raw)) + sum_p log prior_p(theta_p) ) / N

[UPSTREAM] parameterisation (gpytorch constraints): lengthscale = softplus(raw) (Positive), outputscale = softplus(raw)
(Positive), noise = 1e-4 + softplus(raw) (GreaterThan(1e-4)), constant mean unconstrained.  Starting point: the median of
each parameter's prior (exp(loc) for a log-normal, concentration / rate for a gamma), 1.0 / 0.0 without a prior --
BoTorch's own initial values are not visible from the reference, so the fitted optimum (not the path) is what can agree.
"""
import math
from dataclasses import dataclass, replace
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
from scipy.optimize import minimize

from . import kernels as K
from .model import DeviceGPState, SingleTaskGPSpec, standardize_stats


# ------------------------------------------------------------------------------------------------
# priors (priors/mapper.py:9-61 -> gpytorch.priors)
# ------------------------------------------------------------------------------------------------
@dataclass
class LogNormalPrior:
    loc: float
    scale: float

    def log_prob(self, x):
        return -math.log(x) - math.log(self.scale * math.sqrt(2 * math.pi)) - (math.log(x) - self.loc) ** 2 / (2 * self.scale ** 2)

    def dlog_prob(self, x):
        return -1.0 / x - (math.log(x) - self.loc) / (self.scale ** 2 * x)

    def start(self):
        return math.exp(self.loc)


@dataclass
class GammaPrior:
    concentration: float
    rate: float

    def log_prob(self, x):
        a, b = self.concentration, self.rate
        return a * math.log(b) + (a - 1.0) * math.log(x) - b * x - math.lgamma(a)

    def dlog_prob(self, x):
        return (self.concentration - 1.0) / x - self.rate

    def start(self):
        # [UPSTREAM] BoTorch initialises a Gamma-distributed hyper-parameter at the prior MODE where it exists (e.g. the
        # GammaPrior(1.1, 0.05) noise level of MixedSingleTaskGP starts at 2.0, not at its mean of 22, from where the
        # marginal likelihood is flat and the fit ends in the all-noise solution)
        if self.concentration > 1.0:
            return (self.concentration - 1.0) / self.rate
        return self.concentration / self.rate


@dataclass
class NormalPrior:
    loc: float
    scale: float

    def log_prob(self, x):
        return -math.log(self.scale * math.sqrt(2 * math.pi)) - (x - self.loc) ** 2 / (2 * self.scale ** 2)

    def dlog_prob(self, x):
        return -(x - self.loc) / self.scale ** 2

    def start(self):
        return self.loc


def DimensionalityScaledLogNormalPrior(d: int, loc: float = math.sqrt(2.0), loc_scaling: float = 0.5,
                                       scale: float = math.sqrt(3.0), scale_scaling: float = 0.0) -> LogNormalPrior:
    """map_DimensionalityScaledLogNormalPrior (priors/mapper.py:53-61; defaults data_models/priors/normal.py:37-49)."""
    return LogNormalPrior(loc + math.log(d) * loc_scaling, (scale ** 2 + math.log(d) * scale_scaling) ** 0.5)


def HVARFNER_NOISE_PRIOR() -> LogNormalPrior:          # data_models/priors/api.py:50
    return LogNormalPrior(-4.0, 1.0)


def THREESIX_LENGTHSCALE_PRIOR() -> GammaPrior:        # data_models/priors/api.py:30-32
    return GammaPrior(3.0, 6.0)


def THREESIX_NOISE_PRIOR() -> GammaPrior:
    return GammaPrior(1.1, 0.05)


def THREESIX_SCALE_PRIOR() -> GammaPrior:
    return GammaPrior(2.0, 0.15)


# ------------------------------------------------------------------------------------------------
# parameter vector <-> kernel tree
# ------------------------------------------------------------------------------------------------
def _softplus(x):
    return x if x > 30.0 else math.log1p(math.exp(x))


def _inv_softplus(y):
    return y if y > 30.0 else math.log(math.expm1(y))


def _sigmoid(x):
    return 1.0 / (1.0 + math.exp(-x))


class _Layout:
    """Walks the kernel tree in the order K.flatten does and records, for every leaf, where its lengthscales live in the
    natural-parameter vector, and for every flattened term which ScaleKernel nodes multiply into its coefficient."""

    def __init__(self, kernel):
        self.leaves: List[object] = []
        self.scales: List[K.ScaleKernel] = []
        self.terms: List[Tuple[List[int], List[int]]] = []      # (scale node ids, leaf ids) per flattened term
        self.terms = self._rec(kernel)
        self.ls_slices: List[Optional[slice]] = []
        n = 0
        for lf in self.leaves:
            if isinstance(lf, K.TanimotoKernel):
                self.ls_slices.append(None)
                continue
            k = len(lf.categorical_features) if isinstance(lf, K.HammingDistanceKernel) else len(lf.active_dims)
            self.ls_slices.append(slice(n, n + k))
            n += k
        self.n_ls = n

    def _rec(self, k):
        if isinstance(k, (K.RBFKernel, K.MaternKernel, K.HammingDistanceKernel, K.TanimotoKernel)):
            self.leaves.append(k)
            return [([], [len(self.leaves) - 1])]
        if isinstance(k, K.ScaleKernel):
            self.scales.append(k)
            sid = len(self.scales) - 1
            return [([sid] + s, f) for s, f in self._rec(k.base_kernel)]
        if isinstance(k, K.AdditiveKernel):
            out = []
            for c in k.kernels:
                out += self._rec(c)
            return out
        if isinstance(k, K.MultiplicativeKernel):
            acc = [([], [])]
            for c in k.kernels:
                terms = self._rec(c)
                acc = [(s1 + s2, f1 + f2) for s1, f1 in acc for s2, f2 in terms]
            return acc
        raise NotImplementedError(type(k).__name__)

    def n_free_ls(self, leaf_idx):
        """Free lengthscale parameters of a leaf: 1 when isotropic (one value broadcast), else one per slot."""
        lf = self.leaves[leaf_idx]
        n = len(list(lf.lengthscale))
        if isinstance(lf, K.HammingDistanceKernel) and n > 1:
            return len(lf.categorical_features)   # an ARD vector over the one-hot columns: the first F entries are used
        return n


def _rebuild(kernel, ls_values: Dict[int, List[float]], scale_values: List[float], counters=None):
    """Copy of the tree with new lengthscales (per leaf, in flatten order) and outputscales (per ScaleKernel, in order)."""
    counters = counters if counters is not None else {"leaf": 0, "scale": 0}
    if isinstance(kernel, (K.RBFKernel, K.MaternKernel, K.HammingDistanceKernel, K.TanimotoKernel)):
        i = counters["leaf"]
        counters["leaf"] += 1
        if isinstance(kernel, K.TanimotoKernel):
            return kernel
        return replace(kernel, lengthscale=list(ls_values[i]))
    if isinstance(kernel, K.ScaleKernel):
        i = counters["scale"]
        counters["scale"] += 1
        return K.ScaleKernel(_rebuild(kernel.base_kernel, ls_values, scale_values, counters), float(scale_values[i]))
    if isinstance(kernel, K.AdditiveKernel):
        return K.AdditiveKernel([_rebuild(c, ls_values, scale_values, counters) for c in kernel.kernels])
    if isinstance(kernel, K.MultiplicativeKernel):
        return K.MultiplicativeKernel([_rebuild(c, ls_values, scale_values, counters) for c in kernel.kernels])
    raise NotImplementedError(type(kernel).__name__)


@dataclass
class FitResult:
    spec: SingleTaskGPSpec
    loss: float                 # -(mll + log priors) / N at the optimum
    mll: float                  # log marginal likelihood (sum over the N points)
    n_iterations: int
    n_evaluations: int
    message: str


def mll_and_grad(X, spec: SingleTaskGPSpec, device=None, state: Optional[DeviceGPState] = None):
    """(mll, d/d noise, d/d mean_const, d/d lengthscale slots, d/d term coefficients) from the device for the
    hyper-parameters in `spec`.  Raises NotPSDError when the training Gram matrix cannot be factorised.
    `state`: a DeviceGPState on the same (X, y, kernel tree) whose hyper-parameters are overwritten in place
    (set_hyperparameters) instead of building and destroying a state per evaluation."""
    import ctypes as C

    from . import _lib as L

    own = state is None
    st = DeviceGPState(X, [spec], device=device) if own else state.set_hyperparameters(0, spec)
    try:
        st.factorize()
        lay = _Layout(spec.kernel)
        n_terms = len(K.flatten(spec.kernel).terms)
        mll, dn, dm = C.c_double(0), C.c_double(0), C.c_double(0)
        dls = (C.c_double * max(lay.n_ls, 1))()
        dco = (C.c_double * n_terms)()
        with torch.cuda.device(st.device):
            L.check(st.lib.bo_mll_forward_backward(st.handle, 0, C.byref(mll), C.byref(dn), C.byref(dm), dls, lay.n_ls, dco,
                                                   n_terms, C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        return mll.value, dn.value, dm.value, np.array(dls[: lay.n_ls]), np.array(dco[:n_terms])
    finally:
        if own:
            st.close()


def fit_gp(X, y, kernel, in_offset=None, in_scale=None, noise_prior=None, lengthscale_priors=None, outputscale_priors=None,
           min_noise: float = 1e-4, fit_mean: bool = True, options: Optional[dict] = None, device=None) -> FitResult:
    """Fit lengthscales, outputscales, noise and constant mean of `kernel` on (X, y).

    kernel               spec tree whose lengthscale lists give the parameter SHAPES (1 value = isotropic, one per active dim
                         = ARD) -- the values are ignored unless no prior provides a starting point
    lengthscale_priors   {leaf index in flatten order: prior} (kernels/mapper.py passes `lengthscale_prior` per leaf)
    outputscale_priors   {ScaleKernel index in tree order: prior} (mapper.py:168-188)
    noise_prior          likelihood.noise_covar.noise_prior (single_task_gp.py:68); default none
    options              scipy L-BFGS-B options, e.g. {"maxiter": 200} (`training_specs`)
    """
    X = np.ascontiguousarray(np.asarray(X, dtype=np.float64))
    y = np.ascontiguousarray(np.asarray(y, dtype=np.float64)).reshape(-1)
    N = X.shape[0]
    y_mean, y_std = standardize_stats(y)
    lay = _Layout(kernel)
    lengthscale_priors = lengthscale_priors or {}
    outputscale_priors = outputscale_priors or {}
    # free parameters: [lengthscales of every leaf (as given: 1 or per-dim)] + [outputscales] + [noise] + [mean]
    blocks = []   # (kind, index, count)
    for i, lf in enumerate(lay.leaves):
        if lay.ls_slices[i] is not None:
            blocks.append(("ls", i, lay.n_free_ls(i)))
    for i in range(len(lay.scales)):
        blocks.append(("os", i, 1))
    blocks.append(("noise", 0, 1))
    if fit_mean:
        blocks.append(("mean", 0, 1))
    n_raw = sum(b[2] for b in blocks)

    def start_value(prior, fallback):
        return float(prior.start()) if prior is not None else float(fallback)

    raw0 = []
    for kind, i, cnt in blocks:
        if kind == "ls":
            v = start_value(lengthscale_priors.get(i), 1.0)
            raw0 += [_inv_softplus(v)] * cnt
        elif kind == "os":
            raw0.append(_inv_softplus(start_value(outputscale_priors.get(i), 1.0)))
        elif kind == "noise":
            raw0.append(_inv_softplus(max(start_value(noise_prior, 1e-2) - min_noise, 1e-6)))
        else:
            raw0.append(0.0)
    raw0 = np.asarray(raw0, dtype=np.float64)
    state = {"n_eval": 0, "last_mll": float("nan"), "st": None}

    def unpack(raw):
        ls_values, scale_values, pos = {}, [], 0
        noise, mean = None, 0.0
        for kind, i, cnt in blocks:
            seg = raw[pos:pos + cnt]
            pos += cnt
            if kind == "ls":
                ls_values[i] = [_softplus(float(v)) for v in seg]
            elif kind == "os":
                scale_values.append(_softplus(float(seg[0])))
            elif kind == "noise":
                noise = min_noise + _softplus(float(seg[0]))
            else:
                mean = float(seg[0])
        return ls_values, scale_values, noise, mean

    def make_spec(raw):
        ls_values, scale_values, noise, mean = unpack(raw)
        return SingleTaskGPSpec(kernel=_rebuild(kernel, ls_values, scale_values), y=y, in_offset=in_offset, in_scale=in_scale,
                                mean_const=mean, noise=noise, y_mean=y_mean, y_std=y_std)

    def loss_and_grad(raw):
        from ._lib import NotPSDError

        state["n_eval"] += 1
        ls_values, scale_values, noise, mean = unpack(raw)
        try:
            spec_k = make_spec(raw)
            if state["st"] is None:     # one device state for the whole fit: later evaluations only rewrite its hyper-parameters
                state["st"] = DeviceGPState(X, [spec_k], device=device)
            mll, d_noise, d_mean, d_ls, d_coef = mll_and_grad(X, spec_k, device=device, state=state["st"])
        except NotPSDError:
            return 1e10, np.zeros(n_raw)
        if not math.isfinite(mll):
            return 1e10, np.zeros(n_raw)
        state["last_mll"] = mll
        total = mll
        grad = np.zeros(n_raw)
        # d mll / d outputscale_s = sum over the flattened terms that contain the node: d_coef_t * coef_t / outputscale_s
        coefs = [float(np.prod([scale_values[s] for s in sids])) if sids else 1.0 for sids, _ in lay.terms]
        pos = 0
        for kind, i, cnt in blocks:
            if kind == "ls":
                sl = lay.ls_slices[i]
                g_slots = d_ls[sl]
                vals = ls_values[i]
                g_nat = [float(g_slots.sum())] if cnt == 1 and len(g_slots) != 1 else [float(v) for v in g_slots[:cnt]]
                prior = lengthscale_priors.get(i)
                for k in range(cnt):
                    gk = g_nat[k]
                    if prior is not None:
                        total += prior.log_prob(vals[k])
                        gk += prior.dlog_prob(vals[k])
                    grad[pos + k] = gk * _sigmoid(float(raw[pos + k]))     # d softplus / d raw
            elif kind == "os":
                v = scale_values[i]
                gk = sum(d_coef[t] * coefs[t] / v for t, (sids, _) in enumerate(lay.terms) if i in sids)
                prior = outputscale_priors.get(i)
                if prior is not None:
                    total += prior.log_prob(v)
                    gk += prior.dlog_prob(v)
                grad[pos] = gk * _sigmoid(float(raw[pos]))
            elif kind == "noise":
                gk = d_noise
                if noise_prior is not None:
                    total += noise_prior.log_prob(noise)
                    gk += noise_prior.dlog_prob(noise)
                grad[pos] = gk * _sigmoid(float(raw[pos]))
            else:
                grad[pos] = d_mean
            pos += cnt
        return -total / N, -grad / N

    try:
        res = minimize(loss_and_grad, raw0, jac=True, method="L-BFGS-B", options=dict(options or {"maxiter": 200}))
    finally:
        if state["st"] is not None:
            state["st"].close()
    best = res.x if math.isfinite(res.fun) and res.fun < 1e9 else raw0
    spec = make_spec(best)
    return FitResult(spec=spec, loss=float(res.fun), mll=float(state["last_mll"]), n_iterations=int(res.nit),
                     n_evaluations=int(state["n_eval"]), message=str(res.message))


def single_task_gp_factory(kernel_factory, in_offset=None, in_scale=None, noise_prior=None, lengthscale_prior_factory=None,
                           options: Optional[dict] = None):
    """`surrogate_factory` for everest_b200.strategy: fits one SingleTaskGP per output column like
    BotorchSurrogates.fit does for a list of SingleTaskGPSurrogates (surrogates/botorch_surrogates.py:43-78).
    kernel_factory(d) -> kernel tree; lengthscale_prior_factory(d_leaf) -> prior (default: Hvarfner, the BoFire default of
    SingleTaskGPSurrogate, data_models/surrogates/single_task_gp.py:109-115)."""
    def make(X, Y):
        d = X.shape[1]
        specs = []
        for m in range(Y.shape[1]):
            kern = kernel_factory(d)
            lay = _Layout(kern)
            lpf = lengthscale_prior_factory or (lambda n: DimensionalityScaledLogNormalPrior(n))
            lps = {i: lpf(len(lf.active_dims)) for i, lf in enumerate(lay.leaves) if isinstance(lf, (K.RBFKernel, K.MaternKernel))}
            res = fit_gp(X, Y[:, m], kern, in_offset=in_offset, in_scale=in_scale,
                         noise_prior=noise_prior if noise_prior is not None else HVARFNER_NOISE_PRIOR(),
                         lengthscale_priors=lps, options=options)
            specs.append(res.spec)
        return specs
    return make

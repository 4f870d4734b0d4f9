"""Device-resident GP state: what ``BotorchSurrogates.compatibilize`` (surrogates/botorch_surrogates.py:79-128)
hands to the acquisition function in the reference -- a ModelListGP of independent single-output exact GPs,
each with a Normalize input transform, a Standardize outcome transform, a constant mean and a kernel tree
(surrogates/single_task_gp.py:39-71, mixed_single_task_gp.py:46-112, mixed_tanimoto_gp.py:43-224).

Hyper-parameter fitting (fit_gpytorch_mll) is outside this path: the state takes fitted values.
"""
import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _lib as L
from . import kernels as K


@dataclass
class SingleTaskGPSpec:
    """One output of the model list.

    kernel       kernel tree on the columns of the (BoFire-transformed) input
    y            [N] raw training targets
    in_offset/in_scale  Normalize transform on the ordinal columns: x' = (x - offset) / scale
                 (surrogates/utils.py:144-154: bounds = feature bounds U data range); None = identity
    mean_const   gpytorch ConstantMean value (standardised space)
    noise        likelihood noise variance (standardised space)
    y_mean/y_std Standardize(m=1) statistics; computed from y when omitted (unbiased std, floor 1e-8)
    """
    kernel: K.AnyKernel
    y: Sequence[float]
    in_offset: Optional[Sequence[float]] = None
    in_scale: Optional[Sequence[float]] = None
    mean_const: float = 0.0
    noise: float = 1e-4
    y_mean: Optional[float] = None
    y_std: Optional[float] = None


def standardize_stats(y):
    y = np.asarray(y, dtype=np.float64)
    s = float(y.std(ddof=1)) if y.size > 1 else 1.0
    if not (s >= 1e-8):
        s = 1.0
    return float(y.mean()), s


def normalize_bounds(X, lower=None, upper=None, columns=None):
    """Normalize(d, bounds, indices) as built by get_scaler (surrogates/utils.py:103-164): bounds are the
    union of the feature bounds and the data range; only `columns` are transformed."""
    X = np.asarray(X, dtype=np.float64)
    d = X.shape[1]
    off, scl = np.zeros(d), np.ones(d)
    cols = range(d) if columns is None else columns
    for j in cols:
        lo = X[:, j].min() if lower is None else min(lower[j], X[:, j].min())
        hi = X[:, j].max() if upper is None else max(upper[j], X[:, j].max())
        off[j] = lo
        scl[j] = (hi - lo) if hi > lo else 1.0
    return off, scl


def _dev_ptr(t: torch.Tensor):
    return C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class DeviceGPState:
    """Owns the bo_state handle.  One handle per strategy instance; calls are serialised per handle."""

    def __init__(self, X_train, outputs: List[SingleTaskGPSpec], device=None):
        if not torch.cuda.is_available():
            raise L.EverestError("everest_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.lib = L.load()
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        X = np.ascontiguousarray(np.asarray(X_train, dtype=np.float64))
        if X.ndim != 2:
            raise ValueError("X_train must be [N, d]")
        self.N, self.d = X.shape
        self.M = len(outputs)
        self.outputs = outputs
        keep = []
        outs = (L.OutputModel * self.M)()
        for m, spec in enumerate(outputs):
            self._fill_output(outs[m], m, spec, keep)
        cfg = L.StateConfig(self.N, self.d, self.M, X.ctypes.data_as(L.c_double_p), outs)
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            L.check(self.lib.bo_state_create(C.byref(cfg), C.byref(h)))
        self._h = h
        self.jitter = None
        self.factorized = False

    def _fill_output(self, o, m, spec, keep):
        """SingleTaskGPSpec -> bo_output_model; `keep` collects the arrays the struct points into."""
        flat = K.flatten(spec.kernel)
        leaves = (L.KernelLeaf * len(flat.leaves))(*[K.leaf_to_c(lf, keep) for lf in flat.leaves])
        terms = (L.KernelTerm * len(flat.terms))()
        for t, (coef, facs) in enumerate(flat.terms):
            terms[t].coef = float(coef)
            terms[t].n_factors = len(facs)
            for i, f in enumerate(facs):
                terms[t].factors[i] = int(f)
        y = np.ascontiguousarray(np.asarray(spec.y, dtype=np.float64))
        if y.shape != (self.N,):
            raise ValueError(f"output {m}: y must have shape [{self.N}]")
        ym, ys = (spec.y_mean, spec.y_std) if spec.y_mean is not None and spec.y_std is not None else standardize_stats(y)
        off = np.ascontiguousarray(np.zeros(self.d) if spec.in_offset is None else np.asarray(spec.in_offset, dtype=np.float64))
        scl = np.ascontiguousarray(np.ones(self.d) if spec.in_scale is None else np.asarray(spec.in_scale, dtype=np.float64))
        keep += [leaves, terms, y, off, scl]
        o.n_leaves, o.leaves, o.n_terms, o.terms = len(flat.leaves), leaves, len(flat.terms), terms
        o.in_offset = off.ctypes.data_as(L.c_double_p)
        o.in_scale = scl.ctypes.data_as(L.c_double_p)
        o.mean_const, o.noise, o.y_mean, o.y_std = float(spec.mean_const), float(spec.noise), float(ym), float(ys)
        o.y = y.ctypes.data_as(L.c_double_p)

    def set_hyperparameters(self, m: int, spec: SingleTaskGPSpec):
        """New hyper-parameter VALUES for output m (same kernel tree / columns / transforms / targets) without rebuilding the
        state: bo_state_set_hyperparameters.  `factorize()` must follow.  This is what one step of the marginal-likelihood
        optimiser changes (fit.py); every large device buffer of the handle is reused."""
        keep = []
        o = L.OutputModel()
        self._fill_output(o, m, spec, keep)
        with torch.cuda.device(self.device):
            L.check(self.lib.bo_state_set_hyperparameters(self.handle, int(m), C.byref(o), _stream()))
        self.outputs = list(self.outputs)
        self.outputs[m] = spec
        self.factorized = False
        self._active_acqf = None
        return self

    # -- lifetime ------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self.lib.bo_state_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        if not self._h:
            raise L.EverestError("state already destroyed")
        return self._h

    # -- factorisation -------------------------------------------------------------------------
    def factorize(self):
        info = (C.c_int32 * self.M)()
        jit = (C.c_double * self.M)()
        with torch.cuda.device(self.device):
            L.check(self.lib.bo_state_factorize(self.handle, info, jit, _stream()))
        self.jitter = list(jit)
        self.factorized = True
        return self

    # -- posterior -----------------------------------------------------------------------------
    def _as_dev(self, X):
        X = torch.as_tensor(X, dtype=torch.double)
        return X.to(self.device).contiguous()

    def posterior(self, X, observation_noise=False):
        """model.posterior(X).mean / .variance, [n, M] each (botorch.py:174-194)."""
        Xd = self._as_dev(X)
        n = Xd.shape[0]
        mean = torch.empty(n, self.M, dtype=torch.double, device=self.device)
        var = torch.empty(n, self.M, dtype=torch.double, device=self.device)
        with torch.cuda.device(self.device):
            L.check(self.lib.bo_posterior_marginal(self.handle, _dev_ptr(Xd), n, int(bool(observation_noise)),
                                                   _dev_ptr(mean), _dev_ptr(var), _stream()))
        return mean, var

    def posterior_joint(self, X):
        """mean [n, M], covariance [M, n, n] of one point set."""
        Xd = self._as_dev(X)
        n = Xd.shape[0]
        mean = torch.empty(n, self.M, dtype=torch.double, device=self.device)
        cov = torch.empty(self.M, n, n, dtype=torch.double, device=self.device)
        with torch.cuda.device(self.device):
            L.check(self.lib.bo_posterior_joint(self.handle, _dev_ptr(Xd), n, _dev_ptr(mean), _dev_ptr(cov), _stream()))
        return mean, cov

    def predict(self, X):
        """BotorchStrategy._predict: (preds, stds) as numpy, observation noise included."""
        mean, var = self.posterior(X, observation_noise=True)
        return mean.cpu().numpy(), np.sqrt(var.cpu().numpy())

    # -- introspection (tests) -----------------------------------------------------------------
    def debug_get(self, name, m=0, capacity=None, dtype=torch.double):
        cap = int(capacity if capacity is not None else max(self.N * (self.N + 16), 1 << 16))
        buf = torch.empty(cap, dtype=torch.double, device=self.device)
        n = C.c_int64(0)
        with torch.cuda.device(self.device):
            L.check(self.lib.bo_debug_get(self.handle, name.encode(), int(m), _dev_ptr(buf), cap, C.byref(n), _stream()))
        torch.cuda.synchronize(self.device)
        if dtype == torch.int32:
            return buf.view(torch.int32)[: n.value].clone()
        return buf[: n.value].clone()

    def launch_count(self):
        return int(self.lib.bo_launch_count(self.handle))

    def set_timing(self, enabled=True):
        L.check(self.lib.bo_set_timing(self.handle, int(enabled)))

    def last_timing(self, name):
        ms = C.c_double(0.0)
        cnt = self.lib.bo_last_timing(self.handle, name.encode(), C.byref(ms))
        if cnt < 0:
            L.check(cnt)
        return ms.value, cnt

    @classmethod
    def from_botorch(cls, model, device=None, X_train=None):
        """Rebuild the device state from a fitted BoTorch ModelListGP / SingleTaskGP / MixedSingleTaskGP (hyper-parameters,
        transforms, train data), the hand-over point of BotorchSurrogates.compatibilize.  See INTEGRATION.md."""
        from .bofire_adapter import state_from_botorch_model

        return state_from_botorch_model(model, device=device, X_train=X_train)

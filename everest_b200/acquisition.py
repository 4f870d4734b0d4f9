"""Acquisition functions with the constructor / forward contract BoFire relies on.

Reference call sites (arguments mirrored here):
  * qNoisyExpectedHypervolumeImprovement(model, ref_point, X_baseline, prune_baseline, objective,
    cache_root, X_pending, constraints, eta, alpha)          strategies/predictives/qnehvi.py:39-52, mobo.py:72-90
  * qExpectedHypervolumeImprovement(model, ref_point, partitioning, objective, X_pending)   qehvi.py:67-76
  * get_acquisition_function("qLogEI", model, objective, X_observed, ..., mc_samples)       sobo.py:64-89
forward(X[b, q, d]) -> [b] is the L1 boundary of SURVEY.md 8b (called from calc_acquisition
botorch.py:223, optimize_acqf's raw-sample screen and optimize_acqf_discrete botorch.py:461).
"""
import ctypes as C
import os
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _lib as L
from . import sampling
from .model import DeviceGPState, _dev_ptr, _stream
from .objectives import MultiObjective, ObjectiveSpec, OutputConstraint, ScalarObjective, c_array


class _DeviceAcquisition:
    """Shared forward plumbing: X -> device, base samples of the q new points, bo_acqf_forward."""

    def __init__(self, model: DeviceGPState, mc_samples: int, seed: Optional[int]):
        if not model.factorized:
            model.factorize()
        self.model = model
        self.S = int(mc_samples)
        # BoTorch draws the sampler seed from torch's global RNG (seeded by BotorchStrategy, botorch.py:86)
        self.seed = int(torch.randint(0, 1000000, (1,)).item()) if seed is None else int(seed)
        self._zq = {}
        self.nb = 0
        self.last_info = None
        self.X_pending = None

    # [UPSTREAM] botorch's @concatenate_pending_points: acquisition functions without a cached baseline score the joint
    # batch cat(X, X_pending); the NEHVI family overrides set_X_pending and folds the points into its baseline instead
    _PENDING_CONCAT = True
    _force_fallback = False    # tests: re-score every q-batch through the fallback
    last_resampled = 0
    nb = 0

    def set_X_pending(self, X_pending=None):
        if X_pending is None:
            self.X_pending = None
        else:
            self.X_pending = torch.as_tensor(X_pending, dtype=torch.double).reshape(-1, self.model.d).to(self.model.device)

    def _with_pending(self, Xd):
        if self._PENDING_CONCAT and self.X_pending is not None and self.X_pending.shape[0] > 0:
            return torch.cat([Xd, self.X_pending.unsqueeze(0).expand(Xd.shape[0], -1, -1)], dim=1).contiguous()
        return Xd

    def _claim(self):
        """The handle holds ONE prepared acquisition function: remember which object prepared it last."""
        self.model._active_acqf = self

    def activate(self):
        """Make this the prepared acquisition function of its DeviceGPState again (the handle holds one at a time):
        optimize_acqf_list walks a list of acquisition functions built on the same model.  Re-runs the prepare call with
        the stored arguments (same base samples, same pruned baseline): results are identical to the first preparation."""
        if getattr(self.model, "_active_acqf", None) is not self:
            self._reprepare()
        return self

    def _reprepare(self):
        raise NotImplementedError

    def _check_active(self):
        if getattr(self.model, "_active_acqf", None) is not self:
            raise L.EverestError("another acquisition function has been prepared on this DeviceGPState since this one "
                                 "was built; rebuild it (one prepared acquisition function per state)")

    def set_option(self, name: str, value: float):
        self._check_active()
        L.check(self.model.lib.bo_acqf_set_option(self.model.handle, name.encode(), float(value)))

    def base_samples_q(self, q: int) -> torch.Tensor:
        """[S, q, M] base samples of the new points: the last q points of a fresh (n_b + q)-point Sobol
        draw with the sampler's seed ([UPSTREAM] NormalMCSampler._update_base_samples)."""
        if q not in self._zq:
            z = sampling.base_samples_device(self.nb + q, self.model.M, self.S, self.seed, self.model.device)
            self._zq[q] = z[:, self.nb:, :].contiguous()
        return self._zq[q]

    def set_base_samples_q(self, q: int, zq):
        zq = torch.as_tensor(zq, dtype=torch.double)
        if tuple(zq.shape) != (self.S, q, self.model.M):
            raise ValueError(f"zq must be [{self.S}, {q}, {self.model.M}]")
        self._zq[q] = zq.to(self.model.device).contiguous()

    def forward(self, X) -> torch.Tensor:
        X = torch.as_tensor(X, dtype=torch.double)
        if X.dim() == 2:
            X = X.unsqueeze(0)
        if X.dim() != 3 or X.shape[-1] != self.model.d:
            raise ValueError(f"X must be [b, q, {self.model.d}]")
        self._check_active()
        on_cpu = X.device.type == "cpu"
        Xd = self._with_pending(X.to(self.model.device).contiguous())
        b, q, _ = Xd.shape
        out = torch.empty(b, dtype=torch.double, device=self.model.device)
        info = torch.zeros(b, dtype=torch.int32, device=self.model.device)
        if b == 0:                       # empty t-batch: BoTorch returns an empty tensor
            self.last_info = info
            return out.cpu() if on_cpu else out
        zq = self.base_samples_q(q)
        with torch.cuda.device(self.model.device):
            L.check(self.model.lib.bo_acqf_forward(self.model.handle, _dev_ptr(Xd), b, q, _dev_ptr(zq), _dev_ptr(out),
                                                   _dev_ptr(info), _stream()))
        self.last_info = info
        # [UPSTREAM] sample_cached_cholesky: q-batches whose q x q conditional root exhausted the jitter ladder are re-scored
        # from the joint posterior over (X_baseline, X).  bo_acqf_forward does it by itself (option "joint_fallback"): the
        # counter of exhausted ladders is read while the MC kernels still run.  The explicit call below is the test hook that
        # pushes EVERY q-batch through the fallback.
        self.last_resampled = int(self.model.lib.bo_acqf_last_resampled(self.model.handle))
        if self._force_fallback:
            n = C.c_int32(0)
            with torch.cuda.device(self.model.device):
                L.check(self.model.lib.bo_acqf_resample_flagged(self.model.handle, _dev_ptr(Xd), b, q, _dev_ptr(zq), _dev_ptr(out),
                                                                _dev_ptr(info), C.byref(n), _stream()))
            self.last_resampled = int(n.value)
        return out.cpu() if on_cpu else out

    def __call__(self, X):
        """forward(X); when X requires grad the result carries the analytic gradient (BoTorch's acqfs are
        differentiable torch modules: gen_candidates_scipy back-propagates through them)."""
        if isinstance(X, torch.Tensor) and X.requires_grad and torch.is_grad_enabled():
            return _AcqfAutograd.apply(X, self)
        return self.forward(X)

    def forward_backward(self, X):
        """(values [b], d values[i] / d X[i] as [b, q, d]) through bo_acqf_forward_backward: the analytic adjoint of
        the whole chain HVI / LogEI -> MC samples -> conditional root -> posterior -> kernel."""
        X = torch.as_tensor(X, dtype=torch.double)
        if X.dim() == 2:
            X = X.unsqueeze(0)
        if X.dim() != 3 or X.shape[-1] != self.model.d:
            raise ValueError(f"X must be [b, q, {self.model.d}]")
        self._check_active()
        on_cpu = X.device.type == "cpu"
        q_in = X.shape[1]
        Xd = self._with_pending(X.detach().to(self.model.device).contiguous())
        b, q, _ = Xd.shape
        out = torch.empty(b, dtype=torch.double, device=self.model.device)
        dX = torch.empty_like(Xd)
        info = torch.zeros(b, dtype=torch.int32, device=self.model.device)
        if b == 0:
            self.last_info = info
            dX = dX[:, :q_in].contiguous()
            return (out.cpu(), dX.cpu()) if on_cpu else (out, dX)
        zq = self.base_samples_q(q)
        with torch.cuda.device(self.model.device):
            L.check(self.model.lib.bo_acqf_forward_backward(self.model.handle, _dev_ptr(Xd), b, q, _dev_ptr(zq),
                                                            _dev_ptr(out), _dev_ptr(dX), _dev_ptr(info), _stream()))
        self.last_info = info
        if q != q_in:
            dX = dX[:, :q_in].contiguous()   # the pending points are constants
        return (out.cpu(), dX.cpu()) if on_cpu else (out, dX)

    def optimize(self, X0, lower_bounds, upper_bounds, maxiter: int = 2000, history: int = 10, pgtol: float = 1e-5,
                 ftol: float = 2.220446049250313e-09):
        """On-device multi-start refinement through bo_acqf_optimize (csrc/lbfgs.cu): every restart X0[i] [q, d] runs its
        own box-constrained L-BFGS with the analytic gradient, no host round trip per iteration.  The defaults are scipy's
        L-BFGS-B defaults (m = 10, pgtol = 1e-5, ftol = factr * eps = 2.2e-9), which is what BoTorch's gen_candidates_scipy
        runs with.  Returns (X [r, q, d], values [r], stats dict) on the device."""
        X = torch.as_tensor(X0, dtype=torch.double)
        if X.dim() == 2:
            X = X.unsqueeze(0)
        if X.dim() != 3 or X.shape[-1] != self.model.d:
            raise ValueError(f"X0 must be [r, q, {self.model.d}]")
        self._check_active()
        d = self.model.d
        q_free = X.shape[1]
        lb = np.ascontiguousarray(np.broadcast_to(np.asarray(torch.as_tensor(lower_bounds).cpu(), dtype=np.float64), (d,)))
        ub = np.ascontiguousarray(np.broadcast_to(np.asarray(torch.as_tensor(upper_bounds).cpu(), dtype=np.float64), (d,)))
        stats = (C.c_int32 * 4)()
        # the loop replays a CUDA graph of one evaluation; the legacy default stream cannot be captured, so the call runs
        # on a side stream ordered after the caller's stream (the C call returns synchronised: it reads its statistics back)
        dev = self.model.device
        # EVEREST_LBFGS_GRAPH=1 replays a CUDA graph of one evaluation; the legacy default stream cannot be captured, so the
        # call then runs on a side stream ordered after the caller's (creating the first non-default torch stream of a
        # process costs up to 0.5 s on this platform -- measured in bench.py -- which is why plain launches on the caller's
        # stream are the default)
        use_side = os.environ.get("EVEREST_LBFGS_GRAPH", "0") not in ("", "0")
        cur = torch.cuda.current_stream(dev)
        side = cur
        if use_side:
            side = getattr(self.model, "_side_stream", None)
            if side is None:
                side = self.model._side_stream = torch.cuda.Stream(device=dev)
            side.wait_stream(cur)
        with torch.cuda.device(dev), torch.cuda.stream(side):
            Xd = self._with_pending(X.detach().to(dev)).contiguous().clone()
            r, q_tot, _ = Xd.shape
            out = torch.empty(r, dtype=torch.double, device=dev)
            zq = self.base_samples_q(q_tot)
            L.check(self.model.lib.bo_acqf_optimize(self.model.handle, _dev_ptr(Xd), r, q_tot, q_free,
                                                    lb.ctypes.data_as(L.c_double_p), ub.ctypes.data_as(L.c_double_p), _dev_ptr(zq),
                                                    int(maxiter), int(history), float(pgtol), float(ftol), _dev_ptr(out), stats,
                                                    C.c_void_p(side.cuda_stream)))
        if use_side:
            cur.wait_stream(side)
        info = {"n_acqf_evals": int(stats[0]) * r, "n_steps": int(stats[0]), "nit": int(stats[1]), "n_converged": int(stats[2]),
                "n_budget": int(stats[3])}
        return Xd[:, :q_free].contiguous(), out, info

    def forward_host(self, X: np.ndarray) -> np.ndarray:
        """Same call with HOST buffers through bo_acqf_forward_host (pinned staging + H2D + D2H inside)."""
        X = np.ascontiguousarray(X, dtype=np.float64)
        if X.ndim == 2:
            X = X[None]
        b, q, d = X.shape
        if d != self.model.d:
            raise ValueError(f"X must be [b, q, {self.model.d}]")
        self._check_active()
        if self._PENDING_CONCAT and self.X_pending is not None and self.X_pending.shape[0] > 0:
            P = np.broadcast_to(self.X_pending.cpu().numpy()[None], (b, self.X_pending.shape[0], d))
            X = np.ascontiguousarray(np.concatenate([X, P], axis=1))
            q = X.shape[1]
        out = np.empty(b, dtype=np.float64)
        zq = self.base_samples_q(q)
        with torch.cuda.device(self.model.device):
            L.check(self.model.lib.bo_acqf_forward_host(self.model.handle, X.ctypes.data_as(C.c_void_p), b, q,
                                                        _dev_ptr(zq), out.ctypes.data_as(C.c_void_p), _stream()))
        return out


    # -- packed wire format (fingerprint columns as bits) ---------------------------------------------------------
    def pack_layout(self):
        """(dense_cols, bit_cols): which columns of a candidate row travel as float64 and which as bits (bo_pack_layout)."""
        nd, nb = C.c_int32(0), C.c_int32(0)
        lib, h = self.model.lib, self.model.handle
        L.check(lib.bo_pack_layout(h, C.byref(nd), C.byref(nb), None, None))
        dc = np.zeros(max(nd.value, 1), dtype=np.int32)
        bc = np.zeros(max(nb.value, 1), dtype=np.int32)
        L.check(lib.bo_pack_layout(h, C.byref(nd), C.byref(nb), dc.ctypes.data_as(C.c_void_p), bc.ctypes.data_as(C.c_void_p)))
        return dc[: nd.value], bc[: nb.value]

    def pack_rows(self, X: np.ndarray):
        """X [..., d] float64 -> (dense [rows, n_dense] float64, bits [rows, ceil(n_bits / 64)] uint64) with the host packer of
        the library (bo_pack_rows_host); raises ValueError when a fingerprint column holds anything but 0 / 1."""
        X = np.ascontiguousarray(X, dtype=np.float64).reshape(-1, self.model.d)
        dc, bc = self.pack_layout()
        dense = np.empty((X.shape[0], len(dc)), dtype=np.float64)
        bits = np.empty((X.shape[0], (len(bc) + 63) // 64), dtype=np.uint64)
        L.check(self.model.lib.bo_pack_rows_host(self.model.handle, X.ctypes.data_as(C.c_void_p), X.shape[0],
                                                 dense.ctypes.data_as(C.c_void_p), bits.ctypes.data_as(C.c_void_p)))
        return dense, bits

    def forward_host_packed(self, dense: np.ndarray, bits: np.ndarray, q: int = 1) -> np.ndarray:
        """forward over candidates kept in the packed wire format (bo_acqf_forward_host_packed): rows = b * q."""
        dense = np.ascontiguousarray(dense, dtype=np.float64)
        bits = np.ascontiguousarray(bits, dtype=np.uint64)
        rows = dense.shape[0] if dense.size else bits.shape[0]
        if rows % q or (bits.size and bits.shape[0] != rows):
            raise ValueError("dense / bits must hold b * q rows each")
        if self._PENDING_CONCAT and self.X_pending is not None and self.X_pending.shape[0] > 0:
            raise NotImplementedError("pending points with the packed wire format: pack them into the rows of every q-batch")
        self._check_active()
        b = rows // q
        out = np.empty(b, dtype=np.float64)
        zq = self.base_samples_q(q)
        with torch.cuda.device(self.model.device):
            L.check(self.model.lib.bo_acqf_forward_host_packed(self.model.handle, dense.ctypes.data_as(C.c_void_p),
                                                               bits.ctypes.data_as(C.c_void_p), b, q, _dev_ptr(zq),
                                                               out.ctypes.data_as(C.c_void_p), _stream()))
        return out


class _AcqfAutograd(torch.autograd.Function):
    """torch.autograd bridge: forward = bo_acqf_forward_backward, backward = chain rule with the stored dX."""

    @staticmethod
    def forward(ctx, X, acqf):
        squeeze = X.dim() == 2
        vals, dX = acqf.forward_backward(X)
        ctx.save_for_backward(dX)
        ctx.squeeze = squeeze
        return vals

    @staticmethod
    def backward(ctx, grad_out):
        (dX,) = ctx.saved_tensors
        g = grad_out.to(dX.device).view(-1, 1, 1) * dX
        return (g[0] if ctx.squeeze else g), None


class qNoisyExpectedHypervolumeImprovement(_DeviceAcquisition):
    def __init__(self, model: DeviceGPState, ref_point: Sequence[float], X_baseline, objective: MultiObjective,
                 constraints: Optional[List[OutputConstraint]] = None, eta=None, prune_baseline: bool = False,
                 alpha: float = 0.0, cache_root: bool = True, X_pending=None, mc_samples: int = 512,
                 seed: Optional[int] = None, prune_samples: int = 2048, base_samples_baseline=None):
        super().__init__(model, mc_samples, seed)
        if not (0.0 <= float(alpha) <= 0.5):
            raise ValueError("alpha must be in [0, 0.5]")      # data_models/strategies/predictives/qnehvi.py:19
        self.alpha = float(alpha)
        if not cache_root:
            raise NotImplementedError("cache_root=False (joint re-sampling of the baseline) is not accelerated")
        self.ref_point = [float(v) for v in ref_point]
        self.objective = objective
        if len(self.ref_point) != len(objective.ops):
            raise ValueError("ref_point and objective must have the same number of outcomes")
        self.constraints = list(constraints) if constraints else []
        if eta is not None and self.constraints:
            etas = [float(eta)] * len(self.constraints) if np.ndim(eta) == 0 else [float(e) for e in eta]
            self.constraints = [OutputConstraint(c.idx, c.sign, c.tp, e) for c, e in zip(self.constraints, etas)]
        self._obj_c, self._n_obj = c_array(list(objective.ops), L.ObjectiveOp)
        self._con_c, self._n_con = c_array(self.constraints, L.ConstraintOp)
        self._ref_c = (C.c_double * len(self.ref_point))(*self.ref_point)

        Xb = torch.as_tensor(X_baseline, dtype=torch.double)
        if Xb.dim() != 2:
            raise ValueError("X_baseline must be [n, d]")
        self.prune_idx = None
        if prune_baseline and Xb.shape[0] > 0:
            self.prune_idx = self._prune(Xb, prune_samples)
            Xb = Xb[self.prune_idx.cpu()]
        self._Xb_pruned = Xb
        self.X_pending = None if X_pending is None else torch.as_tensor(X_pending, dtype=torch.double).reshape(-1, model.d)
        self._prepare(self.X_pending, base_samples_baseline)

    _PENDING_CONCAT = False

    def _prepare(self, X_pending=None, base_samples_baseline=None):
        """NoisyExpectedHypervolumeMixin._set_cell_bounds: baseline = pruned observations + pending points."""
        model = self.model
        Xb = self._Xb_pruned
        if X_pending is not None:
            Xb = torch.cat([Xb, torch.as_tensor(X_pending, dtype=torch.double).reshape(-1, model.d).cpu()], dim=0)
        self.X_baseline = Xb
        self.nb = Xb.shape[0]
        self._zq = {}
        zb = base_samples_baseline
        if zb is None:
            zb = sampling.base_samples_device(self.nb, model.M, self.S, self.seed, model.device)
        zb = torch.as_tensor(zb, dtype=torch.double).to(model.device).contiguous()
        if tuple(zb.shape) != (self.S, self.nb, model.M):
            raise ValueError(f"base_samples_baseline must be [{self.S}, {self.nb}, {model.M}]")
        self._zb = zb
        Xbd = Xb.to(model.device).contiguous()
        info = (C.c_int32 * model.M)()
        maxc = C.c_int32(0)
        with torch.cuda.device(model.device):
            # BoFire's `alpha`: approximate binary partitioning for more than two objectives ([UPSTREAM]
            # NondominatedPartitioning(alpha)); two objectives are always decomposed exactly, as in BoTorch
            L.check(model.lib.bo_acqf_set_option(model.handle, b"partition_alpha", float(getattr(self, "alpha", 0.0))))
            L.check(model.lib.bo_nehvi_prepare(model.handle, _dev_ptr(Xbd), self.nb, _dev_ptr(zb), self.S, self._obj_c,
                                               self._n_obj, self._con_c, self._n_con, self._ref_c, info, C.byref(maxc),
                                               _stream()))
        self.max_cells = int(maxc.value)
        self._claim()
        self._after_prepare()

    def _after_prepare(self):
        pass

    def _reprepare(self):
        self._prepare(self.X_pending, self._zb)

    def set_X_pending(self, X_pending=None):
        """[UPSTREAM] qNEHVI.set_X_pending with cache_pending=True: the pending points join the baseline and the
        per-sample box decompositions are rebuilt (no new pruning)."""
        self.X_pending = None if X_pending is None else torch.as_tensor(X_pending, dtype=torch.double).reshape(-1, self.model.d)
        self._prepare(self.X_pending)

    def _prune(self, Xb, prune_samples, seed_offset=7919):
        """[UPSTREAM] prune_inferior_points_multi_objective: keep points with non-zero probability of being
        non-dominated and better than the reference point under `prune_samples` joint posterior samples."""
        model = self.model
        n = Xb.shape[0]
        z = sampling.base_samples_device(n, model.M, prune_samples, self.seed + seed_offset, model.device)
        self._z_prune = z
        counts = torch.zeros(n, dtype=torch.int32, device=model.device)
        info = (C.c_int32 * model.M)()
        Xd = Xb.to(model.device).contiguous()
        with torch.cuda.device(model.device):
            L.check(model.lib.bo_prune_counts(model.handle, _dev_ptr(Xd), n, _dev_ptr(z), prune_samples, self._obj_c,
                                              self._n_obj, self._con_c, self._n_con, self._ref_c, _dev_ptr(counts), info,
                                              _stream()))
        self.prune_counts = counts
        return torch.nonzero(counts > 0).view(-1)

    # introspection used by the parity tests
    def cell_bounds(self):
        """(lower, upper) as [S, C, m] CPU tensors and the per-sample cell counts."""
        m = len(self.ref_point)
        st = self.model
        nc = st.debug_get("ncells", dtype=torch.int32).cpu()
        cap_guess = 64 * max(64, 8 * (self.nb + 1) * m) * m * self.S
        lo = st.debug_get("cell_lo", capacity=cap_guess)
        up = st.debug_get("cell_up", capacity=cap_guess)
        cap = lo.numel() // (m * self.S)
        lo = lo.view(cap, m, self.S).permute(2, 0, 1).cpu()
        up = up.view(cap, m, self.S).permute(2, 0, 1).cpu()
        return lo, up, nc


class qExpectedHypervolumeImprovement(_DeviceAcquisition):
    """Fixed partitioning of the observed front (qehvi.py:37-77). `partitioning_Y` are the objective values
    (maximisation frame, already multiplied by the ref-point mask) BoFire hands to NondominatedPartitioning."""

    def __init__(self, model: DeviceGPState, ref_point, partitioning_Y, objective: MultiObjective, mc_samples: int = 512,
                 seed: Optional[int] = None, X_pending=None, alpha: float = 0.0):
        super().__init__(model, mc_samples, seed)
        if not (0.0 <= float(alpha) <= 0.5):
            raise ValueError("alpha must be in [0, 0.5]")
        self.alpha = float(alpha)      # [UPSTREAM] NondominatedPartitioning(ref_point, Y, alpha) of get_acquisition_function
        self.set_X_pending(X_pending)
        self.ref_point = [float(v) for v in ref_point]
        self.objective = objective
        self._Y = torch.as_tensor(partitioning_Y, dtype=torch.double).reshape(-1, len(self.ref_point)).to(model.device).contiguous()
        self._obj_c, self._n_obj = c_array(list(objective.ops), L.ObjectiveOp)
        self._ref_c = (C.c_double * len(self.ref_point))(*self.ref_point)
        self._reprepare()

    def _reprepare(self):
        model, Y = self.model, self._Y
        maxc = C.c_int32(0)
        with torch.cuda.device(model.device):
            L.check(model.lib.bo_acqf_set_option(model.handle, b"partition_alpha", float(getattr(self, "alpha", 0.0))))
            L.check(model.lib.bo_ehvi_prepare(model.handle, _dev_ptr(Y), Y.shape[0], self.S, self._obj_c, self._n_obj,
                                              self._ref_c, C.byref(maxc), _stream()))
        self.max_cells = int(maxc.value)
        self._claim()
        self._after_prepare()

    def _after_prepare(self):
        pass


class qLogNoisyExpectedHypervolumeImprovement(qNoisyExpectedHypervolumeImprovement):
    """MoboStrategy's default (mobo.py:72-90, data_models/strategies/predictives/mobo.py): same construction as qNEHVI,
    log-space smoothed value ([UPSTREAM] _compute_log_qehvi; tau_relu = 1e-6, tau_max = 1e-2, fat = True)."""

    def __init__(self, *args, tau_relu: float = 1e-6, tau_max: float = 1e-2, **kwargs):
        self._taus = (float(tau_relu), float(tau_max))
        super().__init__(*args, **kwargs)

    def _after_prepare(self):
        self.set_option("log_hvi", 1)
        self.set_option("tau_relu", self._taus[0])
        self.set_option("tau_max", self._taus[1])


class qLogExpectedHypervolumeImprovement(qExpectedHypervolumeImprovement):
    def __init__(self, *args, tau_relu: float = 1e-6, tau_max: float = 1e-2, **kwargs):
        self._taus = (float(tau_relu), float(tau_max))
        super().__init__(*args, **kwargs)

    def _after_prepare(self):
        self.set_option("log_hvi", 1)
        self.set_option("tau_relu", self._taus[0])
        self.set_option("tau_max", self._taus[1])


def _feasible(constraints, Y: torch.Tensor) -> torch.Tensor:
    """[UPSTREAM] compute_feasibility_indicator: every output constraint c(y) <= 0."""
    ok = torch.ones(Y.shape[:-1], dtype=torch.bool, device=Y.device)
    for c in constraints or []:
        ok = ok & (c(Y) <= 0)
    return ok


def estimate_objective_lower_bound(model: DeviceGPState, objective, X, generator=None) -> float:
    """[UPSTREAM] botorch.acquisition.utils._estimate_objective_lower_bound: minus get_infeasible_cost at 32 random convex
    combinations of X -- objective(mean - 6 sd), minimum over the points, clamped at 0 from above.  BoTorch draws the
    weights from torch's global generator; so does this unless `generator` is given."""
    X = torch.as_tensor(X, dtype=torch.double).cpu()
    w = torch.rand(32, X.shape[-2], dtype=torch.double, generator=generator)
    w = w / w.sum(dim=0, keepdim=True)
    mean, var = model.posterior(w @ X)
    lb = objective((mean - 6.0 * var.clamp_min(0.0).sqrt()).cpu())
    return float(lb.min().clamp_max(0.0))


def best_feasible_objective(model: DeviceGPState, objective, constraints, Y: torch.Tensor, X_baseline, generator=None) -> float:
    """[UPSTREAM] compute_best_feasible_objective on posterior means Y [n, M] (the incumbent of qEI / qLogEI / qPI as
    get_acquisition_function builds it): the best objective among the points that satisfy every output constraint; when
    none does, the pessimistic lower bound above."""
    Y = Y.cpu()
    obj = objective(Y)
    if not constraints:
        return float(obj.max())
    ok = _feasible(constraints, Y)
    if bool(ok.any()):
        return float(obj[ok].max())
    return estimate_objective_lower_bound(model, objective, X_baseline, generator=generator)


class _ScalarAcquisition(_DeviceAcquisition):
    """Single-objective MC acquisition functions of SoboStrategy (sobo.py:51-90): qLogEI, qEI, qSR, qUCB, qPI with a
    fixed incumbent and -- when `X_baseline` is given -- the noisy variants qLogNEI (SoboStrategy's default) / qNEI, whose
    incumbent is the best baseline objective of each MC sample under the cached-root joint posterior."""

    _VARIANT = L.ACQF_QLOGEI
    _NOISY = False

    def __init__(self, model: DeviceGPState, objective: ScalarObjective, best_f: Optional[float] = None, X_baseline=None,
                 mc_samples: int = 512, seed: Optional[int] = None, constraints: Optional[List[OutputConstraint]] = None,
                 eta=None, X_pending=None, prune_baseline: bool = True, prune_samples: int = 2048, param: float = 0.0,
                 cache_root: bool = True, lb_generator=None):
        super().__init__(model, mc_samples, seed)
        self._lb_generator = lb_generator
        if not cache_root:
            raise NotImplementedError("cache_root=False (joint re-sampling of the baseline) is not accelerated")
        self.objective = objective
        self.constraints = list(constraints) if constraints else []
        if eta is not None and self.constraints:
            etas = [float(eta)] * len(self.constraints) if np.ndim(eta) == 0 else [float(e) for e in eta]
            self.constraints = [OutputConstraint(c.idx, c.sign, c.tp, e) for c, e in zip(self.constraints, etas)]
        self._obj_c, self._n_obj = c_array(list(objective.ops), L.ObjectiveOp)
        self._con_c, self._n_con = c_array(self.constraints, L.ConstraintOp)
        self.param = float(param)
        Xbd, zb = None, None
        self.prune_idx = None
        if self._NOISY:
            if X_baseline is None:
                raise ValueError("the noisy variants need X_baseline (X_observed)")
            Xb = torch.as_tensor(X_baseline, dtype=torch.double)
            if prune_baseline and Xb.shape[0] > 0:
                self.prune_idx = self._prune(Xb, prune_samples)
                Xb = Xb[self.prune_idx.cpu()]
            if Xb.shape[0] == 0:
                raise ValueError("the noisy variants need at least one baseline point")
            self.X_baseline = Xb
            self.nb = Xb.shape[0]
            zb = sampling.base_samples_device(self.nb, model.M, self.S, self.seed, model.device)
            self._zb = zb
            Xbd = Xb.to(model.device).contiguous()
            self.best_f = float("nan")
        else:
            if best_f is None:
                raise ValueError("best_f is required")
            self.best_f = float(best_f)
        self.set_X_pending(X_pending)    # scored jointly with X ([UPSTREAM] @concatenate_pending_points)
        self._Xbd = Xbd
        self._reprepare()

    def _reprepare(self):
        model, Xbd, zb = self.model, self._Xbd, (self._zb if self._NOISY else None)
        info = (C.c_int32 * model.M)()
        with torch.cuda.device(model.device):
            L.check(model.lib.bo_scalar_prepare(model.handle, self._VARIANT, self.param, self.S, self.objective.combine_code,
                                                self._obj_c, self._n_obj, self._con_c, self._n_con,
                                                0.0 if self._NOISY else self.best_f,
                                                _dev_ptr(Xbd) if Xbd is not None else None, self.nb,
                                                _dev_ptr(zb) if zb is not None else None, info, _stream()))
            if self._NOISY and self._n_con:
                # [UPSTREAM] compute_best_feasible_objective: infeasible baseline samples are -inf unless some MC sample
                # has no feasible baseline point at all -- then BoTorch substitutes its pessimistic lower bound
                cnt = C.c_int32(0)
                L.check(model.lib.bo_scalar_baseline_best(model.handle, float("-inf"), C.byref(cnt), _stream()))
                self.n_all_infeasible = int(cnt.value)
                if cnt.value > 0:
                    if getattr(self, "_infeasible_value", None) is None:
                        self._infeasible_value = estimate_objective_lower_bound(model, self.objective, self.X_baseline,
                                                                                generator=self._lb_generator)
                    L.check(model.lib.bo_scalar_baseline_best(model.handle, self._infeasible_value, C.byref(cnt), _stream()))
        self._claim()

    def _prune(self, Xb, prune_samples, seed_offset=7919):
        """[UPSTREAM] prune_inferior_points: keep the points that are the best one in at least one joint sample."""
        model = self.model
        n = Xb.shape[0]
        z = sampling.base_samples_device(n, model.M, prune_samples, self.seed + seed_offset, model.device)
        counts = torch.zeros(n, dtype=torch.int32, device=model.device)
        info = (C.c_int32 * model.M)()
        Xd = Xb.to(model.device).contiguous()
        with torch.cuda.device(model.device):
            L.check(model.lib.bo_prune_counts_scalar(model.handle, _dev_ptr(Xd), n, _dev_ptr(z), prune_samples,
                                                     self.objective.combine_code, self._obj_c, self._n_obj, self._con_c,
                                                     self._n_con, _dev_ptr(counts), info, _stream()))
        self.prune_counts = counts
        return torch.nonzero(counts > 0).view(-1)


class qLogExpectedImprovement(_ScalarAcquisition):
    def __init__(self, model: DeviceGPState, best_f: float, objective: ScalarObjective, mc_samples: int = 512,
                 seed: Optional[int] = None, constraints=None, eta=None, X_pending=None):
        super().__init__(model, objective, best_f=best_f, mc_samples=mc_samples, seed=seed, constraints=constraints,
                         eta=eta, X_pending=X_pending)


class qExpectedImprovement(_ScalarAcquisition):
    _VARIANT = L.ACQF_QEI

    def __init__(self, model, best_f, objective, mc_samples=512, seed=None, constraints=None, eta=None, X_pending=None):
        super().__init__(model, objective, best_f=best_f, mc_samples=mc_samples, seed=seed, constraints=constraints,
                         eta=eta, X_pending=X_pending)


class qSimpleRegret(_ScalarAcquisition):
    _VARIANT = L.ACQF_QSR

    def __init__(self, model, objective, mc_samples=512, seed=None, X_pending=None):
        super().__init__(model, objective, best_f=0.0, mc_samples=mc_samples, seed=seed, X_pending=X_pending)


class qUpperConfidenceBound(_ScalarAcquisition):
    _VARIANT = L.ACQF_QUCB

    def __init__(self, model, beta, objective, mc_samples=512, seed=None, X_pending=None):
        super().__init__(model, objective, best_f=0.0, mc_samples=mc_samples, seed=seed, X_pending=X_pending, param=beta)


class qProbabilityOfImprovement(_ScalarAcquisition):
    _VARIANT = L.ACQF_QPI

    def __init__(self, model, best_f, objective, tau=1e-3, mc_samples=512, seed=None, constraints=None, eta=None,
                 X_pending=None):
        super().__init__(model, objective, best_f=best_f, mc_samples=mc_samples, seed=seed, constraints=constraints,
                         eta=eta, X_pending=X_pending, param=tau)


class qNoisyExpectedImprovement(_ScalarAcquisition):
    _VARIANT = L.ACQF_QEI
    _NOISY = True

    def __init__(self, model, X_baseline, objective, mc_samples=512, seed=None, constraints=None, eta=None, X_pending=None,
                 prune_baseline=True, prune_samples=2048, cache_root=True, lb_generator=None):
        super().__init__(model, objective, X_baseline=X_baseline, mc_samples=mc_samples, seed=seed, constraints=constraints,
                         eta=eta, X_pending=X_pending, prune_baseline=prune_baseline, prune_samples=prune_samples,
                         cache_root=cache_root, lb_generator=lb_generator)


class qLogNoisyExpectedImprovement(qNoisyExpectedImprovement):
    _VARIANT = L.ACQF_QLOGEI


def get_acquisition_function(acquisition_function_name: str, model: DeviceGPState, objective, X_observed,
                             X_pending=None, constraints=None, eta=None, mc_samples: int = 512, seed=None,
                             ref_point=None, Y=None, alpha: float = 0.0, prune_baseline: bool = True, beta: float = 0.2,
                             tau: float = 1e-3, cache_root: bool = True, **kwargs):
    """Mirror of botorch.acquisition.factory.get_acquisition_function as BoFire calls it (sobo.py:64-89,
    mobo.py:72-90): every name the Sobo / Mobo data models can select."""
    name = acquisition_function_name
    if name in ("qEI", "qLogEI", "qPI"):
        # [UPSTREAM] the non-noisy variants take the best FEASIBLE objective of the posterior mean at the observed points
        # (compute_best_feasible_objective); `eta` does not enter the hard feasibility test
        mean, _ = model.posterior(X_observed)
        best_f = best_feasible_objective(model, objective, constraints, mean, X_observed, generator=kwargs.pop("lb_generator", None))
        if name == "qLogEI":
            return qLogExpectedImprovement(model, best_f, objective, mc_samples=mc_samples, seed=seed,
                                           constraints=constraints, eta=eta, X_pending=X_pending)
        if name == "qEI":
            return qExpectedImprovement(model, best_f, objective, mc_samples=mc_samples, seed=seed, constraints=constraints,
                                        eta=eta, X_pending=X_pending)
        return qProbabilityOfImprovement(model, best_f, objective, tau=tau, mc_samples=mc_samples, seed=seed,
                                         constraints=constraints, eta=eta, X_pending=X_pending)
    if name == "qSR":
        return qSimpleRegret(model, objective, mc_samples=mc_samples, seed=seed, X_pending=X_pending)
    if name == "qUCB":
        return qUpperConfidenceBound(model, beta, objective, mc_samples=mc_samples, seed=seed, X_pending=X_pending)
    if name in ("qNEI", "qLogNEI"):
        cls = qNoisyExpectedImprovement if name == "qNEI" else qLogNoisyExpectedImprovement
        return cls(model, X_observed, objective, mc_samples=mc_samples, seed=seed, constraints=constraints, eta=eta,
                   X_pending=X_pending, prune_baseline=prune_baseline, cache_root=cache_root, **kwargs)
    if name in ("qNEHVI", "qLogNEHVI"):
        if ref_point is None:
            raise ValueError("`ref_point` must be specified")
        cls = qNoisyExpectedHypervolumeImprovement if name == "qNEHVI" else qLogNoisyExpectedHypervolumeImprovement
        return cls(model, ref_point, X_observed, objective, constraints=constraints, eta=eta, prune_baseline=prune_baseline,
                   alpha=alpha, cache_root=cache_root, X_pending=X_pending, mc_samples=mc_samples, seed=seed, **kwargs)
    if name in ("qEHVI", "qLogEHVI"):
        if ref_point is None or Y is None:
            raise ValueError("`ref_point` and `Y` must be specified")
        if constraints:
            raise NotImplementedError("output constraints with qEHVI / qLogEHVI are not accelerated; use qNEHVI / qLogNEHVI")
        # [UPSTREAM] the partitioning is built from the objective values of the observations
        Yobj = objective(torch.as_tensor(Y, dtype=torch.double))
        cls = qExpectedHypervolumeImprovement if name == "qEHVI" else qLogExpectedHypervolumeImprovement
        return cls(model, ref_point, Yobj, objective, mc_samples=mc_samples, seed=seed, X_pending=X_pending, alpha=alpha)
    raise NotImplementedError(f"Unknown / non-accelerated acquisition function {acquisition_function_name}")

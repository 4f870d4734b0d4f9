"""Problem dictionary (everest_b200.configs) -> oracle objects, so that the CUDA path and the CPU oracle
are built from exactly the same inputs."""
import torch

from everest_b200 import kernels as K
from everest_b200.objectives import MultiObjective, ScalarObjective
from oracle import bo_oracle as O

DT = torch.float64


def kernel_to_oracle(k):
    if isinstance(k, K.RBFKernel):
        return O.RBF(list(k.active_dims), list(k.lengthscale))
    if isinstance(k, K.MaternKernel):
        return O.Matern(k.nu, list(k.active_dims), list(k.lengthscale))
    if isinstance(k, K.HammingDistanceKernel):
        return O.Hamming(sorted(k.categorical_features.items()), list(k.lengthscale))
    if isinstance(k, K.TanimotoKernel):
        return O.Tanimoto(list(k.active_dims))
    if isinstance(k, K.ScaleKernel):
        return O.Scale(kernel_to_oracle(k.base_kernel), k.outputscale)
    if isinstance(k, K.AdditiveKernel):
        return O.Add([kernel_to_oracle(c) for c in k.kernels])
    if isinstance(k, K.MultiplicativeKernel):
        return O.Mul([kernel_to_oracle(c) for c in k.kernels])
    raise TypeError(type(k))


def op_to_oracle(op):
    if op.kind in ("max", "min", "close_to_target", "min_sigmoid", "max_sigmoid"):
        return (op.kind, op.idx, op.p0, op.p1)
    return (op.kind, op.idx, op.p0, op.p1, op.p2)


def oracle_gp(problem):
    X = torch.as_tensor(problem["X"], dtype=DT)
    outs = []
    for o in problem["outputs"]:
        y = torch.as_tensor(o["y"], dtype=DT)
        ym, ys = O.standardize_stats(y)
        outs.append(O.GPOutput(kernel=kernel_to_oracle(o["kernel"]), in_offset=torch.as_tensor(problem["in_offset"], dtype=DT),
                               in_scale=torch.as_tensor(problem["in_scale"], dtype=DT), mean_const=o["mean_const"],
                               noise=o["noise"], y=y, y_mean=ym, y_std=ys))
    return O.GPOracle(X, outs).factorize()


def oracle_acqf(problem, gp, prune_baseline=True, prune_samples=2048, constraints=None):
    obj = problem["objective"]
    if problem["acqf"] == "qnehvi":
        cons = [(c.idx, c.sign, c.tp, c.eta) for c in constraints] if constraints else None
        return O.QNEHVIOracle(gp, problem["ref_point"], problem["X"], [op_to_oracle(o) for o in obj.ops], constraints=cons,
                              mc_samples=problem["S"], seed=problem["sampler_seed"], prune_baseline=prune_baseline,
                              prune_samples=prune_samples, prune_seed=problem["sampler_seed"] + 7919)
    if problem["acqf"] == "qlogei":
        if obj.combine == "single":
            spec = ("single", op_to_oracle(obj.ops[0]))
        else:
            spec = (obj.combine, [(op_to_oracle(o), o.w) for o in obj.ops])
        return O.QLogEIOracle(gp, spec, problem["X"], mc_samples=problem["S"], seed=problem["sampler_seed"])
    raise ValueError(problem["acqf"])


def device_cells_for_oracle(acq_d, ref_point):
    """(lower, upper) [S, C, m] of the DEVICE's per-sample box decompositions, padded to rectangular with empty cells at
    the reference point -- what QNEHVIOracle(cell_bounds=...) takes.  The oracle's own pure-Python Lacour decomposition of
    512 fronts with thousands of cells takes minutes; the parity tests pin the device cell lists bit for bit against it at
    small sizes (tests/test_gpu_parity.py::test_box_decomposition_bit_exact), the full-size tests and bench.py's
    cpu_baseline leg inject them."""
    lo, up, nc = acq_d.cell_bounds()
    C = int(nc.max())
    ref = torch.tensor(ref_point, dtype=DT)
    lo, up = lo[:, :C].clone(), up[:, :C].clone()
    pad = torch.arange(C).unsqueeze(0) >= nc.unsqueeze(1)
    lo[pad] = ref
    up[pad] = ref
    return lo, up


def oracle_qnehvi_on_device_baseline(problem, gp, acq_d, inject_cells=False):
    """QNEHVIOracle on the baseline the device pruned to (same points, no second pruning pass), optionally with the
    device's cell lists."""
    idx = acq_d.prune_idx.cpu()
    cells = device_cells_for_oracle(acq_d, problem["ref_point"]) if inject_cells else None
    return O.QNEHVIOracle(gp, problem["ref_point"], torch.as_tensor(problem["X"])[idx],
                          [op_to_oracle(o) for o in problem["objective"].ops], mc_samples=problem["S"],
                          seed=problem["sampler_seed"], prune_baseline=False, cell_bounds=cells)


class OracleAcqfAdapter:
    """The CPU oracle behind the interface everest_b200.optim drives (forward, forward_backward through torch autograd,
    model.device): lets the parity tests run the SAME gen_candidates_scipy / optimize_acqf code on the oracle -- what
    BoTorch does when it back-propagates through its own acquisition function."""

    class _Model:
        device = torch.device("cpu")

    def __init__(self, acq_o, d):
        self.acq_o = acq_o
        self.model = self._Model()
        self.model.d = d
        self.X_pending = None

    def __call__(self, X):
        with torch.no_grad():
            return self.acq_o.forward(torch.as_tensor(X, dtype=DT).cpu())

    def forward_backward(self, X):
        Xr = torch.as_tensor(X, dtype=DT).cpu().detach().clone().requires_grad_(True)
        v = self.acq_o.forward(Xr)
        (g,) = torch.autograd.grad(v.sum(), Xr)
        return v.detach(), g

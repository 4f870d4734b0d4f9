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

"""Generate golden vectors from the reference's OWN in-tree arithmetic.

Run in the build container only (needs /root/reference):

    python tests/golden/gen_reference_golden.py

botorch / gpytorch / linear_operator / formulaic are not installed, so they are replaced
by inert stub modules: enough for ``import bofire...`` to succeed, never called for
arithmetic.  Everything written to ``reference_golden.json`` is computed by reference
code that lives under /root/reference/bofire:

  * batch_tanimoto_sim / BitDistance._sim      kernels/fingerprint_kernels/base_fingerprint_kernel.py:36-53,63-87
  * HammingKernelWithOneHots.forward           kernels/categorical.py:43-70
      (its OneHotToNumeric dependency is BoTorch's; a 3-line argmax stand-in is injected)
  * get_objective_callable, get_multiobjective_objective, additive / multiplicative
    objectives, constrained_objective2botorch  utils/torch_tools.py:258-337,384-450,662-727
  * get_ref_point_mask, infer_ref_point        utils/multiobjective.py:18-55,133-159
  * DTLZ2 / Himmelblau / Detergent _f          benchmarks/multi.py:95-132, single.py:409-446, detergent.py:10-88

plus the known-answer tables copied as DATA (inputs + expected answers, no code) from
tests/bofire/utils/test_multiobjective.py:75-266 and tests/bofire/kernels/test_categorical.py.
"""

import importlib.abc
import importlib.machinery
import json
import os
import sys
import types

import numpy as np
import pandas as pd
import torch

REF = "/root/reference"
STUB_ROOTS = {"botorch", "gpytorch", "linear_operator", "formulaic", "cvxpy", "cyipopt", "entmoot",
              "xgboost", "multiprocess", "plotly", "shap", "pymoo", "gurobipy", "pyomo"}


class _Auto(type):
    def __getattr__(cls, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _mk(name)


def _mk(name):
    return _Auto(name, (), {"__init__": lambda self, *a, **k: None,
                            "__class_getitem__": classmethod(lambda c, i: c)})


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        v = _mk(name)
        setattr(self, name, v)
        return v


class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path, target=None):
        if fullname.split(".")[0] in STUB_ROOTS:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)

    def create_module(self, spec):
        m = _StubModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


def install_stubs():
    sys.meta_path.insert(0, _Finder())
    sys.path.insert(0, REF)
    import gpytorch.kernels
    import gpytorch.kernels.kernel
    import botorch.models.transforms.input as bti

    class _KernelBase(torch.nn.Module):
        """Holds what gpytorch.kernels.Kernel would hold for the in-tree forward()s."""

        def __init__(self, ard_num_dims=None, **kw):
            super().__init__()
            self.ard_num_dims = ard_num_dims
            self.lengthscale = torch.ones(1, 1 if ard_num_dims is None else ard_num_dims, dtype=torch.double)
            self.distance_module = None

    class _OneHotToNumeric:
        def __init__(self, dim, categorical_features):
            self.cf = categorical_features

        def __call__(self, X):
            return torch.stack([X[..., s : s + c].argmax(dim=-1) for s, c in self.cf.items()], dim=-1).to(X)

    gpytorch.kernels.Kernel = _KernelBase
    gpytorch.kernels.kernel.Kernel = _KernelBase
    bti.OneHotToNumeric = _OneHotToNumeric


def main():
    install_stubs()
    from bofire.benchmarks.api import DTLZ2, Detergent, Himmelblau
    from bofire.data_models.domain.api import Domain, Outputs
    from bofire.data_models.features.api import ContinuousInput, ContinuousOutput
    from bofire.data_models.objectives.api import (
        CloseToTargetObjective, MaximizeObjective, MaximizeSigmoidObjective, MinimizeObjective,
        MinimizeSigmoidObjective, TargetObjective)
    from bofire.kernels.categorical import HammingKernelWithOneHots
    from bofire.kernels.fingerprint_kernels.base_fingerprint_kernel import BitDistance
    from bofire.kernels.fingerprint_kernels.tanimoto_kernel import TanimotoKernel
    from bofire.utils import multiobjective as mo
    from bofire.utils import torch_tools as tt

    g = {}
    rng = np.random.default_rng(20261018)

    # ---- Tanimoto ------------------------------------------------------------------
    x1 = (rng.random((9, 96)) < 0.15).astype(np.float64)
    x2 = (rng.random((6, 96)) < 0.15).astype(np.float64)
    x1[3] = 0.0  # all-zero fingerprint
    x2[2] = 0.0
    x2[4] = x1[5]  # identical pair
    sim = BitDistance()._sim(torch.tensor(x1), torch.tensor(x2), postprocess=False)
    tk = TanimotoKernel()
    sim_k = tk.forward(torch.tensor(x1), torch.tensor(x2))
    assert torch.equal(sim, sim_k)
    g["tanimoto"] = dict(x1=x1.tolist(), x2=x2.tolist(), K=sim.tolist(),
                         K11=BitDistance()._sim(torch.tensor(x1), torch.tensor(x1), postprocess=False).tolist())

    # ---- Hamming with one-hots -------------------------------------------------------
    cases = []
    for cat, ard, ls in [({0: 3}, None, None), ({0: 2, 2: 4}, None, None),
                         ({0: 2, 2: 4}, 6, [1.5, 3.0, 0.0, 0.0, 0.0, 0.0]),
                         ({0: 4, 4: 6}, 10, [0.7, 2.2] + [0.0] * 8)]:
        k = HammingKernelWithOneHots(categorical_features=cat, ard_num_dims=ard)
        if ls is not None:
            k.lengthscale = torch.tensor([ls], dtype=torch.double)
        dim = sum(cat.values())
        n = 7
        X = np.zeros((n, dim))
        for s, c in cat.items():
            codes = rng.integers(0, c, size=n)
            X[np.arange(n), s + codes] = 1.0
        X2 = X[::-1].copy()[:5]
        K = k.forward(torch.tensor(X), torch.tensor(X2))
        cases.append(dict(groups=[[s, c] for s, c in cat.items()],
                          lengthscale=(ls if ls is not None else [1.0]), x1=X.tolist(), x2=X2.tolist(),
                          K=K.tolist()))
    g["hamming"] = cases

    # ---- objective callables (same probe as tests/bofire/utils/test_torch_tools.py:105-139)
    Y = rng.random((50, 3)) * 5
    Yt = torch.tensor(Y)
    objs = [
        ("max", MaximizeObjective(w=1.0, bounds=(0, 1)), [0.0, 1.0]),
        ("max", MaximizeObjective(w=1.0, bounds=(1, 3)), [1.0, 3.0]),
        ("min", MinimizeObjective(w=1.0, bounds=(0, 1)), [0.0, 1.0]),
        ("min", MinimizeObjective(w=1.0, bounds=(1, 4)), [1.0, 4.0]),
        ("close_to_target", CloseToTargetObjective(w=1.0, target_value=2.0, exponent=1.0), [2.0, 1.0]),
        ("close_to_target", CloseToTargetObjective(w=1.0, target_value=2.5, exponent=2.0), [2.5, 2.0]),
        ("min_sigmoid", MinimizeSigmoidObjective(w=1.0, steepness=2.0, tp=1.5), [2.0, 1.5]),
        ("max_sigmoid", MaximizeSigmoidObjective(w=1.0, steepness=0.5, tp=2.0), [0.5, 2.0]),
        ("target", TargetObjective(w=1.0, target_value=2.0, steepness=3.0, tolerance=0.5), [2.0, 0.5, 3.0]),
    ]
    oc = []
    for kind, o, params in objs:
        for idx in (0, 2):
            f = tt.get_objective_callable(idx=idx, objective=o, x_adapt=Yt[:, idx])
            oc.append(dict(op=[kind, idx] + params, out=f(Yt, None).tolist()))
    g["objective_callables"] = dict(Y=Y.tolist(), cases=oc)

    # ---- multi-objective objective + additive / multiplicative + constraints ----------
    of = [
        ContinuousOutput(key="a", objective=MaximizeObjective(w=0.5, bounds=(0, 2))),
        ContinuousOutput(key="b", objective=MinimizeObjective(w=1.0, bounds=(0, 1))),
        ContinuousOutput(key="c", objective=CloseToTargetObjective(w=0.7, target_value=2.0, exponent=2.0)),
    ]
    outputs = Outputs(features=of)
    exp = pd.DataFrame(Y, columns=["a", "b", "c"])
    for k_ in ["a", "b", "c"]:
        exp[f"valid_{k_}"] = 1
    mobj = tt.get_multiobjective_objective(outputs, exp)
    addobj = tt.get_additive_botorch_objective(outputs, exp)
    mulobj = tt.get_multiplicative_botorch_objective(
        Outputs(features=[ContinuousOutput(key="a", objective=MaximizeObjective(w=0.5, bounds=(0, 2))),
                          ContinuousOutput(key="b", objective=MaximizeSigmoidObjective(w=1.0, steepness=2.0, tp=1.0)),
                          ContinuousOutput(key="c", objective=MinimizeSigmoidObjective(w=0.7, steepness=1.0, tp=3.0))]),
        exp, adapt_weights_to_1_inf=False)
    g["multiobjective"] = dict(
        ops=[["max", 0, 0.0, 2.0], ["min", 1, 0.0, 1.0], ["close_to_target", 2, 2.0, 2.0]],
        out=mobj(Yt, None).tolist(),
        additive=dict(weights=[0.5, 1.0, 0.7], out=addobj(Yt, None).tolist()),
        multiplicative=dict(ops=[["max", 0, 0.0, 2.0], ["max_sigmoid", 1, 2.0, 1.0], ["min_sigmoid", 2, 1.0, 3.0]],
                            weights=[0.5, 1.0, 0.7], out=mulobj(Yt, None).tolist()))
    cons_out = Outputs(features=[
        ContinuousOutput(key="a", objective=MaximizeObjective(w=1.0)),
        ContinuousOutput(key="b", objective=MaximizeSigmoidObjective(w=1.0, steepness=4.0, tp=1.5)),
        ContinuousOutput(key="c", objective=MinimizeSigmoidObjective(w=1.0, steepness=0.5, tp=2.5))])
    cons, etas = tt.get_output_constraints(cons_out, exp)
    g["constraints"] = dict(ops=[[1, -1.0, 1.5, 1.0 / 4.0], [2, 1.0, 2.5, 1.0 / 0.5]], etas=list(etas),
                            values=[c(Yt).tolist() for c in cons])

    # ---- ref points (KAT tables from tests/bofire/utils/test_multiobjective.py) ------
    if1, if2 = ContinuousInput(key="if1", bounds=(0, 1)), ContinuousInput(key="if2", bounds=(0, 1))
    of1 = ContinuousOutput(objective=MaximizeObjective(w=1), key="of1")
    of2 = ContinuousOutput(objective=MinimizeObjective(w=1), key="of2")
    of3 = ContinuousOutput(objective=MaximizeObjective(w=1), key="of3")
    of7 = ContinuousOutput(objective=CloseToTargetObjective(w=1, target_value=5, exponent=1), key="of7")
    base = {"if1": [3.0, 4.0, 5.0, 6.0], "if2": [10.0, 7.0, 8.0, 12.0]}
    tables = [
        (Domain(inputs=[if1, if2], outputs=[of1, of2]), {"of1": [1.0, 10.0, 4.0, 5.0], "of2": [5.0, 3.0, 2.0, 5.0]},
         [["max", 0, 0.0, 1.0], ["min", 1, 0.0, 1.0]], [1, 2]),
        (Domain(inputs=[if1, if2], outputs=[of1, of3]), {"of1": [1.0, 10.0, 4.0, 5.0], "of3": [5.0, 3.0, 2.0, 5.0]},
         [["max", 0, 0.0, 1.0], ["max", 1, 0.0, 1.0]], [1, 3]),
        (Domain(inputs=[if1, if2], outputs=[of1, of2, of7]),
         {"of1": [1.0, 10.0, 4.0, 5.0], "of2": [5.0, 3.0, 2.0, 5.0], "of7": [10.0, 0.0, 30.0, 6.0]},
         [["max", 0, 0.0, 1.0], ["min", 1, 0.0, 1.0], ["close_to_target", 2, 5.0, 1.0]], [1, 2, 3]),
    ]
    rp = []
    for dom, cols, ops, pareto_idx in tables:
        d_ = dict(base)
        d_.update(cols)
        for k_ in cols:
            d_[f"valid_{k_}"] = [1, 1, 1, 1]
        df = pd.DataFrame.from_dict(d_)
        mask = mo.get_ref_point_mask(dom)
        r_masked = mo.infer_ref_point(dom, df, return_masked=True)
        r_plain = mo.infer_ref_point(dom, df, return_masked=False)
        keys = list(cols.keys())
        rp.append(dict(Y=np.array([cols[k_] for k_ in keys]).T.tolist(), ops=ops, mask=mask.tolist(),
                       ref_masked=[float(r_masked[k_]) for k_ in keys], ref_plain=[float(r_plain[k_]) for k_ in keys],
                       expected_pareto_idx=pareto_idx))
    g["ref_points"] = rp

    # ---- benchmark functions ------------------------------------------------------
    X6 = rng.random((12, 6))
    dt = DTLZ2(dim=6, num_objectives=4)
    ydt = dt._f(pd.DataFrame(X6, columns=dt.domain.inputs.get_keys()))
    g["dtlz2_6d_4obj"] = dict(X=X6.tolist(), Y=ydt[[f"f_{i}" for i in range(4)]].values.tolist())
    X2 = rng.random((12, 2)) * 12 - 6
    hb = Himmelblau()
    g["himmelblau"] = dict(X=X2.tolist(), Y=hb._f(pd.DataFrame(X2, columns=["x_1", "x_2"]))["y"].values.tolist())
    det = Detergent()
    X5 = rng.random((8, 5)) * np.array([0.2, 0.3, 0.18, 0.06, 0.04]) + np.array([0, 0, 0.02, 0, 0])
    g["detergent"] = dict(X=X5.tolist(), coef=det.coef.tolist(),
                          Y=det._f(pd.DataFrame(X5, columns=[f"x{i+1}" for i in range(5)])).values.tolist())

    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_golden.json")
    with open(out, "w") as fh:
        json.dump(g, fh)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()

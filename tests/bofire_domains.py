"""BoFire `Domain`s built from the reference's OWN data-model classes (bofire.data_models: pure pydantic / pandas).

The package comes from `baseline/_ref` (the --no-deps pip install of the unmodified reference; it travels to
the GPU box) or from an installed BoFire.  `bofire.strategies` / `bofire.benchmarks` need botorch and are NOT importable:
the benchmark functions below are the numpy restatements of everest_b200/benchmarks.py (pinned bit for bit to the reference's
`_f` by tests/golden/reference_golden.json)."""
import os
import sys
import warnings

import numpy as np
import pandas as pd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_REF = os.path.join(ROOT, "baseline", "_ref")


def have_bofire() -> bool:
    if os.path.isdir(os.path.join(_REF, "bofire")) and _REF not in sys.path:
        sys.path.insert(0, _REF)
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            import bofire.data_models.domain.api  # noqa: F401
            import bofire.data_models.strategies.predictives.qnehvi  # noqa: F401
        return True
    except Exception:
        return False


def detergent_domain():
    """benchmarks/detergent.py:58-79."""
    from bofire.data_models.constraints.linear import LinearInequalityConstraint
    from bofire.data_models.domain.domain import Domain
    from bofire.data_models.features.continuous import ContinuousInput, ContinuousOutput

    return Domain.from_lists(
        inputs=[ContinuousInput(key="x1", bounds=[0.0, 0.2]), ContinuousInput(key="x2", bounds=[0.0, 0.3]),
                ContinuousInput(key="x3", bounds=[0.02, 0.2]), ContinuousInput(key="x4", bounds=[0.0, 0.06]),
                ContinuousInput(key="x5", bounds=[0.0, 0.04])],
        outputs=[ContinuousOutput(key=f"y{i + 1}") for i in range(5)],
        constraints=[LinearInequalityConstraint(features=["x1", "x2", "x3", "x4", "x5"], coefficients=[-1] * 5, rhs=-0.2),
                     LinearInequalityConstraint(features=["x1", "x2", "x3", "x4", "x5"], coefficients=[1] * 5, rhs=0.4)])


def detergent_f(domain, candidates: pd.DataFrame) -> pd.DataFrame:
    """Benchmark.f(..., return_complete=True): inputs + outputs + valid_<key> columns."""
    from everest_b200 import benchmarks as B

    X = candidates[domain.inputs.get_keys()].reset_index(drop=True)
    Y = pd.DataFrame(B.detergent(X.values), columns=domain.outputs.get_keys())
    out = pd.concat([X, Y], axis=1)
    for k in domain.outputs.get_keys():
        out[f"valid_{k}"] = 1
    return out


def himmelblau_domain():
    """benchmarks/single.py:377-407."""
    from bofire.data_models.domain.domain import Domain
    from bofire.data_models.features.continuous import ContinuousInput, ContinuousOutput
    from bofire.data_models.objectives.api import MinimizeObjective

    return Domain.from_lists(
        inputs=[ContinuousInput(key="x_1", bounds=[-6, 6]), ContinuousInput(key="x_2", bounds=[-6, 6])],
        outputs=[ContinuousOutput(key="y", objective=MinimizeObjective())])


def himmelblau_f(domain, candidates: pd.DataFrame) -> pd.DataFrame:
    from everest_b200 import benchmarks as B

    X = candidates[domain.inputs.get_keys()].reset_index(drop=True)
    out = X.copy()
    out["y"] = B.himmelblau(X.values)
    out["valid_y"] = 1
    return out


def mixed_domain(n_choose_k: bool = False):
    """Two continuous inputs, a discrete one and two categoricals (one with a forbidden level), two objectives + one
    sigmoid output constraint."""
    from bofire.data_models.constraints.api import LinearInequalityConstraint
    from bofire.data_models.domain.domain import Domain
    from bofire.data_models.features.api import CategoricalInput, ContinuousInput, ContinuousOutput, DiscreteInput
    from bofire.data_models.objectives.api import MaximizeObjective, MaximizeSigmoidObjective, MinimizeObjective

    return Domain.from_lists(
        inputs=[ContinuousInput(key="a", bounds=[0.0, 1.0]), ContinuousInput(key="b", bounds=[-1.0, 3.0]),
                DiscreteInput(key="n", values=[1.0, 2.0, 5.0]),
                CategoricalInput(key="solvent", categories=["w", "e", "m"], allowed=[True, True, False]),
                CategoricalInput(key="cat", categories=["p", "q"])],
        outputs=[ContinuousOutput(key="y1", objective=MaximizeObjective()), ContinuousOutput(key="y2", objective=MinimizeObjective()),
                 ContinuousOutput(key="y3", objective=MaximizeSigmoidObjective(steepness=10.0, tp=0.2))],
        constraints=[LinearInequalityConstraint(features=["a", "b"], coefficients=[1.0, 1.0], rhs=3.5)])


def mixed_f(domain, candidates: pd.DataFrame) -> pd.DataFrame:
    X = candidates[domain.inputs.get_keys()].reset_index(drop=True)
    a, b, n = X["a"].values.astype(float), X["b"].values.astype(float), X["n"].values.astype(float)
    s = (X["solvent"] == "e").values.astype(float)
    c = (X["cat"] == "q").values.astype(float)
    out = X.copy()
    out["y1"] = np.sin(3 * a) + 0.3 * b + 0.5 * s - 0.1 * n
    out["y2"] = (a - 0.3) ** 2 + 0.2 * b * c + 0.05 * n
    out["y3"] = 0.5 * a + 0.1 * b + 0.2 * c
    for k in ("y1", "y2", "y3"):
        out[f"valid_{k}"] = 1
    return out

"""The drop-in behind BoFire's own API on the device (everest_b200/bofire_strategy.py): strategy DATA MODELS + Domain in,
DataFrames in and out.  The README loop of BASELINE config 1 (README.md:78-106) literally: RandomStrategy init, then
QnehviStrategy tell / ask; MoboStrategy with BoFire's literal MixedSingleTaskGPSurrogate default on a mixed space;
SoboStrategy (qLogNEI default) with the drop-in's predictions pinned against the CPU oracle built from the fitted specs."""
import warnings

import numpy as np
import pandas as pd
import pytest
import torch

from tests import bofire_domains as BD
from tests import problems as P

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not BD.have_bofire(), reason="bofire.data_models not importable")]


@pytest.fixture(autouse=True)
def _quiet():
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        yield


def test_readme_detergent_loop_from_domain_and_dataframes():
    import everest_b200.bofire_strategy as strategies
    from bofire.data_models.strategies.predictives.qnehvi import QnehviStrategy
    from bofire.data_models.strategies.random import RandomStrategy

    domain = BD.detergent_domain()
    sampler = strategies.map(RandomStrategy(domain=domain, seed=0))
    initial_samples = sampler.ask(2)
    experiments = BD.detergent_f(domain, initial_samples)
    mobo_strategy = strategies.map(QnehviStrategy(domain=domain, seed=0, num_raw_samples=256, maxiter=200))
    with pytest.raises(ValueError):
        mobo_strategy.ask(1)                      # "Not enough experiments available to execute the strategy."
    for _ in range(4):
        mobo_strategy.tell(experiments=experiments)
        candidates = mobo_strategy.ask(candidate_count=1)
        assert len(candidates) == 1
        for k in domain.inputs.get_keys():
            assert k in candidates.columns
        for k in domain.outputs.get_keys():
            assert {f"{k}_pred", f"{k}_sd", f"{k}_des"} <= set(candidates.columns)
            assert float(candidates[f"{k}_sd"].iloc[0]) > 0
        assert domain.constraints.is_fulfilled(candidates, tol=1e-6).all()
        experiments = BD.detergent_f(domain, candidates)
    assert len(mobo_strategy.experiments) == 5 and mobo_strategy.is_fitted
    vals = mobo_strategy.calc_acquisition(mobo_strategy.experiments.iloc[:3])
    assert vals.shape == (3,) and np.all(np.isfinite(vals))
    # pending candidates: add_pending joins the baseline of the next acquisition function
    c2 = mobo_strategy.ask(1, add_pending=True)
    assert mobo_strategy.num_candidates == 1
    _, X_pending = mobo_strategy.get_acqf_input_tensors()
    assert X_pending.shape == (1, 5) and np.allclose(X_pending.numpy()[0], c2[domain.inputs.get_keys()].values[0])


def test_mobo_mixed_space_literal_mixed_single_task_gp():
    """MoboStrategy data-model defaults on a space with categoricals: MixedSingleTaskGPSurrogate per output
    (data_models/strategies/predictives/botorch.py:219-228), qLogNEHVI, exhaustive categorical combinations, a sigmoid
    output constraint, a linear inequality, candidate_count = 2."""
    import everest_b200.bofire_strategy as strategies
    from bofire.data_models.strategies.predictives.mobo import MoboStrategy
    from bofire.data_models.strategies.random import RandomStrategy
    from everest_b200 import acquisition as A
    from everest_b200 import kernels as K

    domain = BD.mixed_domain()
    samples = strategies.map(RandomStrategy(domain=domain, seed=2)).ask(14)
    samples.loc[:2, "solvent"] = "m"   # measured before the level was forbidden: BoFire's strict validation wants every level seen
    experiments = BD.mixed_f(domain, samples)
    dm = MoboStrategy(domain=domain, seed=3, num_raw_samples=64, num_restarts=2, maxiter=30,
                      ref_point={"y1": -1.0, "y2": 2.0})
    assert all(type(s).__name__ == "MixedSingleTaskGPSurrogate" for s in dm.surrogate_specs.surrogates)
    strat = strategies.map(dm, fit_options={"maxiter": 30})
    strat.tell(experiments)
    for key, spec in strat.fitted_specs.items():
        flat = K.flatten(spec.kernel)
        assert len(flat.leaves) == 4 and len(flat.terms) == 3 and spec.noise >= 1e-4
    acqf = strat._get_acqfs(2)[0]
    assert isinstance(acqf, A.qLogNoisyExpectedHypervolumeImprovement) and len(acqf.constraints) == 1
    cand = strat.ask(2)
    assert len(cand) == 2 and set(cand["solvent"]) <= {"w", "e"} and set(cand["cat"]) <= {"p", "q"}
    assert set(cand["n"]) <= {1.0, 2.0, 5.0}
    assert ((cand["a"] + cand["b"]) <= 3.5 + 1e-6).all()
    pred = strat.predict(experiments)
    assert list(pred.columns) == ["y1_pred", "y2_pred", "y3_pred", "y1_sd", "y2_sd", "y3_sd", "y1_des", "y2_des", "y3_des"]
    # an exact GP with small noise reproduces its training targets
    assert np.allclose(pred["y1_pred"].values, experiments["y1"].values, atol=0.2)


def test_sobo_predictions_match_oracle_built_from_fitted_specs():
    """SoboStrategy (qLogNEI default) on Himmelblau through DataFrames; the drop-in's `predict` (Inputs.transform -> device
    posterior with observation noise) must equal the CPU oracle GP built from the SAME fitted hyper-parameters to 1e-9."""
    import everest_b200.bofire_strategy as strategies
    from bofire.data_models.strategies.predictives.sobo import SoboStrategy
    from bofire.data_models.strategies.random import RandomStrategy
    from everest_b200 import acquisition as A
    from oracle import bo_oracle as O

    domain = BD.himmelblau_domain()
    experiments = BD.himmelblau_f(domain, strategies.map(RandomStrategy(domain=domain, seed=1)).ask(30))
    strat = strategies.map(SoboStrategy(domain=domain, seed=4, num_raw_samples=128, maxiter=60), fit_options={"maxiter": 60})
    strat.tell(experiments)
    acqf = strat._get_acqfs(1)[0]
    assert isinstance(acqf, A.qLogNoisyExpectedImprovement)
    cand = strat.ask(1)
    assert -6.0 <= float(cand["x_1"].iloc[0]) <= 6.0 and -6.0 <= float(cand["x_2"].iloc[0]) <= 6.0
    assert {"y_pred", "y_sd", "y_des"} <= set(cand.columns)
    spec = strat.fitted_specs["y"]
    X = torch.as_tensor(experiments[domain.inputs.get_keys()].values, dtype=torch.double)
    y = torch.as_tensor(spec.y, dtype=torch.double)
    gp = O.GPOracle(X, [O.GPOutput(kernel=P.kernel_to_oracle(spec.kernel), in_offset=torch.as_tensor(spec.in_offset),
                                   in_scale=torch.as_tensor(spec.in_scale), mean_const=spec.mean_const, noise=spec.noise, y=y,
                                   y_mean=spec.y_mean, y_std=spec.y_std)]).factorize()
    grid = pd.DataFrame({"x_1": np.linspace(-5.5, 5.5, 23), "x_2": np.linspace(5.0, -4.0, 23)})
    pred = strat.predict(grid)
    mean, cov = gp.posterior(torch.as_tensor(grid.values, dtype=torch.double), observation_noise=True)   # [n, M], [M, n, n]
    mu_o = mean[:, 0].numpy()
    var_o = torch.diagonal(cov[0]).numpy()
    assert np.allclose(pred["y_pred"].values, mu_o, rtol=1e-9, atol=1e-9 * np.abs(mu_o).max())
    assert np.allclose(pred["y_sd"].values ** 2, var_o, rtol=1e-8, atol=1e-9 * var_o.max())
    # MinimizeObjective desirability = -(y - 0) / (1 - 0)  (get_objective / Outputs.__call__)
    assert np.allclose(pred["y_des"].values, -pred["y_pred"].values)


def test_qehvi_additive_sobo_and_nchoosek_through_the_data_models():
    """Further strategy data models on the device: QehviStrategy (fixed partitioning of the observed front), AdditiveSoboStrategy
    (weighted sum objective + sigmoid output constraint), and a domain with an NChooseK constraint (nonlinear constraint
    callables + RandomStrategy as the generator of feasible raw samples, SLSQP one restart at a time: botorch.py:117-121,
    250-265)."""
    import everest_b200.bofire_strategy as strategies
    from bofire.data_models.constraints.api import NChooseKConstraint
    from bofire.data_models.domain.domain import Domain
    from bofire.data_models.features.api import ContinuousInput, ContinuousOutput
    from bofire.data_models.objectives.api import MaximizeObjective, MaximizeSigmoidObjective, MinimizeObjective
    from bofire.data_models.strategies.predictives.qehvi import QehviStrategy
    from bofire.data_models.strategies.predictives.sobo import AdditiveSoboStrategy, SoboStrategy
    from bofire.data_models.strategies.random import RandomStrategy
    from everest_b200 import acquisition as A

    # --- qEHVI on two objectives
    dom = Domain.from_lists(
        inputs=[ContinuousInput(key=f"x{i}", bounds=[0, 1]) for i in range(3)],
        outputs=[ContinuousOutput(key="y1", objective=MaximizeObjective()), ContinuousOutput(key="y2", objective=MinimizeObjective())])

    def f(domain, c):
        X = c[domain.inputs.get_keys()].reset_index(drop=True)
        out = X.copy()
        out["y1"] = np.sin(3 * X["x0"]) + X["x1"]
        out["y2"] = (X["x0"] - 0.4) ** 2 + X["x2"]
        if "y3" in domain.outputs.get_keys():
            out["y3"] = 0.5 * X["x1"] + 0.2 * X["x2"]
        for k in domain.outputs.get_keys():
            out[f"valid_{k}"] = 1
        return out

    exp = f(dom, strategies.map(RandomStrategy(domain=dom, seed=0)).ask(12))
    strat = strategies.map(QehviStrategy(domain=dom, seed=1, num_raw_samples=64, num_restarts=2, maxiter=30, num_sobol_samples=64,
                                         ref_point={"y1": -0.5, "y2": 2.5}), fit_options={"maxiter": 40})
    strat.tell(exp)
    acqf = strat._get_acqfs(1)[0]
    assert isinstance(acqf, A.qExpectedHypervolumeImprovement) and strat.get_adjusted_refpoint() == [-0.5, -2.5]
    cand = strat.ask(1)
    assert len(cand) == 1 and all(0.0 <= float(cand[f"x{i}"].iloc[0]) <= 1.0 for i in range(3))

    # --- additive SOBO with an output constraint, qLogNEI default
    dom2 = Domain.from_lists(
        inputs=[ContinuousInput(key=f"x{i}", bounds=[0, 1]) for i in range(3)],
        outputs=[ContinuousOutput(key="y1", objective=MaximizeObjective(w=1.0)), ContinuousOutput(key="y2", objective=MinimizeObjective(w=0.5)),
                 ContinuousOutput(key="y3", objective=MaximizeSigmoidObjective(steepness=20.0, tp=0.2))])
    exp2 = f(dom2, strategies.map(RandomStrategy(domain=dom2, seed=2)).ask(14))
    s2 = strategies.map(AdditiveSoboStrategy(domain=dom2, seed=3, num_raw_samples=64, num_restarts=2, maxiter=30,
                                             use_output_constraints=True), fit_options={"maxiter": 40})
    s2.tell(exp2)
    obj, cons = s2._get_objective_and_constraints()
    assert obj.combine == "additive" and [(o.kind, o.idx, o.w) for o in obj.ops] == [("max", 0, 1.0), ("min", 1, 0.5)]
    assert len(cons) == 1 and cons[0].idx == 2
    c2 = s2.ask(2)
    assert len(c2) == 2 and {"y1_pred", "y2_pred", "y3_pred", "y3_des"} <= set(c2.columns)

    # --- NChooseK: at most 2 of the 3 inputs non-zero
    dom3 = Domain.from_lists(
        inputs=[ContinuousInput(key=f"x{i}", bounds=[0, 1]) for i in range(3)],
        outputs=[ContinuousOutput(key="y1", objective=MaximizeObjective())],
        constraints=[NChooseKConstraint(features=["x0", "x1", "x2"], min_count=1, max_count=2, none_also_valid=False)])
    exp3 = f(dom3, strategies.map(RandomStrategy(domain=dom3, seed=4)).ask(12))
    assert ((exp3[["x0", "x1", "x2"]].values > 0).sum(axis=1) <= 2).all()
    s3 = strategies.map(SoboStrategy(domain=dom3, seed=5, num_raw_samples=32, num_restarts=2, maxiter=20), fit_options={"maxiter": 40})
    s3.tell(exp3)
    assert s3._get_optimizer_options()["batch_limit"] == 1
    c3 = s3.ask(1)
    x = c3[["x0", "x1", "x2"]].values[0]
    assert (np.abs(x) > 1e-3).sum() <= 2 and (x >= -1e-9).all() and (x <= 1 + 1e-9).all()


def test_fully_combinatorial_space_through_the_data_models():
    """All inputs categorical / discrete (botorch.py:425-467): the choices are enumerated as DataFrames by BoFire's own
    `get_categorical_combinations`, the measured ones are dropped, and optimize_acqf_discrete picks unique candidates."""
    import everest_b200.bofire_strategy as strategies
    from bofire.data_models.domain.domain import Domain
    from bofire.data_models.features.api import CategoricalInput, ContinuousOutput, DiscreteInput
    from bofire.data_models.objectives.api import MaximizeObjective
    from bofire.data_models.strategies.predictives.sobo import SoboStrategy

    dom = Domain.from_lists(
        inputs=[CategoricalInput(key="c", categories=["a", "b", "c"]), DiscreteInput(key="n", values=[1.0, 2.0, 4.0, 8.0])],
        outputs=[ContinuousOutput(key="y", objective=MaximizeObjective())])
    rows = [("a", 1.0), ("b", 2.0), ("c", 4.0), ("a", 8.0), ("b", 8.0), ("c", 1.0)]
    exp = pd.DataFrame({"c": [r[0] for r in rows], "n": [r[1] for r in rows]})
    exp["y"] = [0.1, 0.5, 0.9, 0.4, 0.7, 0.2]
    exp["valid_y"] = 1
    strat = strategies.map(SoboStrategy(domain=dom, seed=0), fit_options={"maxiter": 40})
    strat.tell(exp)
    cand = strat.ask(3)
    got = set(zip(cand["c"], cand["n"]))
    assert len(got) == 3 and not (got & set(rows))                     # unique, none of them measured before
    assert set(cand["c"]) <= {"a", "b", "c"} and set(cand["n"]) <= {1.0, 2.0, 4.0, 8.0}
    assert {"y_pred", "y_sd", "y_des"} <= set(cand.columns)

"""Host logic of the BoFire drop-in (everest_b200/bofire_strategy.py) against the reference's own data models: column layout,
linear / inter-point / nonlinear constraints, fixed features and categorical combinations, kernel trees and priors of the
surrogate data models, RandomStrategy samples.  No GPU: nothing here touches DeviceGPState."""
import warnings

import numpy as np
import pytest
import torch

from tests import bofire_domains as BD

pytestmark = pytest.mark.skipif(not BD.have_bofire(), reason="bofire.data_models not importable (baseline/_ref absent)")


@pytest.fixture(autouse=True)
def _quiet():
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        yield


def test_map_refuses_strategies_outside_the_path():
    import everest_b200.bofire_strategy as S

    class FooStrategy:
        pass

    with pytest.raises(NotImplementedError):
        S.map(FooStrategy())


def test_linear_constraints_match_torch_tools_convention():
    """utils/torch_tools.py:45-100: `sum c x <= rhs` -> (indices, -c, -rhs); fixed features move to the rhs."""
    import everest_b200.bofire_strategy as S
    from bofire.data_models.constraints.api import LinearEqualityConstraint, LinearInequalityConstraint
    from bofire.data_models.domain.domain import Domain
    from bofire.data_models.features.continuous import ContinuousInput, ContinuousOutput

    dom = Domain.from_lists(
        inputs=[ContinuousInput(key="x1", bounds=[0, 1]), ContinuousInput(key="x2", bounds=[0.5, 0.5]), ContinuousInput(key="x3", bounds=[0, 2])],
        outputs=[ContinuousOutput(key="y")],
        constraints=[LinearInequalityConstraint(features=["x1", "x2", "x3"], coefficients=[1.0, 2.0, 3.0], rhs=4.0),
                     LinearEqualityConstraint(features=["x1", "x3"], coefficients=[1.0, 1.0], rhs=1.5)])
    (idx, coef, rhs), = S.get_linear_constraints(dom, "LinearInequalityConstraint")
    assert idx.tolist() == [0, 2] and coef.tolist() == [-1.0, -3.0] and rhs == pytest.approx(-(4.0 - 2.0 * 0.5))
    (idx, coef, rhs), = S.get_linear_constraints(dom, "LinearEqualityConstraint")
    assert idx.tolist() == [0, 2] and coef.tolist() == [-1.0, -1.0] and rhs == -1.5


def test_interpoint_and_nonlinear_constraints():
    import everest_b200.bofire_strategy as S
    from bofire.data_models.constraints.api import InterpointEqualityConstraint, NChooseKConstraint, ProductInequalityConstraint
    from bofire.data_models.domain.domain import Domain
    from bofire.data_models.features.continuous import ContinuousInput, ContinuousOutput

    dom = Domain.from_lists(
        inputs=[ContinuousInput(key=f"x{i}", bounds=[0, 1]) for i in range(4)], outputs=[ContinuousOutput(key="y")],
        constraints=[InterpointEqualityConstraint(feature="x1", multiplicity=2),
                     NChooseKConstraint(features=["x0", "x1", "x2"], min_count=1, max_count=2, none_also_valid=False),
                     ProductInequalityConstraint(features=["x2", "x3"], exponents=[1.0, 2.0], rhs=0.5, sign=1)])
    ip = S.get_interpoint_constraints(dom, 4)
    assert [c[0].tolist() for c in ip] == [[[0, 1], [1, 1]], [[2, 1], [3, 1]]]
    assert S.get_interpoint_constraints(dom, 1) == []
    nl = S.get_nonlinear_constraints(dom)
    assert len(nl) == 3 and all(intra for _, intra in nl)
    x = torch.tensor([[0.0, 0.4, 0.0, 0.9], [0.3, 0.4, 0.5, 0.9], [0.0, 0.0, 0.0, 0.1]], dtype=torch.double)
    vals = torch.stack([fn(x) for fn, _ in nl], dim=1)
    # row 0: 1 active feature (feasible for both NChooseK sides), product 0 <= 0.5;  row 1: 3 active (max violated);
    # row 2: none active (min violated)
    assert (vals[0] >= -1e-6).all()
    assert vals[1, 0] < 0 and vals[2, 1] < 0
    assert vals[1, 2] == pytest.approx(0.5 - 0.5 * 0.81)


def test_layout_fixed_features_and_combinations():
    import everest_b200.bofire_strategy as S
    from bofire.data_models.strategies.predictives.mobo import MoboStrategy

    dom = BD.mixed_domain()
    strat = S.map(MoboStrategy(domain=dom, seed=1))
    assert strat._features2idx == {"a": (0,), "b": (1,), "n": (2,), "cat": (3, 4), "solvent": (5, 6, 7)}
    combos = strat.get_categorical_combinations()
    # 2 allowed solvents x 2 cats x 3 discrete values; the forbidden solvent level never appears
    assert len(combos) == 12
    for ff in combos:
        assert set(ff) == {2, 3, 4, 5, 6, 7}
        assert ff[7] == 0.0 and ff[5] + ff[6] == 1.0 and ff[3] + ff[4] == 1.0 and ff[2] in (1.0, 2.0, 5.0)
    bounds, generator, nonlinear, fixed, fixed_list = strat._setup_ask()
    assert bounds.shape == (2, 8) and generator is None and nonlinear is None and fixed is None and len(fixed_list) == 12
    assert bounds[:, 1].tolist() == [-1.0, 3.0]


def test_mixed_single_task_gp_tree_is_botorchs():
    """Scale(K_c + Scale(K_h)) + Scale(K_c * K_h) on the one-hot layout, Matern-5/2 ARD on the ordinal columns, one Hamming
    lengthscale per categorical FEATURE, the continuous kernel's prior on both continuous factors."""
    import everest_b200.bofire_strategy as S
    from bofire.data_models.strategies.predictives.mobo import MoboStrategy
    from everest_b200 import fit as F
    from everest_b200 import kernels as K

    dom = BD.mixed_domain()
    dm = MoboStrategy(domain=dom, seed=1)
    strat = S.map(dm)
    sf = S._SurrogateFit(dm.surrogate_specs.surrogates[0], dom.inputs, strat.input_preprocessing_specs, strat._features2idx, 8)
    spec, ls, os_ = sf.kernel()
    flat = K.flatten(spec)
    assert len(flat.leaves) == 4 and len(flat.terms) == 3
    c1, h1, c2, h2 = flat.leaves
    assert isinstance(c1, K.MaternKernel) and c1.nu == 2.5 and list(c1.active_dims) == [0, 1, 2]
    assert isinstance(h1, K.HammingDistanceKernel) and h1.categorical_features == {3: 2, 5: 3} and len(h1.lengthscale) == 2
    assert [sorted(f) for _, f in flat.terms] == [[0], [1], [2, 3]]
    assert set(ls) == {0, 2} and isinstance(ls[0], F.GammaPrior) and (ls[0].concentration, ls[0].rate) == (3.0, 6.0)
    lay = F._Layout(spec)
    assert len(lay.scales) == 3 and lay.n_ls == 3 + 2 + 3 + 2
    off, scl = sf.scaler(BD.mixed_f(dom, S.map(_random(dom)).ask(6)))
    assert off[3:].tolist() == [0.0] * 5 and scl[3:].tolist() == [1.0] * 5     # one-hot columns are never scaled
    assert off[1] == -1.0 and scl[1] == 4.0 and scl[2] == 4.0 and off[2] == 1.0


def _random(dom, seed=5):
    from bofire.data_models.strategies.random import RandomStrategy

    return RandomStrategy(domain=dom, seed=seed)


def test_single_task_gp_default_kernel_and_priors():
    import everest_b200.bofire_strategy as S
    from bofire.data_models.strategies.predictives.qnehvi import QnehviStrategy
    from everest_b200 import fit as F
    from everest_b200 import kernels as K

    dom = BD.detergent_domain()
    dm = QnehviStrategy(domain=dom, seed=0)
    strat = S.map(dm)
    assert strat.num_sobol_samples == 512 and strat.num_raw_samples == 1024 and strat.num_restarts == 8 and strat.alpha == 0.0
    sf = S._SurrogateFit(dm.surrogate_specs.surrogates[2], dom.inputs, {}, strat._features2idx, 5)
    spec, ls, os_ = sf.kernel()
    assert isinstance(spec, K.RBFKernel) and list(spec.active_dims) == [0, 1, 2, 3, 4] and len(spec.lengthscale) == 5 and os_ == {}
    ref = F.DimensionalityScaledLogNormalPrior(5)       # Hvarfner: loc sqrt(2) + log(5) / 2, scale sqrt(3)
    assert (ls[0].loc, ls[0].scale) == (ref.loc, ref.scale)
    noise = S.map_prior(dm.surrogate_specs.surrogates[2].noise_prior)
    assert (noise.loc, noise.scale) == (-4.0, 1.0)


def test_random_strategy_respects_constraints():
    import everest_b200.bofire_strategy as S
    from bofire.data_models.constraints.api import LinearEqualityConstraint, NChooseKConstraint
    from bofire.data_models.domain.domain import Domain
    from bofire.data_models.features.continuous import ContinuousInput, ContinuousOutput

    dom = BD.detergent_domain()
    c = S.map(_random(dom)).ask(16)
    assert list(c.columns) == dom.inputs.get_keys() and len(c) == 16
    assert dom.constraints.is_fulfilled(c).all()
    s = c.sum(axis=1)
    assert (s >= 0.2 - 1e-9).all() and (s <= 0.4 + 1e-9).all() and c.drop_duplicates().shape[0] == 16
    dom2 = Domain.from_lists(
        inputs=[ContinuousInput(key=f"x{i}", bounds=[0, 1]) for i in range(4)], outputs=[ContinuousOutput(key="y")],
        constraints=[LinearEqualityConstraint(features=["x0", "x1", "x2", "x3"], coefficients=[1.0] * 4, rhs=1.0),
                     NChooseKConstraint(features=["x0", "x1", "x2", "x3"], min_count=1, max_count=2, none_also_valid=False)])
    c2 = S.map(_random(dom2, seed=7)).ask(6)
    assert np.allclose(c2.sum(axis=1), 1.0) and ((c2.values > 1e-9).sum(axis=1) <= 2).all()
    assert dom2.constraints.is_fulfilled(c2, tol=1e-6).all()


def test_objectives_and_output_constraints_from_domain():
    import everest_b200.bofire_strategy as S

    dom = BD.mixed_domain()
    exp = BD.mixed_f(dom, S.map(_random(dom)).ask(5))
    mo = S.get_multiobjective_objective(dom.outputs, exp)
    assert [(op.kind, op.idx) for op in mo.ops] == [("max", 0), ("min", 1)]
    cons = S.get_output_constraints(dom.outputs, exp)
    assert len(cons) == 1 and (cons[0].idx, cons[0].sign, cons[0].tp, cons[0].eta) == (2, -1.0, 0.2, 0.1)
    assert S.get_ref_point_mask(dom).tolist() == [1.0, -1.0]

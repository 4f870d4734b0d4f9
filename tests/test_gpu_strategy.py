"""Strategy-level mirror on the device (everest_b200/strategy.py): the README loop of BASELINE config 1 through
QnehviStrategy.tell / ask, MoboStrategy (qLogNEHVI default) and SoboStrategy (qLogNEI default), mixed and fully
combinatorial spaces with candidate_count > 1 (sequential greedy with pending points), and X_pending semantics against
the oracle."""
import numpy as np
import pytest
import torch

from everest_b200 import acquisition as A
from everest_b200 import benchmarks as B
from everest_b200 import configs as Cf
from everest_b200 import kernels as K
from everest_b200 import optim
from everest_b200.model import SingleTaskGPSpec
from everest_b200.objectives import MaximizeObjective, MinimizeObjective, MultiObjective, ScalarObjective
from everest_b200.strategy import InputSpace, MoboStrategy, QnehviStrategy, SoboStrategy
from oracle import bo_oracle as O
from tests import problems as P

pytestmark = pytest.mark.gpu
DT = torch.float64


def rbf_factory(d, ls, lo, hi, noise=1e-4):
    def make(X, Y):
        return [SingleTaskGPSpec(kernel=K.RBFKernel(list(range(d)), [ls] * d), y=Y[:, m], in_offset=lo, in_scale=hi - lo,
                                 mean_const=0.0, noise=noise) for m in range(Y.shape[1])]
    return make


def test_detergent_readme_loop_through_qnehvi_strategy():
    p = Cf.detergent_qnehvi(N=2, S=128, raw=256)
    lo, hi = p["bounds"]
    space = InputSpace(bounds=p["bounds"], inequality_constraints=p["inequality_constraints"])
    strat = QnehviStrategy(space, rbf_factory(5, 0.5, lo, hi), objective=p["objective"], n_mc_samples=128,
                           num_restarts=8, num_raw_samples=256, maxiter=200, seed=0)
    with pytest.raises(ValueError):
        strat.ask(1)                                  # "No experiments have been provided yet."
    strat.tell(p["X"], p["Y"])
    for it in range(3):
        cand, preds, stds = strat.ask(1)
        assert cand.shape == (1, 5) and preds.shape == (1, 5) and stds.shape == (1, 5) and np.all(stds > 0)
        assert np.all(cand >= lo - 1e-9) and np.all(cand <= hi + 1e-9) and 0.2 - 1e-8 <= cand.sum() <= 0.4 + 1e-8
        strat.tell(cand, B.detergent(cand))
    assert strat.X.shape == (5, 5)
    acq_vals = strat.calc_acquisition(strat.X[:3])
    assert acq_vals.shape == (3,) and np.all(np.isfinite(acq_vals))


def test_mobo_default_qlognehvi_and_candidate_count_two():
    p = Cf.zdt1_qnehvi(N=48, S=32, raw=64, d=4, q=2)
    lo, hi = p["bounds"]
    strat = MoboStrategy(InputSpace(bounds=p["bounds"]), rbf_factory(4, 0.6, lo, hi), objective=p["objective"],
                         ref_point=[1.0, 5.0], n_mc_samples=32, num_restarts=4, num_raw_samples=64, maxiter=50, seed=1)
    strat.tell(p["X"], p["Y"])
    acqf = strat._get_acqfs(2)[0]
    assert isinstance(acqf, A.qLogNoisyExpectedHypervolumeImprovement)
    cand, preds, _ = strat.ask(2)
    assert cand.shape == (2, 4) and preds.shape == (2, 2) and np.all(cand >= 0) and np.all(cand <= 1)
    # pending candidates join the baseline of the next acquisition function (get_acqf_input_tensors, botorch.py:713-722)
    strat.add_candidates(cand)
    acqf2 = strat._get_acqfs(1)[0]
    assert acqf2.nb == acqf.nb + 2


def test_sobo_default_qlognei_mixed_space_sequential():
    """2 continuous + one-hot(3): optimize_acqf_mixed over the 3 categorical combinations, candidate_count = 2 picks
    sequentially with the first point pending."""
    rng = np.random.default_rng(0)
    n = 40
    cat = np.eye(3)[rng.integers(0, 3, n)]
    Xc = rng.random((n, 2))
    X = np.concatenate([Xc, cat], axis=1)
    y = -((Xc[:, 0] - 0.3) ** 2 + (Xc[:, 1] - 0.7) ** 2) + 0.2 * cat[:, 1] + 0.01 * rng.normal(size=n)

    def factory(X_, Y_):
        kern = K.AdditiveKernel([K.ScaleKernel(K.RBFKernel([0, 1], [0.4, 0.4]), 1.0),
                                 K.ScaleKernel(K.HammingDistanceKernel({2: 3}, [1.0]), 0.5)])
        return [SingleTaskGPSpec(kernel=kern, y=Y_[:, 0], mean_const=0.0, noise=1e-3)]

    space = InputSpace(bounds=np.array([[0.0] * 5, [1.0] * 5]), categorical_groups={2: 3})
    strat = SoboStrategy(space, factory, ScalarObjective([MaximizeObjective(0)], "single"), n_mc_samples=64,
                         num_restarts=4, num_raw_samples=64, maxiter=50, seed=2)
    strat.tell(X, y[:, None])
    assert isinstance(strat._get_acqfs(1)[0], A.qLogNoisyExpectedImprovement)
    cand, preds, _ = strat.ask(2)
    assert cand.shape == (2, 5)
    assert np.allclose(cand[:, 2:].sum(axis=1), 1.0) and set(np.unique(cand[:, 2:])) <= {0.0, 1.0}
    assert not np.allclose(cand[0], cand[1])


def test_fully_combinatorial_space_discrete_greedy():
    """All inputs categorical / discrete: optimize_acqf_discrete over the unmeasured combinations (botorch.py:425-467)."""
    rng = np.random.default_rng(1)
    levels = [(a, b) for a in range(3) for b in (1.0, 2.0, 3.0, 4.0)]
    seen = [levels[i] for i in rng.choice(len(levels), 6, replace=False)]
    X = np.array([list(np.eye(3)[a]) + [b] for a, b in seen])
    y = np.array([0.5 * a - (b - 2.5) ** 2 for a, b in seen])

    def factory(X_, Y_):
        kern = K.MultiplicativeKernel([K.HammingDistanceKernel({0: 3}, [1.0]), K.RBFKernel([3], [0.5])])
        return [SingleTaskGPSpec(kernel=kern, y=Y_[:, 0], in_offset=np.array([0, 0, 0, 1.0]), in_scale=np.array([1, 1, 1, 3.0]),
                                 mean_const=0.0, noise=1e-3)]

    space = InputSpace(bounds=np.array([[0, 0, 0, 1.0], [1, 1, 1, 4.0]]), categorical_groups={0: 3},
                       discrete_values={3: [1.0, 2.0, 3.0, 4.0]})
    strat = SoboStrategy(space, factory, ScalarObjective([MaximizeObjective(0)], "single"), acquisition_function="qLogEI",
                         n_mc_samples=64, seed=3)
    strat.tell(X, y[:, None])
    cand, _, _ = strat.ask(2)
    rows = {tuple(r) for r in X}
    assert cand.shape == (2, 4) and all(tuple(c) not in rows for c in cand) and tuple(cand[0]) != tuple(cand[1])


@pytest.mark.parametrize("name", ["qLogEI", "qNEI", "qEHVI"])
def test_x_pending_is_scored_jointly_like_the_oracle(name):
    """[UPSTREAM] @concatenate_pending_points: value(X | pending) = value(cat(X, pending)) for the acquisition functions
    without a cached-baseline merge; gradients only flow to X."""
    p = Cf.zdt1_qnehvi(N=60, S=32, raw=6, d=4, q=2)
    gp = P.oracle_gp(p)
    st = Cf.build_state(p)
    # candidates and the pending point around the incumbent so that the improvement is not identically zero
    g_ = torch.Generator().manual_seed(0)
    best = torch.as_tensor(p["X"][int(torch.as_tensor(p["Y"])[:, 1].argmin())], dtype=DT)
    pts = (best.view(1, 1, -1) + 0.15 * torch.randn(3, 2, p["d"], dtype=DT, generator=g_)).clamp(0.0, 1.0)
    X, Xp = pts[:2].contiguous(), pts[2, :1].contiguous()
    if name == "qEHVI":
        Yobj = -torch.as_tensor(p["Y"], dtype=DT)
        ops = [P.op_to_oracle(o) for o in p["objective"].ops]
        acq_o = O.QEHVIOracle(gp, p["ref_point"], Yobj, ops, mc_samples=32, seed=9)
        acq_d = A.qExpectedHypervolumeImprovement(st, p["ref_point"], Yobj, p["objective"], mc_samples=32, seed=9, X_pending=Xp)
        ref = acq_o.forward(torch.cat([X, Xp.unsqueeze(0).expand(2, -1, -1)], dim=1))
    else:
        obj = ScalarObjective([MinimizeObjective(1)], "single")
        spec = ("single", P.op_to_oracle(obj.ops[0]))
        acq_o = O.QScalarOracle(gp, name, spec, p["X"], mc_samples=32, seed=5, prune_samples=64, prune_seed=5 + 7919, X_pending=Xp)
        kw = dict(prune_samples=64) if name == "qNEI" else {}
        acq_d = A.get_acquisition_function(name, st, obj, p["X"], X_pending=Xp, mc_samples=32, seed=5, **kw)
        ref = acq_o.forward(X)
    v, g = acq_d.forward_backward(X.to(st.device))
    assert g.shape == X.shape
    scale = max(1.0, float(ref.abs().max())) if name.startswith("qLog") else max(float(ref.abs().max()), 1e-12)
    assert float((v.cpu() - ref).abs().max()) < 1e-7 * scale
    acq_d.set_X_pending(None)
    v0 = acq_d(X.to(st.device)).cpu()
    assert float((v0 - v.cpu()).abs().max()) > 0          # the pending point mattered

"""End-to-end ask() pieces on the device path: the L2 boundary of SURVEY.md 8b (BotorchStrategy._optimize_acqf_continuous,
strategies/predictives/botorch.py:326-406) with the linear constraints of the Detergent README loop (BASELINE config 1)
and the analytic gradient; the chosen candidate's value is re-scored by the CPU oracle."""
import numpy as np
import pytest
import torch

from everest_b200 import benchmarks as B
from everest_b200 import configs as Cf
from everest_b200 import multiobjective as MO
from everest_b200 import optim
from tests import problems as P

pytestmark = pytest.mark.gpu
DT = torch.float64


def test_detergent_readme_loop_with_linear_constraints():
    """README.md:78-106 -- 2 random feasible experiments, then 4 x (tell, ask(1)) with QnehviStrategy; every candidate
    must satisfy 0.2 <= sum x <= 0.4 and the bounds, its acquisition value must match the oracle's, and refinement never
    returns less than the best screened raw sample."""
    p = Cf.detergent_qnehvi(N=2, S=128, raw=256)
    ineq = p["inequality_constraints"]
    bounds = torch.as_tensor(p["bounds"])
    X, Y = p["X"].copy(), p["Y"].copy()
    assert np.all(X.sum(1) >= 0.2 - 1e-12) and np.all(X.sum(1) <= 0.4 + 1e-12)
    for it in range(4):
        p["X"], p["Y"] = X, Y
        for m in range(5):
            p["outputs"][m]["y"] = Y[:, m]
        p["ref_point"] = Y.min(axis=0).tolist()
        st = Cf.build_state(p)
        acq = Cf.build_acqf(p, st, prune_samples=256)
        cand, val = optim.optimize_acqf(acq, bounds, q=1, num_restarts=8, raw_samples=p["raw_samples"],
                                        options={"maxiter": 200}, seed=it, inequality_constraints=ineq)
        assert cand.shape == (1, 5) and cand.device.type == "cpu"
        x = cand[0].numpy()
        assert np.all(x >= p["bounds"][0] - 1e-9) and np.all(x <= p["bounds"][1] + 1e-9)
        assert 0.2 - 1e-8 <= x.sum() <= 0.4 + 1e-8
        # screened raw samples come from the polytope and the refined value is at least the best of them
        X_raw = optim.sample_q_batches_from_polytope(p["raw_samples"], 1, bounds, ineq, None, seed=it)
        assert float(val) >= float(acq(X_raw.to(st.device)).max()) - 1e-15
        # the oracle agrees on the value at the chosen candidate
        acq_o = P.oracle_acqf(p, P.oracle_gp(p), prune_samples=256)
        v_o = float(acq_o.forward(cand.unsqueeze(0))[0])
        assert abs(float(val) - v_o) <= 1e-8 * max(abs(v_o), 1e-12)
        X = np.concatenate([X, x[None]], axis=0)
        Y = np.concatenate([Y, B.detergent(x[None])], axis=0)
    assert X.shape == (6, 5)


def test_analytic_and_fd_refinement_reach_the_same_optimum():
    """Smooth single-objective problem (qLogEI on Himmelblau, Matern-5/2): both gradient sources must converge to the
    same candidate -- "matched candidates" between the autograd-style and the finite-difference refinement."""
    p = Cf.himmelblau_qlogei(N=120, S=64, raw=128)
    st = Cf.build_state(p)
    acq = Cf.build_acqf(p, st)
    bounds = torch.as_tensor(p["bounds"])
    X_ic, Y_ic, _, _ = optim.gen_batch_initial_conditions(acq, bounds, 1, 4, 128, seed=0)
    Xa, Ya, ia = optim.gen_candidates_scipy(X_ic, acq, bounds[0], bounds[1], options={"maxiter": 200})
    Xf, Yf, if_ = optim.gen_candidates_scipy(X_ic, acq, bounds[0], bounds[1], options={"maxiter": 200, "gradient": "fd"})
    assert bool((Ya >= Y_ic - 1e-12).all())
    assert float((Ya - Yf).abs().max()) < 1e-6 * float(Ya.abs().max())
    assert float((Xa - Xf).abs().max()) < 1e-3 * 12.0
    assert ia["n_acqf_evals"] < if_["n_acqf_evals"]


def test_result_metrics_after_a_loop():
    """get_pareto_front / compute_hypervolume (utils/multiobjective.py:58-130) on the device after a short loop."""
    p = Cf.zdt1_qnehvi(N=64, S=32, raw=64, d=4, q=2)
    st = Cf.build_state(p)
    acq = Cf.build_acqf(p, st, prune_samples=128)
    cand, _ = optim.optimize_acqf(acq, torch.as_tensor(p["bounds"]), q=2, num_restarts=4, raw_samples=64,
                                  options={"maxiter": 30}, seed=0)
    Ynew = B.zdt1(cand.numpy())
    Yall = np.concatenate([p["Y"], Ynew], axis=0)
    hv0 = MO.compute_hypervolume(p["objective"], p["Y"], [1.0, 5.0])
    hv1 = MO.compute_hypervolume(p["objective"], Yall, [1.0, 5.0])
    assert hv1 >= hv0 - 1e-12
    front = MO.get_pareto_front(p["objective"], Yall)
    assert len(front) >= 1 and front.max() < Yall.shape[0]


def test_device_base_samples_match_the_host_engine():
    """bo_sobol_scramble / bo_sobol_normal vs torch SobolEngine + erfinv on the host: the uniform points are bit-identical
    (integer pipeline), the normal draws agree to the last ulps of the two erfinv implementations."""
    from everest_b200 import sampling

    for n_points, M, S, seed in [(1, 1, 8, 0), (7, 2, 64, 5), (300, 3, 33, 99), (2000, 2, 128, 1234)]:
        zh = sampling.base_samples(n_points, M, S, seed)
        zd = sampling.base_samples_device(n_points, M, S, seed, "cuda:0").cpu()
        assert zd.shape == zh.shape == (S, n_points, M)
        diff = (zd - zh).abs()
        assert bool(torch.isfinite(zd).all())
        # the uniform points are identical; erfinv(2v - 1) is ill-conditioned in the tails (d z / d x ~ exp(z^2 / 2)), so the
        # few-ulp difference between CUDA's and torch's erfinv grows from 1e-15 at z ~ 1 to ~1e-11 at |z| ~ 6
        assert float(diff.max()) < 1e-9, float(diff.max())
        assert float(diff[zh.abs() < 3.0].max()) < 1e-13, float(diff[zh.abs() < 3.0].max())
    assert sampling.base_samples_device(0, 2, 8, 1, "cuda:0").shape == (8, 0, 2)
    # uniform raw samples of optimize_acqf: bit-identical to the host engine (including torch's float32 first point)
    for dim, n, seed in [(1, 5, 0), (120, 257, 7), (960, 64, 123)]:
        u_h = torch.quasirandom.SobolEngine(dim, scramble=True, seed=seed).draw(n, dtype=DT)
        u_d = sampling.sobol_uniform_device(dim, n, seed, "cuda:0").cpu()
        assert torch.equal(u_d, u_h)
    b_ = torch.tensor([[0.0, -1.0, 2.0], [1.0, 1.0, 5.0]], dtype=DT)
    assert torch.equal(optim.draw_sobol_samples(b_, 33, 2, seed=4, device="cuda:0").cpu(), optim.draw_sobol_samples(b_, 33, 2, seed=4))


@pytest.mark.parametrize("kind", ["qnehvi", "qlogei"])
def test_refined_candidates_match_the_oracle_driven_optimiser(kind):
    """"Matched candidates" (north star): gen_candidates_scipy (L-BFGS-B over all restarts, botorch.py:384-405) run twice from
    IDENTICAL initial conditions -- once on the device acquisition function with its analytic adjoint kernels, once on the
    CPU oracle with torch autograd (what BoTorch does) -- must arrive at the same candidates.  Also with two linear
    inequality constraints (SLSQP, the Detergent case)."""
    from everest_b200 import optim
    from tests import problems as P

    if kind == "qnehvi":
        p = Cf.zdt1_qnehvi(N=64, S=32, raw=64, d=4, q=2)
    else:
        p = Cf.himmelblau_qlogei(N=80, S=64, raw=64)
    gp = P.oracle_gp(p)
    st = Cf.build_state(p)
    acq_o = P.oracle_acqf(p, gp, prune_samples=128) if kind == "qnehvi" else P.oracle_acqf(p, gp)
    acq_d = Cf.build_acqf(p, st, prune_samples=128) if kind == "qnehvi" else Cf.build_acqf(p, st)
    adapter = P.OracleAcqfAdapter(acq_o, p["d"])
    bounds = torch.as_tensor(p["bounds"])
    torch.manual_seed(3)
    X_ic, Y_ic, _, _ = optim.gen_batch_initial_conditions(acq_d, bounds, p["q"], 4, 64, seed=11)
    width = float((bounds[1] - bounds[0]).max())
    # (1) the two L-BFGS-B trajectories stay together: after 8 iterations from the same start the iterates agree to 1e-6
    X8d, _, _ = optim.gen_candidates_scipy(X_ic, acq_d, bounds[0], bounds[1], options={"maxiter": 8})
    X8o, _, _ = optim.gen_candidates_scipy(X_ic, adapter, bounds[0], bounds[1], options={"maxiter": 8})
    assert float((X8d - X8o).abs().max()) < 1e-6 * width
    # (2) at convergence both arrive at the same optimum: the same acquisition value to 1e-7, whichever path scores whichever
    # candidate, and the same coordinates wherever the value depends on them (a point of a q-batch that adds nothing to the
    # hypervolume improvement sits on a flat direction: its final position is decided by the last ulps of the line search)
    opts = {"maxiter": 200}
    Xd, Yd, info_d = optim.gen_candidates_scipy(X_ic, acq_d, bounds[0], bounds[1], options=opts)
    Xo, Yo, info_o = optim.gen_candidates_scipy(X_ic, adapter, bounds[0], bounds[1], options=opts)
    scale = max(1.0, float(Yo.abs().max())) if kind == "qlogei" else float(Yo.abs().max())
    assert float((Yd - Yo).abs().max()) < 1e-7 * scale, (info_d, info_o)
    assert float((adapter(Xd) - Yo).abs().max()) < 1e-7 * scale and float((acq_d(Xo.to(st.device)).cpu() - Yd).abs().max()) < 1e-7 * scale
    _, g_o = adapter.forward_backward(Xo)
    sensitive = g_o.abs() > 1e-3 * float(g_o.abs().max())
    if bool(sensitive.any()):
        assert float((Xd - Xo)[sensitive].abs().max()) < 1e-4 * width
    close = (Xd - Xo).abs().amax(dim=(1, 2)) < 1e-5 * width
    assert int(close.sum()) >= X_ic.shape[0] - 1, (Xd - Xo).abs().amax(dim=(1, 2))
    assert bool((Yd >= Y_ic - 1e-12).all())
    if kind == "qnehvi":
        # SLSQP with linear inequality constraints 0.5 <= sum_j x_j <= 2.5 on every point
        d = p["d"]
        ineq = [(torch.arange(d), torch.ones(d, dtype=torch.double), 0.5), (torch.arange(d), -torch.ones(d, dtype=torch.double), -2.5)]
        X0 = optim.sample_q_batches_from_polytope(3, p["q"], bounds, ineq, None, seed=2, n_burnin=512, n_thinning=4)
        Xd2, Yd2, _ = optim.gen_candidates_scipy(X0, acq_d, bounds[0], bounds[1], options={"maxiter": 100}, inequality_constraints=ineq)
        Xo2, Yo2, _ = optim.gen_candidates_scipy(X0, adapter, bounds[0], bounds[1], options={"maxiter": 100}, inequality_constraints=ineq)
        assert float((Xd2 - Xo2).abs().max()) < 1e-5 * width
        assert float((Yd2 - Yo2).abs().max()) < 1e-6 * float(Yo2.abs().max())
        assert bool(optim.linear_feasibility(Xd2, ineq, None).all())

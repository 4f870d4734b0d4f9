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

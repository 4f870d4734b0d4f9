"""GPU parity (value AND analytic gradient) of every acquisition function the Sobo / Mobo data models can select
(data_models/acquisition_functions/acquisition_function.py:21-95; built at sobo.py:64-89 and mobo.py:72-90) against the
float64 CPU oracle: qLogNEHVI (MoboStrategy's default), qLogEHVI, qLogNEI (SoboStrategy's default), qNEI, qEI, qLogEI,
qSR, qUCB, qPI.  Tolerances: values 1e-8 (relative to the largest value of the call; absolute for log-space values),
gradients 1e-6 relative to the largest gradient entry."""
import pytest
import torch

from everest_b200 import acquisition as A
from everest_b200 import configs as Cf
from everest_b200 import kernels as K
from everest_b200.objectives import MaximizeObjective, MinimizeObjective, OutputConstraint, ScalarObjective
from oracle import bo_oracle as O
from tests import problems as P

pytestmark = pytest.mark.gpu
DT = torch.float64


def value_and_grad(acq_o, X):
    Xr = X.detach().clone().requires_grad_(True)
    v = acq_o.forward(Xr)
    v.sum().backward()
    return v.detach(), Xr.grad.detach()


def compare(acq_d, acq_o, X, st, log_space, val_tol=1e-8, grad_tol=1e-6):
    v_o, g_o = value_and_grad(acq_o, X)
    v_f = acq_d(X.to(st.device)).cpu()
    v_d, g_d = acq_d.forward_backward(X.to(st.device))
    v_d, g_d = v_d.cpu(), g_d.cpu()
    assert bool(torch.isfinite(v_o).all()) and bool(torch.isfinite(v_f).all())
    well = torch.ones_like(v_o, dtype=torch.bool)
    if log_space:
        # log(sum_odd - sum_even) far in the fat tails (improvement ~ e^-40) cancels to the last few digits on BOTH paths
        # (the inclusion-exclusion terms agree to ~1e-13 there); such q-batches are compared loosely
        well = v_o > -30.0
        assert bool(well.any())
        assert float((v_f - v_o)[~well].abs().max() if bool((~well).any()) else 0.0) < 5e-2
    scale = max(1.0, float(v_o[well].abs().max())) if log_space else max(float(v_o.abs().max()), 1e-300)
    assert float((v_f - v_o)[well].abs().max()) < val_tol * scale
    assert float((v_d - v_f)[well].abs().max()) <= 1e-11 * max(1.0, float(v_o[well].abs().max()))
    gs = float(g_o[well].abs().max())
    assert gs > 0
    err = float((g_d - g_o)[well].abs().max())
    assert err < grad_tol * gs, (err, gs)


@pytest.mark.parametrize("kind,q", [("zdt1", 1), ("zdt1", 3), ("dtlz2", 2), ("dtlz2", 4)])
def test_qlognehvi_value_and_gradient(kind, q):
    p = Cf.zdt1_qnehvi(N=80, S=32, raw=8, d=5, q=q) if kind == "zdt1" else Cf.dtlz2_qnehvi(N=50, S=16, raw=6, d=5, m_obj=3, q=q)
    gp = P.oracle_gp(p)
    st = Cf.build_state(p)
    ops = [P.op_to_oracle(o) for o in p["objective"].ops]
    acq_o = O.QLogNEHVIOracle(gp, p["ref_point"], p["X"], ops, mc_samples=p["S"], seed=p["sampler_seed"],
                              prune_baseline=True, prune_samples=128, prune_seed=p["sampler_seed"] + 7919)
    acq_d = A.get_acquisition_function("qLogNEHVI", st, p["objective"], p["X"], ref_point=p["ref_point"],
                                       mc_samples=p["S"], seed=p["sampler_seed"], prune_samples=128)
    assert isinstance(acq_d, A.qLogNoisyExpectedHypervolumeImprovement)
    assert acq_d.prune_idx.cpu().tolist() == acq_o.prune_idx.tolist()
    X = Cf.candidates(p)
    compare(acq_d, acq_o, X, st, log_space=True)
    # exp(qLogNEHVI) is the smoothed version of qNEHVI: close to it, never far above
    plain = A.qNoisyExpectedHypervolumeImprovement(st, p["ref_point"], p["X"], p["objective"], prune_baseline=True,
                                                   mc_samples=p["S"], seed=p["sampler_seed"], prune_samples=128)
    v_plain = plain(X.to(st.device)).cpu()
    with pytest.raises(Exception):
        acq_d(X.to(st.device))     # the state now holds `plain`: the stale object fails loudly instead of silently
    big = v_plain > 1e-3
    if bool(big.any()):
        v_log = value_and_grad(acq_o, X)[0].exp()
        assert float(((v_log[big] - v_plain[big]).abs() / v_plain[big]).max()) < 0.2


def test_qlognehvi_with_constraint_and_pending():
    p = Cf.zdt1_qnehvi(N=60, S=24, raw=6, d=5, q=2)
    y3 = p["X"][:, 2] + 0.1 * p["X"][:, 3]
    p["outputs"].append(dict(kernel=K.RBFKernel(list(range(5)), [0.7] * 5), y=y3, noise=1e-3, mean_const=0.1))
    cons = [OutputConstraint(2, 1.0, 0.6, 0.25)]
    gp = P.oracle_gp(p)
    st = Cf.build_state(p)
    Xp = Cf.candidates(p, 2)[0]
    ops = [P.op_to_oracle(o) for o in p["objective"].ops]
    acq_o = O.QLogNEHVIOracle(gp, p["ref_point"], p["X"], ops, constraints=[(2, 1.0, 0.6, 0.25)], mc_samples=p["S"],
                              seed=p["sampler_seed"], prune_baseline=True, prune_samples=128,
                              prune_seed=p["sampler_seed"] + 7919, X_pending=Xp)
    acq_d = A.qLogNoisyExpectedHypervolumeImprovement(st, p["ref_point"], p["X"], p["objective"], constraints=cons,
                                                      prune_baseline=True, X_pending=Xp, mc_samples=p["S"],
                                                      seed=p["sampler_seed"], prune_samples=128)
    compare(acq_d, acq_o, Cf.candidates(p), st, log_space=True)


def test_qlognehvi_gradient_both_adjoint_kernels():
    """Few q-batches (< 148: refinement) run the sample-split / cell-lane adjoint (mc_loghvi_grad_cl_kernel + finishing kernel),
    many q-batches the one-CTA-per-q-batch kernel: both against oracle autograd, and against each other on shared q-batches
    (the lanes' running sums are merged in a different order: agreement to rounding, not bit for bit)."""
    p = Cf.zdt1_qnehvi(N=80, S=32, raw=160, d=5, q=2)
    p["outputs"].append(dict(kernel=K.RBFKernel(list(range(5)), [0.7] * 5), y=p["X"][:, 2] + 0.1 * p["X"][:, 3], noise=1e-3,
                             mean_const=0.1))
    cons = [OutputConstraint(2, 1.0, 0.6, 0.25)]
    gp = P.oracle_gp(p)
    st = Cf.build_state(p)
    ops = [P.op_to_oracle(o) for o in p["objective"].ops]
    acq_o = O.QLogNEHVIOracle(gp, p["ref_point"], p["X"], ops, constraints=[(2, 1.0, 0.6, 0.25)], mc_samples=p["S"],
                              seed=p["sampler_seed"], prune_baseline=True, prune_samples=128, prune_seed=p["sampler_seed"] + 7919)
    acq_d = A.qLogNoisyExpectedHypervolumeImprovement(st, p["ref_point"], p["X"], p["objective"], constraints=cons,
                                                      prune_baseline=True, mc_samples=p["S"], seed=p["sampler_seed"],
                                                      prune_samples=128)
    X = Cf.candidates(p)
    assert X.shape[0] == 160
    compare(acq_d, acq_o, X, st, log_space=True)                    # 160 q-batches: one CTA per q-batch
    compare(acq_d, acq_o, X[:7], st, log_space=True)                # 7 q-batches: samples split, cells over lanes
    v_all, g_all = acq_d.forward_backward(X.to(st.device))
    v_few, g_few = acq_d.forward_backward(X[:7].to(st.device))
    assert float((v_few - v_all[:7]).abs().max()) <= 1e-11 * max(1.0, float(v_all.abs().max()))
    assert float((g_few - g_all[:7]).abs().max()) <= 1e-9 * float(g_all.abs().max())


def test_qlogehvi_value_and_gradient():
    p = Cf.zdt1_qnehvi(N=80, S=32, raw=8, d=5, q=2)
    gp = P.oracle_gp(p)
    st = Cf.build_state(p)
    ops = [P.op_to_oracle(o) for o in p["objective"].ops]
    Yobj = -torch.as_tensor(p["Y"], dtype=DT)
    acq_o = O.QLogEHVIOracle(gp, p["ref_point"], Yobj, ops, mc_samples=32, seed=9)
    acq_d = A.get_acquisition_function("qLogEHVI", st, p["objective"], p["X"], ref_point=p["ref_point"], Y=p["Y"],
                                       mc_samples=32, seed=9)
    compare(acq_d, acq_o, Cf.candidates(p), st, log_space=True)


@pytest.mark.parametrize("name", ["qEI", "qLogEI", "qSR", "qUCB", "qPI", "qNEI", "qLogNEI"])
@pytest.mark.parametrize("q", [1, 3])
def test_sobo_acquisition_functions_value_and_gradient(name, q):
    p = Cf.zdt1_qnehvi(N=70, S=64, raw=8, d=4, q=q)
    obj = ScalarObjective([MaximizeObjective(0, w=0.3), MinimizeObjective(1, w=0.7)], "additive")
    gp = P.oracle_gp(p)
    st = Cf.build_state(p)
    spec = ("additive", [(P.op_to_oracle(o), o.w) for o in obj.ops])
    tau = 0.05  # qPI: wide enough that the sigmoid has a usable gradient on this problem
    acq_o = O.QScalarOracle(gp, name, spec, p["X"], mc_samples=64, seed=5, beta=0.3, tau=tau, prune_samples=128,
                            prune_seed=5 + 7919)
    kw = dict(prune_samples=128) if name in ("qNEI", "qLogNEI") else {}
    acq_d = A.get_acquisition_function(name, st, obj, p["X"], mc_samples=64, seed=5, beta=0.3, tau=tau, **kw)
    if name in ("qNEI", "qLogNEI"):
        assert acq_d.prune_idx.cpu().tolist() == acq_o.prune_idx.tolist() and acq_d.nb == acq_o.nb
    elif name not in ("qSR", "qUCB"):
        assert abs(acq_d.best_f - acq_o.best_f) < 1e-9 * max(1.0, abs(acq_o.best_f))
    # candidates around the incumbent so that the improvement-based utilities are not identically zero
    g = torch.Generator().manual_seed(0)
    best = torch.as_tensor(p["X"][int(torch.as_tensor(p["Y"])[:, 1].argmin())], dtype=DT)
    X = (best.view(1, 1, -1) + 0.15 * torch.randn(8, q, p["d"], dtype=DT, generator=g)).clamp(0.0, 1.0)
    compare(acq_d, acq_o, X, st, log_space=name.startswith("qLog"), val_tol=1e-7 if name.startswith("qLog") else 1e-8)


@pytest.mark.parametrize("name", ["qEI", "qLogEI", "qPI", "qLogNEI"])
def test_sobo_acquisition_functions_with_output_constraint(name):
    p = Cf.zdt1_qnehvi(N=60, S=32, raw=8, d=4, q=2)
    obj = ScalarObjective([MinimizeObjective(1)], "single")
    cons = [OutputConstraint(0, 1.0, 0.5, 0.1)]      # y0 <= 0.5, eta = 0.1
    gp = P.oracle_gp(p)
    st = Cf.build_state(p)
    spec = ("single", P.op_to_oracle(obj.ops[0]))
    acq_o = O.QScalarOracle(gp, name, spec, p["X"], mc_samples=32, seed=3, tau=0.05, prune_samples=128, prune_seed=3 + 7919,
                            constraints=[(0, 1.0, 0.5, 0.1)])
    kw = dict(prune_samples=128) if name == "qLogNEI" else {}
    acq_d = A.get_acquisition_function(name, st, obj, p["X"], constraints=cons, mc_samples=32, seed=3, tau=0.05, **kw)
    g = torch.Generator().manual_seed(1)
    best = torch.as_tensor(p["X"][int(torch.as_tensor(p["Y"])[:, 1].argmin())], dtype=DT)
    X = (best.view(1, 1, -1) + 0.15 * torch.randn(6, 2, p["d"], dtype=DT, generator=g)).clamp(0.0, 1.0)
    compare(acq_d, acq_o, X, st, log_space=name.startswith("qLog"), val_tol=1e-7 if name.startswith("qLog") else 1e-8)


@pytest.mark.parametrize("name", ["qLogEI", "qEI", "qLogNEI", "qNEI"])
@pytest.mark.parametrize("tp", [0.5, -1.0])
def test_constrained_incumbent_is_the_best_feasible_point(name, tp):
    """[UPSTREAM] compute_best_feasible_objective / prune_inferior_points with output constraints (SoboStrategy hands its
    sigmoid / target outputs over as constraints, sobo.py:120-150): the incumbent is the best FEASIBLE point -- on ZDT1 the
    points with the best objective (smallest y1) have y0 near 1 and violate y0 <= 0.5 -- and infeasible samples cannot
    survive the pruning.  tp = -1: no baseline point is feasible, BoTorch's pessimistic lower bound (objective of
    mean - 6 sd at 32 random convex combinations, shared generator) becomes the incumbent."""
    p = Cf.zdt1_qnehvi(N=60, S=32, raw=8, d=4, q=2)
    obj = ScalarObjective([MinimizeObjective(1)], "single")
    cons = [OutputConstraint(0, 1.0, tp, 0.1)]
    gp = P.oracle_gp(p)
    st = Cf.build_state(p)
    spec = ("single", P.op_to_oracle(obj.ops[0]))
    noisy = name in ("qLogNEI", "qNEI")
    acq_o = O.QScalarOracle(gp, name, spec, p["X"], mc_samples=32, seed=3, prune_samples=128, prune_seed=3 + 7919,
                            constraints=[(0, 1.0, tp, 0.1)], lb_generator=torch.Generator().manual_seed(77))
    kw = dict(prune_samples=128) if noisy else {}
    acq_d = A.get_acquisition_function(name, st, obj, p["X"], constraints=cons, mc_samples=32, seed=3,
                                       lb_generator=torch.Generator().manual_seed(77), **kw)
    unconstrained = float(obj(st.posterior(p["X"])[0].cpu()).max())
    if noisy:
        assert acq_d.prune_idx.cpu().tolist() == acq_o.prune_idx.tolist()
        bfs = st.debug_get("best_f_s").cpu()
        assert float((bfs - acq_o.best_f_s).abs().max()) < 1e-9 * max(1.0, float(acq_o.best_f_s.abs().max()))
        assert (acq_d.n_all_infeasible > 0) == (tp < 0)
        if tp < 0:   # every sample takes the same pessimistic value (BoTorch: objective(mean - 6 sd), clamped at 0)
            assert float(bfs.max()) == float(bfs.min()) <= 0.0
    else:
        assert abs(acq_d.best_f - acq_o.best_f) < 1e-9 * max(1.0, abs(acq_o.best_f))
        assert acq_d.best_f < unconstrained - 1e-3 if tp > 0 else acq_d.best_f <= 0.0
    g = torch.Generator().manual_seed(1)
    X = torch.rand(6, 2, p["d"], dtype=DT, generator=g)
    v_o = acq_o.forward(X)
    v_d = acq_d(X.to(st.device)).cpu()
    scale = max(1.0, float(v_o.abs().max())) if name.startswith("qLog") else max(float(v_o.abs().max()), 1e-300)
    assert float((v_d - v_o).abs().max()) < 1e-7 * scale


def test_factory_errors_are_loud():
    p = Cf.zdt1_qnehvi(N=40, S=16, raw=4, d=4, q=1)
    st = Cf.build_state(p)
    obj = ScalarObjective([MinimizeObjective(1)], "single")
    with pytest.raises(NotImplementedError):
        A.get_acquisition_function("qKG", st, obj, p["X"])
    with pytest.raises(ValueError):
        A.get_acquisition_function("qNEHVI", st, p["objective"], p["X"])          # ref_point missing
    with pytest.raises(ValueError):                                              # constraints need a utility >= 0
        A._ScalarAcquisition.__init__(A.qSimpleRegret.__new__(A.qSimpleRegret), st, obj, best_f=0.0, mc_samples=16,
                                      constraints=[OutputConstraint(0, 1.0, 0.5, 0.1)])
    with pytest.raises(ValueError):
        A.qProbabilityOfImprovement(st, 0.0, obj, tau=0.0)                       # tau must be > 0


def test_activate_and_optimize_acqf_list_on_one_state():
    """Several acquisition functions built on ONE DeviceGPState (the handle holds one prepared function at a time):
    activate() re-prepares with the stored arguments and reproduces the values bit for bit; optimize_acqf_list
    (botorch.py:337-356) generates one candidate per function with the earlier candidates pending."""
    from everest_b200 import _lib as L
    from everest_b200 import optim

    p = Cf.zdt1_qnehvi(N=40, S=32, raw=16, d=3, q=1)
    st = Cf.build_state(p)
    obj = ScalarObjective([MinimizeObjective(1)], "single")
    a1 = A.get_acquisition_function("qLogNEI", st, obj, p["X"], mc_samples=32, seed=5)
    X = Cf.candidates(p)[:8].to(st.device)
    v1 = a1(X).clone()
    a2 = A.get_acquisition_function("qUCB", st, obj, p["X"], mc_samples=32, seed=6, beta=0.3)
    v2 = a2(X).clone()
    with pytest.raises(L.EverestError):
        a1(X)                                   # a2 owns the handle now
    assert torch.equal(a1.activate()(X), v1)
    assert torch.equal(a2.activate()(X), v2)
    bounds = torch.tensor([[0.0] * 3, [1.0] * 3])
    cands, vals = optim.optimize_acqf_list([a1, a2], bounds, num_restarts=3, raw_samples=32, options={"maxiter": 30}, seed=2)
    assert cands.shape == (2, 3) and vals.shape == (2,) and bool(torch.isfinite(vals).all())
    # the second value is a2 at its candidate with the first candidate pending
    a2.activate()
    a2.set_X_pending(cands[:1])
    assert abs(float(a2(cands[1:2].unsqueeze(0).to(st.device))[0]) - float(vals[1])) <= 1e-9 * max(1.0, abs(float(vals[1])))
    a2.set_X_pending(None)
    # NEHVI-family functions re-activate too (pending points live in the baseline)
    n1 = Cf.build_acqf(p, st, prune_samples=64)
    Xq = Cf.candidates(p)[:4].to(st.device)
    w1 = n1(Xq).clone()
    a1.activate()
    assert torch.equal(n1.activate()(Xq), w1)

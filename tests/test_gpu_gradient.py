"""GPU parity of the analytic adjoint (bo_acqf_forward_backward, csrc/grad.cu) against torch autograd through
the float64 CPU oracle -- the same gradient BoTorch's gen_candidates_scipy obtains by back-propagating through
the acquisition function (reached from strategies/predictives/botorch.py:384-405).

Tolerance: gradients are float64 on both sides; 1e-6 relative to the largest gradient entry of the call (the
adjoint runs through (K + s2 I)^-1 explicitly, the oracle through two triangular solves)."""
import pytest
import torch

from everest_b200 import acquisition as A
from everest_b200 import configs as Cf
from everest_b200 import kernels as K
from everest_b200.objectives import (MaximizeObjective, MinimizeObjective, OutputConstraint, ScalarObjective)
from oracle import bo_oracle as O
from tests import problems as P

pytestmark = pytest.mark.gpu
DT = torch.float64
GRAD_TOL = 1e-6


def oracle_value_and_grad(acq_o, X, **kw):
    Xr = X.detach().clone().requires_grad_(True)
    v = acq_o.forward(Xr, **kw)
    v.sum().backward()
    return v.detach(), Xr.grad.detach()


def check(acq_d, acq_o, X, st, val_tol=1e-8, grad_tol=GRAD_TOL):
    v_o, g_o = oracle_value_and_grad(acq_o, X)
    v_d, g_d = acq_d.forward_backward(X.to(st.device))
    v_f = acq_d(X.to(st.device))
    vs = max(float(v_o.abs().max()), 1e-300)
    assert float((v_d.cpu() - v_o).abs().max()) < val_tol * vs
    # the value returned next to the gradient is the forward value (other kernel, same arithmetic)
    assert float((v_d - v_f).abs().max()) <= 1e-12 * vs
    gs = float(g_o.abs().max())
    assert gs > 0
    assert float((g_d.cpu() - g_o).abs().max()) < grad_tol * gs, (float((g_d.cpu() - g_o).abs().max()), gs)
    return v_d, g_d


@pytest.mark.parametrize("kind,q", [("zdt1", 1), ("zdt1", 3), ("zdt1", 4), ("dtlz2", 2), ("dtlz2", 5)])
def test_qnehvi_gradient_matches_oracle_autograd(kind, q):
    if kind == "zdt1":
        p = Cf.zdt1_qnehvi(N=96, S=32, raw=10, d=6, q=q)
    else:
        p = Cf.dtlz2_qnehvi(N=60, S=16, raw=8, d=5, m_obj=3, q=q)
    gp = P.oracle_gp(p)
    st = Cf.build_state(p)
    acq_o = P.oracle_acqf(p, gp, prune_samples=128)
    acq_d = Cf.build_acqf(p, st, prune_samples=128)
    assert acq_d.nb == acq_o.nb and acq_d.nb > 0
    check(acq_d, acq_o, Cf.candidates(p), st)


def test_qnehvi_gradient_with_constraint_and_pending():
    p = Cf.zdt1_qnehvi(N=60, S=24, raw=8, d=5, q=2)
    y3 = p["X"][:, 2] + 0.1 * p["X"][:, 3]
    p["outputs"].append(dict(kernel=K.RBFKernel(list(range(5)), [0.7] * 5), y=y3, noise=1e-3, mean_const=0.1))
    cons = [OutputConstraint(2, 1.0, 0.6, 0.25)]
    gp = P.oracle_gp(p)
    st = Cf.build_state(p)
    Xp = Cf.candidates(p, 2)[0]
    ops = [P.op_to_oracle(o) for o in p["objective"].ops]
    acq_o = O.QNEHVIOracle(gp, p["ref_point"], p["X"], ops, constraints=[(2, 1.0, 0.6, 0.25)], mc_samples=p["S"],
                           seed=p["sampler_seed"], prune_baseline=True, prune_samples=128,
                           prune_seed=p["sampler_seed"] + 7919, X_pending=Xp)
    acq_d = A.qNoisyExpectedHypervolumeImprovement(st, p["ref_point"], p["X"], p["objective"], constraints=cons,
                                                   prune_baseline=True, X_pending=Xp, mc_samples=p["S"],
                                                   seed=p["sampler_seed"], prune_samples=128)
    check(acq_d, acq_o, Cf.candidates(p), st)


def test_qnehvi_gradient_empty_baseline():
    p = Cf.zdt1_qnehvi(N=40, S=16, raw=6, d=4, q=2)
    gp = P.oracle_gp(p)
    st = Cf.build_state(p)
    ops = [P.op_to_oracle(o) for o in p["objective"].ops]
    X0 = torch.zeros(0, p["d"], dtype=DT)
    acq_o = O.QNEHVIOracle(gp, p["ref_point"], X0, ops, mc_samples=16, seed=3, prune_baseline=False)
    acq_d = A.qNoisyExpectedHypervolumeImprovement(st, p["ref_point"], X0, p["objective"], prune_baseline=False,
                                                   mc_samples=16, seed=3)
    check(acq_d, acq_o, Cf.candidates(p), st)


def test_qehvi_gradient():
    p = Cf.zdt1_qnehvi(N=96, S=32, raw=10, d=6, q=3)
    gp = P.oracle_gp(p)
    st = Cf.build_state(p)
    Yobj = -torch.as_tensor(p["Y"], dtype=DT)
    acq_o = O.QEHVIOracle(gp, p["ref_point"], Yobj, [P.op_to_oracle(o) for o in p["objective"].ops], mc_samples=32, seed=9)
    acq_d = A.qExpectedHypervolumeImprovement(st, p["ref_point"], Yobj, p["objective"], mc_samples=32, seed=9)
    check(acq_d, acq_o, Cf.candidates(p, 7), st)


@pytest.mark.parametrize("kind", ["himmelblau", "mixed"])
def test_qlogei_gradient(kind):
    p = Cf.himmelblau_qlogei(N=150, S=64, raw=12) if kind == "himmelblau" else \
        Cf.mixed_tanimoto_qlogei(N=130, n_bits=200, S=32, n_choices=9)
    gp = P.oracle_gp(p)
    st = Cf.build_state(p)
    acq_o = P.oracle_acqf(p, gp)
    acq_d = Cf.build_acqf(p, st)
    X = Cf.candidates(p)
    v_o, g_o = oracle_value_and_grad(acq_o, X)
    v_d, g_d = acq_d.forward_backward(X.to(st.device))
    assert float((v_d.cpu() - v_o).abs().max()) < 1e-7 * float(v_o.abs().max())
    if kind == "mixed":
        # one-hot and 0/1 fingerprint columns are discrete on the device path (bit-packed, fixed features of the
        # optimisation, botorch.py:358-378): zero gradient there; the oracle's float64 Tanimoto is smooth in them, so
        # only the continuous columns (0, 1) are compared with autograd
        assert float(g_d[..., 2:].abs().max()) == 0.0
        g_o, g_d = g_o[..., :2], g_d[..., :2]
    gs = float(g_o.abs().max())
    assert gs > 0
    assert float((g_d.cpu() - g_o).abs().max()) < GRAD_TOL * gs


@pytest.mark.parametrize("combine", ["additive", "multiplicative"])
def test_qlogei_q_batch_combined_objective_gradient(combine):
    p = Cf.zdt1_qnehvi(N=50, S=32, raw=6, d=4, q=3)
    if combine == "multiplicative":  # factors must stay positive under the fractional powers
        obj = ScalarObjective([MaximizeObjective(0, -3.0, 2.0, w=0.4), MinimizeObjective(1, 12.0, 13.0, w=0.6)], combine)
    else:
        obj = ScalarObjective([MaximizeObjective(0, w=0.4), MinimizeObjective(1, w=0.6)], combine)
    gp = P.oracle_gp(p)
    st = Cf.build_state(p)
    spec = (combine, [(P.op_to_oracle(o), o.w) for o in obj.ops])
    acq_o = O.QLogEIOracle(gp, spec, p["X"], mc_samples=32, seed=5)
    acq_d = A.get_acquisition_function("qLogEI", st, obj, p["X"], mc_samples=32, seed=5)
    X = Cf.candidates(p)
    v_o, g_o = oracle_value_and_grad(acq_o, X)
    v_d, g_d = acq_d.forward_backward(X.to(st.device))
    assert float((v_d.cpu() - v_o).abs().max()) < 1e-7 * float(v_o.abs().max())
    assert float((g_d.cpu() - g_o).abs().max()) < GRAD_TOL * float(g_o.abs().max())


@pytest.mark.parametrize("nu", [0.5, 1.5, 2.5])
def test_matern_scale_additive_kernel_gradient(nu):
    """Composite tree Scale(Matern) + Scale(RBF * Matern) on disjoint / shared dims: exercises the per-leaf coefficient
    of the sum-of-products tree in kernel_grad_kernel."""
    p = Cf.zdt1_qnehvi(N=70, S=16, raw=6, d=5, q=2)
    for o in p["outputs"]:
        o["kernel"] = K.AdditiveKernel([
            K.ScaleKernel(K.MaternKernel([0, 1, 2, 3, 4], [0.9, 1.1, 0.8, 1.3, 1.0], nu=nu), 0.7),
            K.ScaleKernel(K.MultiplicativeKernel([K.RBFKernel([0, 1], [0.6, 0.9]), K.MaternKernel([2, 3, 4], [1.2], nu=2.5)]), 0.5),
        ])
    gp = P.oracle_gp(p)
    st = Cf.build_state(p)
    acq_o = P.oracle_acqf(p, gp, prune_samples=64)
    acq_d = Cf.build_acqf(p, st, prune_samples=64)
    # nu = 0.5 is not smooth at r = 0: for baseline points that ARE training points the quadratic-expansion distance is
    # rounding noise (1e-16), its square root 1e-8, and the baseline posterior covariance (7e-6 here) inherits that
    # noise on BOTH paths (gpytorch's formula has the same property) -> samples agree to ~1e-6 only
    tol = dict(val_tol=1e-5, grad_tol=1e-4) if nu == 0.5 else {}
    check(acq_d, acq_o, Cf.candidates(p), st, **tol)


def test_gradient_matches_central_differences_and_autograd_bridge():
    p = Cf.zdt1_qnehvi(N=120, S=64, raw=4, d=6, q=2)
    st = Cf.build_state(p)
    acq = Cf.build_acqf(p, st, prune_samples=128)
    X = Cf.candidates(p).to(st.device)
    v, g = acq.forward_backward(X)
    h = 1e-6
    b, q, d = X.shape
    P_ = X.unsqueeze(1).repeat(1, 2 * q * d, 1, 1)
    for j in range(q):
        for a in range(d):
            P_[:, 2 * (j * d + a), j, a] += h
            P_[:, 2 * (j * d + a) + 1, j, a] -= h
    vals = acq(P_.view(-1, q, d)).view(b, q * d, 2)
    fd = ((vals[..., 0] - vals[..., 1]) / (2 * h)).view(b, q, d)
    assert float((fd - g).abs().max()) < 1e-5 * float(g.abs().max())
    # torch.autograd bridge: the call BoTorch-style optimisers make
    Xr = X.clone().requires_grad_(True)
    out = acq(Xr)
    (out * torch.arange(1, b + 1, device=out.device, dtype=DT)).sum().backward()
    assert torch.allclose(Xr.grad, g * torch.arange(1, b + 1, device=g.device, dtype=DT).view(-1, 1, 1), rtol=1e-12, atol=0)
    # determinism
    v2, g2 = acq.forward_backward(X)
    assert torch.equal(g, g2) and torch.equal(v, v2)


def test_full_size_gradient_properties_headline_config():
    """ZDT1-30D, N=2000, q=4, S=512 (BASELINE config 3): the gradient of 8 restarts agrees with central differences
    of the forward kernel (size-independent property; the oracle needs minutes at this size)."""
    p = Cf.zdt1_qnehvi()
    st = Cf.build_state(p)
    acq = Cf.build_acqf(p, st)
    X = Cf.candidates(p, 8).to(st.device)
    v, g = acq.forward_backward(X)
    assert bool(torch.isfinite(g).all()) and float(g.abs().max()) > 0
    b, q, d = X.shape
    h = 1e-6
    cols = [0, 7, 29]
    P_ = X.unsqueeze(1).repeat(1, 2 * q * len(cols), 1, 1)
    for j in range(q):
        for k, a in enumerate(cols):
            P_[:, 2 * (j * len(cols) + k), j, a] += h
            P_[:, 2 * (j * len(cols) + k) + 1, j, a] -= h
    vals = acq(P_.view(-1, q, d)).view(b, q * len(cols), 2)
    fd = ((vals[..., 0] - vals[..., 1]) / (2 * h)).view(b, q, len(cols))
    assert float((fd - g[:, :, cols]).abs().max()) < 1e-4 * float(g.abs().max())


def test_gradient_chunk_loops_large_n_and_large_baseline():
    """N > 2048 exercises the partner-chunk loop of kernel_grad_kernel (and N + n_b + q > one chunk); no pruning keeps a
    baseline too large for the shared-memory copy of L_b^-1 (cond_root streams it, cond_root_bwd runs one warp per CTA)."""
    p = Cf.zdt1_qnehvi(N=2200, S=16, raw=5, d=3, q=2)
    st = Cf.build_state(p)
    gp = P.oracle_gp(p)
    ops = [P.op_to_oracle(o) for o in p["objective"].ops]
    Xb = p["X"][:150]
    acq_o = O.QNEHVIOracle(gp, p["ref_point"], Xb, ops, mc_samples=16, seed=7, prune_baseline=False)
    acq_d = A.qNoisyExpectedHypervolumeImprovement(st, p["ref_point"], Xb, p["objective"], prune_baseline=False,
                                                   mc_samples=16, seed=7)
    assert acq_d.nb == 150
    X = Cf.candidates(p).clone()
    X[..., 1:] *= 0.05           # near ZDT1's Pareto set (x2.. = 0): the improvement over the 150 baseline points is > 0
    check(acq_d, acq_o, X, st, val_tol=1e-7, grad_tol=1e-5)


def test_gradient_q16_and_many_batches():
    """q = 16 (BO_MAX_Q) through the generic posterior GEMM, and b = 70 q-batches (> 64 rows: tensor-pipe path + gemm_nt
    for U) against b = 3 (skinny path): both against oracle autograd."""
    p = Cf.zdt1_qnehvi(N=80, S=8, raw=3, d=3, q=16)
    gp = P.oracle_gp(p)
    st = Cf.build_state(p)
    acq_o = P.oracle_acqf(p, gp, prune_samples=64)
    acq_d = Cf.build_acqf(p, st, prune_samples=64)
    check(acq_d, acq_o, Cf.candidates(p), st, grad_tol=1e-5)
    p2 = Cf.zdt1_qnehvi(N=90, S=8, raw=70, d=4, q=2)
    gp2 = P.oracle_gp(p2)
    st2 = Cf.build_state(p2)
    acq_o2 = P.oracle_acqf(p2, gp2, prune_samples=64)
    acq_d2 = Cf.build_acqf(p2, st2, prune_samples=64)
    X = Cf.candidates(p2)
    v_all, g_all = check(acq_d2, acq_o2, X, st2)
    v_few, g_few = acq_d2.forward_backward(X[:3].to(st2.device))
    assert float((g_few - g_all[:3]).abs().max()) < 1e-9 * float(g_all.abs().max())
    # b = 160 > 148 q-batches: cond_root_bwd keeps its warp-per-q-batch layout (four q-batches per CTA), fewer use a CTA each
    p3 = Cf.zdt1_qnehvi(N=90, S=8, raw=160, d=4, q=2)
    gp3 = P.oracle_gp(p3)
    st3 = Cf.build_state(p3)
    acq_o3 = P.oracle_acqf(p3, gp3, prune_samples=64)
    acq_d3 = Cf.build_acqf(p3, st3, prune_samples=64)
    X3 = Cf.candidates(p3)
    v3, g3 = check(acq_d3, acq_o3, X3, st3)
    v3f, g3f = acq_d3.forward_backward(X3[:5].to(st3.device))
    assert float((g3f - g3[:5]).abs().max()) < 1e-9 * float(g3.abs().max())


def test_failed_conditional_root_poisons_value_and_gradient():
    """A q-batch whose q x q conditional covariance is not p.d. even after jitter 1e-3 is flagged (info) and returns NaN for
    the value and its gradient instead of silently continuing; the other q-batches are unaffected."""
    p = small_two_objective_problem()
    st = Cf.build_state(p)
    acq = Cf.build_acqf(p, st, prune_samples=64)
    X = Cf.candidates(p, 3).clone()
    v_ok, g_ok = acq.forward_backward(X.to(st.device))
    assert bool(torch.isfinite(v_ok).all()) and bool(torch.isfinite(g_ok).all())
    Xbad = X.clone()
    Xbad[1, 1] = Xbad[1, 0]            # duplicate point: singular to rounding -> jitter ladder, still finite
    v, g = acq.forward_backward(Xbad.to(st.device))
    assert bool(torch.isfinite(v[[0, 2]]).all()) and bool(torch.isfinite(g[[0, 2]]).all())
    assert torch.equal(v[[0, 2]], v_ok[[0, 2]])


def small_two_objective_problem():
    return Cf.zdt1_qnehvi(N=60, S=16, raw=3, d=4, q=2)

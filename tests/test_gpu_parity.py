"""Parity of the CUDA path (through the C ABI) against the CPU oracle.  Run on the B200 box:
    python -m pytest tests -m gpu -x -q
Tolerances: posterior 1e-9 relative (north_star); acquisition values 1e-8 relative (different summation
order over cells / MC samples); front / box indexing bit-exact."""
import math

import numpy as np
import pytest
import torch

from everest_b200 import configs as Cf
from everest_b200 import kernels as K
from everest_b200.objectives import (MaximizeObjective, MinimizeObjective, MultiObjective, OutputConstraint,
                                     ScalarObjective)
from oracle import bo_oracle as O
from tests import problems as P

pytestmark = pytest.mark.gpu
DT = torch.float64


def rel_err(a, b, floor=0.0):
    a, b = a.detach().cpu().to(DT), b.detach().cpu().to(DT)
    return float(((a - b).abs() / (b.abs() + floor)).max())


def small_problem(kind, **kw):
    if kind == "zdt1":
        return Cf.zdt1_qnehvi(N=96, S=32, raw=24, d=6, q=3, **kw)
    if kind == "dtlz2":
        return Cf.dtlz2_qnehvi(N=70, S=16, raw=12, d=5, m_obj=3, q=2, **kw)
    if kind == "himmelblau":
        return Cf.himmelblau_qlogei(N=150, S=64, raw=40)
    if kind == "mixed":
        return Cf.mixed_tanimoto_qlogei(N=130, n_bits=200, S=32, n_choices=50)
    raise ValueError(kind)


@pytest.mark.parametrize("kind", ["zdt1", "dtlz2", "himmelblau", "mixed"])
def test_factorization_and_posterior(kind):
    p = small_problem(kind)
    gp = P.oracle_gp(p)
    st = Cf.build_state(p)
    N, ldk = st.N, ((st.N + 15) // 16) * 16
    for m in range(st.M):
        f = gp._fact[m]
        Ld = st.debug_get("L", m).view(N, ldk)[:, :N].cpu()
        assert torch.allclose(Ld, f["L"], rtol=1e-9, atol=1e-11), kind
        Li = st.debug_get("Linv", m).view(N, ldk)[:, :N].cpu()
        scale = float(f["Linv"].abs().max())
        assert float((Li - f["Linv"]).abs().max()) <= 1e-9 * scale
        al = st.debug_get("alpha", m).cpu()
        assert float((al - f["alpha"]).abs().max()) <= 1e-8 * float(f["alpha"].abs().max())
    Xq = Cf.candidates(p, 9).reshape(-1, p["d"])
    mean_o, cov_o = gp.posterior(Xq)
    mean_d, cov_d = st.posterior_joint(Xq)
    assert rel_err(mean_d, mean_o, floor=1e-6) < 1e-9
    vmax = float(cov_o.diagonal(dim1=-1, dim2=-2).abs().max())
    assert float((cov_d.cpu() - cov_o).abs().max()) < 1e-9 * vmax
    mean_m, var_m = st.posterior(Xq, observation_noise=True)
    mo, co = gp.posterior(Xq, observation_noise=True)
    assert rel_err(mean_m, mo, floor=1e-6) < 1e-9
    assert float((var_m.cpu().T - co.diagonal(dim1=-1, dim2=-2)).abs().max()) < 1e-9 * vmax


@pytest.mark.parametrize("nu", [0.5, 1.5, 2.5])
def test_matern_and_scale_kernels(nu):
    p = Cf.himmelblau_qlogei(N=90, S=16, raw=8)
    p["outputs"][0]["kernel"] = K.ScaleKernel(K.MaternKernel([0, 1], [0.3, 0.5], nu=nu), outputscale=1.7)
    gp = P.oracle_gp(p)
    st = Cf.build_state(p)
    Xq = Cf.candidates(p, 8).reshape(-1, 2)
    mo, co = gp.posterior(Xq)
    md, cd = st.posterior_joint(Xq)
    assert rel_err(md, mo, floor=1e-6) < 1e-9
    assert float((cd.cpu() - co).abs().max()) < 1e-9 * float(co.abs().max())


@pytest.mark.parametrize("cards", [(4, 6), (20, 3), (2, 3, 2, 4, 2, 3, 2)])
def test_hamming_leaf_variants_of_the_tree_kernel(cards):
    """The generic cross-covariance kernel evaluates a Hamming leaf three ways: the mismatch pattern read off packed nibble
    codes (up to 6 groups, cardinalities <= 16), per-group compares into the table of the 2^nd patterns (a cardinality above
    16), and the direct sum + exp (more than 6 groups).  All three must reproduce the reference's HammingKernelWithOneHots
    arithmetic (kernels/categorical.py), here inside an additive / multiplicative tree with a Matern leaf."""
    import numpy as np
    rng = np.random.default_rng(5)
    N, nq = 140, 37
    def draw(n):
        cols = [rng.random((n, 2))] + [np.eye(c)[rng.integers(0, c, n)] for c in cards]
        return np.concatenate(cols, axis=1)
    X = draw(N)
    d = X.shape[1]
    cats, start = {}, 2
    for c in cards:
        cats[start] = c
        start += c
    y = np.sin(3 * X[:, 0]) + X[:, 1] ** 2 + 0.3 * X[:, 2] - 0.2 * X[:, 3] + 0.05 * rng.normal(size=N)
    ls = [0.5 + 0.3 * i for i in range(len(cards))]
    kern = K.AdditiveKernel([
        K.ScaleKernel(K.MaternKernel([0, 1], [0.4, 0.6], nu=2.5), 0.8),
        K.ScaleKernel(K.HammingDistanceKernel(cats, ls), 0.5),
        K.MultiplicativeKernel([K.ScaleKernel(K.MaternKernel([0, 1], [0.7, 0.3], nu=2.5), 0.9),
                                K.ScaleKernel(K.HammingDistanceKernel(cats, ls[::-1]), 1.1)])])
    p = dict(name="hamming_variants", d=d, X=X, Y=y[:, None], bounds=None, in_offset=np.zeros(d), in_scale=np.ones(d),
             outputs=[dict(kernel=kern, y=y, noise=1e-2, mean_const=0.0)])
    gp = P.oracle_gp(p)
    st = Cf.build_state(p)
    Xq = torch.as_tensor(draw(nq))
    mo, co = gp.posterior(Xq)
    md, cd = st.posterior_joint(Xq)
    assert rel_err(md, mo, floor=1e-6) < 1e-9
    assert float((cd.cpu() - co).abs().max()) < 1e-9 * float(co.abs().max())


@pytest.mark.parametrize("m_obj,n", [(2, 40), (3, 25), (4, 18)])
def test_box_decomposition_bit_exact(m_obj, n):
    """Identical objective values in -> identical cell lists out (no arithmetic, only comparisons)."""
    from everest_b200 import acquisition as A

    p = Cf.dtlz2_qnehvi(N=40, S=8, raw=4, d=m_obj + 2, m_obj=m_obj, q=1)
    st = Cf.build_state(p)
    g = torch.Generator().manual_seed(11 + m_obj)
    Yobj = torch.rand(n, m_obj, dtype=DT, generator=g)
    Yobj[3] = Yobj[1]  # duplicate point
    ref = [0.1] * m_obj
    acq = A.qExpectedHypervolumeImprovement(st, ref, Yobj, p["objective"], mc_samples=8, seed=1)
    lo = st.debug_get("cell_lo")
    up = st.debug_get("cell_up")
    nc = int(st.debug_get("ncells", dtype=torch.int32)[0])
    cap = lo.numel() // m_obj
    lo = lo.view(cap, m_obj)[:nc].cpu()
    up = up.view(cap, m_obj)[:nc].cpu()
    Pf, _ = O.pareto_front_above_ref(Yobj, torch.tensor(ref, dtype=DT))
    if m_obj == 2:
        lo_o, up_o, _ = O.partition_2d(Pf, torch.tensor(ref, dtype=DT))
    else:
        lo_o, up_o = O.partition_nd(Pf, torch.tensor(ref, dtype=DT))
    assert nc == lo_o.shape[0] == acq.max_cells
    assert torch.equal(lo, lo_o) and torch.equal(up, up_o)


@pytest.mark.parametrize("kind", ["zdt1", "dtlz2"])
def test_qnehvi_prune_cells_and_forward(kind):
    p = small_problem(kind)
    gp = P.oracle_gp(p)
    st = Cf.build_state(p)
    acq_o = P.oracle_acqf(p, gp, prune_samples=256)
    acq_d = Cf.build_acqf(p, st, prune_samples=256)
    # pruning keeps exactly the same baseline points
    assert acq_d.prune_idx.cpu().tolist() == acq_o.prune_idx.tolist()
    assert acq_d.nb == acq_o.nb
    # cached baseline root and samples
    ldlb = ((max(acq_o.nb, 1) + 15) // 16) * 16
    for m in range(st.M):
        Lb = st.debug_get("baseline_L", m).view(acq_o.nb, ldlb)[:, : acq_o.nb].cpu()
        assert float((Lb - acq_o.baseline_L[m]).abs().max()) < 1e-8 * float(acq_o.baseline_L[m].abs().max())
    fb = st.debug_get("samples_b").view(acq_o.S, acq_o.nb, st.M).cpu()
    assert float((fb - acq_o.samples_b).abs().max()) < 1e-8 * float(acq_o.samples_b.abs().max())
    # per-sample cells: same count, same front membership, bounds to rounding
    lo, up, nc = acq_d.cell_bounds()
    assert nc.tolist() == acq_o.n_cells.tolist()
    for s in range(acq_o.S):
        c = int(nc[s])
        assert torch.allclose(lo[s, :c], acq_o.cell_lower[s, :c], rtol=1e-9, atol=1e-10)
        fin = torch.isfinite(acq_o.cell_upper[s, :c])
        assert torch.equal(torch.isfinite(up[s, :c]), fin)
        assert torch.allclose(up[s, :c][fin], acq_o.cell_upper[s, :c][fin], rtol=1e-9, atol=1e-10)
    if kind == "zdt1":
        fi = st.debug_get("front_idx", dtype=torch.int32).view(acq_o.S, acq_o.nb + 1).cpu()
        for s in range(acq_o.S):
            k = len(acq_o.fronts[s])
            assert fi[s, :k].tolist() == acq_o.fronts[s].tolist()  # bit-exact front indexing
    # forward
    X = Cf.candidates(p)
    v_o, parts = acq_o.forward(X, return_parts=True)
    v_d = acq_d(X.to(st.device))
    q, nr = p["q"], acq_o.nb + p["q"]
    root = st.debug_get("root").view(-1)[: X.shape[0] * st.M * q * nr].view(X.shape[0], st.M, q, nr).cpu()
    bl_o = parts["bl"].permute(1, 0, 2, 3)
    br_o = parts["br"].permute(1, 0, 2, 3)
    sc = float(br_o.abs().max())
    assert float((root[..., : acq_o.nb] - bl_o).abs().max()) < 1e-7 * max(sc, float(bl_o.abs().max()))
    assert float((root[..., acq_o.nb:] - br_o).abs().max()) < 1e-7 * sc
    assert int(acq_d.last_info.sum()) == 0 and float(parts["jitter"].abs().sum()) == 0.0
    scale = float(v_o.abs().max())
    assert scale > 0
    assert float((v_d.cpu() - v_o).abs().max()) < 1e-8 * scale
    # determinism: the same call twice is bit-identical; a slice agrees to rounding (Gram partial order)
    assert torch.equal(acq_d(X.to(st.device)), v_d)
    v_half = acq_d(X[: X.shape[0] // 2].to(st.device))
    assert float((v_half.cpu() - v_d.cpu()[: X.shape[0] // 2]).abs().max()) <= 1e-12 * scale
    # host-buffer entry point agrees bit for bit with the device-pointer one
    v_host = acq_d.forward_host(X.numpy())
    assert np.array_equal(v_host, v_d.cpu().numpy())


@pytest.mark.parametrize("q,N", [(1, 77), (2, 130), (4, 200), (8, 96), (3, 50), (5, 141), (16, 64)])
def test_qnehvi_forward_all_q_paths(q, N):
    """q in {1,2,4,8} runs the TMA/mbarrier GEMM with the tensor-pipe Gram epilogue, every other q the generic
    cp.async kernel; N exercises the zero padding up to the 16 / 128 multiples."""
    p = Cf.zdt1_qnehvi(N=N, S=16, raw=37, d=5, q=q)
    gp = P.oracle_gp(p)
    st = Cf.build_state(p)
    acq_o = P.oracle_acqf(p, gp, prune_samples=64)
    acq_d = Cf.build_acqf(p, st, prune_samples=64)
    assert acq_d.prune_idx.cpu().tolist() == acq_o.prune_idx.tolist()
    X = Cf.candidates(p)
    v_o, parts = acq_o.forward(X, return_parts=True)
    v_d = acq_d(X.to(st.device))
    nr = acq_o.nb + q
    root = st.debug_get("root").view(-1)[: X.shape[0] * st.M * q * nr].view(X.shape[0], st.M, q, nr).cpu()
    br_o = parts["br"].permute(1, 0, 2, 3)
    assert float((root[..., acq_o.nb:] - br_o).abs().max()) < 1e-7 * float(br_o.abs().max())
    mu = st.debug_get("mu").view(-1)[: X.shape[0] * q * st.M].view(X.shape[0], q, st.M).cpu()
    assert rel_err(mu, parts["mu"], floor=1e-6) < 1e-9
    assert float((v_d.cpu() - v_o).abs().max()) < 1e-8 * float(v_o.abs().max())


@pytest.mark.parametrize("path", ["tiled", "chunked", "generic"])
@pytest.mark.parametrize("kind,q", [("zdt1", 4), ("dtlz2", 2), ("dtlz2", 7)])
def test_all_three_hvi_kernels_agree_with_oracle(monkeypatch, path, kind, q):
    """EVEREST_MC_PATH forces the tiled (cells resident in smem), chunked (cells streamed, 8 q-batches per
    thread) or generic (one CTA per q-batch) inclusion-exclusion kernel."""
    monkeypatch.setenv("EVEREST_MC_PATH", path)
    p = Cf.zdt1_qnehvi(N=60, S=40, raw=70, d=5, q=q) if kind == "zdt1" else Cf.dtlz2_qnehvi(N=50, S=40, raw=70, d=5, m_obj=3, q=q)
    gp = P.oracle_gp(p)
    st = Cf.build_state(p)
    acq_o = P.oracle_acqf(p, gp, prune_samples=64)
    acq_d = Cf.build_acqf(p, st, prune_samples=64)
    X = Cf.candidates(p)
    v_o = acq_o.forward(X)
    v_d = acq_d(X.to(st.device))
    assert float((v_d.cpu() - v_o).abs().max()) < 1e-8 * float(v_o.abs().max())


@pytest.mark.parametrize("N", [2, 5])
def test_detergent_config1_five_objectives(N):
    """BASELINE config 1 (README loop): 5 inputs, 5 Maximize outputs, N = 2..5 observations, q = 1, Normalize
    input transform, ref point inferred from the data; exercises the 5-objective box decomposition and the
    generic HVI kernel on a launch-bound problem."""
    p = Cf.detergent_qnehvi(N=N, S=64, raw=50)
    gp = P.oracle_gp(p)
    st = Cf.build_state(p)
    acq_o = P.oracle_acqf(p, gp, prune_samples=128)
    acq_d = Cf.build_acqf(p, st, prune_samples=128)
    assert acq_d.prune_idx.cpu().tolist() == acq_o.prune_idx.tolist()
    lo, up, nc = acq_d.cell_bounds()
    assert nc.tolist() == acq_o.n_cells.tolist()
    X = Cf.candidates(p)
    v_o = acq_o.forward(X)
    v_d = acq_d(X.to(st.device))
    assert float(v_o.abs().max()) > 0
    assert float((v_d.cpu() - v_o).abs().max()) < 1e-8 * float(v_o.abs().max())
    preds, stds = st.predict(X[:, 0, :])
    mo, co = gp.posterior(X[:, 0, :], observation_noise=True)
    assert np.allclose(preds, mo.numpy(), rtol=1e-9, atol=1e-12)
    assert np.allclose(stds, co.diagonal(dim1=-1, dim2=-2).T.sqrt().numpy(), rtol=1e-7, atol=1e-12)


def test_single_q_batch_and_single_sample_edge_cases():
    """b = 1, S = 1, empty baseline after pruning (ref point better than every observation)."""
    from everest_b200 import acquisition as A

    p = Cf.zdt1_qnehvi(N=20, S=1, raw=1, d=3, q=2)
    gp = P.oracle_gp(p)
    st = Cf.build_state(p)
    acq_o = P.oracle_acqf(p, gp, prune_samples=32)
    acq_d = Cf.build_acqf(p, st, prune_samples=32)
    X = Cf.candidates(p)
    assert X.shape[0] == 1
    assert abs(float(acq_d(X.to(st.device))[0]) - float(acq_o.forward(X)[0])) <= 1e-8 * max(1e-12, abs(float(acq_o.forward(X)[0])))
    p["ref_point"] = [10.0, 10.0]  # nothing is better than this -> pruning removes every baseline point
    p["S"] = 8
    acq_o2 = P.oracle_acqf(p, gp, prune_samples=32)
    acq_d2 = Cf.build_acqf(p, st, prune_samples=32)
    assert acq_o2.nb == 0 and acq_d2.nb == 0
    X2 = Cf.candidates(p, 5)
    v_o, v_d = acq_o2.forward(X2), acq_d2(X2.to(st.device)).cpu()
    assert float((v_d - v_o).abs().max()) <= 1e-8 * max(float(v_o.abs().max()), 1e-300)


def test_qnehvi_with_output_constraint_and_pending():
    from everest_b200 import acquisition as A

    p = Cf.zdt1_qnehvi(N=60, S=24, raw=16, d=5, q=2)
    # third output used only as a constraint
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
    assert acq_d.prune_idx.cpu().tolist() == acq_o.prune_idx.tolist()
    X = Cf.candidates(p)
    v_o = acq_o.forward(X)
    v_d = acq_d(X.to(st.device))
    assert float((v_d.cpu() - v_o).abs().max()) < 1e-8 * float(v_o.abs().max())


@pytest.mark.parametrize("kind", ["himmelblau", "mixed"])
def test_qlogei_forward(kind):
    p = small_problem(kind)
    gp = P.oracle_gp(p)
    st = Cf.build_state(p)
    acq_o = P.oracle_acqf(p, gp)
    acq_d = Cf.build_acqf(p, st)
    assert abs(acq_d.best_f - acq_o.best_f) < 1e-9 * max(1.0, abs(acq_o.best_f))
    X = Cf.candidates(p)
    v_o = acq_o.forward(X)
    v_d = acq_d(X.to(st.device))
    assert float((v_d.cpu() - v_o).abs().max()) < 1e-7 * float(v_o.abs().max())


def test_qlogei_q_batch_additive_objective():
    from everest_b200 import acquisition as A

    p = Cf.zdt1_qnehvi(N=50, S=32, raw=10, d=4, q=3)
    obj = ScalarObjective([MaximizeObjective(0, w=0.4), MinimizeObjective(1, w=0.6)], "additive")
    gp = P.oracle_gp(p)
    st = Cf.build_state(p)
    spec = ("additive", [(P.op_to_oracle(o), o.w) for o in obj.ops])
    acq_o = O.QLogEIOracle(gp, spec, p["X"], mc_samples=32, seed=5)
    acq_d = A.get_acquisition_function("qLogEI", st, obj, p["X"], mc_samples=32, seed=5)
    X = Cf.candidates(p)
    v_o = acq_o.forward(X)
    v_d = acq_d(X.to(st.device))
    assert float((v_d.cpu() - v_o).abs().max()) < 1e-7 * float(v_o.abs().max())


def test_qehvi_forward():
    from everest_b200 import acquisition as A

    p = small_problem("zdt1")
    gp = P.oracle_gp(p)
    st = Cf.build_state(p)
    Yobj = -torch.as_tensor(p["Y"], dtype=DT)
    acq_o = O.QEHVIOracle(gp, p["ref_point"], Yobj, [P.op_to_oracle(o) for o in p["objective"].ops], mc_samples=32, seed=9)
    acq_d = A.qExpectedHypervolumeImprovement(st, p["ref_point"], Yobj, p["objective"], mc_samples=32, seed=9)
    X = Cf.candidates(p, 12)
    v_o = acq_o.forward(X)
    v_d = acq_d(X.to(st.device))
    assert float((v_d.cpu() - v_o).abs().max()) < 1e-8 * max(float(v_o.abs().max()), 1e-12)


def test_degenerate_q_batch_is_handled_like_the_oracle():
    """Two identical points in a q-batch make the conditional q x q block singular to rounding: the last
    pivot is +-1e-17, so whether the psd_safe_cholesky ladder engages is a coin flip on BOTH paths.  The
    well-posed batches must still agree to 1e-8, the degenerate one stays finite, flags no hard failure
    and agrees to the size of the largest jitter (1e-3 relative is far above it)."""
    p = small_problem("zdt1")
    gp = P.oracle_gp(p)
    st = Cf.build_state(p)
    acq_o = P.oracle_acqf(p, gp, prune_samples=128)
    acq_d = Cf.build_acqf(p, st, prune_samples=128)
    X = Cf.candidates(p, 4).clone()
    X[1, 1] = X[1, 0]
    v_o = acq_o.forward(X)
    v_d = acq_d(X.to(st.device)).cpu()
    assert bool(torch.isfinite(v_d).all()) and int(acq_d.last_info.sum()) == 0
    scale = float(v_o.abs().max())
    keep = torch.tensor([0, 2, 3])
    assert float((v_d[keep] - v_o[keep]).abs().max()) < 1e-8 * scale
    assert abs(float(v_d[1] - v_o[1])) < 1e-3 * scale


def test_jitter_ladder_and_not_psd_error():
    """A (nonsensical but deterministic) slightly NEGATIVE-definite Gram: -c K_rbf with c * lambda_max = 3e-6
    and zero noise fails for jitter 1e-8, 1e-7, 1e-6 and succeeds at 1e-5 on both paths; a strongly negative
    one exhausts the ladder and raises NotPSDError (BoTorch: NotPSDError from psd_safe_cholesky)."""
    from everest_b200 import NotPSDError

    p = Cf.zdt1_qnehvi(N=40, S=16, raw=8, d=4, q=1)
    Xt = torch.as_tensor(p["X"], dtype=DT)
    base = p["outputs"][0]["kernel"]
    lam = float(torch.linalg.eigvalsh(O.eval_kernel(P.kernel_to_oracle(base), Xt, Xt, Xt.mean(0), same=True)).max())
    for o in p["outputs"]:
        o["kernel"] = K.ScaleKernel(base, -3e-6 / lam)
        o["noise"] = 0.0
    gp = P.oracle_gp(p)
    st = Cf.build_state(p)
    assert [f["jitter"] for f in gp._fact] == pytest.approx([1e-5] * 2, rel=1e-12)
    assert st.jitter == pytest.approx([1e-5] * 2, rel=1e-12)
    for m in range(2):
        al = st.debug_get("alpha", m).cpu()
        assert float((al - gp._fact[m]["alpha"]).abs().max()) <= 1e-9 * float(gp._fact[m]["alpha"].abs().max())
    for o in p["outputs"]:
        o["kernel"] = K.ScaleKernel(base, -1.0)
    with pytest.raises(O.NotPSDError):
        P.oracle_gp(p)
    with pytest.raises(NotPSDError):
        Cf.build_state(p)


def test_full_size_properties_headline_config():
    """Config 3 at BASELINE size (N=2000, d=30, q=4, S=512): size-independent properties."""
    p = Cf.zdt1_qnehvi(raw=1024)
    st = Cf.build_state(p)
    N, ldk = st.N, ((st.N + 15) // 16) * 16
    # Cholesky round trip and inverse-root identity on the device factor
    for m in range(st.M):
        Ld = st.debug_get("L", m).view(N, ldk)[:, :N]
        Li = st.debug_get("Linv", m).view(N, ldk)[:, :N]
        eye = torch.eye(N, dtype=DT, device=Ld.device)
        assert float((Li @ Ld - eye).abs().max()) < 1e-8
        assert float(torch.triu(Ld, 1).abs().max()) == 0.0 and float(torch.triu(Li, 1).abs().max()) == 0.0
    # interpolation: posterior mean at the training points reproduces y to noise level, variance ~ noise
    Xt = torch.as_tensor(p["X"][:256])
    mean, var = st.posterior(Xt)
    Y = torch.as_tensor(p["Y"][:256])
    assert float((mean.cpu() - Y).abs().max()) < 5e-2
    assert float(var.min()) > 0 and float(var.max()) < 1e-2
    acq = Cf.build_acqf(p, st)
    X = Cf.candidates(p)
    v = acq(X.to(st.device))
    assert bool(torch.isfinite(v).all()) and float(v.min()) >= 0.0 and float(v.max()) > 0.0
    assert int(acq.last_info.sum()) == 0
    # reproducible bit for bit for a fixed batch shape; independent of the batch it is evaluated in up to the
    # summation order of the Gram partials (the column-block grouping of the GEMM depends on the batch size) and to the
    # GEMM variant the batch size selects (INT8 digit planes for the large batch, FP64 DMMA for the slice: ~1e-12)
    assert torch.equal(acq(X.to(st.device)), v)
    v2 = acq(X[300:700].to(st.device))
    assert float((v2 - v[300:700]).abs().max()) <= 1e-10 * float(v.abs().max())
    # a q-batch whose points are all dominated by the baseline front in every MC sample scores exactly 0
    Xbad = torch.ones(1, p["q"], p["d"], dtype=DT)
    assert float(acq(Xbad.to(st.device))[0]) < 1e-3


def test_optimize_acqf_screen_refine_and_fd_gradient():
    """Finite-difference gradient of the device acquisition value vs autograd through the oracle, and the
    L2 boundary: optimize_acqf returns a [q, d] CPU candidate inside the bounds that is at least as good as the
    best screened raw sample."""
    from everest_b200 import optim

    p = Cf.zdt1_qnehvi(N=64, S=32, raw=64, d=4, q=2)
    gp = P.oracle_gp(p)
    st = Cf.build_state(p)
    acq_o = P.oracle_acqf(p, gp, prune_samples=128)
    acq_d = Cf.build_acqf(p, st, prune_samples=128)
    X = Cf.candidates(p)
    v = acq_d(X.to(st.device)).cpu()
    i = int(torch.argmax(v))
    x0 = X[i : i + 1].clone().requires_grad_(True)
    acq_o.forward(x0).sum().backward()
    g_ad = x0.grad[0]
    h = 1e-6
    g_fd = torch.zeros_like(g_ad)
    pert = []
    for a in range(p["q"]):
        for j in range(p["d"]):
            for sgn in (1.0, -1.0):
                xp = X[i].clone()
                xp[a, j] += sgn * h
                pert.append(xp)
    vals = acq_d(torch.stack(pert).to(st.device)).cpu().view(p["q"], p["d"], 2)
    g_fd = (vals[..., 0] - vals[..., 1]) / (2 * h)
    assert float((g_fd - g_ad).abs().max()) < 1e-5 * max(float(g_ad.abs().max()), 1e-8) + 1e-9
    bounds = torch.as_tensor(p["bounds"])
    cand, val = optim.optimize_acqf(acq_d, bounds, q=p["q"], num_restarts=4, raw_samples=64, options={"maxiter": 30}, seed=3)
    assert cand.shape == (p["q"], p["d"]) and cand.device.type == "cpu"
    assert bool((cand >= bounds[0] - 1e-12).all()) and bool((cand <= bounds[1] + 1e-12).all())
    X_rnd = optim.draw_sobol_samples(bounds, 64, p["q"], seed=3)
    best_raw = float(acq_d(X_rnd.to(st.device)).max())
    assert float(val) >= best_raw - 1e-15
    assert abs(float(acq_d(cand.unsqueeze(0).to(st.device))[0]) - float(val)) < 1e-12
    # mixed branch (botorch.py:358-378): best over a list of fixed-feature dictionaries, q = 1
    p1m = Cf.zdt1_qnehvi(N=64, S=32, raw=64, d=4, q=1)
    acq1m = Cf.build_acqf(p1m, Cf.build_state(p1m), prune_samples=128)
    # initialize_q_batch draws the restarts from torch's GLOBAL generator (like BoTorch): same state for both runs
    torch.manual_seed(11)
    cm, vm = optim.optimize_acqf_mixed(acq1m, bounds, q=1, num_restarts=2, raw_samples=32,
                                       fixed_features_list=[{0: 0.1}, {0: 0.9}], options={"maxiter": 10}, seed=1)
    assert cm.shape == (1, 4) and float(cm[0, 0]) in (0.1, 0.9)
    torch.manual_seed(11)
    singles = [optim.optimize_acqf(acq1m, bounds, 1, 2, 32, fixed_features=ff, options={"maxiter": 10}, seed=1)[1] for ff in ({0: 0.1}, {0: 0.9})]
    assert float(vm) == max(float(s_) for s_ in singles)
    vals_rows = optim.calc_acquisition(acq1m, X_rnd[:5, 0, :])
    val_comb = optim.calc_acquisition(acq_d, X_rnd[0], combined=True)
    assert vals_rows.shape == (5,) and val_comb.shape == (1,)
    # discrete branch (botorch.py:461): arg-max over a choice set
    choices = X_rnd[:, 0, :]
    p1 = Cf.zdt1_qnehvi(N=64, S=32, raw=64, d=4, q=1)
    acq1 = Cf.build_acqf(p1, Cf.build_state(p1), prune_samples=128)
    c, cv = optim.optimize_acqf_discrete(acq1, q=1, choices=choices)
    allv = acq1(choices.unsqueeze(1).to(st.device)).cpu()
    assert float(cv) == float(allv.max()) and torch.equal(c[0], choices[int(torch.argmax(allv))])


def test_multiobjective_utilities_kat_and_oracle():
    """get_pareto_front / infer_ref_point / compute_hypervolume against the reference's known-answer tables
    (tests/bofire/utils/test_multiobjective.py:216-266, via tests/golden/reference_golden.json) and the oracle."""
    import json
    import os

    from everest_b200 import multiobjective as MO
    from everest_b200.objectives import ObjectiveSpec

    G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_golden.json")))
    for case in G["ref_points"]:
        Y = np.array(case["Y"])
        obj = MultiObjective([ObjectiveSpec(o[0], o[1], *o[2:]) for o in case["ops"]])
        assert MO.get_pareto_front(obj, Y).tolist() == case["expected_pareto_idx"]
        assert np.array_equal(MO.get_ref_point_mask(obj), np.array(case["mask"]))
        assert np.array_equal(MO.infer_ref_point(obj, Y, return_masked=True), np.array(case["ref_masked"]))
        assert np.array_equal(MO.infer_ref_point(obj, Y, return_masked=False), np.array(case["ref_plain"]))
        hv = MO.compute_hypervolume(obj, Y, (np.array(case["ref_plain"]) - 1.0 * np.array(case["mask"])).tolist())
        assert hv > 0  # test_compute_hypervolume pins only positivity
    g = torch.Generator().manual_seed(4)
    for m in (2, 3, 4, 5):
        Y = torch.rand(40, m, dtype=DT, generator=g)
        Y[7] = Y[2]
        assert torch.equal(MO.is_non_dominated(Y), O.is_non_dominated(Y))
        assert torch.equal(MO.is_non_dominated(Y, deduplicate=False), O.is_non_dominated(Y, deduplicate=False))
        obj = MultiObjective([MaximizeObjective(i) for i in range(m)])
        ref = [0.05] * m
        hv_d = MO.compute_hypervolume(obj, Y.numpy(), ref)
        hv_o = O.hypervolume(Y[O.is_non_dominated(Y)], torch.tensor(ref, dtype=DT))
        assert abs(hv_d - hv_o) < 1e-12
    assert MO.compute_hypervolume(MultiObjective([MaximizeObjective(0), MaximizeObjective(1)]), np.zeros((3, 2)), [1.0, 1.0]) == 0.0


def test_errors_are_loud():
    from everest_b200 import DeviceGPState, SingleTaskGPSpec, acquisition as A

    X = np.random.default_rng(0).random((10, 3))
    with pytest.raises(ValueError):
        DeviceGPState(X, [SingleTaskGPSpec(kernel=K.RBFKernel([0, 1, 5], [0.5]), y=np.zeros(10))])
    with pytest.raises(ValueError):
        DeviceGPState(X, [SingleTaskGPSpec(kernel=K.TanimotoKernel([0, 1, 2]), y=np.zeros(10))])  # not 0/1 bits
    st = DeviceGPState(X, [SingleTaskGPSpec(kernel=K.RBFKernel([0, 1, 2], [0.5]), y=X[:, 0])] * 2).factorize()
    with pytest.raises(ValueError):      # BoFire's data model bounds alpha to [0, 0.5]
        A.qNoisyExpectedHypervolumeImprovement(st, [0, 0], X, MultiObjective([MaximizeObjective(0), MaximizeObjective(1)]),
                                               alpha=0.7)
    with pytest.raises(NotImplementedError):
        A.qNoisyExpectedHypervolumeImprovement(st, [0, 0], X, MultiObjective([MaximizeObjective(0), MaximizeObjective(1)]),
                                               cache_root=False)
    acq = A.qNoisyExpectedHypervolumeImprovement(st, [0, 0], X, MultiObjective([MaximizeObjective(0), MaximizeObjective(1)]),
                                                 mc_samples=16, seed=0)
    with pytest.raises(ValueError):
        acq(torch.zeros(2, 1, 4, dtype=DT))


@pytest.mark.parametrize("N,q,raw,d", [(300, 4, 100, 6), (257, 1, 333, 3), (640, 2, 70, 5), (130, 8, 40, 4), (1000, 4, 300, 8)])
def test_int8_digit_plane_gemm_matches_fp64_and_oracle(N, q, raw, d):
    """csrc/ozaki.cu: the posterior GEMM as 28 exact INT8 x INT8 -> INT32 products of 7 balanced base-256 digit planes on
    tcgen05 (accumulators in TMEM), recombined in FP64.  Forced on (option 2) for shapes with ragged row / column / K tiles
    and extra columns (mean, baseline rows); compared with the FP64 DMMA kernel (1e-11) and with the oracle (1e-8)."""
    p = Cf.zdt1_qnehvi(N=N, S=16, raw=raw, d=d, q=q)
    gp = P.oracle_gp(p)
    st = Cf.build_state(p)
    acq_o = P.oracle_acqf(p, gp, prune_samples=64)
    acq_d = Cf.build_acqf(p, st, prune_samples=64)
    X = Cf.candidates(p)
    Xd = X.to(st.device)
    nr = acq_o.nb + q
    outs = {}
    for mode in (0, 2):
        acq_d.set_option("ozaki", mode)
        v = acq_d(Xd).cpu()
        root = st.debug_get("root").view(-1)[: X.shape[0] * st.M * q * nr].cpu().clone()
        mu = st.debug_get("mu").view(-1)[: X.shape[0] * q * st.M].cpu().clone()
        outs[mode] = (v, root, mu)
    v_o, parts = acq_o.forward(X, return_parts=True)
    scale = float(v_o.abs().max())
    assert scale > 0
    assert float((outs[2][0] - v_o).abs().max()) < 1e-8 * scale
    # the digit-plane product drops the pairs below 2^-51 of (row scale x column scale): ~10x the rounding of a DGEMM; the
    # conditional roots (Cholesky of a nearly singular q x q block) amplify that like they amplify DGEMM rounding
    assert float((outs[2][0] - outs[0][0]).abs().max()) < 1e-10 * scale
    assert float((outs[2][1] - outs[0][1]).abs().max()) < 1e-8 * float(outs[0][1].abs().max())
    assert float((outs[2][2] - outs[0][2]).abs().max()) < 1e-10 * float(outs[0][2].abs().max())
    assert rel_err(outs[2][2].view(X.shape[0], q, st.M), parts["mu"], floor=1e-6) < 1e-9
    # determinism of the INT8 path
    acq_d.set_option("ozaki", 2)
    assert torch.equal(acq_d(Xd).cpu(), outs[2][0])
    # the three kernel variants (one-pass 128x64, two-pass 128x128, two-pass on CTA pairs) compute the same exact integer
    # sums and differ only in how the FP64 Gram partial sums are grouped
    for tile in (64, 128, 256):
        acq_d.set_option("ozaki_tile", tile)
        v_t = acq_d(Xd).cpu()
        mu_t = st.debug_get("mu").view(-1)[: X.shape[0] * q * st.M].cpu()
        assert float((v_t - outs[0][0]).abs().max()) < 1e-10 * scale, tile
        assert float((mu_t - outs[0][2]).abs().max()) < 1e-10 * float(outs[0][2].abs().max()), tile
        assert torch.equal(acq_d(Xd).cpu(), v_t), tile
    acq_d.set_option("ozaki_tile", 0)
    with pytest.raises(ValueError):
        acq_d.set_option("ozaki", 3)
    with pytest.raises(ValueError):
        acq_d.set_option("ozaki_tile", 32)


def test_int8_per_row_guard_and_empty_batch():
    """Automatic mode: every INT8 call is followed by the per-row guard (csrc/ozaki.cu: digit-plane error estimate
    2 sqrt(G_ii) eps against 1e-10 x the posterior variance of the row); q-batches it flags are redone with the FP64 kernel,
    the rest keep the INT8 result.  Random candidates in 30 dimensions (posterior variance within two orders of the
    prior): almost nothing is flagged and the values match the FP64 kernel's.  An empty t-batch returns an empty tensor."""
    p = Cf.zdt1_qnehvi(N=1000, S=16, raw=1200, d=30, q=4)
    st = Cf.build_state(p)
    acq = Cf.build_acqf(p, st, prune_samples=64)
    Xd = Cf.candidates(p).to(st.device)
    assert Xd.shape[0] * 4 * 1000 >= 1 << 22
    acq.set_option("ozaki", 0)
    v64 = acq(Xd).clone()
    assert st.debug_get("ozaki_check", capacity=16)[0] == 0        # no INT8 call yet
    acq.set_option("ozaki", 1)
    v_auto = acq(Xd)
    state, flagged, batches, flagged_tot, batches_tot, kappa, tol = st.debug_get("ozaki_check", capacity=16).tolist()
    assert state == 1.0 and batches == Xd.shape[0] and flagged <= 0.01 * batches and kappa == 8.0 and tol == 1e-10
    assert float((v_auto - v64).abs().max()) <= 1e-10 * float(v64.abs().max())
    assert torch.equal(acq(Xd), v_auto)
    # a small call stays on the FP64 kernels and does not disturb the state
    v_small = acq(Xd[:16])
    assert float((v_small - v64[:16]).abs().max()) <= 1e-10 * float(v64.abs().max())
    assert st.debug_get("ozaki_check", capacity=16)[0] == state
    # empty t-batch
    empty = acq(Xd[:0])
    assert empty.shape == (0,)
    v0, g0 = acq.forward_backward(Xd[:0])
    assert v0.shape == (0,) and g0.shape == (0, 4, 30)
    # dense data in few dimensions: the posterior variance of every candidate is ~1e-3 of the prior, the digit planes
    # cannot hold 1e-10 of that (measured 4e-10, profiles/r02_guard_calibration.txt) -> the guard sends the state to FP64
    p8 = Cf.zdt1_qnehvi(N=1000, S=16, raw=1200, d=8, q=4)
    st8 = Cf.build_state(p8)
    acq8 = Cf.build_acqf(p8, st8, prune_samples=64)
    X8 = Cf.candidates(p8).to(st8.device)
    acq8.set_option("ozaki", 0)
    v8 = acq8(X8).clone()
    acq8.set_option("ozaki", 1)
    v8a = acq8(X8)
    assert st8.debug_get("ozaki_check", capacity=16)[0] == -1.0
    assert float((v8a - v8).abs().max()) <= 1e-12 * float(v8.abs().max())
    assert torch.equal(acq8(X8), v8a)


def test_int8_guard_redoes_candidates_planted_next_to_training_points():
    """The failure mode of fixed-point digit planes: the absolute error of V is set by the row SCALES, so a candidate
    1e-6 away from a training point -- posterior variance four orders below the prior -- loses relative accuracy.  Such
    q-batches are planted AFTER row 256 (where the round-1 probe never looked) and in a later chunk; the guard must flag
    exactly q-batches that contain them, redo those in FP64, and the posterior variance of every row must then agree with the
    all-FP64 call to 1e-10 relative (the 1e-9 bar of the north star with a 10x margin)."""
    p = Cf.zdt1_qnehvi(N=1000, S=16, raw=1200, d=30, q=4)
    st = Cf.build_state(p)
    acq = Cf.build_acqf(p, st, prune_samples=64)
    X = Cf.candidates(p).clone()
    Xt = torch.as_tensor(p["X"], dtype=DT)
    planted = [300, 301, 777, 1199]
    g = torch.Generator().manual_seed(5)
    for k, bi in enumerate(planted):
        idx = torch.randint(0, Xt.shape[0], (4,), generator=g)
        pts = Xt[idx] + 1e-6 * torch.randn(4, p["d"], dtype=DT, generator=g)
        if k % 2 == 0:
            X[bi] = pts.clamp(0.0, 1.0)            # the whole q-batch sits on training points
        else:
            X[bi, 2] = pts[2].clamp(0.0, 1.0)      # one point of the q-batch only
    Xd = X.to(st.device)
    q, b, M = 4, X.shape[0], st.M
    kmax = 1.0                                      # RBF without outputscale: prior variance 1 in standardised space

    def gram_and_mean():
        G = st.debug_get("Gqq", capacity=b * q * q * M).view(M, b, q, q).clone()
        mu = st.debug_get("mu_raw", capacity=b * q * M).view(M, b * q).clone()
        return G, mu

    acq.set_option("ozaki", 0)
    v64 = acq(Xd).clone()
    G64, mu64 = gram_and_mean()
    var64 = kmax - torch.diagonal(G64, dim1=-2, dim2=-1)               # [M, b, q]
    assert float(var64[:, planted].min()) < 1e-2 * float(var64.median())   # the planted rows really are low-variance rows
    acq.set_option("ozaki", 2)                      # forced INT8, no guard: what the guard protects against
    acq(Xd)
    G8, _ = gram_and_mean()
    err8 = ((torch.diagonal(G8, dim1=-2, dim2=-1) - torch.diagonal(G64, dim1=-2, dim2=-1)).abs() / var64)
    acq.set_option("ozaki", 1)
    v_auto = acq(Xd)
    Ga, mua = gram_and_mean()
    state, flagged, batches, *_ = st.debug_get("ozaki_check", capacity=16).tolist()
    assert state == 1.0 and batches == b
    assert len(planted) <= flagged <= len(planted) + 0.01 * b
    erra = ((torch.diagonal(Ga, dim1=-2, dim2=-1) - torch.diagonal(G64, dim1=-2, dim2=-1)).abs() / var64)
    assert float(erra.max()) <= 1e-10, (float(erra.max()), float(err8.max()))
    assert float(((mua - mu64).abs() / (1.0 + mu64.abs())).max()) <= 1e-10
    # the redone q-batches carry FP64-kernel numbers (a different summation order than the all-FP64 call: the rounding of
    # G ~ 1, 1e-16 sqrt(N), relative to a variance of 1e-4), the forced INT8 path is far off on exactly those rows
    assert float(erra[:, planted].max()) <= 1e-10 < float(err8[:, planted].max())
    ok = torch.isfinite(v64)
    assert bool(ok[[i for i in range(b) if i not in planted]].all())
    assert float((v_auto - v64)[ok].abs().max()) <= 1e-9 * float(v64[ok].abs().max())
    print(f"forced INT8 max rel. variance error {float(err8.max()):.2e} (planted rows {float(err8[:, planted].max()):.2e}), "
          f"guarded {float(erra.max()):.2e}, flagged q-batches {int(flagged)}")
    # a call that is mostly next to the data sends the state to the FP64 kernel for good
    Xnear = (Xt[torch.randint(0, Xt.shape[0], (b * q,), generator=g)] + 1e-6 * torch.randn(b * q, p["d"], dtype=DT, generator=g))
    acq(Xnear.clamp(0.0, 1.0).view(b, q, -1).to(st.device))
    assert st.debug_get("ozaki_check", capacity=16)[0] == -1.0


@pytest.mark.parametrize("kind,alpha", [("dtlz2", 0.01), ("dtlz2", 0.1), ("zdt1", 0.05)])
def test_qnehvi_approximate_partitioning_alpha(kind, alpha):
    """BoFire's `alpha` (data_models/strategies/predictives/qnehvi.py:19 -> qnehvi.py:50): approximate binary partitioning
    ([UPSTREAM] NondominatedPartitioning(alpha)) of every MC sample's front for more than two objectives; exact decomposition
    for two.  Device cell lists must equal the oracle restatement's bit for bit (same traversal order), values to 1e-8."""
    p = small_problem(kind)
    gp = P.oracle_gp(p)
    st = Cf.build_state(p)
    from everest_b200 import acquisition as A

    ops = [P.op_to_oracle(o) for o in p["objective"].ops]
    acq_o = O.QNEHVIOracle(gp, p["ref_point"], p["X"], ops, mc_samples=p["S"], seed=p["sampler_seed"], prune_baseline=True,
                           prune_samples=256, prune_seed=p["sampler_seed"] + 7919, alpha=alpha)
    acq_d = A.qNoisyExpectedHypervolumeImprovement(st, p["ref_point"], torch.as_tensor(p["X"]), p["objective"], prune_baseline=True,
                                                   alpha=alpha, mc_samples=p["S"], seed=p["sampler_seed"], prune_samples=256)
    lo, up, nc = acq_d.cell_bounds()
    assert nc.tolist() == acq_o.n_cells.tolist()
    for s in range(p["S"]):
        c = int(nc[s])
        for dev, ora in ((lo[s, :c], acq_o.cell_lower[s, :c]), (up[s, :c], acq_o.cell_upper[s, :c])):
            assert torch.equal(torch.isinf(dev), torch.isinf(ora)), s
            fin = ~torch.isinf(ora)
            assert torch.allclose(dev[fin], ora[fin], rtol=0.0, atol=1e-10), s
    X = Cf.candidates(p, 16)
    v_o = acq_o.forward(X)
    v_d = acq_d(X.to(st.device)).cpu()
    assert float((v_d - v_o).abs().max()) < 1e-8 * max(float(v_o.abs().max()), 1e-12)
    # dropping cells can only lose hypervolume improvement: the approximate value never exceeds the exact one
    acq_e = A.qNoisyExpectedHypervolumeImprovement(st, p["ref_point"], torch.as_tensor(p["X"]), p["objective"], prune_baseline=True,
                                                   alpha=0.0, mc_samples=p["S"], seed=p["sampler_seed"], prune_samples=256)
    v_e = acq_e(X.to(st.device)).cpu()
    assert bool((v_d <= v_e + 1e-12 * float(v_e.abs().max())).all())
    if len(p["ref_point"]) == 2:
        assert torch.equal(v_d, v_e)


def test_exact_binary_partitioning_equals_local_upper_bound_decomposition():
    """Two different published algorithms (Couckuyt 2012 binary partitioning with threshold 0, Lacour 2017 local upper
    bounds) must tile the same non-dominated region: identical qNEHVI values up to the summation order over the cells."""
    from everest_b200 import acquisition as A

    p = small_problem("dtlz2")
    st = Cf.build_state(p)
    X = Cf.candidates(p, 16).to(st.device)
    acq_l = A.qNoisyExpectedHypervolumeImprovement(st, p["ref_point"], torch.as_tensor(p["X"]), p["objective"], prune_baseline=True,
                                                   mc_samples=p["S"], seed=p["sampler_seed"], prune_samples=256)
    v_l = acq_l(X).cpu()
    n_l = acq_l.cell_bounds()[2]
    acq_b = A.qNoisyExpectedHypervolumeImprovement(st, p["ref_point"], torch.as_tensor(p["X"]), p["objective"], prune_baseline=True,
                                                   mc_samples=p["S"], seed=p["sampler_seed"], prune_samples=256)
    acq_b.alpha = -1.0          # the C ABI's spelling of "binary partitioning, nothing dropped"
    acq_b._reprepare()
    v_b = acq_b(X).cpu()
    n_b = acq_b.cell_bounds()[2]
    assert int(n_b.sum()) != int(n_l.sum())                       # different cell lists ...
    assert float((v_b - v_l).abs().max()) < 1e-12 * float(v_l.abs().max())   # ... same region


def test_joint_resampling_fallback_of_the_cached_root():
    """[UPSTREAM] sample_cached_cholesky falls back to sampling the joint posterior over (X_baseline, X) when the q x q
    conditional root cannot be factorised.  bo_acqf_resample_flagged does the same for the flagged
    q-batches of a forward call.  Forced on every q-batch here: the joint root's last q rows ARE [bl | br] when no jitter
    is needed, so the values must equal the cached-root values to rounding and the oracle's joint computation to 1e-8; a
    normal call re-scores nothing."""
    p = small_problem("zdt1")
    gp = P.oracle_gp(p)
    st = Cf.build_state(p)
    acq_o = P.oracle_acqf(p, gp, prune_samples=128)
    acq_d = Cf.build_acqf(p, st, prune_samples=128)
    X = Cf.candidates(p, 10)
    v_cached = acq_d(X.to(st.device)).cpu()
    assert acq_d.last_resampled == 0 and int(acq_d.last_info.sum()) == 0
    acq_d.set_option("force_joint_fallback", 1)
    acq_d._force_fallback = True
    v_joint = acq_d(X.to(st.device)).cpu()
    acq_d.set_option("force_joint_fallback", 0)
    acq_d._force_fallback = False
    assert acq_d.last_resampled == 10 and acq_d.last_info.cpu().tolist() == [1] * 10
    scale = float(v_cached.abs().max())
    assert float((v_joint - v_cached).abs().max()) < 1e-9 * scale
    v_o = acq_o.forward_joint(X)
    assert float((v_joint - v_o).abs().max()) < 1e-8 * scale
    # and the cached path is back, bit for bit
    assert torch.equal(acq_d(X.to(st.device)).cpu(), v_cached)

"""BoTorch-gated parity harness ("first action on any box: check baseline/_ref/").

The arithmetic below BoFire's call sites lives in botorch / gpytorch, which are NOT installed in the build image and not
in /opt/wheelhouse, so the CPU oracle (oracle/bo_oracle.py) is pinned only on the reference's in-tree arithmetic.  This
module pins it -- and the device path -- on the REAL classes BoFire calls (strategies/predictives/qnehvi.py:39-52,
mobo.py:72-90, sobo.py:64-89, botorch.py:180) wherever botorch can be imported: the whole module is skipped otherwise
(collected and skipped in the build image and on the GPU box).  Put a botorch install on sys.path (e.g. baseline/_ref)
and run

    python -m pytest tests/test_botorch_parity.py -q                 # oracle vs BoTorch on CPU
    python -m pytest tests/test_botorch_parity.py -q -m gpu          # device path vs BoTorch (needs the B200 too)

Base samples are never re-derived: every test reads the base samples BoTorch's own sampler drew and injects them into the
oracle / the device path, so that a mismatch is a mismatch of the arithmetic, not of a seed convention.  The seed
convention itself (flat Sobol dimension = output * n_points + point) has its own test.
"""
import os
import sys

import pytest
import torch

_REF = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")
if os.path.isdir(_REF) and _REF not in sys.path:
    sys.path.insert(0, _REF)

botorch = pytest.importorskip("botorch", reason="botorch is not installed: BoTorch-level parity stays unpinned here")
gpytorch = pytest.importorskip("gpytorch", reason="gpytorch is not installed")

from everest_b200 import configs as Cf  # noqa: E402
from everest_b200 import kernels as K  # noqa: E402
from oracle import bo_oracle as O  # noqa: E402
from tests import problems as P  # noqa: E402

DT = torch.float64
needs_gpu = pytest.mark.gpu


# ----------------------------------------------------------------------------------------------------
# problem dictionary -> the BoTorch model BoFire would hand over (surrogates/single_task_gp.py:48-66,
# botorch_surrogates.py:124-128): SingleTaskGP(covar_module, Normalize, Standardize) per output, ModelListGP
# ----------------------------------------------------------------------------------------------------
def to_gpytorch_kernel(k):
    from gpytorch.kernels import MaternKernel, RBFKernel, ScaleKernel

    if isinstance(k, K.RBFKernel):
        g = RBFKernel(ard_num_dims=len(k.lengthscale) if len(k.lengthscale) > 1 else None, active_dims=tuple(k.active_dims))
        g.lengthscale = torch.tensor(list(k.lengthscale), dtype=DT)
        return g.to(DT)
    if isinstance(k, K.MaternKernel):
        g = MaternKernel(nu=k.nu, ard_num_dims=len(k.lengthscale) if len(k.lengthscale) > 1 else None,
                         active_dims=tuple(k.active_dims))
        g.lengthscale = torch.tensor(list(k.lengthscale), dtype=DT)
        return g.to(DT)
    if isinstance(k, K.ScaleKernel):
        g = ScaleKernel(to_gpytorch_kernel(k.base_kernel)).to(DT)
        g.outputscale = torch.tensor(float(k.outputscale), dtype=DT)
        return g
    if isinstance(k, K.AdditiveKernel):
        from gpytorch.kernels import AdditiveKernel

        return AdditiveKernel(*[to_gpytorch_kernel(c) for c in k.kernels])
    if isinstance(k, K.MultiplicativeKernel):
        from gpytorch.kernels import ProductKernel

        return ProductKernel(*[to_gpytorch_kernel(c) for c in k.kernels])
    raise NotImplementedError(type(k))


def botorch_model(p):
    from botorch.models import ModelListGP, SingleTaskGP
    from botorch.models.transforms import Normalize, Standardize
    from gpytorch.constraints import GreaterThan

    X = torch.as_tensor(p["X"], dtype=DT)
    d = X.shape[1]
    off = torch.as_tensor(p["in_offset"], dtype=DT)
    scl = torch.as_tensor(p["in_scale"], dtype=DT)
    models = []
    for o in p["outputs"]:
        y = torch.as_tensor(o["y"], dtype=DT).unsqueeze(-1)
        m = SingleTaskGP(X, y, covar_module=to_gpytorch_kernel(o["kernel"]),
                         input_transform=Normalize(d, bounds=torch.stack([off, off + scl])),
                         outcome_transform=Standardize(m=1))
        m.likelihood.noise_covar.register_constraint("raw_noise", GreaterThan(1e-9))
        m.likelihood.noise = torch.tensor(float(o["noise"]), dtype=DT)
        m.mean_module.constant = torch.tensor(float(o["mean_const"]), dtype=DT)
        models.append(m.eval())
    return ModelListGP(*models).eval()


def sampler_base_samples(acqf, n_points, M):
    """[S, n_points, M] base samples of an MC acquisition function's sampler after a call."""
    z = acqf.sampler.base_samples.detach().to(DT)
    return z.reshape(z.shape[0], -1, n_points, M)[:, 0]


def small(kind="zdt1", **kw):
    if kind == "zdt1":
        return Cf.zdt1_qnehvi(N=48, S=32, raw=8, d=5, q=2, **kw)
    if kind == "dtlz2":
        return Cf.dtlz2_qnehvi(N=40, S=16, raw=6, d=5, m_obj=3, q=2, **kw)
    return Cf.himmelblau_qlogei(N=80, S=64, raw=16)


def neg_objective():
    from botorch.acquisition.multi_objective.objective import GenericMCMultiOutputObjective

    # MinimizeObjective with bounds (0, 1): -(y - 0) / (1 - 0)   (utils/torch_tools.py:394-398)
    return GenericMCMultiOutputObjective(lambda samples, X=None: -samples)


def sorted_cells(lo, up):
    """rows of [lo | up] sorted lexicographically, empty cells (lo == up in some coordinate) dropped"""
    keep = (up > lo).all(dim=-1)
    rows = torch.cat([lo[keep], up[keep]], dim=-1)
    if rows.shape[0] == 0:
        return rows
    key = rows.clone()
    key[torch.isinf(key)] = 1e300
    order = sorted(range(rows.shape[0]), key=lambda i: tuple(key[i].tolist()))
    return rows[order]


# ----------------------------------------------------------------------------------------------------
# oracle vs BoTorch (CPU)
# ----------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind", ["zdt1", "himmelblau"])
def test_posterior_matches_botorch(kind):
    p = small(kind)
    model = botorch_model(p)
    gp = P.oracle_gp(p)
    Xq = Cf.candidates(p, 6).reshape(-1, p["d"])
    with torch.no_grad():
        post = model.posterior(Xq)
    mean_o, cov_o = gp.posterior(Xq)
    assert float(((post.mean - mean_o).abs() / (mean_o.abs() + 1e-6)).max()) < 1e-9
    var_o = torch.diagonal(cov_o, dim1=-2, dim2=-1).transpose(0, 1)
    assert float(((post.variance - var_o).abs() / var_o).max()) < 1e-9
    with torch.no_grad():
        post_n = model.posterior(Xq, observation_noise=True)
    noise = torch.tensor([o["noise"] for o in p["outputs"]], dtype=DT) * torch.tensor(
        [float(torch.as_tensor(o["y"]).std()) ** 2 for o in p["outputs"]], dtype=DT)
    assert float(((post_n.variance - (var_o + noise)).abs() / (var_o + noise)).max()) < 1e-9


def test_sobol_base_sample_layout_matches_the_sampler():
    """[UPSTREAM] SobolQMCNormalSampler on a ModelListGP posterior: the oracle's claim is flat Sobol dimension
    = output * n_points + point."""
    from botorch.sampling.normal import SobolQMCNormalSampler

    p = small("zdt1")
    model = botorch_model(p)
    Xb = torch.as_tensor(p["X"][:7], dtype=DT)
    sampler = SobolQMCNormalSampler(sample_shape=torch.Size([16]), seed=1234)
    with torch.no_grad():
        sampler(model.posterior(Xb))
    z_bt = sampler.base_samples.detach().to(DT).reshape(16, -1, 7, 2)[:, 0]
    z_or = O.base_samples_points_by_outputs(7, 2, 16, 1234)
    assert float((z_bt - z_or).abs().max()) < 1e-12


@pytest.mark.parametrize("kind", ["zdt1", "dtlz2"])
def test_qnehvi_value_cells_and_cached_root_match_botorch(kind):
    from botorch.acquisition.multi_objective.monte_carlo import qNoisyExpectedHypervolumeImprovement
    from botorch.sampling.normal import SobolQMCNormalSampler

    p = small(kind)
    model = botorch_model(p)
    gp = P.oracle_gp(p)
    ops = [P.op_to_oracle(o) for o in p["objective"].ops]
    # same baseline on both sides: the oracle's pruned set (BoTorch's pruning draws from an unseeded sampler)
    pruned = O.QNEHVIOracle(gp, p["ref_point"], p["X"], ops, mc_samples=p["S"], seed=1, prune_baseline=True, prune_samples=256)
    Xb = torch.as_tensor(p["X"], dtype=DT)[pruned.prune_idx]
    acqf = qNoisyExpectedHypervolumeImprovement(
        model=model, ref_point=p["ref_point"], X_baseline=Xb, prune_baseline=False, objective=neg_objective(), cache_root=True,
        alpha=0.0, sampler=SobolQMCNormalSampler(sample_shape=torch.Size([p["S"]]), seed=p["sampler_seed"]))
    X = Cf.candidates(p, 6)
    with torch.no_grad():
        v_bt = acqf(X)
    nb, q, M = Xb.shape[0], p["q"], len(p["outputs"])
    z = sampler_base_samples(acqf, nb + q, M)
    acq_o = O.QNEHVIOracle(gp, p["ref_point"], Xb, ops, mc_samples=p["S"], seed=p["sampler_seed"], prune_baseline=False,
                           base_samples_baseline=z[:, :nb].contiguous())
    v_o = acq_o.forward(X, zq=z[:, nb:].contiguous())
    assert float((v_o - v_bt).abs().max()) < 1e-8 * float(v_bt.abs().max())
    # cached baseline root and per-sample cell lists (as sets: BoTorch's cell order is its own)
    Lb = acqf._baseline_L.detach().to(DT)
    assert float((Lb.reshape(acq_o.baseline_L.shape) - acq_o.baseline_L).abs().max()) < 1e-8 * float(acq_o.baseline_L.abs().max())
    lo_bt, up_bt = acqf.cell_lower_bounds.detach().to(DT), acqf.cell_upper_bounds.detach().to(DT)
    lo_bt, up_bt = lo_bt.reshape(p["S"], -1, len(ops)), up_bt.reshape(p["S"], -1, len(ops))
    for s_ in range(p["S"]):
        c = int(acq_o.n_cells[s_])
        a = sorted_cells(acq_o.cell_lower[s_, :c], acq_o.cell_upper[s_, :c])
        b = sorted_cells(lo_bt[s_], up_bt[s_])
        assert a.shape == b.shape, (s_, a.shape, b.shape)
        fin = torch.isfinite(a)
        assert torch.equal(fin, torch.isfinite(b)) and float((a[fin] - b[fin]).abs().max()) < 1e-9, s_


def test_pruning_matches_botorch_in_distribution():
    """prune_inferior_points_multi_objective draws from its own unseeded sampler: compare the kept SETS loosely (points
    with a clearly non-zero probability of being non-dominated must be kept by both)."""
    from botorch.acquisition.multi_objective.utils import prune_inferior_points_multi_objective

    p = small("zdt1")
    model = botorch_model(p)
    gp = P.oracle_gp(p)
    X = torch.as_tensor(p["X"], dtype=DT)
    ops = [P.op_to_oracle(o) for o in p["objective"].ops]
    a = O.QNEHVIOracle(gp, p["ref_point"], p["X"], ops, mc_samples=8, seed=1, prune_baseline=True, prune_samples=2048)
    kept_bt = prune_inferior_points_multi_objective(model=model, X=X, ref_point=torch.tensor(p["ref_point"], dtype=DT),
                                                    objective=neg_objective(), num_samples=2048)
    bt_rows = {tuple(r.tolist()) for r in kept_bt}
    or_rows = {tuple(r.tolist()) for r in X[a.prune_idx]}
    inter = len(bt_rows & or_rows)
    assert inter >= 0.8 * max(len(bt_rows), len(or_rows)), (len(bt_rows), len(or_rows), inter)


@pytest.mark.parametrize("m", [2, 3, 4])
def test_box_decomposition_matches_fast_nondominated_partitioning(m):
    from botorch.utils.multi_objective.box_decompositions.non_dominated import FastNondominatedPartitioning
    from botorch.utils.multi_objective.hypervolume import Hypervolume
    from botorch.utils.multi_objective.pareto import is_non_dominated

    g = torch.Generator().manual_seed(m)
    Y = torch.rand(20, m, dtype=DT, generator=g)
    ref = torch.full((m,), 0.05, dtype=DT)
    assert torch.equal(is_non_dominated(Y), O.is_non_dominated(Y))
    part = FastNondominatedPartitioning(ref_point=ref, Y=Y)
    lo_bt, up_bt = part.get_hypercell_bounds()
    Pf, _ = O.pareto_front_above_ref(Y, ref)
    if m == 2:
        lo_o, up_o, _ = O.partition_2d(Pf, ref)
    else:
        lo_o, up_o = O.partition_nd(Pf, ref)
    a, b = sorted_cells(lo_o, up_o), sorted_cells(lo_bt.reshape(-1, m), up_bt.reshape(-1, m))
    assert a.shape == b.shape
    fin = torch.isfinite(a)
    assert torch.equal(fin, torch.isfinite(b)) and float((a[fin] - b[fin]).abs().max()) < 1e-12
    hv_bt = float(Hypervolume(ref_point=ref).compute(Y[is_non_dominated(Y)]))
    assert abs(O.hypervolume(Y[O.is_non_dominated(Y)], ref) - hv_bt) < 1e-12


@pytest.mark.parametrize("name", ["qLogEI", "qEI", "qLogNEI", "qNEI"])
def test_single_objective_acqfs_match_botorch(name):
    from botorch.acquisition import logei, monte_carlo
    from botorch.acquisition.objective import GenericMCObjective
    from botorch.sampling.normal import SobolQMCNormalSampler

    p = small("himmelblau")
    model = botorch_model(p)
    gp = P.oracle_gp(p)
    spec = ("single", P.op_to_oracle(p["objective"].ops[0]))
    lo, hi = p["objective"].ops[0].p0, p["objective"].ops[0].p1
    objective = GenericMCObjective(lambda Y, X=None: -1.0 * ((Y[..., 0] - lo) / (hi - lo)))
    S = p["S"]
    sampler = SobolQMCNormalSampler(sample_shape=torch.Size([S]), seed=7)
    Xo = torch.as_tensor(p["X"], dtype=DT)
    q = 2
    X = Cf.candidates(p, 8).reshape(4, q, p["d"])
    noisy = name in ("qLogNEI", "qNEI")
    if noisy:
        seed_or = O.QScalarOracle(gp, name, spec, Xo, mc_samples=S, seed=7, prune_samples=256)
        Xb = Xo[seed_or.prune_idx]
        cls = logei.qLogNoisyExpectedImprovement if name == "qLogNEI" else monte_carlo.qNoisyExpectedImprovement
        acqf = cls(model=model, X_baseline=Xb, sampler=sampler, objective=objective, prune_baseline=False, cache_root=True)
        with torch.no_grad():
            v_bt = acqf(X)
        nb = Xb.shape[0]
        z = sampler_base_samples(acqf, nb + q, 1)
        acq_o = O.QScalarOracle(gp, name, spec, Xb, mc_samples=S, seed=7, prune_baseline=False)
        acq_o.zb = z[:, :nb].contiguous()
        mean, cov = gp.posterior(Xb)
        fb = mean.unsqueeze(0) + torch.einsum("mij,sjm->sim", acq_o.baseline_L, acq_o.zb)
        acq_o.best_f_s = O.scalar_objective(spec, fb).max(dim=-1).values
        v_o = acq_o.forward(X, zq=z[:, nb:].contiguous())
    else:
        mean, _ = gp.posterior(Xo)
        best_f = float(O.scalar_objective(spec, mean).max())
        cls = logei.qLogExpectedImprovement if name == "qLogEI" else monte_carlo.qExpectedImprovement
        acqf = cls(model=model, best_f=best_f, sampler=sampler, objective=objective)
        with torch.no_grad():
            v_bt = acqf(X)
        z = sampler_base_samples(acqf, q, 1)
        acq_o = O.QScalarOracle(gp, name, spec, Xo, mc_samples=S, seed=7, best_f=best_f)
        v_o = acq_o.forward(X, zq=z)
    scale = max(1.0, float(v_bt.abs().max())) if name.startswith("qLog") else float(v_bt.abs().max())
    assert float((v_o - v_bt).abs().max()) < 1e-7 * scale


def test_qlognehvi_matches_botorch():
    from botorch.acquisition.multi_objective.logei import qLogNoisyExpectedHypervolumeImprovement
    from botorch.sampling.normal import SobolQMCNormalSampler

    p = small("zdt1")
    model = botorch_model(p)
    gp = P.oracle_gp(p)
    ops = [P.op_to_oracle(o) for o in p["objective"].ops]
    pruned = O.QNEHVIOracle(gp, p["ref_point"], p["X"], ops, mc_samples=p["S"], seed=1, prune_baseline=True, prune_samples=256)
    Xb = torch.as_tensor(p["X"], dtype=DT)[pruned.prune_idx]
    acqf = qLogNoisyExpectedHypervolumeImprovement(
        model=model, ref_point=p["ref_point"], X_baseline=Xb, prune_baseline=False, objective=neg_objective(), cache_root=True,
        alpha=0.0, sampler=SobolQMCNormalSampler(sample_shape=torch.Size([p["S"]]), seed=p["sampler_seed"]))
    X = Cf.candidates(p, 6)
    with torch.no_grad():
        v_bt = acqf(X)
    nb, q, M = Xb.shape[0], p["q"], 2
    z = sampler_base_samples(acqf, nb + q, M)
    acq_o = O.QLogNEHVIOracle(gp, p["ref_point"], Xb, ops, mc_samples=p["S"], seed=p["sampler_seed"], prune_baseline=False,
                              base_samples_baseline=z[:, :nb].contiguous())
    v_o = acq_o.forward(X, zq=z[:, nb:].contiguous())
    well = v_bt > -30.0
    # BoTorch pads the per-sample cell lists with zero-volume cells that contribute ~1e-6 relative through the fat tails
    # (oracle/bo_oracle.py header): the tolerance of the log-space value is set accordingly
    assert float((v_o - v_bt)[well].abs().max()) < 1e-4


def test_initialize_q_batch_matches_botorch():
    from botorch.optim.initializers import initialize_q_batch

    from everest_b200 import optim

    g = torch.Generator().manual_seed(0)
    X = torch.rand(40, 2, 3, dtype=DT, generator=g)
    Y = torch.rand(40, dtype=DT, generator=g)
    torch.manual_seed(5)
    out_bt = initialize_q_batch(X=X, acq_vals=Y, n=6, eta=2.0)
    X_bt = out_bt[0] if isinstance(out_bt, tuple) else out_bt
    torch.manual_seed(5)
    X_ev, _ = optim.initialize_q_batch(X, Y, n=6, eta=2.0)
    assert torch.equal(X_bt, X_ev)


# ----------------------------------------------------------------------------------------------------
# device path vs BoTorch (needs botorch AND the B200)
# ----------------------------------------------------------------------------------------------------
@needs_gpu
def test_from_botorch_posterior_on_the_device():
    from everest_b200 import DeviceGPState

    p = small("zdt1")
    model = botorch_model(p)
    st = DeviceGPState.from_botorch(model)
    Xq = Cf.candidates(p, 6).reshape(-1, p["d"])
    with torch.no_grad():
        post = model.posterior(Xq)
    mean_d, var_d = st.posterior(Xq)
    assert float(((mean_d.cpu() - post.mean).abs() / (post.mean.abs() + 1e-6)).max()) < 1e-9
    assert float(((var_d.cpu() - post.variance).abs() / post.variance).max()) < 1e-9


@needs_gpu
def test_device_qnehvi_matches_botorch():
    from botorch.acquisition.multi_objective.monte_carlo import qNoisyExpectedHypervolumeImprovement
    from botorch.sampling.normal import SobolQMCNormalSampler

    from everest_b200 import DeviceGPState
    from everest_b200 import acquisition as A

    p = small("zdt1")
    model = botorch_model(p)
    st = DeviceGPState.from_botorch(model)
    Xb = torch.as_tensor(p["X"], dtype=DT)[:12]
    acqf = qNoisyExpectedHypervolumeImprovement(
        model=model, ref_point=p["ref_point"], X_baseline=Xb, prune_baseline=False, objective=neg_objective(), cache_root=True,
        alpha=0.0, sampler=SobolQMCNormalSampler(sample_shape=torch.Size([p["S"]]), seed=p["sampler_seed"]))
    X = Cf.candidates(p, 6)
    with torch.no_grad():
        v_bt = acqf(X)
    nb, q, M = Xb.shape[0], p["q"], 2
    z = sampler_base_samples(acqf, nb + q, M)
    acq_d = A.qNoisyExpectedHypervolumeImprovement(st, p["ref_point"], Xb, p["objective"], prune_baseline=False,
                                                   mc_samples=p["S"], seed=p["sampler_seed"],
                                                   base_samples_baseline=z[:, :nb].contiguous())
    acq_d.set_base_samples_q(q, z[:, nb:].contiguous())
    v_d = acq_d(X.to(st.device)).cpu()
    assert float((v_d - v_bt).abs().max()) < 1e-8 * float(v_bt.abs().max())

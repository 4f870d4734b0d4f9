"""CPU tests of the host-side mirror: kernel flattening, objective specs against the reference golden
vectors, restart selection, base samples, config builders, C-ABI export list, sharding (gloo, world 2)."""
import ctypes
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

import everest_b200 as E
from everest_b200 import _lib, configs as Cf, distributed as D, kernels as K, objectives as Ob, optim, sampling
from oracle import bo_oracle as O
from tests import problems as P

DT = torch.float64
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_golden.json")))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "everest_b200.h")).read()
    declared = set(re.findall(r"\b(bo_[a-z_]+)\s*\(", hdr))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert _lib.load().bo_version() >= 100


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "everest_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), f


def test_no_gpu_means_loud_failure():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(E.EverestError):
        E.DeviceGPState(np.zeros((4, 2)), [E.SingleTaskGPSpec(kernel=K.RBFKernel([0, 1], [1.0]), y=np.zeros(4))])


def test_flatten_matches_tree_evaluation():
    kc = lambda: K.MaternKernel([0, 1], [0.5, 0.7], nu=2.5)  # noqa: E731
    kh = lambda: K.HammingDistanceKernel({2: 3, 5: 2}, [1.0, 2.0])  # noqa: E731
    km = lambda: K.TanimotoKernel(list(range(7, 19)))  # noqa: E731
    tree = K.AdditiveKernel([
        K.ScaleKernel(K.AdditiveKernel([kc(), K.ScaleKernel(kh(), 0.3)]), 1.5),
        K.ScaleKernel(K.MultiplicativeKernel([kc(), kh(), K.ScaleKernel(km(), 0.8)]), 0.7)])
    flat = K.flatten(tree)
    assert len(flat.leaves) == 5 and len(flat.terms) == 3
    g = torch.Generator().manual_seed(0)
    n, d = 12, 19
    X = torch.rand(n, d, dtype=DT, generator=g)
    X[:, 2:5] = torch.eye(3, dtype=DT)[torch.randint(0, 3, (n,), generator=g)]
    X[:, 5:7] = torch.eye(2, dtype=DT)[torch.randint(0, 2, (n,), generator=g)]
    X[:, 7:] = (torch.rand(n, 12, generator=g) < 0.3).to(DT)
    c = X.mean(0)
    full = O.eval_kernel(P.kernel_to_oracle(tree), X, X, c)
    acc = torch.zeros(n, n, dtype=DT)
    for coef, facs in flat.terms:
        t = torch.full((n, n), coef, dtype=DT)
        for f in facs:
            t = t * O.eval_kernel(P.kernel_to_oracle(flat.leaves[f]), X, X, c)
        acc = acc + t
    assert torch.allclose(acc, full, rtol=1e-13, atol=1e-15)
    with pytest.raises(NotImplementedError):
        K.flatten(object())


def test_objective_specs_match_reference_golden():
    Y = torch.tensor(G["objective_callables"]["Y"], dtype=DT)
    for c in G["objective_callables"]["cases"]:
        kind, idx, *params = c["op"]
        spec = Ob.ObjectiveSpec(kind, idx, *params)
        assert torch.equal(spec(Y), torch.tensor(c["out"], dtype=DT)), c["op"]
    g = G["multiobjective"]
    mo = Ob.MultiObjective([Ob.ObjectiveSpec(o[0], o[1], *o[2:]) for o in g["ops"]])
    assert torch.equal(mo(Y), torch.tensor(g["out"], dtype=DT))
    add = Ob.ScalarObjective([Ob.ObjectiveSpec(o[0], o[1], *o[2:], w=w) for o, w in zip(g["ops"], g["additive"]["weights"])],
                             "additive")
    assert torch.equal(add(Y), torch.tensor(g["additive"]["out"], dtype=DT))
    mg = g["multiplicative"]
    mul = Ob.ScalarObjective([Ob.ObjectiveSpec(o[0], o[1], *o[2:], w=w) for o, w in zip(mg["ops"], mg["weights"])],
                             "multiplicative")
    assert torch.equal(mul(Y), torch.tensor(mg["out"], dtype=DT))
    cons = Ob.constraints_from_sigmoid_objectives([Ob.MaximizeSigmoidObjective(1, 4.0, 1.5), Ob.MinimizeSigmoidObjective(2, 0.5, 2.5)])
    assert [c.eta for c in cons] == G["constraints"]["etas"]
    for c, ref in zip(cons, G["constraints"]["values"]):
        assert torch.equal(c(Y), torch.tensor(ref, dtype=DT))
    with pytest.raises(NotImplementedError):
        Ob.ObjectiveSpec("desirability", 0).to_c()


def test_benchmark_functions_match_reference_golden():
    from everest_b200 import benchmarks as B

    g = G["dtlz2_6d_4obj"]
    assert np.allclose(B.dtlz2(np.array(g["X"]), 4), np.array(g["Y"]), rtol=1e-14, atol=1e-15)
    g = G["himmelblau"]
    assert np.allclose(B.himmelblau(np.array(g["X"])), np.array(g["Y"]), rtol=1e-13, atol=1e-12)
    g = G["detergent"]
    assert np.allclose(B.detergent(np.array(g["X"])), np.array(g["Y"]), rtol=1e-13, atol=1e-13)
    assert np.array_equal(B.DETERGENT_COEF, np.array(g["coef"]))


def test_base_samples_agree_with_oracle_convention():
    z = sampling.base_samples(5, 3, 16, seed=7)
    assert torch.equal(z, O.base_samples_points_by_outputs(5, 3, 16, 7))
    assert sampling.base_samples(0, 2, 8, 1).shape == (8, 0, 2)


def test_initialize_q_batch_semantics():
    g = torch.Generator().manual_seed(0)
    X = torch.rand(64, 2, 3, dtype=DT, generator=g)
    Y = torch.rand(64, dtype=DT, generator=g)
    Xs, idcs = optim.initialize_q_batch(X, Y, n=8, generator=g)
    assert Xs.shape == (8, 2, 3) and len(set(idcs.tolist())) == 8
    assert int(torch.argmax(Y)) in idcs.tolist()  # the arg-max is always kept
    Xall, idall = optim.initialize_q_batch(X, Y, n=64)
    assert torch.equal(Xall, X)
    with pytest.raises(RuntimeError):
        optim.initialize_q_batch(X, Y, n=65)
    Xc, _ = optim.initialize_q_batch(X, torch.ones(64, dtype=DT), n=4, generator=g)  # zero std -> random pick
    assert Xc.shape == (4, 2, 3)
    S = optim.draw_sobol_samples(torch.tensor([[0.0, -1.0], [2.0, 1.0]]), 16, 3, seed=1)
    assert S.shape == (16, 3, 2) and float(S[..., 0].min()) >= 0 and float(S[..., 1].min()) >= -1 and float(S[..., 0].max()) <= 2
    assert torch.equal(optim.apply_fixed_features(S, {1: 0.25})[..., 1], torch.full((16, 3), 0.25, dtype=DT))


def test_gen_candidates_scipy_on_a_known_concave_function():
    class _M:
        device = torch.device("cpu")

    class _Fake:
        model = _M()

        def __call__(self, X):
            return -((X - 0.3) ** 2).sum(dim=(1, 2))

        def forward_backward(self, X):
            return self(X), -2.0 * (X - 0.3)

    g = torch.Generator().manual_seed(1)
    ic = torch.rand(3, 2, 4, dtype=DT, generator=g)
    for mode in ("analytic", "fd"):
        Xf, vals, info = optim.gen_candidates_scipy(ic, _Fake(), torch.zeros(4), torch.ones(4), fixed_features={1: 0.9},
                                                    options={"maxiter": 50, "gradient": mode})
        assert torch.allclose(Xf[..., 1], torch.full((3, 2), 0.9, dtype=DT))          # fixed feature untouched
        assert torch.allclose(Xf[..., [0, 2, 3]], torch.full((3, 2, 3), 0.3, dtype=DT), atol=1e-6)
        assert torch.allclose(vals, torch.full((3,), -0.72, dtype=DT), atol=1e-9)
        # optimum outside the box -> lands on the bound (one-sided differences at the boundary)
        Xb, _, _ = optim.gen_candidates_scipy(ic, _Fake(), torch.full((4,), 0.5), torch.ones(4),
                                              options={"maxiter": 50, "gradient": mode})
        assert torch.allclose(Xb, torch.full((3, 2, 4), 0.5, dtype=DT), atol=1e-9)
    # linear constraints (BoTorch form sum coef x >= rhs, utils/torch_tools.py:45-100) -> SLSQP: maximise -|x - 0.3|^2
    # subject to x0 + x1 >= 1.0 and x2 == x3: optimum x0 = x1 = 0.5, x2 = x3 = 0.3
    ineq = [(torch.tensor([0, 1]), torch.tensor([1.0, 1.0]), 1.0)]
    eq = [(torch.tensor([2, 3]), torch.tensor([1.0, -1.0]), 0.0)]
    Xc, vc, _ = optim.gen_candidates_scipy(ic, _Fake(), torch.zeros(4), torch.ones(4), options={"maxiter": 200},
                                           inequality_constraints=ineq, equality_constraints=eq)
    assert torch.allclose(Xc[..., :2], torch.full((3, 2, 2), 0.5, dtype=DT), atol=1e-5)
    assert torch.allclose(Xc[..., 2:], torch.full((3, 2, 2), 0.3, dtype=DT), atol=1e-5)
    # whole optimize_acqf with constraints: polytope raw samples, feasible result
    cand, val = optim.optimize_acqf(_Fake(), torch.tensor([[0.0] * 4, [1.0] * 4]), q=2, num_restarts=3, raw_samples=32,
                                    options={"maxiter": 100}, seed=0, inequality_constraints=ineq, equality_constraints=eq)
    assert cand.shape == (2, 4) and float((cand[:, 0] + cand[:, 1]).min()) >= 1.0 - 1e-8
    assert float((cand[:, 2] - cand[:, 3]).abs().max()) <= 1e-8 and abs(float(val) + 2 * 0.08) < 1e-6


def test_optimize_acqf_list_sequencing_with_fake_acqfs():
    """optimize_acqf_list (botorch.py:337-356): one candidate per acquisition function, function i sees the candidates of
    0..i-1 as pending points, its own pending points are restored afterwards, every function is activated before use."""
    class _M:
        device = torch.device("cpu")
        d = 3

    log = []

    class _Fake:
        model = _M()

        def __init__(self, name, target, pending=None):
            self.name, self.target, self.X_pending = name, target, pending

        def activate(self):
            log.append(("activate", self.name))
            return self

        def set_X_pending(self, Xp=None):
            self.X_pending = None if Xp is None else torch.as_tensor(Xp, dtype=DT).reshape(-1, 3)
            log.append(("pending", self.name, 0 if self.X_pending is None else self.X_pending.shape[0]))

        def _pen(self, X):
            # repelled from the pending points: the second function must not return the first one's optimum
            if self.X_pending is None:
                return torch.zeros(X.shape[0], dtype=DT), torch.zeros_like(X)
            diff = X[:, :, None, :] - self.X_pending[None, None]
            w = torch.exp(-50.0 * (diff ** 2).sum(-1))
            return w.sum(dim=(1, 2)), (-100.0 * diff * w[..., None]).sum(2)

        def __call__(self, X):
            return -((X - self.target) ** 2).sum(dim=(1, 2)) - self._pen(X)[0]

        def forward_backward(self, X):
            return self(X), -2.0 * (X - self.target) - self._pen(X)[1]

    a = _Fake("a", 0.25)
    b = _Fake("b", 0.25, pending=torch.full((1, 3), 0.9, dtype=DT))
    bounds = torch.tensor([[0.0] * 3, [1.0] * 3])
    cands, vals = optim.optimize_acqf_list([a, b], bounds, num_restarts=4, raw_samples=64, options={"maxiter": 100}, seed=3)
    assert cands.shape == (2, 3) and vals.shape == (2,)
    assert torch.allclose(cands[0], torch.full((3,), 0.25, dtype=DT), atol=1e-5) and abs(float(vals[0])) < 1e-9
    assert float((cands[1] - cands[0]).norm()) > 0.05           # b was repelled from a's candidate
    assert b.X_pending.shape == (1, 3) and a.X_pending is None   # own pending points restored
    assert [e for e in log if e[0] == "activate"] == [("activate", "a"), ("activate", "b")]
    assert ("pending", "b", 2) in log                            # b scored with its own pending point + a's candidate
    with pytest.raises(ValueError):
        optim.optimize_acqf_list([], bounds, 2, 8)
    # with a fixed_features_list every step is an optimize_acqf_mixed(q=1)
    c2, v2 = optim.optimize_acqf_list([a], bounds, 3, 32, fixed_features_list=[{0: 0.0}, {0: 0.25}], options={"maxiter": 50},
                                      seed=1)
    assert abs(float(c2[0, 0]) - 0.25) < 1e-12 and abs(float(v2[0])) < 1e-8


def test_nonlinear_constraints_nchoosek_and_product():
    """Nonlinear inequality constraints as BoFire builds them (utils/torch_tools.py:147-252) through optimize_acqf:
    feasible raw samples from the caller's generator (botorch.py:257-265), SLSQP one restart at a time."""
    class _M:
        device = torch.device("cpu")
        d = 4

    class _Fake:
        model = _M()
        X_pending = None

        def __init__(self, t):
            self.t = t

        def __call__(self, X):
            return -((X - self.t) ** 2).sum(dim=(1, 2))

        def forward_backward(self, X):
            return self(X), -2.0 * (X - self.t)

    bounds = torch.tensor([[0.0] * 4, [1.0] * 4])
    # Product: x0 * x1 <= 0.25 -> the point of the hyperbola closest to (0.8, 0.8) is (0.5, 0.5)
    prod = [optim.product_constraint([0, 1], [1.0, 1.0], rhs=0.25, sign=1)]

    def gen(n, q, seed):
        g = torch.Generator().manual_seed(seed)
        X = torch.rand(n, q, 4, dtype=DT, generator=g)
        X[..., 0] *= 0.5
        X[..., 1] *= 0.5
        return X

    # initialize_q_batch picks the restarts with the global generator (as BoTorch does): pin it, and leave SLSQP's own
    # stopping tolerance room in the coordinates (the value is checked to 1e-6 below; from some restarts SLSQP stops
    # 1.6e-4 away along the flat direction of the constraint)
    torch.manual_seed(0)
    cand, val = optim.optimize_acqf(_Fake(0.8), bounds, q=1, num_restarts=3, raw_samples=32, options={"maxiter": 200}, seed=4,
                                    nonlinear_inequality_constraints=prod, generator=gen)
    assert bool(optim.nonlinear_constraints_satisfied(cand.unsqueeze(0), prod).all())
    assert torch.allclose(cand[0], torch.tensor([0.5, 0.5, 0.8, 0.8], dtype=DT), atol=5e-4)
    assert abs(float(val) + 2 * 0.09) < 1e-6
    # no generator -> the same error BoTorch raises; infeasible start points are rejected
    with pytest.raises(RuntimeError):
        optim.optimize_acqf(_Fake(0.8), bounds, 1, 2, 8, nonlinear_inequality_constraints=prod)
    with pytest.raises(ValueError):
        optim.gen_candidates_scipy(torch.full((1, 1, 4), 0.9, dtype=DT), _Fake(0.8), bounds[0], bounds[1],
                                   nonlinear_inequality_constraints=prod)
    # NChooseK: at most one of x0..x2 non-zero; the generator zeroes the others exactly
    nck = optim.nchoosek_constraints([0, 1, 2], max_count=1, min_count=0)
    assert len(nck) == 1 and nck[0][1] is True
    gen2 = optim.nchoosek_generator(bounds, [([0, 1, 2], 1, 0)])
    Xg = gen2(64, 2, 0)
    assert Xg.shape == (64, 2, 4) and int(((Xg[..., :3] != 0).sum(-1)).max()) <= 1
    assert bool(optim.nonlinear_constraints_satisfied(Xg, nck).all())
    assert not bool(optim.nonlinear_constraints_satisfied(torch.full((1, 1, 4), 0.5, dtype=DT), nck).any())
    cand2, val2 = optim.optimize_acqf(_Fake(0.5), bounds, q=2, num_restarts=3, raw_samples=48, options={"maxiter": 100}, seed=1,
                                      nonlinear_inequality_constraints=nck, generator=gen2)
    assert cand2.shape == (2, 4) and bool(optim.nonlinear_constraints_satisfied(cand2.unsqueeze(0), nck).all())
    assert int(((cand2[:, :3].abs() > 5e-3).sum(-1)).max()) <= 1
    # (SLSQP on the narrow-Gaussian relaxation rarely improves on a start with exact zeros -- the constraint gradient
    # vanishes there; an infeasible refinement is discarded, so the result is at least the best screened raw sample)
    assert abs(float(val2) - float(_Fake(0.5)(cand2.unsqueeze(0))[0])) < 1e-12
    # min_count and both bounds
    both = optim.nchoosek_constraints([0, 1, 2], max_count=2, min_count=1)
    assert len(both) == 2
    Xb = optim.nchoosek_generator(bounds, [([0, 1, 2], 2, 1)])(32, 1, 3)
    nz = (Xb[..., :3] != 0).sum(-1)
    assert int(nz.min()) >= 1 and int(nz.max()) <= 2 and bool(optim.nonlinear_constraints_satisfied(Xb, both).all())


def test_polytope_sampler_and_dense_constraints():
    """sample_q_batches_from_polytope ([UPSTREAM] hit-and-run, reached from botorch.py:384-405 whenever the domain holds
    Linear(In)EqualityConstraints): feasibility, fixed features, inter-point equalities, roughly uniform marginals."""
    b = torch.tensor([[0.0] * 5, [1.0] * 5])
    ineq = [(torch.tensor([0, 1, 2]), torch.tensor([-1.0, -1.0, -1.0]), -1.0), (torch.tensor([3, 4]), torch.tensor([1.0, 1.0]), 0.5)]
    eq = [(torch.tensor([0, 4]), torch.tensor([1.0, -1.0]), 0.0)]
    X = optim.sample_q_batches_from_polytope(512, 2, b, ineq, eq, seed=0)
    assert X.shape == (512, 2, 5) and float(X.min()) >= 0 and float(X.max()) <= 1
    assert float(X[..., :3].sum(-1).max()) <= 1 + 1e-12 and float((X[..., 3] + X[..., 4]).min()) >= 0.5 - 1e-12
    assert float((X[..., 0] - X[..., 4]).abs().max()) < 1e-12
    assert torch.equal(X, optim.sample_q_batches_from_polytope(512, 2, b, ineq, eq, seed=0))  # seeded
    Xf = optim.sample_q_batches_from_polytope(8, 3, b, ineq, None, seed=0, fixed_features={1: 0.25})
    assert torch.equal(Xf[..., 1], torch.full((8, 3), 0.25, dtype=DT))
    ip = [(torch.tensor([[0, 2], [1, 2]]), torch.tensor([1.0, -1.0]), 0.0)]     # x[0, 2] == x[1, 2]
    Xi = optim.sample_q_batches_from_polytope(16, 2, b, ineq, ip, seed=1)
    assert float((Xi[:, 0, 2] - Xi[:, 1, 2]).abs().max()) < 1e-12
    # simplex x0 + x1 + x2 <= 1 in the unit cube: uniform marginal mean of each coordinate is 1/4
    Xs = optim.sample_q_batches_from_polytope(4096, 1, torch.tensor([[0.0] * 3, [1.0] * 3]),
                                              [(torch.tensor([0, 1, 2]), torch.tensor([-1.0] * 3), -1.0)], None, seed=3)
    assert float((Xs.mean(dim=(0, 1)) - 0.25).abs().max()) < 0.02
    A, r = optim.dense_linear_constraints(ineq, 2, 5)
    assert A.shape == (4, 10) and r.tolist() == [-1.0, -1.0, 0.5, 0.5] and A[1, 5:8].tolist() == [-1.0] * 3
    with pytest.raises(RuntimeError):
        optim.dense_linear_constraints([(torch.tensor([7]), torch.tensor([1.0]), 0.0)], 1, 5)
    with pytest.raises(ValueError):   # empty polytope
        optim.sample_q_batches_from_polytope(4, 1, b, [(torch.tensor([0]), torch.tensor([1.0]), 2.0)], None, seed=0)


def test_config_builders_and_oracle_conversion():
    for p in (Cf.zdt1_qnehvi(scale=0.01), Cf.dtlz2_qnehvi(scale=0.02), Cf.himmelblau_qlogei(scale=0.05),
              Cf.detergent_qnehvi(), Cf.mixed_tanimoto_qlogei(scale=0.004, n_bits=128)):
        gp = P.oracle_gp(p)
        mean, cov = gp.posterior(torch.as_tensor(p["X"][:3], dtype=DT))
        assert torch.isfinite(mean).all() and torch.isfinite(cov).all()
        assert Cf.candidates(p, 5).shape[-1] == p["d"]
    full = Cf.zdt1_qnehvi()
    assert full["X"].shape == (2000, 30) and full["q"] == 4 and full["S"] == 512 and full["raw_samples"] == 16384


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 16, 16385):
        for world in (1, 2, 3, 8):
            spans = [D.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from everest_b200 import distributed as D
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % sys.argv[2], rank=int(sys.argv[3]), world_size=2)
g = torch.Generator().manual_seed(3)
X = torch.rand(37, 2, 3, dtype=torch.double, generator=g)
acq = lambda x: (x.sum(dim=(1, 2)) * 1.5 - x[:, 0, 0] ** 2)
full = acq(X)
got = D.sharded_forward(acq, X)
assert torch.equal(got, full), "sharded_forward mismatch"
v, i = D.sharded_argmax(acq, X)
assert i == int(torch.argmax(full)) and v == float(full.max())
# sharded ask: screen in slices, each rank refines its share of the restarts, one arg-max exchange + one broadcast
class _M:
    device = torch.device("cpu")
class _Fake:
    model = _M()
    def __call__(self, X):
        return -((X - 0.3) ** 2).sum(dim=(1, 2)) + 0.1 * torch.cos(9.0 * X[:, 0, 0])
    def forward_backward(self, X):
        Xr = X.clone().requires_grad_(True)
        v = self(Xr); v.sum().backward()
        return v.detach(), Xr.grad
bounds = torch.tensor([[0.0] * 3, [1.0] * 3], dtype=torch.double)
cand, val = D.sharded_optimize_acqf(_Fake(), bounds, q=2, num_restarts=4, raw_samples=64, options={"maxiter": 60}, seed=5)
both = [torch.empty(7, dtype=torch.double) for _ in range(2)]
dist.all_gather(both, torch.cat([cand.reshape(-1), val.reshape(1)]))
assert torch.equal(both[0], both[1]), "ranks disagree on the winning candidate"
from everest_b200 import optim
X_rnd = optim.draw_sobol_samples(bounds, 64, 2, seed=5)
assert float(val) >= float(_Fake()(X_rnd).max()) - 1e-12 and abs(float(_Fake()(cand.unsqueeze(0))[0]) - float(val)) < 1e-12
dist.barrier()
dist.destroy_process_group()
print("rank", sys.argv[3], "ok")
"""


def test_sharded_forward_world_size_2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    port = str(29600 + os.getpid() % 300)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)


def test_bofire_data_model_adapter():
    """map_kernel / objectives_from_outputs against the reference's own data models (importable here; the
    GPU box has no /root/reference, so skip there)."""
    ref = "/root/reference"
    if not os.path.isdir(ref):
        pytest.skip("reference tree not available")
    sys.path.insert(0, ref)
    try:
        import bofire.data_models.kernels.api as dk
        from bofire.data_models.domain.api import Outputs
        from bofire.data_models.features.api import ContinuousOutput
        from bofire.data_models.objectives.api import MaximizeObjective, MaximizeSigmoidObjective, MinimizeObjective
    except Exception as exc:  # pragma: no cover
        pytest.skip(f"bofire data models not importable: {exc}")
    finally:
        sys.path.remove(ref)
    from everest_b200 import bofire_adapter as BA

    layout = {"x1": [0], "x2": [1], "c1": [2, 3, 4], "c2": [5, 6]}
    mapper = lambda feats: [i for f in feats for i in layout[f]]  # noqa: E731
    dm = dk.AdditiveKernel(kernels=[
        dk.ScaleKernel(base_kernel=dk.MaternKernel(ard=True, nu=2.5, features=["x1", "x2"])),
        dk.ScaleKernel(base_kernel=dk.MultiplicativeKernel(kernels=[
            dk.RBFKernel(ard=False, features=["x1", "x2"]), dk.HammingDistanceKernel(ard=True, features=["c1", "c2"])]))])
    hyper = {"kernels.0": {"outputscale": 1.5}, "kernels.0.base_kernel": {"lengthscale": [0.3, 0.4]},
             "kernels.1": {"outputscale": 0.5}, "kernels.1.base_kernel.kernels.0": {"lengthscale": [0.9]},
             "kernels.1.base_kernel.kernels.1": {"lengthscale": [1.0, 2.0, 0, 0, 0]}}
    spec = BA.map_kernel(dm, active_dims=list(range(7)), features_to_idx_mapper=mapper, hyper=hyper)
    flat = K.flatten(spec)
    assert [type(l).__name__ for l in flat.leaves] == ["MaternKernel", "RBFKernel", "HammingDistanceKernel"]
    assert flat.terms == [(1.5, [0]), (0.5, [1, 2])]
    assert flat.leaves[2].categorical_features == {2: 3, 5: 2}  # same one-hot layout as mapper.py:223-245 (offset by active dims)
    with pytest.raises(RuntimeError):
        BA.map_kernel(dk.HammingDistanceKernel(features=["x1"]), list(range(7)), mapper)
    with pytest.raises(NotImplementedError):
        BA.map_kernel(dk.LinearKernel(), list(range(7)), mapper)
    outs = Outputs(features=[ContinuousOutput(key="a", objective=MaximizeObjective(w=1.0, bounds=(0, 2))),
                             ContinuousOutput(key="b", objective=MaximizeSigmoidObjective(w=1.0, steepness=4.0, tp=1.5)),
                             ContinuousOutput(key="c", objective=MinimizeObjective(w=1.0))])
    mo, cons = BA.objectives_from_outputs(outs)
    assert [(o.kind, o.idx, o.p0, o.p1) for o in mo.ops] == [("max", 0, 0.0, 2.0), ("min", 2, 0.0, 1.0)]
    assert [(c.idx, c.sign, c.tp, c.eta) for c in cons] == [(1, -1.0, 1.5, 0.25)]


def test_strategy_host_logic_input_space_and_duplicates():
    """Host pieces of the strategy mirror (everest_b200/strategy.py <-> botorch.py:226-296, 408-467, 696-724)."""
    from everest_b200.strategy import InputSpace, drop_duplicate_rows

    # 2 continuous + one-hot(3) + one-hot(2)
    sp = InputSpace(bounds=np.array([[0.0] * 7, [1.0] * 7]), categorical_groups={2: 3, 5: 2}, fixed_features={1: 0.5},
                    allowed_categories={2: [0, 2]})
    assert sp.d == 7 and not sp.is_fully_combinatorial()
    combos = sp.categorical_combinations()
    assert len(combos) == 2 * 2
    for ff in combos:
        assert ff[1] == 0.5 and sum(ff[c] for c in (2, 3, 4)) == 1.0 and sum(ff[c] for c in (5, 6)) == 1.0 and ff[3] == 0.0
    spc = InputSpace(bounds=np.array([[0.0] * 4, [1.0, 1.0, 1.0, 3.0]]), categorical_groups={0: 3}, discrete_values={3: [1.0, 2.0, 3.0]})
    assert spc.is_fully_combinatorial() and len(spc.categorical_combinations()) == 9
    X = np.array([[0.1, 0.2], [0.3, 0.4], [0.1, 0.2], [0.5, 0.6], [0.3, 0.4]])
    Y = np.arange(5.0)[:, None]
    Xd, Yd = drop_duplicate_rows(X, Y)           # keep="first", original order
    assert Xd.tolist() == [[0.1, 0.2], [0.3, 0.4], [0.5, 0.6]] and Yd[:, 0].tolist() == [0.0, 1.0, 3.0]
    with pytest.raises(ValueError):
        InputSpace(bounds=np.zeros(3))


def test_device_sobol_integer_pipeline_matches_torch_engine():
    """csrc/sobol.cu reproduces torch's SobolEngine(scramble=True) bit for bit; the same integer arithmetic is replayed
    here in torch (the CUDA kernels themselves are checked on the GPU box): GF(2) scramble with the packed
    unit-lower-triangular rows, Gray-code draw, float32 first point."""
    for dim, seed, n in [(5, 3, 40), (257, 11, 17)]:
        ss, shift, rows = sampling.sobol_scramble_inputs(dim, seed)
        v = ss.unsqueeze(-1) & rows.unsqueeze(1)            # [dim, j, p]
        par = torch.zeros_like(v)
        for b in range(30):
            par ^= (v >> b) & 1                              # parity of popc(row & v)
        scr = (par << (29 - torch.arange(30))).sum(-1)
        eng = torch.quasirandom.SobolEngine(dim, scramble=True, seed=seed)
        assert torch.equal(scr, eng.sobolstate) and torch.equal(shift, eng.shift)
        u = eng.draw(n, dtype=DT)
        x = shift.unsqueeze(0).repeat(n, 1)
        for s_ in range(n):
            gray, b = s_ ^ (s_ >> 1), 0
            while gray:
                if gray & 1:
                    x[s_] ^= scr[:, b]
                gray >>= 1
                b += 1
        mine = x.to(DT) * 2.0 ** -30
        mine[0] = (x[0].float() * torch.tensor(2.0 ** -30, dtype=torch.float32)).to(DT)
        assert torch.equal(mine, u)
    with pytest.raises(ValueError):
        sampling.sobol_scramble_inputs(0, 1)


def test_base_samples_fall_back_to_iid_above_sobol_maxdim():
    """[UPSTREAM] get_sampler: IIDNormalSampler above SobolEngine.MAXDIM (pruning > 10600 baseline points with two outputs)."""
    n_points = sampling.MAXDIM // 2 + 5
    z = sampling.base_samples(n_points, 2, 4, seed=3)
    assert z.shape == (4, n_points, 2) and bool(torch.isfinite(z).all())
    assert torch.equal(z, sampling.base_samples(n_points, 2, 4, seed=3))
    assert abs(float(z.mean())) < 0.05 and abs(float(z.std()) - 1.0) < 0.05
    # at and below MAXDIM the Sobol pipeline is untouched
    assert torch.equal(sampling.base_samples(3, 2, 8, seed=1), O.base_samples_points_by_outputs(3, 2, 8, 1))


def test_linear_feasibility_uses_a_relative_tolerance():
    """Mixture constraint sum x = 100: a residual of 1e-6 (SLSQP's stopping accuracy) is feasible, 1e-2 is not."""
    eq = [(torch.arange(3), torch.ones(3, dtype=DT), 100.0)]
    ineq = [(torch.tensor([0]), torch.tensor([1.0], dtype=DT), 10.0)]   # x0 >= 10
    X = torch.tensor([[[30.0, 30.0, 40.0 + 1e-6]], [[30.0, 30.0, 40.01]], [[9.0, 41.0, 50.0]], [[10.0 - 1e-7, 40.0, 50.0]]], dtype=DT)
    ok = optim.linear_feasibility(X, ineq, eq)
    assert ok.tolist() == [True, False, False, True]


def test_oracle_best_feasible_objective_and_constrained_pruning():
    """[UPSTREAM] compute_best_feasible_objective as restated in the oracle: infeasible entries are -inf while every
    leading index keeps a feasible point, otherwise the pessimistic lower bound replaces them."""
    p = Cf.zdt1_qnehvi(N=30, S=8, raw=4, d=3, q=1)
    gp = P.oracle_gp(p)
    spec = ("single", ("min", 1, 0.0, 1.0))
    X = torch.as_tensor(p["X"], dtype=DT)
    mean, _ = gp.posterior(X)
    obj = O.scalar_objective(spec, mean)
    cons = [(0, 1.0, 0.5, 0.1)]
    feas = mean[:, 0] <= 0.5
    assert bool(feas.any()) and not bool(feas.all())
    bf = O.best_feasible_objective(gp, spec, cons, mean, obj, X)
    assert float(bf) == float(obj[feas].max()) < float(obj.max())
    none = [(0, 1.0, -5.0, 0.1)]
    g1, g2 = torch.Generator().manual_seed(5), torch.Generator().manual_seed(5)
    lb = O.objective_lower_bound(gp, spec, X, generator=g1)
    assert lb <= 0.0 and float(O.best_feasible_objective(gp, spec, none, mean, obj, X, generator=g2)) == lb
    a = O.QScalarOracle(gp, "qLogNEI", spec, X, mc_samples=8, seed=2, prune_samples=64, constraints=cons)
    # no pruned-in point may be infeasible in every pruning sample: the survivors are feasible somewhere
    assert len(a.prune_idx) >= 1 and bool(torch.isfinite(a.best_f_s).all())


def test_specs_from_botorch_model_literal_mixed_single_task_gp_with_stand_ins():
    """The attribute walk of bofire_adapter.specs_from_botorch_model on duck-typed stand-ins of the objects BoFire's literal
    MixedSingleTaskGPSurrogate builds (mixed_single_task_gp.py:46-112): Chained(Normalize on the ordinal columns,
    OneHotToNumeric) + Scale(K_c + Scale(CategoricalKernel)) + Scale(K_c * CategoricalKernel) on [ordinal..., integer codes].
    The mapped spec must evaluate, on the ONE-HOT layout, to the same Gram matrix as the stand-in tree on the coded layout."""
    from everest_b200.bofire_adapter import specs_from_botorch_model

    def T(v):
        return torch.as_tensor(v, dtype=torch.double)

    class _K:
        active_dims = None

    def mk(name, **kw):
        return type(name, (_K,), {})().__class__, kw

    def kern(name, **kw):
        obj = type(name, (_K,), {})()
        for k, v in kw.items():
            setattr(obj, k, v)
        return obj

    g = torch.Generator().manual_seed(0)
    N, n_ord = 9, 2
    cats = {2: 3, 5: 2}                       # one-hot blocks of the BoFire layout [a, b | c0 c1 c2 | e0 e1]
    codes = torch.stack([torch.randint(0, 3, (N,), generator=g), torch.randint(0, 2, (N,), generator=g)], dim=1)
    Xo = torch.rand(N, n_ord, dtype=torch.double, generator=g) * T([4.0, 2.0]) + T([-1.0, 3.0])
    onehot = torch.cat([torch.nn.functional.one_hot(codes[:, 0], 3), torch.nn.functional.one_hot(codes[:, 1], 2)], dim=1).double()
    X_bofire = torch.cat([Xo, onehot], dim=1)                         # [N, 7]
    X_coded = torch.cat([Xo, codes.double()], dim=1)                 # what MixedSingleTaskGP is constructed with

    class Normalize:
        indices = torch.tensor([0, 1])
        offset = T([[-1.0, 3.0]])
        coefficient = T([[4.0, 2.0]])

    class OneHotToNumeric:
        dim = 7
        categorical_features = cats

        def untransform(self, X):
            return torch.cat([X[:, :n_ord], torch.nn.functional.one_hot(X[:, n_ord].long(), 3).double(),
                              torch.nn.functional.one_hot(X[:, n_ord + 1].long(), 2).double()], dim=1)

    class Chain(dict):
        pass

    ls_c1, ls_c2, ls_h1, ls_h2 = T([[0.7, 1.3]]), T([[0.4, 0.9]]), T([[0.8, 2.0]]), T([[1.5, 0.6]])
    cont1 = kern("MaternKernel", active_dims=torch.tensor([0, 1]), lengthscale=ls_c1, nu=2.5)
    cont2 = kern("MaternKernel", active_dims=torch.tensor([0, 1]), lengthscale=ls_c2, nu=2.5)
    cat1 = kern("CategoricalKernel", active_dims=torch.tensor([2, 3]), lengthscale=ls_h1)
    cat2 = kern("CategoricalKernel", active_dims=torch.tensor([2, 3]), lengthscale=ls_h2)
    tree = kern("AdditiveKernel", kernels=[
        kern("ScaleKernel", outputscale=T(1.7), base_kernel=kern("AdditiveKernel", kernels=[cont1, kern("ScaleKernel", outputscale=T(0.6), base_kernel=cat1)])),
        kern("ScaleKernel", outputscale=T(0.9), base_kernel=kern("ProductKernel", kernels=[cont2, cat2]))])

    class Const:
        constant = T(0.25)

    class Lik:
        noise = T([3e-3])

    class Std:
        means, stdvs = T([[2.0]]), T([[1.5]])

    class Model:
        training = True
        input_transform = Chain(tf1=Normalize(), tf2=OneHotToNumeric())
        train_inputs = (X_coded,)
        train_targets = T(np.linspace(-1, 1, N))
        covar_module = tree
        mean_module = Const()
        likelihood = Lik()
        outcome_transform = Std()

    X_full, (spec,) = specs_from_botorch_model(Model())
    assert np.array_equal(X_full, X_bofire.numpy())
    assert spec.in_offset.tolist() == [-1.0, 3.0, 0, 0, 0, 0, 0] and spec.in_scale.tolist() == [4.0, 2.0, 1, 1, 1, 1, 1]
    assert spec.noise == 3e-3 and spec.mean_const == 0.25 and (spec.y_mean, spec.y_std) == (2.0, 1.5)
    assert np.allclose(spec.y, np.linspace(-1, 1, N) * 1.5 + 2.0)
    flat = K.flatten(spec.kernel)
    assert [type(lf).__name__ for lf in flat.leaves] == ["MaternKernel", "HammingDistanceKernel", "MaternKernel", "HammingDistanceKernel"]
    assert flat.leaves[1].categorical_features == {2: 3, 5: 2} and list(flat.leaves[1].lengthscale) == [0.8, 2.0]
    # Gram matrix: oracle evaluation of the mapped spec on the normalised one-hot layout vs the coded formula by hand
    Xn = (X_bofire - T(spec.in_offset)) / T(spec.in_scale)
    G = O.eval_kernel(P.kernel_to_oracle(spec.kernel), Xn, Xn, Xn.mean(dim=0), same=True)

    def matern(ls):
        z = Xn[:, :2] / ls
        r = torch.cdist(z, z)
        return (1 + 5 ** 0.5 * r + 5.0 / 3.0 * r ** 2) * torch.exp(-(5 ** 0.5) * r)

    def catk(ls):
        delta = (codes.unsqueeze(1) != codes.unsqueeze(0)).double()
        return torch.exp(-(delta / ls).mean(-1))

    G_ref = 1.7 * (matern(ls_c1) + 0.6 * catk(ls_h1)) + 0.9 * (matern(ls_c2) * catk(ls_h2))
    assert torch.allclose(G, G_ref, rtol=1e-10, atol=1e-12)
    # FilterFeatures without the full training matrix cannot be handed over silently
    class FilterFeatures:
        feature_indices = torch.tensor([0, 1])

    class Sub(Model):
        input_transform = Chain(tcompatibilize=FilterFeatures(), tf2=Normalize())
        covar_module = cont1

    with pytest.raises(ValueError):
        specs_from_botorch_model(Sub())
    Xf, (s2,) = specs_from_botorch_model(Sub(), X_train=X_bofire)
    assert list(s2.kernel.active_dims) == [0, 1] and s2.in_scale.tolist()[:2] == [4.0, 2.0] and Xf.shape == (N, 7)


def test_bench_clock_sampler_brackets_the_timed_region():
    """bench.py's nvidia-smi poller: samples inside [mark_begin, mark_end] are reported (median clock, throttle reasons); a
    region that slipped between two polls falls back to the nearest samples and says so."""
    import datetime
    import importlib.util
    import tempfile
    import time

    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)

    class Done:
        def terminate(self):
            pass

        def wait(self, timeout=None):
            pass

    def sampler(t_begin, t_end, now):
        cs = bench.ClockSampler(0)
        fd, path = tempfile.mkstemp(suffix=".csv")
        os.close(fd)
        with open(path, "w") as f:
            for k in range(10):
                ts = datetime.datetime.fromtimestamp(now + k * 0.05).strftime("%Y/%m/%d %H:%M:%S.%f")[:-3]
                cap = "Active" if k == 3 else "Not Active"
                f.write(f"{ts}, 0, {1900 + k}, 1965, {300 + k}.5, 0x0000000000000004, Not Active, Not Active, Not Active, {cap}\n")
        cs.proc, cs.path, cs.t_begin, cs.t_end = Done(), path, t_begin, t_end
        return cs.stop()

    now = time.time()
    out = sampler(now + 0.12, now + 0.22, now)           # polls 3 and 4 fall inside
    assert out["samples"] == 2 and out["sm_mhz"] == 1903.5 and out["sm_max_mhz"] == 1965.0
    assert out["reasons"] == ["sw_power_cap"] and "nearest_samples_only" not in out
    out = sampler(now + 0.221, now + 0.224, now)         # a 3 ms region between two polls
    assert out["samples"] == 2 and out["nearest_samples_only"] is True and out["sm_mhz"] in (1904.0, 1904.5, 1903.5)
    empty = bench.ClockSampler(0).stop()                 # poller never started
    assert empty["samples"] == 0 and empty["sm_mhz"] is None

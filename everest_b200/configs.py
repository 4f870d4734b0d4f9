"""The five BASELINE.json configurations as neutral problem dictionaries (synthetic data, fixed
hyper-parameters -- BASELINE.md section 2).  `scale` < 1 shrinks N / raw samples / S for parity tests."""
import math

import numpy as np

from . import benchmarks as B
from . import kernels as K
from .objectives import MaximizeObjective, MinimizeObjective, MultiObjective, ScalarObjective


def _sized(full, scale, lo=4):
    return max(lo, int(round(full * scale)))


def zdt1_qnehvi(scale=1.0, N=None, q=4, S=None, raw=None, d=30, seed=0):
    """Config 3 (headline): ZDT1 30-D, 2 objectives (Minimize), N=2000, q=4, 512 MC samples, 16384 raw samples,
    ref point {y1: 1, y2: 5} (tutorials/benchmarks/011-ZDT1.ipynb), RBF-ARD per output."""
    N = N or _sized(2000, scale)
    S = S or _sized(512, scale, 16)
    raw = raw or _sized(16384, scale, 8)
    rng = np.random.default_rng(seed)
    X = rng.random((N, d))
    Y = B.zdt1(X)
    ls = [0.3 * math.sqrt(d)] * d
    outputs = [dict(kernel=K.RBFKernel(list(range(d)), ls), y=Y[:, m], noise=1e-4, mean_const=0.0) for m in range(2)]
    return dict(name="zdt1_30d_qnehvi", d=d, X=X, Y=Y, outputs=outputs, bounds=np.array([[0.0] * d, [1.0] * d]),
                in_offset=np.zeros(d), in_scale=np.ones(d), acqf="qnehvi",
                objective=MultiObjective([MinimizeObjective(0), MinimizeObjective(1)]), ref_point=[-1.0, -5.0],
                q=q, S=S, raw_samples=raw, num_restarts=8, cand_seed=0, sampler_seed=1234)


def dtlz2_qnehvi(scale=1.0, N=None, q=8, S=None, raw=None, d=6, m_obj=4, seed=0):
    """Config 4: DTLZ2 6-D, 4 objectives, N=1000, q=8, ref 1.1 (benchmarks/multi.py:68-70)."""
    N = N or _sized(1000, scale)
    S = S or _sized(512, scale, 16)
    raw = raw or _sized(1024, scale, 8)
    rng = np.random.default_rng(seed)
    X = rng.random((N, d))
    Y = B.dtlz2(X, m_obj)
    ls = [0.4 * math.sqrt(d)] * d
    outputs = [dict(kernel=K.RBFKernel(list(range(d)), ls), y=Y[:, m], noise=1e-4, mean_const=0.0) for m in range(m_obj)]
    return dict(name="dtlz2_6d_4obj_qnehvi", d=d, X=X, Y=Y, outputs=outputs, bounds=np.array([[0.0] * d, [1.0] * d]),
                in_offset=np.zeros(d), in_scale=np.ones(d), acqf="qnehvi",
                objective=MultiObjective([MinimizeObjective(i) for i in range(m_obj)]), ref_point=[-1.1] * m_obj,
                q=q, S=S, raw_samples=raw, num_restarts=8, cand_seed=0, sampler_seed=1234)


def himmelblau_qlogei(scale=1.0, N=None, S=None, raw=None, seed=0):
    """Config 2: Himmelblau, SoboStrategy qLogEI, ScaleKernel(Matern-5/2 ARD), N=500, raw_samples=4096, q=1."""
    N = N or _sized(500, scale)
    S = S or _sized(512, scale, 16)
    raw = raw or _sized(4096, scale, 8)
    rng = np.random.default_rng(seed)
    X = rng.random((N, 2)) * 12.0 - 6.0
    y = B.himmelblau(X)
    kern = K.ScaleKernel(K.MaternKernel([0, 1], [0.2, 0.2], nu=2.5), outputscale=1.0)
    outputs = [dict(kernel=kern, y=y, noise=1e-4, mean_const=0.0)]
    return dict(name="himmelblau_qlogei", d=2, X=X, Y=y[:, None], outputs=outputs,
                bounds=np.array([[-6.0, -6.0], [6.0, 6.0]]), in_offset=np.array([-6.0, -6.0]),
                in_scale=np.array([12.0, 12.0]), acqf="qlogei",
                objective=ScalarObjective([MinimizeObjective(0)], "single"), q=1, S=S, raw_samples=raw,
                num_restarts=8, cand_seed=0, sampler_seed=1234)


def detergent_qnehvi(N=5, S=512, raw=1024, seed=0):
    """Config 1: Detergent README loop -- 5 inputs, 5 Maximize outputs, N=2..5 points, q=1, RBF-ARD."""
    import torch

    from .optim import sample_q_batches_from_polytope

    lo, hi = B.DETERGENT_BOUNDS
    # the two LinearInequalityConstraints of benchmarks/detergent.py:66-77 (0.2 <= sum x <= 0.4) in BoTorch's
    # sum coef x >= rhs form, as get_linear_constraints (utils/torch_tools.py:45-100) hands them to optimize_acqf
    ineq = [(torch.arange(5), torch.ones(5, dtype=torch.double), 0.2),
            (torch.arange(5), -torch.ones(5, dtype=torch.double), -0.4)]
    # initial experiments like RandomStrategy: uniform in the constrained polytope (strategies/random.py:180-353)
    X = sample_q_batches_from_polytope(N, 1, torch.as_tensor(np.stack([lo, hi])), ineq, None, seed=seed, n_burnin=256,
                                       n_thinning=4)[:, 0, :].numpy()
    Y = B.detergent(X)
    outputs = [dict(kernel=K.RBFKernel(list(range(5)), [0.5] * 5), y=Y[:, m], noise=1e-4, mean_const=0.0) for m in range(5)]
    obj = MultiObjective([MaximizeObjective(i) for i in range(5)])
    ref = Y.min(axis=0).tolist()  # infer_ref_point: worst observed objective value per output
    return dict(name="detergent_qnehvi", d=5, X=X, Y=Y, outputs=outputs, bounds=np.stack([lo, hi]), in_offset=lo,
                in_scale=hi - lo, acqf="qnehvi", objective=obj, ref_point=ref, q=1, S=S, raw_samples=raw,
                num_restarts=8, cand_seed=0, sampler_seed=1234, inequality_constraints=ineq)


def mixed_tanimoto_qlogei(scale=1.0, N=None, n_bits=2048, S=None, n_choices=None, seed=0):
    """Config 5: 2 continuous + 2048-bit fingerprint + 2 categoricals (4 and 6 levels, one-hot), N=5000,
    MixedTanimotoGP-style composite kernel (surrogates/mixed_tanimoto_gp.py:165-215):
    (s1 Kc + s2 Km + s3 Kh) + (s4 Kc * s5 Km * s6 Kh); discrete candidate set, q=1."""
    N = N or _sized(5000, scale)
    S = S or _sized(512, scale, 16)
    n_choices = n_choices or _sized(16384, scale, 8)
    rng = np.random.default_rng(seed)

    def draw(n):
        cont = rng.random((n, 2))
        bits = (rng.random((n, n_bits)) < 0.03).astype(np.float64)
        c1 = np.eye(4)[rng.integers(0, 4, n)]
        c2 = np.eye(6)[rng.integers(0, 6, n)]
        return np.concatenate([cont, bits, c1, c2], axis=1)

    X = draw(N)
    w = rng.normal(size=16)
    y = X[:, 2:18] @ w + 0.5 * X[:, 0] - X[:, 1] ** 2 + 0.3 * X[:, 2 + n_bits] + 0.05 * rng.normal(size=N)
    d = X.shape[1]
    cont_dims, bit_dims = [0, 1], list(range(2, 2 + n_bits))
    cats = {2 + n_bits: 4, 2 + n_bits + 4: 6}

    def kc():
        return K.MaternKernel(cont_dims, [0.5, 0.5], nu=2.5)

    def km():
        return K.TanimotoKernel(bit_dims)

    def kh():
        return K.HammingDistanceKernel(cats, [1.0, 2.0])

    kern = K.AdditiveKernel([
        K.AdditiveKernel([K.ScaleKernel(kc(), 0.5), K.ScaleKernel(km(), 1.0), K.ScaleKernel(kh(), 0.3)]),
        K.MultiplicativeKernel([K.ScaleKernel(kc(), 0.7), K.ScaleKernel(km(), 0.9), K.ScaleKernel(kh(), 0.8)])])
    outputs = [dict(kernel=kern, y=y, noise=1e-2, mean_const=0.0)]
    choices = draw(n_choices)
    return dict(name="mixed_tanimoto_qlogei", d=d, X=X, Y=y[:, None], outputs=outputs, bounds=None,
                in_offset=np.zeros(d), in_scale=np.ones(d), acqf="qlogei",
                objective=ScalarObjective([MaximizeObjective(0)], "single"), q=1, S=S, raw_samples=n_choices,
                num_restarts=8, choices=choices, cand_seed=0, sampler_seed=1234)


def build_state(problem, device=None):
    """Problem dictionary -> factorised DeviceGPState."""
    from .model import DeviceGPState, SingleTaskGPSpec

    specs = [SingleTaskGPSpec(kernel=o["kernel"], y=o["y"], in_offset=problem["in_offset"], in_scale=problem["in_scale"],
                              mean_const=o["mean_const"], noise=o["noise"]) for o in problem["outputs"]]
    return DeviceGPState(problem["X"], specs, device=device).factorize()


def build_acqf(problem, state, prune_baseline=True, **kw):
    from . import acquisition as A

    if problem["acqf"] == "qnehvi":
        return A.qNoisyExpectedHypervolumeImprovement(state, problem["ref_point"], problem["X"], problem["objective"],
                                                      prune_baseline=prune_baseline, mc_samples=problem["S"],
                                                      seed=problem["sampler_seed"], **kw)
    if problem["acqf"] == "qlogei":
        return A.get_acquisition_function("qLogEI", state, problem["objective"], problem["X"], mc_samples=problem["S"],
                                          seed=problem["sampler_seed"])
    raise ValueError(problem["acqf"])


def candidates(problem, n=None):
    """raw-sample q-batches [n, q, d] (Sobol, seed fixed) or the discrete choice set."""
    import torch

    from .optim import draw_sobol_samples

    n = n or problem["raw_samples"]
    if problem.get("choices") is not None:
        return torch.as_tensor(problem["choices"][:n], dtype=torch.double).unsqueeze(1)
    return draw_sobol_samples(torch.as_tensor(problem["bounds"]), n, problem["q"], seed=problem["cand_seed"])

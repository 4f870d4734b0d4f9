"""Marginal log likelihood + hyper-parameter gradient on the device (bo_mll_forward_backward, csrc/mll.cu) against torch
autograd through the oracle's restatement of ExactMarginalLogLikelihood, and the fit built on it (everest_b200/fit.py <->
SingleTaskGPSurrogate._fit, surrogates/single_task_gp.py:39-71).  Tolerances: value 1e-10 relative, gradients 1e-7 relative
to the largest entry (the device gradient goes through the explicit inverse, autograd through Cholesky solves)."""
import math

import numpy as np
import pytest
import torch

from everest_b200 import fit as F
from everest_b200 import kernels as K
from everest_b200.model import SingleTaskGPSpec, standardize_stats
from oracle import bo_oracle as O

pytestmark = pytest.mark.gpu
DT = torch.float64


def synth(N=90, d=4, seed=0):
    rng = np.random.default_rng(seed)
    X = rng.random((N, d))
    y = np.sin(3 * X[:, 0]) + 0.5 * X[:, 1] ** 2 - X[:, 2] + 0.05 * rng.normal(size=N)
    return X, y


def oracle_mll_and_grads(X, y, build, params):
    """build(params as tensors) -> (oracle kernel, noise, mean); returns mll and d mll / d params."""
    ts = {k: torch.tensor(v, dtype=DT, requires_grad=True) for k, v in params.items()}
    kern, noise, mean = build(ts)
    ym, ys = standardize_stats(y)
    out = O.GPOutput(kernel=kern, in_offset=torch.zeros(X.shape[1], dtype=DT), in_scale=torch.ones(X.shape[1], dtype=DT),
                     mean_const=mean, noise=noise, y=torch.as_tensor(y, dtype=DT), y_mean=ym, y_std=ys)
    mll = O.log_marginal_likelihood(X, out)
    mll.backward()
    return float(mll), {k: t.grad.clone() for k, t in ts.items()}


@pytest.mark.parametrize("kind", ["rbf_ard", "matern52_scale", "matern32_iso", "matern12_scale"])
def test_mll_value_and_gradient_single_leaf(kind):
    X, y = synth()
    d = X.shape[1]
    ls = [0.7, 1.3, 0.9, 2.1]
    if kind == "rbf_ard":
        spec_k = K.RBFKernel(list(range(d)), ls)
        build = lambda t: (O.RBF(list(range(d)), t["ls"]), t["noise"], t["mean"])                       # noqa: E731
        params = dict(ls=ls, noise=0.03, mean=0.2)
    elif kind == "matern52_scale":
        spec_k = K.ScaleKernel(K.MaternKernel(list(range(d)), ls, nu=2.5), 1.7)
        build = lambda t: (O.Scale(O.Matern(2.5, list(range(d)), t["ls"]), t["os"]), t["noise"], t["mean"])  # noqa: E731
        params = dict(ls=ls, os=1.7, noise=0.03, mean=0.2)
    elif kind == "matern32_iso":
        spec_k = K.MaternKernel(list(range(d)), [0.8], nu=1.5)
        build = lambda t: (O.Matern(1.5, list(range(d)), t["ls"].expand(d)), t["noise"], t["mean"])      # noqa: E731
        params = dict(ls=[0.8], noise=0.03, mean=0.2)
    else:
        spec_k = K.ScaleKernel(K.MaternKernel(list(range(d)), ls, nu=0.5), 0.6)
        build = lambda t: (O.Scale(O.Matern(0.5, list(range(d)), t["ls"]), t["os"]), t["noise"], t["mean"])  # noqa: E731
        params = dict(ls=ls, os=0.6, noise=0.03, mean=0.2)
    mll_o, g_o = oracle_mll_and_grads(X, y, build, params)
    spec = SingleTaskGPSpec(kernel=spec_k, y=y, mean_const=0.2, noise=0.03)
    mll_d, dn, dm, dls, dco = F.mll_and_grad(X, spec)
    assert abs(mll_d - mll_o) < 1e-10 * abs(mll_o)
    assert abs(dn - float(g_o["noise"])) < 1e-7 * abs(float(g_o["noise"]))
    assert abs(dm - float(g_o["mean"])) < 1e-7 * max(abs(float(g_o["mean"])), 1e-3)
    g_ls = g_o["ls"].numpy()
    if kind == "matern32_iso":
        assert abs(dls.sum() - g_ls[0]) < 1e-7 * abs(g_ls[0])
    else:
        assert np.abs(dls - g_ls).max() < 1e-7 * np.abs(g_ls).max()
    if "os" in params:
        assert abs(dco[0] - float(g_o["os"])) < 1e-7 * abs(float(g_o["os"]))


def test_mll_gradient_composite_tree_with_hamming():
    """(s1 Kc + s2 Kh) + s3 (Kc' * Kh): outputscale chain through the flattened term coefficients, lengthscales of two
    continuous leaves and the Hamming leaf (MixedSingleTaskGP-style tree, mixed_single_task_gp.py:90-108)."""
    rng = np.random.default_rng(1)
    N = 70
    Xc = rng.random((N, 2))
    cat = np.eye(3)[rng.integers(0, 3, N)]
    X = np.concatenate([Xc, cat], axis=1)
    y = np.sin(4 * Xc[:, 0]) + Xc[:, 1] + 0.3 * cat[:, 1] + 0.05 * rng.normal(size=N)
    p = dict(l1=[0.5, 0.9], lh=[1.4], l2=[0.7, 1.1], s1=0.8, s2=0.4, s3=1.3, noise=0.02, mean=-0.1)

    def build(t):
        kc1, kh1 = O.Matern(2.5, [0, 1], t["l1"]), O.Hamming([(2, 3)], t["lh"])
        kc2, kh2 = O.RBF([0, 1], t["l2"]), O.Hamming([(2, 3)], t["lh"])
        return O.Add([O.Add([O.Scale(kc1, t["s1"]), O.Scale(kh1, t["s2"])]), O.Scale(O.Mul([kc2, kh2]), t["s3"])]), t["noise"], t["mean"]

    mll_o, g_o = oracle_mll_and_grads(X, y, build, p)
    kern = K.AdditiveKernel([
        K.AdditiveKernel([K.ScaleKernel(K.MaternKernel([0, 1], p["l1"], nu=2.5), p["s1"]),
                          K.ScaleKernel(K.HammingDistanceKernel({2: 3}, p["lh"]), p["s2"])]),
        K.ScaleKernel(K.MultiplicativeKernel([K.RBFKernel([0, 1], p["l2"]), K.HammingDistanceKernel({2: 3}, p["lh"])]), p["s3"])])
    spec = SingleTaskGPSpec(kernel=kern, y=y, mean_const=p["mean"], noise=p["noise"])
    mll_d, dn, dm, dls, dco = F.mll_and_grad(X, spec)
    assert abs(mll_d - mll_o) < 1e-10 * abs(mll_o)
    # slots: leaf0 Matern (2), leaf1 Hamming (1), leaf2 RBF (2), leaf3 Hamming (1); the two Hamming leaves share lh
    assert np.abs(dls[0:2] - g_o["l1"].numpy()).max() < 1e-7 * np.abs(g_o["l1"].numpy()).max()
    assert np.abs(dls[3:5] - g_o["l2"].numpy()).max() < 1e-7 * np.abs(g_o["l2"].numpy()).max()
    assert abs((dls[2] + dls[5]) - float(g_o["lh"])) < 1e-7 * abs(float(g_o["lh"]))
    # flattened terms: s1 Kc1, s2 Kh1, s3 Kc2 Kh2  ->  d / d s_k = d_coef_k
    for k, name in enumerate(("s1", "s2", "s3")):
        assert abs(dco[k] - float(g_o[name])) < 1e-7 * abs(float(g_o[name]))
    assert abs(dn - float(g_o["noise"])) < 1e-7 * abs(float(g_o["noise"]))


def test_fit_recovers_structure_and_lowers_the_loss():
    X, y = synth(N=120, d=4, seed=3)
    d = X.shape[1]
    kern = K.RBFKernel(list(range(d)), [1.0] * d)
    lp = {0: F.DimensionalityScaledLogNormalPrior(d)}
    res = F.fit_gp(X, y, kern, noise_prior=F.HVARFNER_NOISE_PRIOR(), lengthscale_priors=lp, options={"maxiter": 150})
    assert res.n_iterations >= 3 and math.isfinite(res.loss)
    ls = np.array(res.spec.kernel.lengthscale)
    assert ls.shape == (d,) and np.all(ls > 0)
    # x3 does not enter y: its lengthscale ends up the longest; the noise level is found (sd 0.05 on a unit-ish signal)
    assert ls.argmax() == 3
    assert 1e-4 <= res.spec.noise < 0.2
    # the loss at the optimum is below the loss at the starting point (prior medians)
    start = F.fit_gp(X, y, kern, noise_prior=F.HVARFNER_NOISE_PRIOR(), lengthscale_priors=lp, options={"maxiter": 0})
    assert res.loss < start.loss
    # L-BFGS-B stopped on its gradient criterion: the host chain rule (softplus transforms + priors) is consistent with
    # the device gradient
    assert "CONVERGENCE" in res.message
    # the fitted model explains held-in data: posterior mean close to y at the training points
    from everest_b200.model import DeviceGPState
    st = DeviceGPState(X, [res.spec]).factorize()
    mean, _ = st.posterior(X)
    assert float(np.abs(mean.cpu().numpy()[:, 0] - y).max()) < 0.25


def test_strategy_with_fitted_surrogates():
    from everest_b200 import configs as Cf
    from everest_b200.strategy import InputSpace, QnehviStrategy

    p = Cf.zdt1_qnehvi(N=40, S=32, raw=64, d=3, q=1)
    lo, hi = p["bounds"]
    factory = F.single_task_gp_factory(lambda d: K.RBFKernel(list(range(d)), [1.0] * d), in_offset=lo, in_scale=hi - lo,
                                       options={"maxiter": 60})
    strat = QnehviStrategy(InputSpace(bounds=p["bounds"]), factory, objective=p["objective"], ref_point=[1.0, 5.0],
                           n_mc_samples=32, num_restarts=2, num_raw_samples=32, maxiter=30, seed=0)
    strat.tell(p["X"], p["Y"])
    assert all(len(o.kernel.lengthscale) == 3 for o in strat.model.outputs)
    cand, preds, stds = strat.ask(1)
    assert cand.shape == (1, 3) and np.all(np.isfinite(preds)) and np.all(stds > 0)


def test_mll_gradient_large_n_chunk_loop():
    """N > 2048: the partner-chunk loop of mll_grad_kernel."""
    rng = np.random.default_rng(5)
    N, d = 2100, 2
    X = rng.random((N, d))
    y = np.sin(5 * X[:, 0]) + X[:, 1] + 0.1 * rng.normal(size=N)
    ls = [0.3, 0.8]
    build = lambda t: (O.Scale(O.Matern(2.5, [0, 1], t["ls"]), t["os"]), t["noise"], t["mean"])   # noqa: E731
    mll_o, g_o = oracle_mll_and_grads(X, y, build, dict(ls=ls, os=1.2, noise=0.05, mean=0.0))
    spec = SingleTaskGPSpec(kernel=K.ScaleKernel(K.MaternKernel([0, 1], ls, nu=2.5), 1.2), y=y, mean_const=0.0, noise=0.05)
    mll_d, dn, dm, dls, dco = F.mll_and_grad(X, spec)
    assert abs(mll_d - mll_o) < 1e-9 * abs(mll_o)
    assert np.abs(dls - g_o["ls"].numpy()).max() < 1e-6 * np.abs(g_o["ls"].numpy()).max()
    assert abs(dco[0] - float(g_o["os"])) < 1e-6 * abs(float(g_o["os"]))
    assert abs(dn - float(g_o["noise"])) < 1e-6 * abs(float(g_o["noise"]))


def test_in_place_hyperparameter_update_equals_a_fresh_state():
    """bo_state_set_hyperparameters: one handle serves every evaluation of a fit.  After rewriting lengthscales, outputscale,
    noise and constant mean in place, the marginal likelihood, its gradient and the posterior must be IDENTICAL to those of
    a state created from scratch with the same values (same kernels, same buffers: bit for bit); a changed tree is refused."""
    from everest_b200.model import DeviceGPState

    X, y = synth(N=120, d=4, seed=3)
    d = X.shape[1]

    def spec(ls, os_, noise, mean, hls):
        kern = K.AdditiveKernel([K.ScaleKernel(K.MaternKernel([0, 1, 2], ls, nu=2.5), os_), K.RBFKernel([3], [hls])])
        return SingleTaskGPSpec(kernel=kern, y=y, mean_const=mean, noise=noise)

    a = spec([0.7, 1.1, 0.9], 1.3, 0.02, 0.1, 0.8)
    b = spec([0.4, 2.0, 1.5], 0.6, 0.005, -0.3, 1.7)
    st = DeviceGPState(X, [a]).factorize()
    Xq = np.random.default_rng(1).random((17, d))
    mean_a, var_a = st.posterior(Xq)
    in_place = F.mll_and_grad(X, b, state=st)
    mean_b, var_b = st.posterior(Xq)
    fresh_state = DeviceGPState(X, [b]).factorize()
    fresh = F.mll_and_grad(X, b)
    mean_f, var_f = fresh_state.posterior(Xq)
    assert in_place[0] == fresh[0] and in_place[1] == fresh[1] and in_place[2] == fresh[2]
    assert np.array_equal(in_place[3], fresh[3]) and np.array_equal(in_place[4], fresh[4])
    assert torch.equal(mean_b, mean_f) and torch.equal(var_b, var_f)
    assert not torch.equal(mean_a, mean_b)
    # and back again: nothing of the previous values lingers
    back = F.mll_and_grad(X, a, state=st)
    ref = F.mll_and_grad(X, a)
    assert back[0] == ref[0] and np.array_equal(back[3], ref[3])
    with pytest.raises(ValueError):
        st.set_hyperparameters(0, SingleTaskGPSpec(kernel=K.RBFKernel([0, 1, 2, 3], [1.0] * 4), y=y))
    st.close(); fresh_state.close()

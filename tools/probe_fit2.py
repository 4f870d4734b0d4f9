"""Hyper-parameter fit time at BASELINE config 3 size (N = 2000, d = 30, RBF-ARD, Hvarfner priors) and at N = 500."""
import sys, time, numpy as np, torch
sys.path.insert(0, '.')
from everest_b200 import fit as F, kernels as K, configs as Cf
for N in (500, 2000):
    p = Cf.zdt1_qnehvi(N=N)
    X, y = p["X"], p["Y"][:, 1]
    kern = K.RBFKernel(list(range(30)), [1.0] * 30)
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        res = F.fit_gp(X, y, kern, noise_prior=F.HVARFNER_NOISE_PRIOR(), lengthscale_priors={0: F.DimensionalityScaledLogNormalPrior(30)},
                       options={"maxiter": 200})
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print(f"N={N} rep {rep}: fit {dt:.2f} s, {res.n_evaluations} MLL evaluations ({1e3 * dt / res.n_evaluations:.1f} ms each), {res.n_iterations} iterations, "
              f"noise {res.spec.noise:.3g}, mll {res.mll:.4g}, {res.message}")

"""Which pinned global-generator seeds make the device / scipy optimiser comparison of tests/test_gpu_optimizer.py land in the
same local maxima (initialize_q_batch draws from torch's global generator, whose default seed is random per process)."""
import sys, torch
sys.path.insert(0, '.')
from everest_b200 import configs as Cf, optim
p = Cf.himmelblau_qlogei(N=200, S=128, raw=512)
st = Cf.build_state(p)
acq = Cf.build_acqf(p, st)
bnds = torch.as_tensor(p["bounds"])
for seed in range(8):
    torch.manual_seed(seed)
    Xic, Yic, _, _ = optim.gen_batch_initial_conditions(acq, bnds, p["q"], 8, 512, seed=0)
    Xd, Yd, info_d = optim.gen_candidates_device(Xic, acq, bnds[0], bnds[1], options={"maxiter": 200})
    Xs, Ys, info_s = optim.gen_candidates_scipy(Xic, acq, bnds[0], bnds[1], options={"maxiter": 200})
    ok = torch.allclose(Yd, Ys, rtol=1e-7, atol=1e-8) and float((Xd - Xs).abs().max()) < 1e-4 * 12.0
    print(seed, "ok" if ok else "DIFF", "max |dY|", float((Yd - Ys).abs().max()), "max |dX|", float((Xd - Xs).abs().max()), "conv", info_d["n_converged"])

# the pending-points comparison of the same file
from everest_b200 import acquisition as A
p = Cf.himmelblau_qlogei(N=120, S=64, raw=256)
st = Cf.build_state(p)
pend = torch.tensor([[1.0, 2.0], [-3.0, 0.5]], dtype=torch.double)
mean, _ = st.posterior(torch.as_tensor(p["X"]))
best_f = float(p["objective"](mean.cpu()).max())
acq = A.qLogExpectedImprovement(st, best_f, p["objective"], mc_samples=64, seed=5, X_pending=pend)
bnds = torch.as_tensor(p["bounds"])
for seed in range(8):
    torch.manual_seed(seed)
    Xic, Yic, _, _ = optim.gen_batch_initial_conditions(acq, bnds, 1, 5, 256, seed=1)
    Xd, Yd, info = optim.gen_candidates_device(Xic, acq, bnds[0], bnds[1], options={"maxiter": 100})
    Xs, Ys, _ = optim.gen_candidates_scipy(Xic, acq, bnds[0], bnds[1], options={"maxiter": 100})
    ok = torch.allclose(Yd, Ys, rtol=1e-6, atol=1e-7) and float((Xd - Xs).abs().max()) < 2e-3
    print("pending", seed, "ok" if ok else "DIFF", "max |dY|", float((Yd - Ys).abs().max()), "max |dX|", float((Xd - Xs).abs().max()))

import sys, time, cProfile, pstats, torch
sys.path.insert(0, '.')
from everest_b200 import configs as Cf, optim
p = Cf.zdt1_qnehvi()
st = Cf.build_state(p); acq = Cf.build_acqf(p, st)
bnds = torch.as_tensor(p["bounds"])
Xic, Yic, _, _ = optim.gen_batch_initial_conditions(acq, bnds, p["q"], 8, 2048, seed=0)
pr = cProfile.Profile(); pr.enable()
t0 = time.perf_counter()
_, Y, info = optim.gen_candidates_scipy(Xic, acq, bnds[0], bnds[1], options={"maxiter": 20})
t1 = time.perf_counter()
pr.disable()
print(info, f"{t1-t0:.3f}s")
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)

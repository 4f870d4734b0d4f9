"""Timing probe: acquisition set-up (base samples, pruning, cached roots, box decompositions) on BASELINE config 3."""
import sys, time, torch
sys.path.insert(0, '.')
from everest_b200 import configs as Cf, sampling
p = Cf.zdt1_qnehvi()
st = Cf.build_state(p)
def T():
    torch.cuda.synchronize(); return time.perf_counter()
for rep in range(4):
    t0 = T(); acq = Cf.build_acqf(p, st); t1 = T()
    print(f"build {rep}: {t1-t0:.3f}s")
    if rep == 1:
        X = Cf.candidates(p).to(st.device)
        for _ in range(3): acq(X)
        acq.forward_host(Cf.candidates(p).numpy())
        print("after big forward calls")
t0 = T(); z = sampling.base_samples_device(2000, 2, 2048, 8153, st.device); t1 = T(); print(f"prune base samples (device): {t1-t0:.3f}s")
t0 = T(); ss, sh, rows = sampling.sobol_scramble_inputs(4000, 8153); t1 = T(); print(f"  host prep: {t1-t0:.3f}s")
t0 = T(); z = sampling.base_samples(2000, 2, 2048, 8153); t1 = T(); print(f"prune base samples (host): {t1-t0:.3f}s")
import cProfile, pstats
pr = cProfile.Profile(); pr.enable(); acq = Cf.build_acqf(p, st); torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(14)

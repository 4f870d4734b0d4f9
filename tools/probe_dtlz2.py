import sys, time, torch
sys.path.insert(0, '.')
from everest_b200 import configs as Cf
for N, S in ((200, 64), (400, 128), (1000, 512)):
    p = Cf.dtlz2_qnehvi(N=N, S=S, raw=256)
    t0 = time.perf_counter(); st = Cf.build_state(p); torch.cuda.synchronize(); t1 = time.perf_counter()
    acq = Cf.build_acqf(p, st); torch.cuda.synchronize(); t2 = time.perf_counter()
    X = Cf.candidates(p).to(st.device)
    v = acq(X); torch.cuda.synchronize(); t3 = time.perf_counter()
    v = acq(X); torch.cuda.synchronize(); t4 = time.perf_counter()
    print(f"N={N} S={S} factorize {t1-t0:.3f}s prepare {t2-t1:.3f}s nb={acq.nb} max_cells={acq.max_cells} fwd1 {t3-t2:.3f}s fwd2 {t4-t3:.3f}s vmax={float(v.max()):.4g}", flush=True)

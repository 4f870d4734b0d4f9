import sys, time, torch
sys.path.insert(0, '.')
from everest_b200 import configs as Cf
p = Cf.zdt1_qnehvi()
st = Cf.build_state(p); acq = Cf.build_acqf(p, st)
for b in (8, 64, 512, 1928, 4096):
    X = Cf.candidates(p, b)
    Xd = X.to(st.device)
    for _ in range(3): acq(Xd)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10): acq(Xd)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    for _ in range(10): v = acq(X)   # CPU in -> CPU out
    t2 = time.perf_counter()
    st.set_timing(True); acq(Xd); torch.cuda.synchronize()
    parts = {n: round(st.last_timing(n)[0], 3) for n in ("prep", "crosscov", "posterior_gemm", "cond_root", "sample_gemm", "mc_acqf")}
    st.set_timing(False)
    print(f"b={b}: device-in {1e3*(t1-t0)/10:.2f} ms, cpu-in/out {1e3*(t2-t1)/10:.2f} ms, kernels {parts}", flush=True)

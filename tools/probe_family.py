"""Timing probe: raw-sample screen (16384 q-batches) of the other acquisition functions on BASELINE config 3 shapes."""
import sys, time, torch
sys.path.insert(0, '.')
from everest_b200 import configs as Cf, acquisition as A
from everest_b200.objectives import ScalarObjective, MinimizeObjective
p = Cf.zdt1_qnehvi()
st = Cf.build_state(p)
X = Cf.candidates(p).to(st.device)
def timeit(acq, name, n=3):
    acq(X[:256]); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): v = acq(X)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / n
    st.set_timing(True); acq(X); torch.cuda.synchronize()
    mc = st.last_timing("mc_acqf")[0]; st.set_timing(False)
    print(f"{name:12s} {dt*1e3:8.2f} ms/screen  {X.shape[0]/dt/1e3:8.1f} k evals/s   mc kernel {mc:8.2f} ms   finite {bool(torch.isfinite(v).all())}")
for name in ["qNEHVI", "qLogNEHVI"]:
    acq = A.get_acquisition_function(name, st, p["objective"], p["X"], ref_point=p["ref_point"], mc_samples=p["S"], seed=1234)
    timeit(acq, name)
obj = ScalarObjective([MinimizeObjective(1)], "single")
for name in ["qLogNEI", "qLogEI", "qEI"]:
    acq = A.get_acquisition_function(name, st, obj, p["X"], mc_samples=p["S"], seed=1234)
    timeit(acq, name + f"(nb={acq.nb})")

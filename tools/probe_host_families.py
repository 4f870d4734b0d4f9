"""Per-family kernel times of a config-3 screen through the host entry point (piece mode) next to the device-resident call."""
import sys, os, time, torch
sys.path.insert(0, '.')
from everest_b200 import configs as Cf
p = Cf.zdt1_qnehvi()
st = Cf.build_state(p); acq = Cf.build_acqf(p, st)
Xh = Cf.candidates(p).contiguous().numpy()
X = torch.as_tensor(Xh).to(st.device)
fam = ["prep", "crosscov", "ozaki_slice", "posterior_gemm", "ozaki_guard", "cond_root", "sample_gemm", "mc_acqf"]
for _ in range(3): acq.forward_host(Xh); acq(X)
torch.cuda.synchronize()
for name, fn in (("host", lambda: acq.forward_host(Xh)), ("device", lambda: acq(X))):
    ts = []
    for _ in range(5):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    st.set_timing(True); fn(); torch.cuda.synchronize()
    tm = {k: round(st.last_timing(k)[0], 3) for k in fam}
    st.set_timing(False)
    print(name, "wall ms", [round(t, 2) for t in ts], "families", tm, "sum", round(sum(tm.values()), 3))

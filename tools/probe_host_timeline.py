"""Timeline of bo_acqf_forward_host on a config-3 screen (EVEREST_HOST_TRACE) next to the device-resident call and the two chunk sizes."""
import sys, os, time, torch
sys.path.insert(0, '.')
os.environ["EVEREST_HOST_TRACE"] = "1"
from everest_b200 import configs as Cf
p = Cf.zdt1_qnehvi()
st = Cf.build_state(p); acq = Cf.build_acqf(p, st)
Xh = Cf.candidates(p).contiguous().numpy()
X = torch.as_tensor(Xh).to(st.device)
for _ in range(3): acq.forward_host(Xh)
torch.cuda.synchronize()
t0 = time.perf_counter(); acq.forward_host(Xh); t1 = time.perf_counter()
print("e2e call", (t1 - t0) * 1e3, "ms", file=sys.stderr)
for _ in range(2): acq(X)
torch.cuda.synchronize(); t0 = time.perf_counter(); acq(X); torch.cuda.synchronize(); print("device call", (time.perf_counter() - t0) * 1e3, "ms", file=sys.stderr)
st.set_timing(True)
acq(X[:2048]); torch.cuda.synchronize()
print("2048 q-batch chunk:", {k: round(st.last_timing(k)[0], 3) for k in ["crosscov", "posterior_gemm", "cond_root", "sample_gemm", "mc_acqf", "ozaki_guard"]}, file=sys.stderr)
t0 = time.perf_counter(); acq(X[:2048]); torch.cuda.synchronize(); print("2048 q-batches alone", (time.perf_counter() - t0) * 1e3, "ms", file=sys.stderr)
t0 = time.perf_counter(); acq(X[2048:]); torch.cuda.synchronize(); print("14336 q-batches alone", (time.perf_counter() - t0) * 1e3, "ms", file=sys.stderr)

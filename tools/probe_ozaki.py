"""Validation / timing probe of the INT8 digit-plane posterior GEMM (csrc/ozaki.cu) against the FP64 DMMA kernel."""
import sys, time, torch
sys.path.insert(0, '.')
from everest_b200 import configs as Cf
size = sys.argv[1] if len(sys.argv) > 1 else "small"
p = Cf.zdt1_qnehvi(N=300, S=32, raw=700, d=6, q=4) if size == "small" else Cf.zdt1_qnehvi()
st = Cf.build_state(p)
acq = Cf.build_acqf(p, st, prune_samples=256 if size == "small" else 2048)
X = Cf.candidates(p).to(st.device)
b, q, M = X.shape[0], p["q"], st.M
def run(oz):
    acq.set_option("ozaki", 2 * oz)
    v = acq(X).clone()
    torch.cuda.synchronize()
    nr = acq.nb + q
    root = st.debug_get("root", capacity=b * M * q * nr + 16).view(-1)[: b * M * q * nr].clone()
    mu = st.debug_get("mu", capacity=b * q * M + 16).view(-1)[: b * q * M].clone()
    return v, root, mu
v0, r0, m0 = run(0)
v1, r1, m1 = run(1)
print("values   max abs diff", float((v1 - v0).abs().max()), "scale", float(v0.abs().max()), "finite", bool(torch.isfinite(v1).all()))
print("roots    max abs diff", float((r1 - r0).abs().max()), "scale", float(r0.abs().max()))
print("means    max abs diff", float((m1 - m0).abs().max()), "scale", float(m0.abs().max()))
for oz in (0, 1):
    acq.set_option("ozaki", 2 * oz)
    for _ in range(2): acq(X)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): acq(X)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
    st.set_timing(True); acq(X); torch.cuda.synchronize()
    print(f"ozaki={oz}: {dt*1e3:.2f} ms/screen, {b/dt/1e3:.1f} k evals/s; gemm {st.last_timing('posterior_gemm')[0]:.2f} ms, slice {st.last_timing('ozaki_slice')[0]:.2f} ms, crosscov {st.last_timing('crosscov')[0]:.2f} ms")
    st.set_timing(False)

"""Does FP64 ALU work co-run with the persistent INT8 GEMM of a config-3 screen?  Stream A: the screen (crosscov 1.5 ms,
ozaki GEMM 5.8 ms with one 320-thread CTA per SM holding 54 K registers and 178 KB of shared memory, MC ...).  Stream B: a
small-footprint FP64-heavy kernel (torch.lgamma on float64: 128-thread CTAs, a few K registers, no shared memory -- it CAN be
co-resident with a GEMM CTA).  If the pipes overlap, A || B takes about max(A, B); if not, about A + B."""
import sys, time, torch
sys.path.insert(0, '.')
from everest_b200 import configs as Cf
p = Cf.zdt1_qnehvi()
st = Cf.build_state(p)
acq = Cf.build_acqf(p, st)
X = Cf.candidates(p).to(st.device)
s2 = torch.cuda.Stream()
def A(n=1):
    for _ in range(n): acq(X)
def wall(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3
tA = wall(lambda: A())
print(f"A alone (one screen): {tA:.2f} ms")
REP = 150
for n_el in (1 << 22, 1 << 23, 1 << 24):
    x = torch.rand(n_el, dtype=torch.double, device=st.device) + 1.5
    y = torch.empty_like(x)
    def B():
        with torch.cuda.stream(s2):
            for _ in range(REP): torch.sqrt(x, out=y)
    def AB():
        s2.wait_stream(torch.cuda.current_stream())
        A(); B()
        torch.cuda.current_stream().wait_stream(s2)
    def BA():
        s2.wait_stream(torch.cuda.current_stream())
        B(); A()
        torch.cuda.current_stream().wait_stream(s2)
    tB = wall(B)
    print(f"n = 2^{n_el.bit_length()-1}: B alone {tB:.2f} ms | A then B launched: {wall(AB):.2f} ms | B then A launched: {wall(BA):.2f} ms | A + B = {tA + tB:.2f}, max = {max(tA, tB):.2f}")
st.set_timing(True)
x = torch.rand(1 << 24, dtype=torch.double, device=st.device) + 1.5; y = torch.empty_like(x)
with torch.cuda.stream(s2):
    for _ in range(150): torch.sqrt(x, out=y)
acq(X); torch.cuda.synchronize()
print("GEMM time with the co-runner in flight:", st.last_timing("posterior_gemm")[0], "ms; crosscov", st.last_timing("crosscov")[0])
acq(X); torch.cuda.synchronize()
print("GEMM time alone:", st.last_timing("posterior_gemm")[0], "ms; crosscov", st.last_timing("crosscov")[0])

"""Timing probe: analytic forward_backward (per-kernel-family breakdown) vs batched finite differences.
usage: python tools/probe_grad.py [zdt1|dtlz2] [--no-opt]"""
import sys, time, torch
sys.path.insert(0, '.')
from everest_b200 import configs as Cf, optim
name = next((a for a in sys.argv[1:] if not a.startswith("-")), "zdt1")
p = {"zdt1": Cf.zdt1_qnehvi, "dtlz2": Cf.dtlz2_qnehvi}[name]()
st = Cf.build_state(p); acq = Cf.build_acqf(p, st)
print("nb", acq.nb, "max_cells", acq.max_cells)
bnds = torch.as_tensor(p["bounds"])
Xic, Yic, _, _ = optim.gen_batch_initial_conditions(acq, bnds, p["q"], 8, 2048, seed=0)
X = Xic.to(st.device)
for _ in range(3):
    v, g = acq.forward_backward(X)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20):
    v, g = acq.forward_backward(X)
torch.cuda.synchronize()
print(f"forward_backward b=8: {(time.perf_counter()-t0)/20*1e3:.3f} ms")
st.lib.bo_set_timing(st.handle, 1)
v, g = acq.forward_backward(X)
torch.cuda.synchronize()
import ctypes as C
for name in ["prep", "crosscov", "posterior_gemm", "cond_root", "sample_gemm", "mc_grad", "grad_reduce", "cond_root_bwd", "u_gemm", "kernel_grad"]:
    ms = C.c_double(0)
    n = st.lib.bo_last_timing(st.handle, name.encode(), C.byref(ms))
    print(f"  {name:16s} {ms.value*1e3:9.1f} us  ({n} launches)")
st.lib.bo_set_timing(st.handle, 0)
for mode in (() if "--no-opt" in sys.argv else ("analytic", "fd")):
    t0 = time.perf_counter()
    _, Y, info = optim.gen_candidates_scipy(Xic, acq, bnds[0], bnds[1], options={"maxiter": 50, "gradient": mode})
    t1 = time.perf_counter()
    print(mode, info, f"{t1-t0:.3f}s", "start", float(Yic.max()), "best", float(Y.max()), "sum", float(Y.sum()))

"""Where the host-buffer path (bo_acqf_forward_host) spends its time: forced INT8 without guard (ozaki=2) vs automatic mode
(guard + redo), and the guard with a tolerance so loose that nothing is flagged (sync only)."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from everest_b200 import configs as Cf

p = Cf.zdt1_qnehvi()
st = Cf.build_state(p)
acq = Cf.build_acqf(p, st)
Xh = Cf.candidates(p).contiguous().numpy()
Xd = torch.as_tensor(Xh).cuda()

def t_host(n=10):
    for _ in range(3): acq.forward_host(Xh)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): acq.forward_host(Xh)
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3

def t_dev(n=10):
    for _ in range(3): acq(Xd)
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): acq(Xd)
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n

for name, opts in [("forced INT8, no guard", {"ozaki": 2}), ("auto: guard + redo", {"ozaki": 1}),
                   ("auto: guard, tol 1e-4 (no flags)", {"ozaki": 1, "ozaki_guard_tol": 1e-4})]:
    for k, v in opts.items(): acq.set_option(k, v)
    print(f"{name:36s} device {t_dev():.3f} ms   host-buffer {t_host():.3f} ms   flagged {int(st.debug_get('ozaki_check', capacity=16)[1])}")
    acq.set_option("ozaki_guard_tol", 1e-10)

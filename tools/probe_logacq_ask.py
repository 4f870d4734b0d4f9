"""Timing probe: forward_backward and a whole ask() sequence (screen + initialize_q_batch + device refinement) with the
log-space acquisition function MoboStrategy builds by default (qLogNEHVI) next to qNEHVI, on BASELINE config 3 shapes."""
import sys, time, torch
sys.path.insert(0, '.')
from everest_b200 import configs as Cf, optim, acquisition as A
p = Cf.zdt1_qnehvi()
st = Cf.build_state(p)
bnds = torch.as_tensor(p["bounds"])
for name in ["qNEHVI", "qLogNEHVI"]:
    acq = A.get_acquisition_function(name, st, p["objective"], p["X"], ref_point=p["ref_point"], mc_samples=p["S"], seed=1234)
    Xic, Yic, _, _ = optim.gen_batch_initial_conditions(acq, bnds, p["q"], 8, 2048, seed=0)
    X = Xic.to(st.device)
    for _ in range(3):
        v, g = acq.forward_backward(X)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        v, g = acq.forward_backward(X)
    torch.cuda.synchronize()
    fb = (time.perf_counter() - t0) / 20 * 1e3
    for rep in range(2):
        t0 = time.perf_counter()
        cand, val = optim.optimize_acqf(acq, bnds, q=p["q"], num_restarts=8, raw_samples=p["raw_samples"], options={"maxiter": 200}, seed=0)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    print(f"{name:10s} forward_backward(b=8) {fb:7.3f} ms   ask sequence (2nd call) {dt:6.3f} s   value {float(val):.6g}")

"""Which earlier activity of bench.py makes the first device refinement slow?  usage: probe_ask2.py <steps: comma list of host,int8,fp64,timing,sampler>"""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench as B
from everest_b200 import configs as Cf, optim
steps = sys.argv[1].split(",") if len(sys.argv) > 1 else []
p = Cf.himmelblau_qlogei()
st = Cf.build_state(p); acq = Cf.build_acqf(p, st)
Xh = Cf.candidates(p).contiguous(); X = Xh.to(st.device)
for _ in range(3): acq(X)
dev = st.device
if "sampler" in steps:
    s = B.ClockSampler(0); s.start(); time.sleep(0.5); print("clocks", s.stop())
if "host" in steps:
    for _ in range(3): acq.forward_host(Xh.numpy())
if "timing" in steps:
    st.set_timing(True); acq(X); torch.cuda.synchronize(); st.set_timing(False)
if "fp64" in steps: print("fp64 peak", B.measure_fp64_peak(dev))
if "int8" in steps: print("int8 peak", B.measure_int8_peak(dev))
torch.cuda.synchronize()
bnds = torch.as_tensor(p["bounds"])
acq2 = Cf.build_acqf(p, st)
Xic, Yic, _, _ = optim.gen_batch_initial_conditions(acq2, bnds, p["q"], p["num_restarts"], p["raw_samples"], seed=0)
torch.cuda.synchronize()
for rep in range(2):
    t0 = time.perf_counter(); optim.gen_candidates_device(Xic, acq2, bnds[0], bnds[1], options={"maxiter": 200}); torch.cuda.synchronize()
    print(f"[{','.join(steps)}] device refine call {rep}: {time.perf_counter() - t0:.3f} s")

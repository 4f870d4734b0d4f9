"""Where the first ask() of a process spends its time (config 3 or 2): build, screen, device refinement (first and second call)."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from everest_b200 import configs as Cf, optim
name = sys.argv[1] if len(sys.argv) > 1 else "zdt1"
p = {"zdt1": Cf.zdt1_qnehvi, "himmelblau": Cf.himmelblau_qlogei}[name]()
st = Cf.build_state(p)
acq = Cf.build_acqf(p, st)
X = Cf.candidates(p).to(st.device)
for _ in range(3): acq(X)
big = [torch.randn(8192, 8192, dtype=torch.double, device=st.device) for _ in range(2)]; c = big[0] @ big[1]; del big, c   # like bench.py's peak probes
torch.cuda.synchronize()
def T():
    torch.cuda.synchronize(); return time.perf_counter()
bnds = torch.as_tensor(p["bounds"])
t0 = T(); acq2 = Cf.build_acqf(p, st); t1 = T()
Xic, Yic, _, _ = optim.gen_batch_initial_conditions(acq2, bnds, p["q"], p["num_restarts"], p["raw_samples"], seed=0); t2 = T()
print(f"build {t1-t0:.3f} screen {t2-t1:.3f}")
for rep in range(3):
    t0 = T(); _, Y, info = optim.gen_candidates_device(Xic, acq2, bnds[0], bnds[1], options={"maxiter": 200}); t1 = T()
    print(f"device refine call {rep}: {t1-t0:.3f} s, {info['n_steps']} steps")
t0 = T(); optim.gen_candidates_scipy(Xic, acq2, bnds[0], bnds[1], options={"maxiter": 200}); t1 = T()
print(f"scipy refine: {t1-t0:.3f} s")

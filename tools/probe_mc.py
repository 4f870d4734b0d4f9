"""MC/HVI tiled kernel experiment switches (EVEREST_MC_STAGE / EVEREST_MC_PREFETCH): config-3 screen values saved for a
bitwise comparison across processes, kernel time from the state's per-stage CUDA events."""
import os, sys, torch
sys.path.insert(0, '.')
from everest_b200 import configs as Cf
tag = sys.argv[1]
dtlz2 = len(sys.argv) > 2 and sys.argv[2] == "dtlz2"   # config 4 (many-cell kernel) instead of config 3 (tiled kernel)
p = Cf.dtlz2_qnehvi() if dtlz2 else Cf.zdt1_qnehvi()
st = Cf.build_state(p)
acq = Cf.build_acqf(p, st) if dtlz2 else Cf.build_acqf(p, st, prune_samples=2048)
X = Cf.candidates(p).to(st.device)
for _ in range(3): v = acq(X)
torch.cuda.synchronize()
st.set_timing(True)
ts = []
for _ in range(5):
    v = acq(X); torch.cuda.synchronize()
    ts.append(st.last_timing('mc_acqf')[0])
st.set_timing(False)
torch.save(v.cpu(), f"gpurun_out/mc_{tag}.pt")
print(tag, "stage", os.environ.get("EVEREST_MC_STAGE"), "prefetch", os.environ.get("EVEREST_MC_PREFETCH"),
      "mc_acqf ms", " ".join(f"{t:.3f}" for t in ts))
if tag != "base":
    v0 = torch.load("gpurun_out/mc_base.pt")
    print(tag, "bit-identical to base:", bool(torch.equal(v0, v.cpu())))

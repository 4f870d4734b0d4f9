"""End-to-end (host buffers) time of a config-3 screen through bo_acqf_forward_host for the chunk plan selected by
EVEREST_HOST_FIRST_DIV / EVEREST_HOST_THREE; values compared bitwise across plans through a saved file."""
import os, sys, time, torch
sys.path.insert(0, '.')
from everest_b200 import configs as Cf
tag = sys.argv[1]
p = Cf.zdt1_qnehvi()
st = Cf.build_state(p)
acq = Cf.build_acqf(p, st, prune_samples=2048)
Xh = Cf.candidates(p).cpu().numpy()
for _ in range(3): v = acq.forward_host(Xh)
ts = []
for _ in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10): v = acq.forward_host(Xh)
    ts.append((time.perf_counter() - t0) / 10 * 1e3)
print(tag, "first_div", os.environ.get("EVEREST_HOST_FIRST_DIV"), "three", os.environ.get("EVEREST_HOST_THREE"),
      "e2e ms/screen", " ".join(f"{t:.3f}" for t in ts))
v = torch.from_numpy(v)
if tag == "d8": torch.save(v, "gpurun_out/host_d8.pt")
else: print(tag, "bit-identical to d8:", bool(torch.equal(torch.load("gpurun_out/host_d8.pt"), v)))

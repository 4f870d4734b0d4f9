"""Timeline of the host-buffer path on BASELINE config 5 (run with EVEREST_HOST_TRACE=1): float64 rows packed inside the call
vs pre-packed buffers vs the float64 wire format (EVEREST_HOST_PACK=0 in a second process)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from everest_b200 import configs as Cf
p = Cf.mixed_tanimoto_qlogei()
st = Cf.build_state(p); acq = Cf.build_acqf(p, st)
Xh = Cf.candidates(p).contiguous().numpy()
Xd = torch.as_tensor(Xh).cuda()
for _ in range(3): acq(Xd)
torch.cuda.synchronize()
t0 = time.perf_counter(); dense, bits = acq.pack_rows(Xh); t1 = time.perf_counter()
print(f"bo_pack_rows_host (8 threads): {1e3 * (t1 - t0):.2f} ms for {Xh.nbytes / 1e6:.0f} MB")
for name, fn in (("float64 in (packs inside)" if os.environ.get("EVEREST_HOST_PACK", "1") != "0" else "float64 wire format", lambda: acq.forward_host(Xh)),
                 ("pre-packed", lambda: acq.forward_host_packed(dense, bits))):
    for _ in range(2): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): fn()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
    print(f"{name:32s} {1e3 * dt:.2f} ms per screen = {Xh.shape[0] / dt / 1e6:.3f} M evals/s")
    if os.environ.get("EVEREST_HOST_TRACE"):
        print("--- trace of one call ---", file=sys.stderr); fn()

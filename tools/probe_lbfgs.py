"""On-device multi-start refinement (bo_acqf_optimize) against scipy L-BFGS-B through the same device gradients.
usage: python tools/probe_lbfgs.py [zdt1|himmelblau|dtlz2] [maxiter]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from everest_b200 import configs as Cf, optim  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "zdt1"
maxiter = int(sys.argv[2]) if len(sys.argv) > 2 else 200
p = {"zdt1": Cf.zdt1_qnehvi, "dtlz2": Cf.dtlz2_qnehvi, "himmelblau": Cf.himmelblau_qlogei}[name]()
st = Cf.build_state(p)
acq = Cf.build_acqf(p, st)
bnds = torch.as_tensor(p["bounds"])
Xic, Yic, _, _ = optim.gen_batch_initial_conditions(acq, bnds, p["q"], p["num_restarts"], min(p["raw_samples"], 4096), seed=0)
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    Xd, Yd, info_d = optim.gen_candidates_device(Xic, acq, bnds[0], bnds[1], options={"maxiter": maxiter})
    torch.cuda.synchronize(); t1 = time.perf_counter()
    Xs, Ys, info_s = optim.gen_candidates_scipy(Xic, acq, bnds[0], bnds[1], options={"maxiter": maxiter})
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"[{name}] device {t1 - t0:.4f} s {info_d}")
    print(f"[{name}] scipy  {t2 - t1:.4f} s {info_s}")
print("start  ", Yic.numpy())
print("device ", Yd.numpy())
print("scipy  ", Ys.numpy())
print("max |X_dev - X_scipy| per restart", (Xd - Xs).abs().amax(dim=(1, 2)).numpy())
# tight tolerances: do both reach the same local maxima?
Xd2, Yd2, i2 = optim.gen_candidates_device(Xic, acq, bnds[0], bnds[1], options={"maxiter": 2000, "pgtol": 1e-9, "ftol": 1e-15})
Xs2, Ys2, i3 = optim.gen_candidates_scipy(Xd2, acq, bnds[0], bnds[1], options={"maxiter": 2000})
print("tight device", i2, Yd2.numpy())
print("scipy polished from there", i3, Ys2.numpy(), "moved", (Xd2 - Xs2).abs().amax(dim=(1, 2)).numpy())

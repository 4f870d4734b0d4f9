import os, sys, torch
sys.path.insert(0, "/root/repo")
from everest_b200 import configs as Cf
p = Cf.zdt1_qnehvi(); st = Cf.build_state(p); acq = Cf.build_acqf(p, st)
Xh = Cf.candidates(p).contiguous().numpy()
for mode in (2, 1):
    acq.set_option("ozaki", mode)
    for i in range(4):
        if i == 3: sys.stderr.write(f"--- ozaki={mode}\n")
        os.environ["X"] = "1"
        acq.forward_host(Xh)

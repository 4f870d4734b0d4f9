"""One BASELINE workload, a few forward screens -- the short program ncu wraps (tools/ncu_capture.sh).
usage: python tools/probe_forward.py zdt1|dtlz2|himmelblau|mixed [n_screens]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from everest_b200 import configs as Cf  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "zdt1"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
p = {"zdt1": Cf.zdt1_qnehvi, "dtlz2": Cf.dtlz2_qnehvi, "himmelblau": Cf.himmelblau_qlogei, "mixed": Cf.mixed_tanimoto_qlogei}[name]()
st = Cf.build_state(p)
acq = Cf.build_acqf(p, st)
X = Cf.candidates(p).to(st.device)
for _ in range(n):
    v = acq(X)
torch.cuda.synchronize()
print(name, "ok", float(v.max()), "launches", st.launch_count())

"""Debug probe: device binary partitioning vs oracle restatement on one MC sample."""
import sys, torch
sys.path.insert(0, '.')
from everest_b200 import configs as Cf, acquisition as A
from oracle import bo_oracle as O
from tests import problems as P
p = Cf.dtlz2_qnehvi(N=70, S=16, raw=12, d=5, m_obj=3, q=2)
gp = P.oracle_gp(p); st = Cf.build_state(p)
ops = [P.op_to_oracle(o) for o in p["objective"].ops]
alpha = 0.01
acq_o = O.QNEHVIOracle(gp, p["ref_point"], p["X"], ops, mc_samples=p["S"], seed=p["sampler_seed"], prune_baseline=True,
                       prune_samples=256, prune_seed=p["sampler_seed"] + 7919, alpha=alpha)
acq_d = A.qNoisyExpectedHypervolumeImprovement(st, p["ref_point"], torch.as_tensor(p["X"]), p["objective"], prune_baseline=True,
                                               alpha=alpha, mc_samples=p["S"], seed=p["sampler_seed"], prune_samples=256)
lo, up, nc = acq_d.cell_bounds()
print("counts equal", nc.tolist() == acq_o.n_cells.tolist())
for s in range(3):
    c = int(nc[s])
    a = torch.cat([lo[s, :c], up[s, :c]], dim=1); b = torch.cat([acq_o.cell_lower[s, :c], acq_o.cell_upper[s, :c]], dim=1)
    print("sample", s, "cells", c, "equal in order", torch.equal(a, b))
    sa = sorted(map(tuple, a.tolist())); sb = sorted(map(tuple, b.tolist()))
    print("  equal as sets", sa == sb, "max abs diff sorted", max(abs(x - y) for r1, r2 in zip(sa, sb) for x, y in zip(r1, r2) if x == x and abs(x) != float('inf') and abs(y) != float('inf')))
    if not torch.equal(a, b):
        for i in range(min(c, 6)):
            print("   dev", [round(v, 4) for v in a[i].tolist()], " ora", [round(v, 4) for v in b[i].tolist()])
    front_o = acq_o.obj_b[s][acq_o.fronts[s]]
    print("  oracle front rows", acq_o.fronts[s].tolist())

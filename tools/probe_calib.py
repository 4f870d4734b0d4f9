"""Determinism / accuracy probe of the INT8 posterior GEMM on the bench workloads (run on the GPU box)."""
import sys, types
import torch
sys.path.insert(0, ".")
from everest_b200 import configs as Cf

for wl in (sys.argv[1:] or ["mixed", "zdt1"]):
    p = {"mixed": Cf.mixed_tanimoto_qlogei, "zdt1": Cf.zdt1_qnehvi, "dtlz2": Cf.dtlz2_qnehvi, "himmelblau": Cf.himmelblau_qlogei}[wl]()
    dev = torch.device("cuda", 0)
    st = Cf.build_state(p, device=dev)
    acq = Cf.build_acqf(p, st)
    X = Cf.candidates(p).contiguous().to(dev)
    acq.set_option("ozaki", 0)
    ref = acq(X).clone()
    acq.set_option("ozaki", 2)
    outs = [acq(X).clone() for _ in range(6)]
    sc = ref.abs().max()
    print(wl, "forced INT8 vs FP64:", [float((o - ref).abs().max() / sc) for o in outs])
    print(wl, "bitwise repeatable:", [bool(torch.equal(o, outs[0])) for o in outs])
    acq.set_option("ozaki", 1)
    outs = [acq(X).clone() for _ in range(3)]
    print(wl, "auto vs FP64:", [float((o - ref).abs().max() / sc) for o in outs], st.debug_get("ozaki_check", capacity=16).tolist())
    # half batches (different tile counts)
    h = X.shape[0] // 2 + 37
    acq.set_option("ozaki", 2)
    o2 = acq(X[:h]).clone()
    print(wl, "sub-batch forced vs FP64:", float((o2 - ref[:h]).abs().max() / sc))

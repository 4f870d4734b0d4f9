"""Run-to-run determinism probe: forward_backward and the device L-BFGS refinement repeated on identical inputs."""
import sys, torch
sys.path.insert(0, '.')
from everest_b200 import configs as Cf, optim
from tests.test_gpu_optimizer import _setup
for name, p in (("himmelblau", Cf.himmelblau_qlogei(N=200, S=128, raw=512)), ("zdt1", Cf.zdt1_qnehvi(N=300, S=64, raw=256, d=8, q=3))):
    st, acq, bnds, Xic, Yic = _setup(p)
    X = Xic.to(st.device)
    v0, g0 = acq.forward_backward(X)
    nv = ng = 0
    for _ in range(300):
        v, g = acq.forward_backward(X)
        nv += int(not torch.equal(v, v0)); ng += int(not torch.equal(g, g0))
    print(name, "forward_backward mismatches of 300: values", nv, "gradients", ng, "max |dg|", float((g - g0).abs().max()))
    outs = []
    for _ in range(8):
        Xd, Yd, info = optim.gen_candidates_device(Xic, acq, bnds[0], bnds[1], options={"maxiter": 200})
        outs.append((Xd.clone(), Yd.clone(), info.get("n_iter"), info.get("n_evals")))
    for k, (Xd, Yd, ni, ne) in enumerate(outs):
        print("  run", k, "dY", float((Yd - outs[0][1]).abs().max()), "dX", float((Xd - outs[0][0]).abs().max()), ni, ne)

"""Timing probe: one marginal-likelihood evaluation at BASELINE config 3 size (N=2000, d=30)."""
import sys, time, ctypes as C, numpy as np, torch
sys.path.insert(0, '.')
from everest_b200 import fit as F, kernels as K, configs as Cf, _lib as L
from everest_b200.model import DeviceGPState, SingleTaskGPSpec
p = Cf.zdt1_qnehvi()
X, y = p["X"], p["Y"][:, 1]
spec = SingleTaskGPSpec(kernel=K.RBFKernel(list(range(30)), [1.6] * 30), y=y, noise=1e-3)
def T():
    torch.cuda.synchronize(); return time.perf_counter()
for rep in range(3):
    t0 = T(); st = DeviceGPState(X, [spec]); t1 = T(); st.factorize(); t2 = T()
    mll, dn, dm = C.c_double(0), C.c_double(0), C.c_double(0)
    dls = (C.c_double * 30)(); dco = (C.c_double * 1)()
    L.check(st.lib.bo_mll_forward_backward(st.handle, 0, C.byref(mll), C.byref(dn), C.byref(dm), dls, 30, dco, 1, None)); t3 = T()
    L.check(st.lib.bo_mll_forward_backward(st.handle, 0, C.byref(mll), C.byref(dn), C.byref(dm), dls, 30, dco, 1, None)); t4 = T()
    st.close(); t5 = T()
    print(f"create {t1-t0:.4f} factorize {t2-t1:.4f} mll(first, builds Kinv) {t3-t2:.4f} mll(again) {t4-t3:.4f} close {t5-t4:.4f}")

"""Where the construction time of a solver goes (config C, pinned host inputs)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, ctypes as C
import __graft_entry__ as g; g.build()
from cmfpy_b200 import _lib
lib = _lib.load()
N, T, K, L = 1024, 1 << 20, 32, 64
dev = torch.device("cuda", 0)
Xh = torch.empty((N, T), dtype=torch.float32, pin_memory=True); Xh.copy_(torch.rand((N, T), device=dev))
W0 = np.random.rand(L, N, K).astype(np.float32) * 0.03; H0 = np.random.rand(K, T).astype(np.float32) * 0.03
torch.cuda.synchronize()
def tic(): torch.cuda.synchronize(); return time.perf_counter()
t0 = tic()
h = C.c_void_p()
p = _lib.Params(n_features=N, n_components=K, maxlag=L, t_local=T, t_global=T, t_offset=0, device=0, precision=1, stream=None, denominators=2)
_lib.check(lib.cmf_mu_create(C.byref(h), C.byref(p))); t1 = tic()
_lib.check(lib.cmf_mu_set_data(h, Xh.data_ptr(), 0, 0, T, T)); t2 = tic()
_lib.check(lib.cmf_mu_set_factors(h, W0.ctypes.data, H0.ctypes.data, 0, 0, T)); t3 = tic()
_lib.check(lib.cmf_mu_recon(h)); t4 = tic()
loss = (C.c_double * 1)()
_lib.check(lib.cmf_mu_step(h, 1, loss, None)); t5 = tic()
_lib.check(lib.cmf_mu_step(h, 1, loss, None)); t6 = tic()
print("create %.3f  set_data %.3f  set_factors %.3f  first recon %.3f  first step %.3f  second step %.3f" % (t1-t0, t2-t1, t3-t2, t4-t3, t5-t4, t6-t5))

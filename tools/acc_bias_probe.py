"""Measures the signed relative error of the tensor-core contractions against a
float64 evaluation on the SAME TF32-rounded operands (isolates accumulation
behaviour from operand rounding)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as g; g.build()
from cmfpy_b200.algs.mult import MultUpdate
from cmfpy_b200.model import ModelDimensions
from oracle import cmf_oracle as o
from tests.cases import make_inputs

def rn(a):
    a = np.ascontiguousarray(a, dtype=np.float32); u = a.view(np.uint32).astype(np.uint64)
    u = (u + 0xFFF + ((u >> 13) & 1)) & 0xFFFFE000
    return u.astype(np.uint32).view(np.float32).reshape(a.shape)

N, T, K, L = 1024, 4096, 32, 64
X, W0, H0 = make_inputs(N, T, K, L, "planted", seed=1)
Xq, Wq, Hq = rn(X).astype(np.float64), rn(W0).astype(np.float64), rn(H0).astype(np.float64)
est = o.cmf_predict(Wq, Hq)
estq = rn(est.astype(np.float32)).astype(np.float64)
numH, denH = o.h_terms(Xq, estq, Wq)
numW, denW = o.w_terms(Xq, estq, Hq, L)
for gram in ("0", "1"):
    os.environ["CMF_GRAM"] = gram
    alg = MultUpdate(X, ModelDimensions(X, maxlag=L, n_components=K), initW=W0, initH=H0, tol=0, precision="tf32")
    e = alg.est
    nW, dW = alg._compute_mult_W()
    nH, dH = alg._compute_mult_H()
    def stat(name, a, b):
        m = b > 1e-6 * b.max()
        r = (a[m] - b[m]) / b[m]
        print("gram=%s %-6s mean signed rel err %+.3e   rms %.3e" % (gram, name, r.mean(), np.sqrt((r ** 2).mean())))
    stat("est", e, est); stat("numW", nW, numW); stat("denW", dW, denW); stat("numH", nH, numH); stat("denH", dH, denH)
    alg.close()

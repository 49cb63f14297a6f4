"""Device HALS at BASELINE config B (N=256, T=65536, K=8, L=32): seconds per update()."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as g; g.build()
from cmfpy_b200.algs import ALGORITHMS
from cmfpy_b200.model import ModelDimensions
from tests.cases import make_inputs
N, T, K, L = 256, 65536, 8, 32
X, W0, H0 = make_inputs(N, T, K, L, "planted", seed=1)
alg = ALGORITHMS["hals"](X, ModelDimensions(X, maxlag=L, n_components=K), initW=W0, initH=H0, tol=0)
l0 = alg.loss
alg.update()
t0 = time.perf_counter()
ls = [alg.update() for _ in range(5)]
dt = (time.perf_counter() - t0) / 5
print("HALS config B: %.4f s per update (%d launches), loss %.5f -> %.5f" % (dt, alg.launch_count, l0, ls[-1]))
mu = ALGORITHMS["mult"](X, ModelDimensions(X, maxlag=L, n_components=K), initW=W0, initH=H0, tol=0, precision="tf32")
lm = mu.update_many(6)
print("MU after 6 updates: loss %.5f" % lm[-1])

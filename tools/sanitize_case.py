"""Smallest run that touches every kernel family, for compute-sanitizer (one tool per gpurun call):
    compute-sanitizer --tool memcheck|racecheck|synccheck python tools/sanitize_case.py [modes]
Two MU iterations per mode on a golden-sized problem, checked against the oracle so that a clean sanitizer log
belongs to a run whose results are right."""
import sys
import numpy as np
sys.path.insert(0, ".")
import __graft_entry__ as ge
ge.build()
from cmfpy_b200.algs.mult import MultUpdate
from cmfpy_b200.model import ModelDimensions
from oracle import cmf_oracle
from tests.cases import make_inputs

modes = sys.argv[1].split(",") if len(sys.argv) > 1 else ["fp32", "tf32", "tf32g", "tf32x3", "tf32x3g"]
for (N, T, K, L) in [(40, 600, 30, 9), (37, 301, 5, 7)]:
    X, W0, H0 = make_inputs(N, T, K, L, "planted", seed=3)
    ref = cmf_oracle.MultUpdateOracle(X.astype(np.float64), L, K, initW=W0.astype(np.float64), initH=H0.astype(np.float64), tol=0)
    ref_hist = [ref.loss] + [ref.update() for _ in range(2)]
    for m in modes:
        prec, den = (m[:-1], "gram") if m.endswith("g") else (m, "direct")
        alg = MultUpdate(X, ModelDimensions(X, maxlag=L, n_components=K), initW=W0, initH=H0, tol=0, precision=prec,
                         denominators=den)
        hist = [alg.loss] + [alg.update() for _ in range(2)]
        err = max(abs(a - b) / b for a, b in zip(hist, ref_hist))
        print("N=%d T=%d K=%d L=%d %-8s %-24s max rel err %.2e" % (N, T, K, L, m, alg.path_name, err), flush=True)
        assert err < (5e-3 if prec == "tf32" else 1e-4)
        alg.close()
print("done")

# per-case max relative loss-trajectory error for each precision mode (GPU box)
import sys, numpy as np
sys.path.insert(0, '.')
import __graft_entry__ as ge
ge.build()
from tests.test_parity_gpu import _inputs, _solver, _supported, ALL_CASES
modes = sys.argv[1].split(',') if len(sys.argv) > 1 else ["fp32", "tf32", "tf32g", "tf32x3"]
only = sys.argv[2].split(',') if len(sys.argv) > 2 else None
for name in ALL_CASES:
    if only and name not in only:
        continue
    g, X, W0, H0 = _inputs(name)
    N, T, K, L = (int(v) for v in g["shape"])
    n_iter = int(g["n_iter"])
    out = []
    for m in modes:
        if not _supported(m, N, K, L):
            out.append("%s: n/a" % m); continue
        alg = _solver(X, W0, H0, L, K, m)
        hist = np.array([alg.loss] + alg.update_many(n_iter))
        rel = np.abs(hist - g["loss_hist"]) / g["loss_hist"]
        out.append("%s: %.2e" % (m, rel.max()))
        alg.close()
    print("%-10s N=%d T=%d K=%d L=%d it=%d  " % (name, N, T, K, L, n_iter) + "  ".join(out), flush=True)

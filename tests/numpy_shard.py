"""A NumPy stand-in for cmfpy_b200.dist.DeviceShard (same phase interface), built
on the CPU oracle.  Test infrastructure: lets the multi-rank orchestration of
ShardedMultUpdate (halo exchange, W-term all-reduce, loss reduction) run under
the gloo backend on CPU."""
import numpy as np
import torch

from oracle import cmf_oracle as o


class NumpyShard:
    def __init__(self, X_local, N, T, K, L, t_offset, t_local):
        self.N, self.T, self.K, self.L = N, T, K, L
        self.t_offset, self.t_local, self.h = t_offset, t_local, L - 1
        X_local = np.asarray(X_local, dtype=np.float64)
        self.Xext = np.zeros((N, t_local + self.h))
        self.Xext[:, :X_local.shape[1]] = X_local          # own columns + static right halo
        self.n_valid = min(T - t_offset, t_local + self.h)  # columns that exist globally
        self.numden = torch.zeros(2 * L * N * K, dtype=torch.float32)
        self.sumsq = torch.zeros(1, dtype=torch.float64)
        self.launches = 0

    # data / factors
    def data_stats(self):
        X = self.Xext[:, :self.t_local]
        return float((X ** 2).sum()), bool((X < 0).any())

    def set_norm_x(self, v):
        self.norm_x = v

    def row_stats(self):
        X = self.Xext[:, :self.t_local]
        return np.stack([X.sum(axis=1), (X ** 2).sum(axis=1), np.abs(X).sum(axis=1)])

    def scale_rows(self, scale):
        self.Xext *= np.asarray(scale)[:, None]

    def set_factors(self, W0, H0):
        self.W = np.asarray(W0, dtype=np.float64).copy()
        self.Hwin = np.zeros((self.K, self.h + self.t_local + self.h))
        self.Hwin[:, self.h:self.h + self.t_local] = np.asarray(H0, dtype=np.float64)

    # halos, time-major (L-1) x K like the device engine
    def halo_buffers(self):
        mk = lambda: torch.zeros((max(self.h, 1), self.K), dtype=torch.float32)
        return mk(), mk(), mk(), mk()

    def halo_export(self, left_edge, right_edge):
        h, own = self.h, self.Hwin[:, self.h:self.h + self.t_local]
        if h:
            left_edge[:h] = torch.from_numpy(own[:, :h].T.astype(np.float32))
            right_edge[:h] = torch.from_numpy(own[:, -h:].T.astype(np.float32))

    def halo_import(self, left_halo, right_halo):
        h = self.h
        if not h:
            return
        self.Hwin[:, :h] = 0 if left_halo is None else left_halo[:h].numpy().T
        self.Hwin[:, h + self.t_local:] = 0 if right_halo is None else right_halo[:h].numpy().T

    # phases
    def recon(self):
        est = o.cmf_predict(self.W, self.Hwin)[:, self.h:]       # own + right halo columns
        est[:, self.n_valid:] = 0
        self.est = est
        d = est[:, :self.t_local] - self.Xext[:, :self.t_local]
        self.sumsq[0] = float((d ** 2).sum())
        self.launches += 1

    def w_terms(self):
        L, h, Tl = self.L, self.h, self.t_local
        num = np.zeros_like(self.W)
        den = np.zeros_like(self.W)
        for l in range(L):
            Hl = self.Hwin[:, h - l:h - l + Tl].T
            num[l] = self.Xext[:, :Tl] @ Hl
            den[l] = self.est[:, :Tl] @ Hl
        self.numden[:] = torch.from_numpy(np.concatenate([num.ravel(), den.ravel()]).astype(np.float32))
        self.launches += 1

    def w_terms_tensor(self):
        return self.numden

    def w_apply(self):
        n = self.W.size
        num = self.numden[:n].numpy().astype(np.float64).reshape(self.W.shape)
        den = self.numden[n:].numpy().astype(np.float64).reshape(self.W.shape)
        self.W = self.W * num / (den + o.EPSILON)

    def h_step(self):
        Tl, h = self.t_local, self.h
        nH = np.zeros((self.K, Tl))
        dH = np.zeros((self.K, Tl))
        for l in range(self.L):
            nH += self.W[l].T @ self.Xext[:, l:l + Tl]
            dH += self.W[l].T @ self.est[:, l:l + Tl]
        self.Hwin[:, h:h + Tl] *= nH / (dH + o.EPSILON)

    def resid_sumsq_tensor(self):
        return self.sumsq

    def step_fused(self, n):
        out = []
        for _ in range(n):
            self.w_terms(); self.w_apply(); self.recon(); self.h_step(); self.recon()
            out.append(float(np.sqrt(self.sumsq[0].item()) / self.norm_x))
        return out

    # read-back / bookkeeping
    def get_W(self):
        return self.W.copy()

    def get_H(self):
        return self.Hwin[:, self.h:self.h + self.t_local].copy()

    def launch_count(self):
        return self.launches

    def path_name(self):
        return "numpy-oracle"

    def set_profiling(self, on):
        pass

    def kernel_ms(self):
        return dict(recon=0.0, w_terms=0.0, h_terms=0.0, elementwise=0.0)

    def close(self):
        pass

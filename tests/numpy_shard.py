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


class NumpyPeerShard(NumpyShard):
    """NumpyShard plus the engine-side transport interface (`peer_export / peer_attach / peer_detach /
    step_sharded`) of cmfpy_b200.dist.DeviceShard.  The device engine moves the data with its own kernels over NVLink
    peer memory; this stand-in moves the same data with gloo collectives INSIDE the engine, in the protocol's order -
    reduce-scatter of slice r of the W terms in rank order, update of that slice, all-gather of W; halo columns to both
    neighbours; one residual sum per rank and step, added in rank order - so the host side of transport="peer"
    (blob exchange, attach handshake, the step_sharded loop, detach on close) runs on CPU."""

    BLOB = 512                                               # cmfpy_b200._lib.CMF_PEER_BLOB_BYTES

    def __init__(self, *a, group=None, fail_attach_on=None, **kw):
        super().__init__(*a, **kw)
        self.group, self.fail_attach_on = group, fail_attach_on
        self.attached = False

    def peer_export(self):
        import torch.distributed as dist
        return bytes([dist.get_rank(self.group) % 256]) * self.BLOB

    def peer_attach(self, rank, world, blobs):
        assert len(blobs) == world and all(len(b) == self.BLOB for b in blobs)
        assert [b[0] for b in blobs] == [r % 256 for r in range(world)]
        if self.fail_attach_on == rank:
            raise RuntimeError("simulated cudaIpcOpenMemHandle failure")
        self.rank, self.world, self.attached = rank, world, True

    def peer_detach(self):
        self.attached = False

    def step_sharded(self, n):
        import torch.distributed as dist
        assert self.attached
        losses = []
        for _ in range(n):
            self.w_terms()
            # W exchange: every rank owns slice r of the elements; partial sums are added in rank order
            parts = [torch.zeros_like(self.numden) for _ in range(self.world)]
            dist.all_gather(parts, self.numden, group=self.group)
            total = parts[0].clone()
            for p in parts[1:]:
                total += p
            self.numden[:] = total
            self.w_apply()
            self.recon()
            self.h_step()
            # halo pushes
            sl, sr, rl, rr = self.halo_buffers()
            self.halo_export(sl, sr)
            if self.h:
                ops = []
                if self.rank + 1 < self.world:
                    ops += [dist.P2POp(dist.isend, sr, self.rank + 1, self.group), dist.P2POp(dist.irecv, rr, self.rank + 1, self.group)]
                if self.rank > 0:
                    ops += [dist.P2POp(dist.isend, sl, self.rank - 1, self.group), dist.P2POp(dist.irecv, rl, self.rank - 1, self.group)]
                for req in dist.batch_isend_irecv(ops):
                    req.wait()
                self.halo_import(rl if self.rank > 0 else None, rr if self.rank + 1 < self.world else None)
            self.recon()
            # loss ring: one value per rank, added in rank order
            ring = [torch.zeros_like(self.sumsq) for _ in range(self.world)]
            dist.all_gather(ring, self.sumsq, group=self.group)
            losses.append(float(np.sqrt(sum(float(v.item()) for v in ring)) / self.norm_x))
        return losses

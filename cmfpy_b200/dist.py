"""Time-sharded MU solver: one process per GPU, `torch.distributed` for the
plumbing.

Rank g owns the global columns [t_offset, t_offset + t_local) of X, est and H;
W and the W-step numerator/denominator are replicated.  Per iteration:

    W terms (local partial sums)  -> all-reduce(sum) of 2*L*N*K fp32
    W update (identical on every rank)
    reconstruction                (H unchanged: no exchange needed)
    H terms + H update            (local; needs the static right halo of X and
                                   est on own + L-1 columns)
    halo exchange of L-1 H columns with both neighbours
    reconstruction + local sum of squared residuals -> all-reduce of 1 double

The orchestration is written against a small "shard engine" interface so that
the host logic can be exercised with the gloo backend on CPU (tests supply a
NumPy engine); `DeviceShard` is the real engine, a thin wrapper over the C ABI.
With world size 1 the fused single-GPU iteration (cmf_mu_step) is used.
"""
import ctypes as C

import numpy as np

from . import _lib


class _CudaView:
    """Exposes a raw device pointer through __cuda_array_interface__ so that
    torch can wrap library-owned memory without a copy."""

    def __init__(self, ptr, count, typestr):
        self.__cuda_array_interface__ = {
            "shape": (int(count),), "typestr": typestr, "data": (int(ptr), False),
            "version": 3, "strides": None}


class DeviceShard:
    """One time shard on one GPU, behind libcmf_b200 (the C ABI)."""

    def __init__(self, X, N, T, K, L, t_offset, t_local, precision, device, stream_ptr, denominators="direct"):
        import torch
        self._torch = torch
        self._lib = _lib.load()
        self._h = C.c_void_p()
        self.N, self.T, self.K, self.L = N, T, K, L
        self.t_offset, self.t_local, self.device = t_offset, t_local, device
        p = _lib.Params(n_features=N, n_components=K, maxlag=L, t_local=t_local, t_global=T,
                        t_offset=t_offset, device=device, precision=_lib.PRECISIONS[precision],
                        stream=stream_ptr, denominators=_lib.DENOMINATORS[denominators])
        import time
        t0 = time.perf_counter()
        _lib.check(self._lib.cmf_mu_create(C.byref(self._h), C.byref(p)))
        t1 = time.perf_counter()
        ptr, dt, mem, ld, ncols = self._describe(X)
        if not (t_local <= ncols <= t_local + L - 1):
            raise ValueError("X must hold t_local .. t_local+L-1 columns, got %d" % ncols)
        _lib.check(self._lib.cmf_mu_set_data(self._h, ptr, dt, mem, ld, ncols))
        # host seconds of the (synchronous) construction calls: what bench.py's end-to-end figure is made of
        self.timings = {"create": t1 - t0, "set_data": time.perf_counter() - t1}
        self._keepalive = None

    def _describe(self, a):
        """(pointer, dtype code, memory space, leading dimension, n_cols) of a
        2-D numpy array or torch tensor (row-major)."""
        torch = self._torch
        if isinstance(a, torch.Tensor):
            if a.dim() != 2 or a.stride(1) != 1:
                raise ValueError("expected a row-major 2-D tensor")
            dt = {torch.float32: _lib.CMF_F32, torch.float64: _lib.CMF_F64}[a.dtype]
            mem = _lib.CMF_DEVICE if a.is_cuda else _lib.CMF_HOST
            return a.data_ptr(), dt, mem, a.stride(0), a.shape[1]
        a = np.asarray(a)
        if a.ndim != 2 or a.strides[1] != a.itemsize:
            raise ValueError("expected a row-major 2-D array")
        return a.ctypes.data, _lib.np_dtype_code(a), _lib.CMF_HOST, a.strides[0] // a.itemsize, a.shape[1]

    # -- data / factors --------------------------------------------------------
    def data_stats(self):
        ss, neg = C.c_double(0), C.c_int(0)
        _lib.check(self._lib.cmf_mu_data_stats(self._h, C.byref(ss), C.byref(neg)))
        return ss.value, bool(neg.value)

    def set_norm_x(self, v):
        _lib.check(self._lib.cmf_mu_set_norm_x(self._h, float(v)))

    def row_stats(self):
        """Per-feature (sum x, sum x^2, sum |x|) over the owned columns, stacked as a 3 x N array."""
        out = np.empty((3, self.N), dtype=np.float64)
        _lib.check(self._lib.cmf_mu_row_stats(self._h, out[0].ctypes.data, out[1].ctypes.data, out[2].ctypes.data))
        return out

    def scale_rows(self, scale):
        scale = np.ascontiguousarray(scale, dtype=np.float64)
        _lib.check(self._lib.cmf_mu_scale_rows(self._h, scale.ctypes.data))

    def set_factors(self, W0, H0):
        torch = self._torch
        if isinstance(W0, torch.Tensor):
            W0c = W0.contiguous()
            hp, dt, mem, ldh, ncols = self._describe(H0)
            if W0c.dtype != H0.dtype or W0c.is_cuda != H0.is_cuda:
                raise ValueError("W0 and H0 must share dtype and device")
            wp = W0c.data_ptr()
        else:
            W0c = np.ascontiguousarray(W0)
            H0 = np.asarray(H0, dtype=W0c.dtype)
            hp, dt, mem, ldh, ncols = self._describe(H0)
            wp = W0c.ctypes.data
        if tuple(W0c.shape) != (self.L, self.N, self.K) or ncols != self.t_local:
            raise ValueError("initW must be L x N x K and initH K x t_local")
        _lib.check(self._lib.cmf_mu_set_factors(self._h, wp, hp, dt, mem, ldh))

    # -- phases ---------------------------------------------------------------
    def recon(self):
        _lib.check(self._lib.cmf_mu_recon(self._h))

    def recon_loss(self):
        _lib.check(self._lib.cmf_mu_recon_loss(self._h))

    def w_terms(self):
        _lib.check(self._lib.cmf_mu_w_terms(self._h))

    def w_terms_tensor(self):
        ptr, cnt = C.c_void_p(), C.c_longlong(0)
        _lib.check(self._lib.cmf_mu_w_terms_buffer(self._h, C.byref(ptr), C.byref(cnt)))
        return self._torch.as_tensor(_CudaView(ptr.value, 2 * cnt.value, "<f4"),
                                     device=self._torch.device("cuda", self.device))

    def w_apply(self):
        _lib.check(self._lib.cmf_mu_w_apply(self._h))

    def h_step(self):
        _lib.check(self._lib.cmf_mu_h_step(self._h))

    def needs_mid_recon(self):
        v = C.c_int(1)
        _lib.check(self._lib.cmf_mu_needs_mid_recon(self._h, C.byref(v)))
        return bool(v.value)

    def halo_buffers(self):
        n, ld = C.c_int(0), C.c_int(0)
        _lib.check(self._lib.cmf_mu_halo_width(self._h, C.byref(n), C.byref(ld)))
        dev = self._torch.device("cuda", self.device)
        mk = lambda: self._torch.zeros((max(n.value, 1), ld.value), dtype=self._torch.float32, device=dev)
        return mk(), mk(), mk(), mk()          # send_left, send_right, recv_left, recv_right

    def halo_export(self, left_edge, right_edge):
        _lib.check(self._lib.cmf_mu_halo_export(self._h, left_edge.data_ptr(), right_edge.data_ptr()))

    def halo_import(self, left_halo, right_halo):
        _lib.check(self._lib.cmf_mu_halo_import(
            self._h, left_halo.data_ptr() if left_halo is not None else None,
            right_halo.data_ptr() if right_halo is not None else None))

    def resid_sumsq_tensor(self):
        ptr = C.c_void_p()
        _lib.check(self._lib.cmf_mu_resid_sumsq_buffer(self._h, C.byref(ptr)))
        return self._torch.as_tensor(_CudaView(ptr.value, 1, "<f8"),
                                     device=self._torch.device("cuda", self.device))

    def step_fused(self, n):
        losses = np.empty(n, dtype=np.float64)
        _lib.check(self._lib.cmf_mu_step(self._h, n, losses.ctypes.data_as(C.POINTER(C.c_double)), None))
        return [float(x) for x in losses]

    # -- collectives over peer memory (CUDA IPC, NVLink P2P) -----------------------
    def peer_export(self):
        blob = (C.c_ubyte * _lib.CMF_PEER_BLOB_BYTES)()
        _lib.check(self._lib.cmf_mu_peer_export(self._h, blob))
        return bytes(blob)

    def peer_attach(self, rank, world, blobs):
        buf = b"".join(blobs)
        assert len(buf) == world * _lib.CMF_PEER_BLOB_BYTES
        _lib.check(self._lib.cmf_mu_peer_attach(self._h, rank, world, buf))

    def peer_detach(self):
        _lib.check(self._lib.cmf_mu_peer_detach(self._h))

    def step_sharded(self, n):
        losses = np.empty(n, dtype=np.float64)
        _lib.check(self._lib.cmf_mu_step_sharded(self._h, n, losses.ctypes.data_as(C.POINTER(C.c_double))))
        return [float(x) for x in losses]

    # -- read-back --------------------------------------------------------------
    def _host_out(self, out, shape):
        """A caller-supplied float32 C-contiguous array (e.g. the NumPy view of a pinned torch tensor: the copy is
        then a direct DMA instead of a staged copy into freshly faulted pages), or a new one."""
        if out is None:
            return np.empty(shape, dtype=np.float32)
        if out.shape != shape or out.dtype != np.float32 or not out.flags.c_contiguous:
            raise ValueError("out must be a C-contiguous float32 array of shape %s" % (shape,))
        return out

    def get_W(self, out=None):
        out = self._host_out(out, (self.L, self.N, self.K))
        _lib.check(self._lib.cmf_mu_get_W(self._h, out.ctypes.data, _lib.CMF_F32, _lib.CMF_HOST))
        return out

    def get_H(self, out=None):
        out = self._host_out(out, (self.K, self.t_local))
        _lib.check(self._lib.cmf_mu_get_H(self._h, out.ctypes.data, _lib.CMF_F32, _lib.CMF_HOST, self.t_local))
        return out

    def get_est(self):
        out = np.empty((self.N, self.t_local), dtype=np.float32)
        _lib.check(self._lib.cmf_mu_get_est(self._h, out.ctypes.data, _lib.CMF_F32, _lib.CMF_HOST, self.t_local))
        return out

    # -- bookkeeping --------------------------------------------------------------
    def launch_count(self):
        n = C.c_longlong(0)
        _lib.check(self._lib.cmf_mu_launch_count(self._h, C.byref(n)))
        return n.value

    def path_name(self):
        return self._lib.cmf_mu_path_name(self._h).decode()

    def set_profiling(self, on):
        _lib.check(self._lib.cmf_mu_set_profiling(self._h, int(on)))

    def kernel_ms(self):
        out = (C.c_float * 4)()
        _lib.check(self._lib.cmf_mu_kernel_ms(self._h, out))
        return dict(recon=out[0], w_terms=out[1], h_terms=out[2], elementwise=out[3])

    def launch_table(self):
        """{kernel label: (launches, total ms)} since set_profiling(2)."""
        return _lib.launch_table(self._lib, self._h)

    def close(self):
        if self._h is not None and self._h.value:
            self._lib.cmf_mu_destroy(self._h)
            self._h = C.c_void_p()


class ShardedMultUpdate:
    """MU solver over a time shard per rank (duck type of reference MultUpdate:
    `update`, `loss`, `converged`, plus `update_many`)."""

    def __init__(self, X_local, N, T, K, L, t_offset, t_local, initW, initH,
                 precision="fp32", device=0, group=None, tol=1e-5, patience=3,
                 engine=None, denominators="auto", transport="auto", normalize=None):
        import torch
        import torch.distributed as dist
        self._torch, self._dist = torch, dist
        self.group = group
        self.world = dist.get_world_size(group) if group is not None else 1
        self.rank = dist.get_rank(group) if group is not None else 0
        self.N, self.T, self.K, self.L = N, T, K, L
        self.t_offset, self.t_local = t_offset, t_local
        self.tol, self.patience, self.precision = tol, patience, precision
        if self.world > 1 and t_local < L - 1:
            raise ValueError("each shard needs at least L-1 columns")
        self.profiled_ms = dict(recon=0.0, w_terms=0.0, h_terms=0.0, elementwise=0.0)
        self._profiling = False

        if engine is None:
            self.torch_stream = torch.cuda.Stream(device=device)
            self.engine = DeviceShard(X_local, N, T, K, L, t_offset, t_local, precision, device,
                                      self.torch_stream.cuda_stream, denominators)
        else:
            self.torch_stream = None
            self.engine = engine
        eng = self.engine
        if normalize is not None:
            # dataset normalisation (songbird.py:18-19 / maze.py:71-72 / vox_celeb.py:100-102) on the shards: the
            # per-feature sums are all-reduced, every rank scales its own columns
            from .common import row_scales
            st = eng.row_stats()
            if self.world > 1:
                dev = "cpu" if self.torch_stream is None else self.torch_stream.device
                t = torch.as_tensor(st, device=dev)
                self._all_reduce(t)
                st = self._to_host(t).numpy()
            self.row_scale = row_scales(normalize, st[0], st[1], st[2], T)
            eng.scale_rows(self.row_scale)
        ss, self.has_negative = eng.data_stats()
        if self.world > 1:
            t = self._host_scalar(ss)
            self._all_reduce(t)
            ss = float(self._to_host(t).item())
        self.normX = float(np.sqrt(ss))
        eng.set_norm_x(self.normX)
        import time
        t0 = time.perf_counter()
        eng.set_factors(initW, initH)
        if hasattr(eng, "timings"):
            eng.timings["set_factors"] = time.perf_counter() - t0
        if self.world > 1:
            self._sl, self._sr, self._rl, self._rr = eng.halo_buffers()
            if self.torch_stream is not None:       # buffers were zero-filled on torch's current stream
                torch.cuda.current_stream(self.torch_stream.device).synchronize()
            self._exchange_halos()
        # the initial loss; est itself is only stored where an MU step will read it (not on the full Gram route)
        t0 = time.perf_counter()
        (eng.recon_loss if hasattr(eng, "recon_loss") else eng.recon)()
        if hasattr(eng, "timings"):
            self._torch.cuda.synchronize()
            eng.timings["first_loss"] = time.perf_counter() - t0
        self._loss = None
        # Collectives: "peer" = the library's own kernels over NVLink peer memory (the W-term all-reduce
        # fused with the W update, halo pushes, a loss ring; no host-side collective per iteration);
        # "nccl" = torch.distributed collectives between the phases.  "auto" takes the peer path when
        # the engine offers it and every rank could map its peers (CMF_TRANSPORT overrides).
        import os
        transport = os.environ.get("CMF_TRANSPORT", transport)
        if transport not in ("auto", "peer", "nccl"):
            raise ValueError("transport must be auto, peer or nccl")
        self.transport = "nccl"
        if self.world > 1 and transport != "nccl" and hasattr(eng, "peer_export") and self.world <= 8:
            self._attach_peers(required=(transport == "peer"))
        elif self.world > 1 and transport == "peer":
            raise RuntimeError("transport='peer' needs the device engine and at most 8 ranks")

    def _attach_peers(self, required):
        dist, eng = self._dist, self.engine
        err = None
        try:
            blob = eng.peer_export()
        except Exception as e:           # noqa: BLE001 - reported to all ranks below
            blob, err = b"", e
        blobs = [None] * self.world
        dist.all_gather_object(blobs, blob, group=self.group)
        ok = all(len(b) == _lib.CMF_PEER_BLOB_BYTES for b in blobs)
        if ok:
            try:
                eng.peer_attach(self.rank, self.world, blobs)
            except Exception as e:       # noqa: BLE001
                ok, err = False, e
        flags = [None] * self.world
        dist.all_gather_object(flags, ok, group=self.group)      # also the barrier between attach and first use
        if all(flags):
            self.transport = "peer"
            return
        if ok:
            eng.peer_detach()
        if required:
            raise RuntimeError("peer-memory transport unavailable on some rank: %s" % (err,))

    # -- collectives ----------------------------------------------------------
    def _stream_ctx(self):
        import contextlib
        if self.torch_stream is None:
            return contextlib.nullcontext()
        return self._torch.cuda.stream(self.torch_stream)

    def _host_scalar(self, v):
        torch = self._torch
        dev = "cpu" if self.torch_stream is None else self.torch_stream.device
        return torch.tensor([v], dtype=torch.float64, device=dev)

    def _all_reduce(self, t):
        with self._stream_ctx():
            self._dist.all_reduce(t, op=self._dist.ReduceOp.SUM, group=self.group)

    def _to_host(self, t):
        """Device -> host on the solver's stream (a .item() on another stream
        would race with the collective that produced `t`)."""
        with self._stream_ctx():
            out = t.cpu()
        if self.torch_stream is not None:
            self.torch_stream.synchronize()
        return out

    def _exchange_halos(self):
        """Send my first/last L-1 H columns to the left/right neighbour and
        receive theirs; the global boundaries get zeros."""
        if self.L == 1:
            return
        dist, eng = self._dist, self.engine
        eng.halo_export(self._sl, self._sr)
        left = self.rank - 1 if self.rank > 0 else None
        right = self.rank + 1 if self.rank < self.world - 1 else None
        ops = []
        gl = (lambda r: dist.get_global_rank(self.group, r)) if self.group is not None else (lambda r: r)
        if right is not None:
            ops.append(dist.P2POp(dist.isend, self._sr, gl(right), self.group))
            ops.append(dist.P2POp(dist.irecv, self._rr, gl(right), self.group))
        if left is not None:
            ops.append(dist.P2POp(dist.isend, self._sl, gl(left), self.group))
            ops.append(dist.P2POp(dist.irecv, self._rl, gl(left), self.group))
        with self._stream_ctx():
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        eng.halo_import(self._rl if left is not None else None,
                        self._rr if right is not None else None)

    # -- solver interface -------------------------------------------------------
    def update_many(self, n):
        eng = self.engine
        if n <= 0:
            return []
        if self.world == 1:
            losses = eng.step_fused(n)
            self._loss = losses[-1]
            return losses
        if self.transport == "peer":
            losses = eng.step_sharded(n)
            self._loss = losses[-1]
            return losses
        torch = self._torch
        sums = []
        marks = []
        mid_recon = eng.needs_mid_recon() if hasattr(eng, "needs_mid_recon") else True

        def mark():
            if self._profiling and self.torch_stream is not None:
                e = torch.cuda.Event(enable_timing=True)
                e.record(self.torch_stream)
                marks.append(e)

        for _ in range(n):
            mark()
            eng.w_terms()
            mark()
            self._all_reduce(eng.w_terms_tensor())
            eng.w_apply()
            mark()
            if mid_recon:
                eng.recon()
            mark()
            eng.h_step()
            mark()
            self._exchange_halos()
            mark()
            (eng.recon_loss if hasattr(eng, "recon_loss") else eng.recon)()
            mark()
            with self._stream_ctx():
                s = eng.resid_sumsq_tensor().clone()
            self._all_reduce(s)
            sums.append(s)
        with self._stream_ctx():
            cat = torch.cat(sums)
        allsums = self._to_host(cat)
        losses = [float(np.sqrt(v) / self.normX) for v in allsums.tolist()]
        self._loss = losses[-1]
        if marks:
            self.profiled_ms = dict(recon=0.0, w_terms=0.0, h_terms=0.0, elementwise=0.0, comm=0.0)
            for i in range(0, len(marks), 7):
                m = marks[i:i + 7]
                dt = [m[k].elapsed_time(m[k + 1]) for k in range(6)]
                self.profiled_ms["w_terms"] += dt[0]
                self.profiled_ms["comm"] += dt[1] + dt[4]      # all-reduce (+W update), halo exchange
                self.profiled_ms["recon"] += dt[2] + dt[5]
                self.profiled_ms["h_terms"] += dt[3]           # H terms + H update
        return losses

    def update(self):
        return self.update_many(1)[0]

    @property
    def loss(self):
        if self._loss is None:
            with self._stream_ctx():
                s = self.engine.resid_sumsq_tensor().clone()
            if self.world > 1:
                self._all_reduce(s)
            self._loss = float(np.sqrt(float(self._to_host(s).item())) / self.normX)
        return self._loss

    def converged(self, loss_hist):
        d = np.diff(loss_hist[-self.patience:])
        return bool(np.all(np.abs(d) < self.tol))

    # -- read-back ----------------------------------------------------------------
    def W_host(self, out=None):
        return self.engine.get_W(out) if out is not None else self.engine.get_W()

    def H_local_host(self, out=None):
        return self.engine.get_H(out) if out is not None else self.engine.get_H()

    def est_local_host(self):
        return self.engine.get_est()

    # -- bookkeeping ----------------------------------------------------------------
    @property
    def launch_count(self):
        return self.engine.launch_count()

    @property
    def path_name(self):
        return self.engine.path_name()

    def set_profiling(self, on):
        self._profiling = bool(on)
        self.engine.set_profiling(on)

    def launch_table(self):
        return self.engine.launch_table() if hasattr(self.engine, "launch_table") else {}

    def kernel_ms(self):
        """Device time per phase during the last update_many (CUDA events)."""
        if self.world == 1 or self.transport == "peer":
            return self.engine.kernel_ms()
        return dict(self.profiled_ms)

    def close(self):
        if self.transport == "peer":
            # Safe without a host barrier: when cmf_mu_step_sharded has returned here, every peer has
            # already issued its last store into this rank's memory (the end-of-call barrier kernel is
            # each rank's final peer access, and this rank saw all of them).
            self.engine.peer_detach()
            self.transport = "nccl"
        self.engine.close()

"""Multi-GPU parity (needs >= 2 B200s; skipped otherwise): the time-sharded
solver over NCCL must reproduce the single-GPU solver on the same inputs."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _ngpu():
    from cmfpy_b200 import _lib
    return _lib.device_count()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, shape, precision, n_iter, q, transport="auto"):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from cmfpy_b200.dist import ShardedMultUpdate
        from tests.cases import make_inputs
        N, T, K, L = shape
        X, W0, H0 = make_inputs(N, T, K, L, "planted", seed=5)
        Tl = T // world
        t0 = rank * Tl
        ncols = min(Tl + L - 1, T - t0)
        prec, den = (precision[:-1], "gram") if precision.endswith("g") else (precision, "direct")
        alg = ShardedMultUpdate(np.ascontiguousarray(X[:, t0:t0 + ncols]), N, T, K, L, t_offset=t0, t_local=Tl,
                                initW=W0, initH=np.ascontiguousarray(H0[:, t0:t0 + Tl]), precision=prec,
                                device=rank, group=dist.group.WORLD, tol=0, denominators=den, transport=transport)
        assert alg.transport == transport
        hist = [alg.loss] + alg.update_many(n_iter // 2) + alg.update_many(n_iter - n_iter // 2)
        H, W = alg.H_local_host(), alg.W_host()
        out = [None] * world
        dist.all_gather_object(out, (H, W, hist))
        if rank == 0:
            q.put(out)
        alg.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("transport", ["peer", "nccl"])
@pytest.mark.parametrize("precision,shape", [("fp32", (96, 2048, 5, 12)), ("fp32", (64, 1024, 32, 33)),
                                             ("tf32", (200, 4096, 32, 64)), ("tf32", (128, 2048, 30, 9)),
                                             ("tf32g", (200, 4096, 32, 64)), ("tf32g", (96, 2048, 5, 12)),
                                             ("tf32g", (256, 2048, 128, 16)), ("tf32x3", (200, 4096, 32, 64)),
                                             ("tf32x3g", (200, 4096, 32, 64)), ("fp32", (40, 600, 3, 1))])
def test_sharded_equals_single_gpu(built_lib, precision, shape, transport):
    """transport "peer": the library's own collectives over NVLink peer memory (fused all-reduce + W update,
    halo pushes, loss ring); "nccl": torch.distributed collectives between the phases."""
    world = min(_ngpu(), 4)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    import torch.multiprocessing as mp
    from cmfpy_b200.algs.mult import MultUpdate
    from cmfpy_b200.model import ModelDimensions
    from tests.cases import make_inputs
    N, T, K, L = shape
    n_iter = 6
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, shape, precision, n_iter, q, transport))
             for r in range(world)]
    for p in procs:
        p.start()
    out = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    X, W0, H0 = make_inputs(N, T, K, L, "planted", seed=5)
    prec, den = (precision[:-1], "gram") if precision.endswith("g") else (precision, "direct")
    ref = MultUpdate(X, ModelDimensions(X, maxlag=L, n_components=K), initW=W0, initH=H0, tol=0, precision=prec,
                     denominators=den)
    ref_hist = [ref.loss] + ref.update_many(n_iter)
    tol = 2e-5 if precision in ("fp32", "tf32x3", "tf32x3g") else 2e-4
    H = np.concatenate([o[0] for o in out], axis=1)
    for o in out:
        assert np.abs(np.array(o[2]) - ref_hist).max() / ref_hist[-1] < tol
        assert np.abs(o[1] - ref.W).max() <= 50 * tol * np.abs(ref.W).max()
        assert np.array_equal(o[1], out[0][1]), "W must be bit-identical across ranks"
    assert np.abs(H - ref.H).max() <= 50 * tol * np.abs(ref.H).max()
    ref.close()


# ---- the same sharded solve from ONE process: CMF(..., devices=[...]) --------------------------------
@pytest.mark.parametrize("precision,shape", [("fp32", (96, 2048, 5, 12)), ("tf32x3", (200, 4096, 32, 64)),
                                             ("tf32x3", (130, 3001, 8, 20)), ("tf32", (128, 2048, 30, 9))])
def test_cmf_fit_on_several_devices(built_lib, precision, shape):
    """`devices` travels through alg_opts like every solver option (reference model.py:80-82, :146); the fit over
    all GPUs must reproduce the single-GPU fit, and the float64 oracle where the mode is parity-grade."""
    n = _ngpu()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    from cmfpy_b200 import CMF
    from oracle import cmf_oracle as o
    from tests.cases import make_inputs
    N, T, K, L = shape
    X, W0, H0 = make_inputs(N, T, K, L, "planted", seed=5)
    devices = list(range(min(n, 4)))
    if T // len(devices) < L:
        devices = devices[:2]
    many = CMF(K, L, n_iter_max=8, verbose=False, tol=0, initW=W0, initH=H0, precision=precision, devices=devices)
    many.fit(X)
    one = CMF(K, L, n_iter_max=8, verbose=False, tol=0, initW=W0, initH=H0, precision=precision, devices=[0])
    one.fit(X)
    tol = 5e-3 if precision == "tf32" else 2e-5
    a, b = np.array(many.loss_hist), np.array(one.loss_hist)
    assert len(a) == 9 and np.abs(a - b).max() <= tol * b.max()
    assert np.abs(many.motifs - one.motifs).max() <= 50 * tol * np.abs(one.motifs).max()
    assert np.abs(many.factors - one.factors).max() <= 50 * tol * np.abs(one.factors).max()
    if precision != "tf32":
        ref = o.MultUpdateOracle(X.astype(np.float64), L, K, initW=W0.astype(np.float64), initH=H0.astype(np.float64), tol=0)
        ref_hist = np.array([ref.loss] + [ref.update() for _ in range(8)])
        assert (np.abs(a - ref_hist) / ref_hist).max() <= 1e-4


def test_multi_device_random_init_and_early_stop(built_lib):
    """rand_init (reference base.py:78-88) summed over the shards: <X, est0> = ||est0||^2; converged() fires."""
    n = _ngpu()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    from cmfpy_b200.algs.mult import MultUpdate
    from cmfpy_b200.model import ModelDimensions
    from tests.cases import make_inputs
    X, _, _ = make_inputs(40, 1200, 3, 9, "planted", seed=9)
    alg = MultUpdate(X, ModelDimensions(X, maxlag=9, n_components=3), seed=1, devices=[0, 1], tol=1e-3)
    assert type(alg).__name__ == "MultiGpuMultUpdate"
    est = alg.est
    assert abs((X * est).sum() / (est ** 2).sum() - 1.0) < 1e-4
    hist = [alg.loss]
    for _ in range(2000):
        hist.append(alg.update())
        if alg.converged(hist):
            break
    assert 3 < len(hist) < 2000
    assert np.all(np.diff(hist) < 1e-6)
    alg.close()

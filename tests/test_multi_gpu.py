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

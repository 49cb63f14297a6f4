"""World-size-2 (and 3) tests of the time-sharded driver on CPU with the gloo
backend: the orchestration of cmfpy_b200.dist.ShardedMultUpdate (halo exchange
order, W-term all-reduce, loss reduction, boundary zeros) must reproduce the
unsharded reference algorithm."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import cmf_oracle as o


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, shape, n_iter, out, normalize=None):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from cmfpy_b200.dist import ShardedMultUpdate
        from tests.numpy_shard import NumpyShard
        N, T, K, L = shape
        rng = np.random.default_rng(0)
        X, W0, H0 = rng.random((N, T)), rng.random((L, N, K)), rng.random((K, T))
        Tl = T // world
        t0 = rank * Tl
        ncols = min(Tl + L - 1, T - t0)
        eng = NumpyShard(X[:, t0:t0 + ncols], N, T, K, L, t0, Tl)
        alg = ShardedMultUpdate(None, N, T, K, L, t_offset=t0, t_local=Tl, initW=W0, initH=H0[:, t0:t0 + Tl],
                                group=dist.group.WORLD, engine=eng, tol=0, normalize=normalize)
        l0 = alg.loss
        losses = alg.update_many(n_iter - 1) + [alg.update()]
        H = alg.H_local_host()
        W = alg.W_host()
        gathered = [None] * world
        dist.all_gather_object(gathered, (H, W, [l0] + losses))
        if rank == 0:
            out.put(gathered)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,shape", [(2, (6, 64, 3, 5)), (3, (5, 90, 2, 9)), (2, (4, 40, 2, 1))])
def test_sharded_driver_matches_unsharded_oracle(world, shape):
    n_iter = 4
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, shape, n_iter, q)) for r in range(world)]
    for p in procs:
        p.start()
    gathered = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    N, T, K, L = shape
    rng = np.random.default_rng(0)
    X, W0, H0 = rng.random((N, T)), rng.random((L, N, K)), rng.random((K, T))
    ref = o.MultUpdateOracle(X, L, K, initW=W0, initH=H0, tol=0)
    ref_hist = [ref.loss] + [ref.update() for _ in range(n_iter)]
    H = np.concatenate([g[0] for g in gathered], axis=1)
    for g in gathered:                       # W is replicated and identical on every rank
        np.testing.assert_allclose(g[1], ref.W, rtol=2e-5)
        np.testing.assert_allclose(g[2], ref_hist, rtol=1e-6)
    np.testing.assert_allclose(H, ref.H, rtol=2e-5)


def test_single_rank_uses_fused_step():
    from cmfpy_b200.dist import ShardedMultUpdate
    from tests.numpy_shard import NumpyShard
    rng = np.random.default_rng(1)
    N, T, K, L = 5, 50, 2, 4
    X, W0, H0 = rng.random((N, T)), rng.random((L, N, K)), rng.random((K, T))
    alg = ShardedMultUpdate(None, N, T, K, L, 0, T, W0, H0, engine=NumpyShard(X, N, T, K, L, 0, T), tol=0)
    ref = o.MultUpdateOracle(X, L, K, initW=W0, initH=H0, tol=0)
    assert alg.loss == pytest.approx(ref.loss, rel=1e-12)
    got = alg.update_many(3)
    exp = [ref.update() for _ in range(3)]
    np.testing.assert_allclose(got, exp, rtol=1e-6)
    assert not alg.converged([1.0, 0.5, 0.5])              # tol = 0: strict "<" never fires
    alg.tol = 1e-5
    assert not alg.converged([1.0, 0.5, 0.4])
    assert alg.converged([1.0, 0.5, 0.5 + 1e-9, 0.5 + 2e-9])


@pytest.mark.parametrize("mode", ["l2", "l1", "std"])
def test_sharded_normalisation_matches_global(mode):
    """Dataset normalisation on shards (songbird.py:18-19, maze.py:71-72, vox_celeb.py:100-102): the per-feature
    sums are all-reduced, so every shard scales with the GLOBAL row norms and the trajectory equals the one of the
    unsharded oracle on the normalised matrix."""
    from cmfpy_b200.common import row_scales
    world, shape, n_iter = 2, (6, 64, 3, 5), 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, shape, n_iter, q, mode)) for r in range(world)]
    for p in procs:
        p.start()
    gathered = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    N, T, K, L = shape
    rng = np.random.default_rng(0)
    X, W0, H0 = rng.random((N, T)), rng.random((L, N, K)), rng.random((K, T))
    Xn = X * row_scales(mode, X.sum(1), (X ** 2).sum(1), np.abs(X).sum(1), T)[:, None]
    ref = o.MultUpdateOracle(Xn, L, K, initW=W0, initH=H0, tol=0)
    ref_hist = [ref.loss] + [ref.update() for _ in range(n_iter)]
    for H, W, hist in gathered:
        np.testing.assert_allclose(hist, ref_hist, rtol=1e-6)


def _peer_required_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from cmfpy_b200.dist import ShardedMultUpdate
        from tests.numpy_shard import NumpyShard
        N, T, K, L = 4, 40, 2, 3
        rng = np.random.default_rng(0)
        X, W0, H0 = rng.random((N, T)), rng.random((L, N, K)), rng.random((K, T))
        Tl, t0 = T // world, rank * (T // world)
        eng = NumpyShard(X[:, t0:min(T, t0 + Tl + L - 1)], N, T, K, L, t0, Tl)
        try:
            ShardedMultUpdate(None, N, T, K, L, t_offset=t0, t_local=Tl, initW=W0, initH=H0[:, t0:t0 + Tl],
                              group=dist.group.WORLD, engine=eng, tol=0, transport="peer")
            res = "no error"
        except RuntimeError as e:
            res = str(e)
        if rank == 0:
            out.put(res)
    finally:
        dist.destroy_process_group()


def test_peer_transport_must_be_available_when_required():
    """transport="peer" is a demand, not a hint: an engine without peer memory (here the NumPy stand-in) must not
    fall back to another transport silently."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_peer_required_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    msg = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert "peer" in msg and msg != "no error"


def _peer_worker(rank, world, port, shape, n_iter, out, fail_attach_on):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from cmfpy_b200.dist import ShardedMultUpdate
        from tests.numpy_shard import NumpyPeerShard
        N, T, K, L = shape
        rng = np.random.default_rng(0)
        X, W0, H0 = rng.random((N, T)), rng.random((L, N, K)), rng.random((K, T))
        Tl = T // world
        t0 = rank * Tl
        eng = NumpyPeerShard(X[:, t0:min(T, t0 + Tl + L - 1)], N, T, K, L, t0, Tl, group=dist.group.WORLD,
                             fail_attach_on=fail_attach_on)
        alg = ShardedMultUpdate(None, N, T, K, L, t_offset=t0, t_local=Tl, initW=W0, initH=H0[:, t0:t0 + Tl],
                                group=dist.group.WORLD, engine=eng, tol=0, transport="auto")
        transport = alg.transport
        l0 = alg.loss
        losses = alg.update_many(n_iter - 1) + [alg.update()]
        gathered = [None] * world
        dist.all_gather_object(gathered, (alg.H_local_host(), alg.W_host(), [l0] + losses, transport, eng.attached))
        alg.close()
        if rank == 0:
            out.put((gathered, eng.attached))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("fail_attach_on", [None, 1])
def test_peer_transport_host_logic(fail_attach_on):
    """transport="auto" with an engine that offers peer memory: every rank exports a blob, all blobs are gathered,
    every rank attaches, and the iterations run through step_sharded.  When ONE rank cannot map its peers
    (fail_attach_on), ALL ranks must fall back to the collective transport together - a split decision would deadlock."""
    world, shape, n_iter = 3, (5, 90, 2, 9), 4
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_peer_worker, args=(r, world, port, shape, n_iter, q, fail_attach_on)) for r in range(world)]
    for p in procs:
        p.start()
    gathered, attached_after_close = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    N, T, K, L = shape
    rng = np.random.default_rng(0)
    X, W0, H0 = rng.random((N, T)), rng.random((L, N, K)), rng.random((K, T))
    ref = o.MultUpdateOracle(X, L, K, initW=W0, initH=H0, tol=0)
    ref_hist = [ref.loss] + [ref.update() for _ in range(n_iter)]
    want = "peer" if fail_attach_on is None else "nccl"
    for H, W, hist, transport, attached in gathered:
        assert transport == want and attached == (want == "peer")
        np.testing.assert_allclose(hist, ref_hist, rtol=1e-6)
        np.testing.assert_allclose(W, ref.W, rtol=1e-5, atol=1e-7)
    assert attached_after_close is False
    np.testing.assert_allclose(np.concatenate([g[0] for g in gathered], axis=1), ref.H, rtol=1e-5, atol=1e-7)


def test_tail_aware_shard_ranges():
    """shard_ranges / tail_handicap (algs/multi_gpu.py): contiguous cover of [0, T), the last shard shorter by about
    the estimated cost of the end-of-data corrections, every shard a multiple of 256 when T / G is, and no change
    where the estimate does not apply (direct denominators, one shard, shards too short to give anything up)."""
    from cmfpy_b200.algs.multi_gpu import shard_ranges, tail_handicap
    T = 1 << 20
    for G in (1, 2, 3, 4, 8):
        h = tail_handicap(1024, 32, 64, T, G, gram=True)
        r = shard_ranges(T, G, h)
        assert r[0][0] == 0 and r[-1][1] == T and all(a[1] == b[0] for a, b in zip(r, r[1:]))
        sizes = [b - a for a, b in r]
        if G == 1:
            assert h == 0 and sizes == [T]
            continue
        assert h == 7104                                   # 1.5 * 256 * 148 / 8 feature tiles
        assert len(set(sizes[:-1])) <= 2 and sizes[-1] < min(sizes[:-1])
        assert abs((sizes[0] - sizes[-1]) - h) <= 256 * G  # the others are `h` columns longer, up to the alignment
        if T % (G * 256) == 0:
            assert all(s % 256 == 0 for s in sizes)
    assert tail_handicap(1024, 32, 64, T, 8, gram=False) == 0
    assert shard_ranges(10, 3) == [(0, 4), (4, 7), (7, 10)]
    assert shard_ranges(1000, 4, 300) == [(0, 250), (250, 500), (500, 750), (750, 1000)]        # 300 / 4 < 256: none
    assert shard_ranges(4096, 2, 4096) == [(0, 2048), (2048, 4096)]                              # would take > 1/4

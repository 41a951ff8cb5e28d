"""world_size-2 check of the data-parallel contract on CPU (gloo): batch shards from
``shard_bounds`` + a sum all-reduce of dW reproduce the single-process full-batch result
(SURVEY.md §8e).  The oracle stands in for the kernels; the GPU path of the same logic
(mgb_allreduce_sum over NCCL) is exercised by bench.py --gpus N."""
import os
import socket

import numpy as np
import pytest

from mygrad_b200.dist import scale_for_mean_loss, shard_bounds
from oracle import conv_pool_oracle as oc

CFG = dict(stride=1, padding=1, dilation=1, pool=(2, 2), pool_stride=2)


def _inputs():
    rng = np.random.default_rng(0)
    return rng.standard_normal((7, 3, 8, 8)), rng.standard_normal((4, 3, 3, 3))


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, q):
    import torch
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        x, w = _inputs()
        lo, hi = shard_bounds(x.shape[0], rank, world)
        out, dx, dw = oc.layer_step(x[lo:hi], w, **CFG)
        t = torch.from_numpy(dw.copy())
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        # a mean-normalised loss divides by the local batch: rescale before the sum
        mean_piece = torch.from_numpy(dw / (hi - lo) * scale_for_mean_loss(hi - lo, x.shape[0]))
        dist.all_reduce(mean_piece, op=dist.ReduceOp.SUM)
        q.put((rank, lo, hi, out, dx, t.numpy(), mean_piece.numpy()))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_two_rank_shards_match_full_batch():
    import torch.multiprocessing as mp

    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted((q.get(timeout=150) for _ in range(world)), key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0

    x, w = _inputs()
    out, dx, dw = oc.layer_step(x, w, **CFG)
    assert [(r[1], r[2]) for r in results] == [(0, 4), (4, 7)]
    np.testing.assert_array_equal(np.concatenate([r[3] for r in results]), out)   # per-sample outputs
    np.testing.assert_allclose(np.concatenate([r[4] for r in results]), dx, rtol=1e-12, atol=1e-12)
    for r in results:
        np.testing.assert_allclose(r[5], dw, rtol=1e-10, atol=1e-10)              # summed dW on every rank
        np.testing.assert_allclose(r[6], dw / x.shape[0], rtol=1e-10, atol=1e-10)


# ---- cross-rank batchnorm moments: the algebra of mgb_batchnorm_merge (csrc/batchnorm.cu) ----------

def _merge_phases(moments, all_reduce):
    """NumPy restatement of the three phases around two sum all-reduces (include/mgb.h):
    moments = [mean, var, count] per channel of THIS rank; returns the full-batch [mean, var, count]"""
    mean, var, n = moments
    buf = np.concatenate([n, n * mean])                                   # phase 0
    buf = all_reduce(buf)
    gn, gsum = np.split(buf, 2)
    gmean = np.where(gn > 0, gsum / np.maximum(gn, 1), 0.0)               # phase 1
    m2 = var * n + n * (mean - gmean) ** 2
    m2 = all_reduce(m2)
    return gmean, np.where(gn > 0, m2 / np.maximum(gn, 1), 0.0), gn       # phase 2


def _bn_worker(rank: int, world: int, port: int, q):
    import torch
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        x = np.random.default_rng(3).standard_normal((7, 5, 4, 3)) * 2 + 10
        lo, hi = shard_bounds(x.shape[0], rank, world)
        xs = x[lo:hi]
        local = (xs.mean(axis=(0, 2, 3)), xs.var(axis=(0, 2, 3)), np.full(5, float(xs.size // 5)))

        def all_reduce(a):
            t = torch.from_numpy(np.ascontiguousarray(a))
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            return t.numpy()

        q.put((rank, *_merge_phases(local, all_reduce)))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_two_rank_batchnorm_moments_match_full_batch():
    import torch.multiprocessing as mp

    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_bn_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=150) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    x = np.random.default_rng(3).standard_normal((7, 5, 4, 3)) * 2 + 10
    for _, mean, var, n in results:
        np.testing.assert_allclose(mean, x.mean(axis=(0, 2, 3)), rtol=1e-13)
        np.testing.assert_allclose(var, x.var(axis=(0, 2, 3)), rtol=1e-11)
        np.testing.assert_array_equal(n, np.full(5, 7 * 12.0))


def test_collective_choice_by_message_size():
    """DataParallel.fuses(): peer-memory kernels for latency-sized messages only"""
    from mygrad_b200.dist import DataParallel

    dp = DataParallel(rank=0, world_size=8, local_rank=0)
    assert not dp.fuses(1024)                       # no peer group connected -> NCCL
    dp.peer_slot_bytes = 8 << 20
    assert dp.fuses(128 * 64 * 9 * 4)               # dW of config 2 (295 KB)
    assert not dp.fuses(256 * 128 * 25 * 8)         # dW of config 3 (6.5 MB): one-shot loses, NCCL
    assert not dp.fuses(0)
    assert not DataParallel(rank=0, world_size=1, local_rank=0).fuses(1024)
    dp.peer_slot_bytes = 0                          # do not let close() touch a device

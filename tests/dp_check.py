"""Multi-GPU data-parallel parity check, run under torchrun (one process per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29611 tests/dp_check.py

Every rank holds the full seeded batch for CHECKING only; it computes on its own shard.  The result
to match is the single-process oracle on the full batch: y / pooled / dX equal to the
oracle's slices, dW (fused split-K + cross-rank sum over peer memory, and the plain all-reduce path)
equal to the full-batch dW and bit-identical on all ranks, batchnorm with cross-rank statistics equal
to the full-batch batchnorm.  Prints "DP_CHECK OK" on rank 0.
"""
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import torch
import torch.distributed as dist

import mygrad_b200 as mg
from conftest import assert_close
from mygrad_b200 import Tensor
from mygrad_b200.dist import DataParallel
from mygrad_b200.nnet.layers import batchnorm, conv_nd, max_pool
from oracle import conv_pool_oracle as oc


def same_on_all_ranks(arr: np.ndarray, what: str):
    box = [None] * dist.get_world_size()
    dist.all_gather_object(box, arr.tobytes())
    assert all(b == box[0] for b in box), f"{what}: ranks disagree bitwise"


def world_gt1(dp) -> bool:
    return dp.world_size > 1


def conv_case(dp, dtype, n_total, fused: bool, channels=(16, 32)):
    """channels = (C, F): (16, 32) runs the generic MMA tiles, (3, 32) the thin-layer kernels (conv_thin.cuh),
    (64, 128) the tensor-core kernels (row-halo wgrad / TMA-fed float64) — each with its own split-K partials that
    the fused peer-memory kernel has to sum across ranks"""
    rank, world = dp.rank, dp.world_size
    rng = np.random.default_rng(11)
    x = rng.standard_normal((n_total, channels[0], 10, 10)).astype(dtype)
    w = rng.standard_normal((channels[1], channels[0], 3, 3)).astype(dtype)
    lo, hi = dp.shard(n_total)
    xt, wt = Tensor(x[lo:hi]), Tensor(w)
    y = conv_nd(xt, wt, stride=1, padding=1)
    out = max_pool(y, (2, 2), 2)
    if fused:
        with dp.fused_wgrad():
            out.backward()
        assert dp.fuses(wt.grad.nbytes), f"peer group not connected: {dp.peer_error}"
        assert wt.grad.reduced, "the fused wgrad must mark dW as already summed over the ranks"
        dp.allreduce([wt.grad])      # must be a no-op for a reduced buffer (the library tracks it, not the caller)
    else:
        out.backward()
        dp.allreduce([wt.grad])
    want_out, want_dx, want_dw = oc.layer_step(x.astype(np.float64), w.astype(np.float64), stride=1, padding=1,
                                               dilation=1, pool=(2, 2), pool_stride=2)
    tag = f"{np.dtype(dtype).name} N={n_total} C,F={channels} fused={fused} rank={rank}"
    if hi > lo:
        assert_close(out.numpy(), want_out[lo:hi], dtype, err_msg=f"out {tag}")
        assert_close(xt.grad.get(), want_dx[lo:hi], dtype, err_msg=f"dx {tag}")
    dw = wt.grad.get()
    assert_close(dw, want_dw, dtype, err_msg=f"dw {tag}")
    same_on_all_ranks(dw, f"dw {tag}")


def batchnorm_case(dp, dtype):
    rng = np.random.default_rng(12)
    n_total = 9
    x = (rng.standard_normal((n_total, 5, 6, 7)) * 2 + 1).astype(dtype)
    g = rng.standard_normal(x.shape).astype(dtype)
    gamma = (rng.standard_normal(5) + 1).astype(dtype)
    beta = rng.standard_normal(5).astype(dtype)
    lo, hi = dp.shard(n_total)
    xt, gt, bt = Tensor(x[lo:hi]), Tensor(gamma), Tensor(beta)
    y = batchnorm(xt, gamma=gt, beta=bt, eps=1e-5, dp=dp)
    y.backward(g[lo:hi])
    if world_gt1(dp):
        assert gt.grad.reduced and bt.grad.reduced and not xt.grad.reduced
    dp.allreduce([gt.grad, bt.grad])   # already global: must not be multiplied by world_size
    x64, g64 = x.astype(np.float64), g.astype(np.float64)
    want_y, _, _ = oc.batchnorm_fwd(x64, gamma, beta, 1e-5)
    want_dx, want_dg, want_db = oc.batchnorm_bwd(g64, x64, gamma, 1e-5)
    tag = f"batchnorm {np.dtype(dtype).name} rank={dp.rank}"
    assert_close(y.numpy(), want_y[lo:hi], dtype, scale=float(np.abs(want_y).max()) * 4, err_msg=tag + " y")
    assert_close(xt.grad.get(), want_dx[lo:hi], dtype, scale=float(np.abs(want_dx).max()) * 50, err_msg=tag + " dx")
    n = x.size / 5
    assert_close(gt.grad.get(), want_dg, dtype, scale=float(np.abs(g).max()) * n, err_msg=tag + " dgamma")
    assert_close(bt.grad.get(), want_db, dtype, scale=float(np.abs(g).max()) * n, err_msg=tag + " dbeta")
    same_on_all_ranks(gt.grad.get(), tag + " dgamma")


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dp = DataParallel(rank, world, local).init()
    if rank == 0:
        print(f"peer slot bytes: {dp.peer_slot_bytes} ({dp.peer_error})", flush=True)
    for dtype in (np.float32, np.float64):
        for n_total in (8, 3, 1):                  # even shards, uneven shards, an EMPTY shard on the last ranks
            for fused in ((True, False) if dp.peer_slot_bytes else (False,)):
                conv_case(dp, dtype, n_total, fused)
        for channels in ((3, 32), (64, 128)):
            for fused in ((True, False) if dp.peer_slot_bytes else (False,)):
                conv_case(dp, dtype, 5, fused, channels)
        batchnorm_case(dp, dtype)
    # many back-to-back collectives: slot parity / epoch protocol
    buf = mg.asdevice(np.full(1000, float(rank + 1), np.float32))
    for _ in range(64):
        dp.allreduce([buf])
        buf *= 1.0 / world
    mg.synchronize()
    expect = np.float32(sum(range(1, world + 1)) / world)
    got = buf.get()
    assert np.allclose(got[:1], got) and abs(float(got[0]) - float(expect)) < 1e-3 * float(expect) + 1.0, got[:4]
    dist.barrier()
    if rank == 0:
        print("DP_CHECK OK", flush=True)
    dp.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""Latency of the two parameter-gradient collectives, under torchrun (one process per GPU):
the one-kernel peer-memory all-reduce (mgb_peer_allreduce_sum) vs ncclAllReduce (mgb_allreduce_sum).
Prints one JSON line on rank 0: microseconds per call (device time, max over ranks)."""
import ctypes as C, json, os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch
import torch.distributed as dist
import mygrad_b200 as mg
from mygrad_b200._lib import check, lib
from mygrad_b200.dist import DataParallel

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dp = DataParallel(rank, world, local).init()
st = C.c_void_p(mg.current_stream())
out = {"n_gpus": world, "peer_slot_bytes": dp.peer_slot_bytes, "sizes": {}}
for name, n, dt in (("dW C2 f32 295 KB", 128 * 64 * 9, np.float32), ("dW C3 f64 6.5 MB", 256 * 128 * 25, np.float64),
                    ("bias 128 f32", 128, np.float32)):
    buf = mg.asdevice(np.ones(n, dt))
    res = {}
    for kind in ("peer", "nccl"):
        if kind == "peer" and not dp.peer_slot_bytes:
            continue
        fn = (lambda: check(lib.mgb_peer_allreduce_sum(C.c_void_p(buf.ptr), C.c_size_t(n), buf.code, st))) if kind == "peer" \
            else (lambda: check(lib.mgb_allreduce_sum(C.c_void_p(buf.ptr), C.c_size_t(n), buf.code, st)))
        for _ in range(10):
            fn(); buf *= 1.0 / world
        mg.synchronize(); dist.barrier(); mg.synchronize()
        a, b = C.c_void_p(), C.c_void_p()
        check(lib.mgb_event_create(C.byref(a))); check(lib.mgb_event_create(C.byref(b)))
        check(lib.mgb_event_record(a, st))
        K = 100
        for _ in range(K):
            fn()
        check(lib.mgb_event_record(b, st)); check(lib.mgb_event_synchronize(b))
        ms = C.c_float(0); check(lib.mgb_event_elapsed_ms(a, b, C.byref(ms)))
        t = torch.tensor([ms.value / K * 1e3], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res[kind + "_us"] = round(float(t.item()), 2)
        buf.fill(1.0)
    out["sizes"][name] = res
if rank == 0:
    print(json.dumps(out), flush=True)
dist.barrier()
dp.close()
dist.destroy_process_group()

"""Pinned-memory PCIe bandwidth of the box (the ceiling of bench.py's e2e number): H2D alone, D2H alone,
both directions at once.  usage: python scripts/pcie_bw.py [MB]"""
import ctypes as C, json, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import mygrad_b200 as mg
from mygrad_b200._lib import check, lib

mb = int(sys.argv[1]) if len(sys.argv) > 1 else 256
n = mb * (1 << 20) // 4
h_in, h_out = mg.pinned_empty((n,), np.float32), mg.pinned_empty((n,), np.float32)
h_in[...] = 1.0
d_in, d_out = mg.empty((n,), np.float32), mg.full((n,), 2.0, np.float32)
s1, s2 = C.c_void_p(), C.c_void_p()
check(lib.mgb_stream_create(C.byref(s1))); check(lib.mgb_stream_create(C.byref(s2)))
nbytes = C.c_size_t(n * 4)

def h2d(): check(lib.mgb_memcpy_h2d(C.c_void_p(d_in.ptr), C.c_void_p(h_in.ctypes.data), nbytes, s1))
def d2h(): check(lib.mgb_memcpy_d2h(C.c_void_p(h_out.ctypes.data), C.c_void_p(d_out.ptr), nbytes, s2))
def sync(): check(lib.mgb_stream_synchronize(s1)); check(lib.mgb_stream_synchronize(s2))

def timed(fns, reps=5):
    for f in fns: f()
    sync()
    t0 = time.perf_counter()
    for _ in range(reps):
        for f in fns: f()
    sync()
    return (time.perf_counter() - t0) / reps

out = {"mb": mb}
out["h2d_gbs"] = n * 4 / timed([h2d]) / 1e9
out["d2h_gbs"] = n * 4 / timed([d2h]) / 1e9
t = timed([h2d, d2h])
out["bidir_each_gbs"] = n * 4 / t / 1e9
print(json.dumps({k: round(v, 2) if isinstance(v, float) else v for k, v in out.items()}))

"""Time conv fwd / dgrad kernels alone under several environment settings, in one process.
usage: python scripts/phase_ab.py "VAR=a,VAR2=b" "VAR=c" ...   (each argument is one setting)"""
import ctypes as C, os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import bench
import mygrad_b200 as mg
from mygrad_b200._lib import check, lib, ConvDesc

name = os.environ.get("WORKLOAD", "C2")
cfg = bench.WORKLOADS[name]
geo = bench.geometry(cfg)
dt = bench.NP_DTYPE[cfg["dtype"]]
xh, wh = bench.make_inputs(cfg)
x, w = mg.asdevice(xh), mg.asdevice(wh)
nd = len(cfg["x"]) - 2
tup = lambda v: (v,) * nd if isinstance(v, int) else tuple(v)
cd = ConvDesc()
cd.dtype, cd.nd = (0 if cfg["dtype"] == "f32" else 1), nd
cd.N, cd.C, cd.F = cfg["x"][0], cfg["x"][1], cfg["w"][0]
for i in range(nd):
    cd.X[i], cd.W[i] = cfg["x"][2 + i], cfg["w"][2 + i]
    cd.stride[i], cd.pad[i], cd.dil[i] = tup(cfg["stride"])[i], tup(cfg["padding"])[i], tup(cfg["dilation"])[i]
y = mg.empty((cfg["x"][0], cfg["w"][0], *geo["G"]), dt)
gy = mg.asdevice(np.random.default_rng(2).standard_normal(y.shape).astype(dt))   # seed-2 normal upstream grad
dx, dw = mg.empty(cfg["x"], dt), mg.empty(cfg["w"], dt)
wss = []
for which in range(3):
    n = C.c_size_t(0)
    check(lib.mgb_conv_workspace_bytes(C.byref(cd), which, C.byref(n)))
    wss.append((mg.empty((max(8, n.value) + 7) // 8, np.float64), n.value))
P = lambda a: C.c_void_p(a.ptr)
stream = C.c_void_p(mg.current_stream())

def timed(fn, reps=10):
    fn(); fn()
    a, b = C.c_void_p(), C.c_void_p()
    check(lib.mgb_event_create(C.byref(a))); check(lib.mgb_event_create(C.byref(b)))
    check(lib.mgb_event_record(a, stream))
    for _ in range(reps):
        fn()
    check(lib.mgb_event_record(b, stream))
    ms = C.c_float(0)
    check(lib.mgb_event_synchronize(b)); check(lib.mgb_event_elapsed_ms(a, b, C.byref(ms)))
    return ms.value / reps

fwd = lambda: check(lib.mgb_conv_fwd(C.byref(cd), P(x), P(w), P(y), P(wss[0][0]), wss[0][1], stream))
dgr = lambda: check(lib.mgb_conv_dgrad(C.byref(cd), P(gy), P(w), P(dx), P(wss[1][0]), wss[1][1], stream))
wgr = lambda: check(lib.mgb_conv_wgrad(C.byref(cd), P(gy), P(x), P(dw), P(wss[2][0]), wss[2][1], stream))
if os.environ.get("ONESHOT"):          # for ncu: one fwd + one dgrad per setting, nothing else
    for setting in sys.argv[1:]:
        kv = dict(s.split("=") for s in setting.split(",") if s)
        os.environ.update(kv)
        fwd(); dgr(); mg.synchronize()
        for k in kv:
            os.environ.pop(k, None)
    sys.exit(0)
for rep in range(2):
    for setting in sys.argv[1:]:
        kv = dict(s.split("=") for s in setting.split(",") if s)
        for k, v in kv.items():
            os.environ[k] = v
        print(f"{name} [{setting}] fwd {timed(fwd):.4f} ms  dgrad {timed(dgr):.4f} ms  wgrad {timed(wgr):.4f} ms (incl. plane prep)", flush=True)
        for k in kv:
            os.environ.pop(k, None)

"""A/B a library environment switch inside ONE process (same box, same clocks):
usage: python scripts/step_ab.py ENVVAR valA valB [workload] [steps]"""
import os, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench
import mygrad_b200 as mg
from mygrad_b200 import Tensor
from mygrad_b200.nnet.layers import conv_nd, max_pool

var, va, vb = sys.argv[1:4]
name = sys.argv[4] if len(sys.argv) > 4 else "C2"
K = int(sys.argv[5]) if len(sys.argv) > 5 else 30
cfg = bench.WORKLOADS[name]
xh, wh = bench.make_inputs(cfg)
x, w = mg.asdevice(xh), mg.asdevice(wh)
kw = dict(stride=cfg["stride"], padding=cfg["padding"], dilation=cfg["dilation"])

def step():
    xt, wt = Tensor(x, copy=False), Tensor(w, copy=False)
    y = conv_nd(xt, wt, **kw)
    o = max_pool(y, cfg["pool"], 2) if cfg["pool"] else y
    o.backward()

def run(v):
    os.environ[var] = v
    for _ in range(3):
        step()
    mg.synchronize()
    t0 = time.perf_counter()
    for _ in range(K):
        step()
    mg.synchronize()
    return 1e3 * (time.perf_counter() - t0) / K

for rep in range(3):
    a = run(va); b = run(vb)
    print(f"{name} {var}={va}: {a:.3f} ms/step   {var}={vb}: {b:.3f} ms/step")

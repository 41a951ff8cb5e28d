"""How long does the HOST take to issue one C2 step (no sync) vs the device time of the step?
usage: python scripts/step_issue_time.py [workload] [steps]"""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench
import mygrad_b200 as mg
from mygrad_b200 import Tensor
from mygrad_b200.nnet.layers import conv_nd, max_pool

name = sys.argv[1] if len(sys.argv) > 1 else "C2"
K = int(sys.argv[2]) if len(sys.argv) > 2 else 30
cfg = bench.WORKLOADS[name]
xh, wh = bench.make_inputs(cfg)
x, w = mg.asdevice(xh), mg.asdevice(wh)
kw = dict(stride=cfg["stride"], padding=cfg["padding"], dilation=cfg["dilation"])

def step():
    xt, wt = Tensor(x, copy=False), Tensor(w, copy=False)
    y = conv_nd(xt, wt, **kw)
    o = max_pool(y, cfg["pool"], 2) if cfg["pool"] else y
    o.backward()

for _ in range(3):
    step()
mg.synchronize()
t0 = time.perf_counter()
for _ in range(K):
    step()
t1 = time.perf_counter()
mg.synchronize()
t2 = time.perf_counter()
print(f"{name}: host issue {1e3*(t1-t0)/K:.3f} ms/step, total {1e3*(t2-t0)/K:.3f} ms/step")
# one step, fully synchronous per call, to see the host cost alone
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(K):
    step()
pr.disable(); mg.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(25)

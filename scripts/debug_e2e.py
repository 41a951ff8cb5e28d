import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mygrad_b200 as mg
from mygrad_b200 import Tensor
from mygrad_b200.nnet.layers import conv_nd, max_pool
dt = np.float32
xs, wsh = (256, 64, 56, 56), (128, 64, 3, 3)
px, pw = mg.pinned_empty(xs, dt), mg.pinned_empty(wsh, dt)
px[...] = np.random.default_rng(0).standard_normal(xs).astype(dt); pw[...] = np.random.default_rng(1).standard_normal(wsh).astype(dt)
xd, wd = mg.empty(xs, dt), mg.empty(wsh, dt)
po = mg.pinned_empty((256, 128, 28, 28), dt); pdx = mg.pinned_empty(xs, dt); pdw = mg.pinned_empty(wsh, dt)
def T(): mg.synchronize(); return time.perf_counter()
for it in range(6):
    t0 = T()
    xd.set_from_host(px, sync=False); wd.set_from_host(pw, sync=False)
    t1 = T()
    xt, wt = Tensor(xd, copy=False), Tensor(wd, copy=False)
    o = max_pool(conv_nd(xt, wt, stride=1, padding=1), (2, 2), 2)
    t2 = T()
    o.backward()
    t3 = T()
    o.data.get(po, sync=False); xt.grad.get(pdx, sync=False); wt.grad.get(pdw, sync=False)
    t4 = T()
    print(f"it{it}: h2d {1e3*(t1-t0):.2f} fwd {1e3*(t2-t1):.2f} bwd {1e3*(t3-t2):.2f} d2h {1e3*(t4-t3):.2f} ms")
# without intermediate syncs
for it in range(4):
    t0 = T()
    xd.set_from_host(px, sync=False); wd.set_from_host(pw, sync=False)
    xt, wt = Tensor(xd, copy=False), Tensor(wd, copy=False)
    o = max_pool(conv_nd(xt, wt, stride=1, padding=1), (2, 2), 2)
    o.backward()
    th = time.perf_counter()
    o.data.get(po, sync=False); xt.grad.get(pdx, sync=False); wt.grad.get(pdw, sync=False)
    t4 = T()
    print(f"nosync it{it}: host-issue {1e3*(th-t0):.2f} total {1e3*(t4-t0):.2f} ms")

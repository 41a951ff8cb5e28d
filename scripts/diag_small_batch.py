"""bisect: C2-geometry layer at small batch vs the oracle (dX / dW errors seen at N = 2)"""
import os, sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import mygrad_b200 as mg
from mygrad_b200 import Tensor
from mygrad_b200.nnet.layers import conv_nd, max_pool
from oracle import conv_pool_oracle as oc

def run(n, env):
    for k in ("MGB_CONV_HALO", "MGB_CONV_F32_PATH", "MGB_WGRAD_F32_PATH"):
        os.environ.pop(k, None)
    os.environ.update(env)
    x = np.random.default_rng(7).standard_normal((n, 64, 56, 56)).astype(np.float32)
    w = np.random.default_rng(1).standard_normal((128, 64, 3, 3)).astype(np.float32)
    xt, wt = Tensor(x), Tensor(w)
    y = conv_nd(xt, wt, stride=1, padding=1)
    o = max_pool(y, (2, 2), 2)
    o.backward()
    want_o, want_dx, want_dw = oc.layer_step(x, w, stride=1, padding=1, dilation=1, pool=(2, 2), pool_stride=2)
    dy = oc.max_pool_bwd(np.ones(o.shape, np.float32), y.numpy(), (2, 2), 2).astype(np.float32)
    e = lambda a, b: float(np.abs(a.astype(np.float64) - b).max() / np.abs(b).max())
    print(f"N={n} {env}: out {e(o.numpy(), want_o):.2e} dy-exact {np.array_equal(y.grad.get(), dy)} "
          f"dx {e(xt.grad.get(), want_dx):.2e} dw {e(wt.grad.get(), want_dw):.2e}", flush=True)
    dxg = xt.grad.get()
    bad = np.argwhere(np.abs(dxg - want_dx) > 1e-3 * np.abs(want_dx).max())
    if len(bad):
        print("   bad dx entries:", len(bad), "first", bad[:3].tolist(), "last", bad[-3:].tolist(),
              "samples", sorted(set(bad[:, 0].tolist())), "rows", sorted(set(bad[:, 2].tolist()))[:12], "cols", sorted(set(bad[:, 3].tolist()))[:12])

for n in (2, 1, 3, 4, 8):
    run(n, {})
run(2, {"MGB_CONV_HALO": "0"})
run(2, {"MGB_CONV_F32_PATH": "cpasync"})
run(2, {"MGB_CONV_F32_PATH": "mma"})
run(2, {"MGB_WGRAD_F32_PATH": "mma"})

"""Random conv_nd shapes against the NumPy oracle: forward, dX and dW, float32 and float64, 1-3 spatial dims,
random stride / padding / dilation, channel counts on both sides of every kernel-family threshold (thin layers,
mma.sync tiles, tcgen05 halo / row-halo / folded kernels, TMA-fed DMMA).  usage: python scripts/fuzz_conv.py [cases] [seed] [big]"""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np

import mygrad_b200 as mg
from mygrad_b200 import Tensor
from mygrad_b200.nnet.layers import conv_nd
from oracle import conv_pool_oracle as oc

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 200
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
big = len(sys.argv) > 3 and sys.argv[3] == "big"       # larger extents: more tiles, more K splits, ragged boxes
rng = np.random.default_rng(seed)
CH = [1, 2, 3, 5, 8, 16, 24, 32, 48, 64, 96]
FS = [1, 4, 16, 32, 33, 48, 64, 96, 100, 128, 160]
bad = 0
ran = 0
t0 = time.time()
for it in range(cases):
    nd = int(rng.choice([1, 2, 2, 2, 3]))
    dtype = np.float32 if rng.random() < 0.6 else np.float64
    C, F = int(rng.choice(CH)), int(rng.choice(FS))
    N = int(rng.integers(1, 6))
    W = tuple(int(rng.integers(1, 4 if nd > 1 else 8)) for _ in range(nd))
    s = tuple(int(rng.choice([1, 1, 1, 2, 3])) for _ in range(nd))
    d = tuple(int(rng.choice([1, 1, 1, 2])) for _ in range(nd))
    p = tuple(int(rng.integers(0, 3)) for _ in range(nd))
    # output extent first, input extent from it: always a valid geometry
    hi = {1: 600, 2: 90, 3: 22}[nd] if big else {1: 200, 2: 40, 3: 12}[nd]
    G = tuple(int(rng.integers(1, hi)) for _ in range(nd))
    X = tuple((g - 1) * si + (w - 1) * di + 1 - 2 * pi for g, si, w, di, pi in zip(G, s, W, d, p))
    if any(x < 1 for x in X) or N * C * int(np.prod(X)) > (2e7 if big else 4e6) or N * F * int(np.prod(G)) > (2e7 if big else 4e6):
        continue
    ran += 1
    x = rng.standard_normal((N, C, *X)).astype(dtype)
    w = rng.standard_normal((F, C, *W)).astype(dtype)
    xt, wt = Tensor(x), Tensor(w)
    y = conv_nd(xt, wt, stride=s, padding=p, dilation=d)
    x64, w64 = x.astype(np.float64), w.astype(np.float64)
    want = oc.conv_fwd(x64, w64, s, p, d)
    g = rng.standard_normal(want.shape).astype(dtype)
    y.backward(g)
    g64 = g.astype(np.float64)
    rtol = 1e-5 if dtype == np.float32 else 1e-10
    for name, got, ref in (("fwd", y.numpy(), want), ("dx", xt.grad.get(), oc.conv_dx(g64, x.shape, w64, s, p, d)),
                           ("dw", wt.grad.get(), oc.conv_dw(g64, x64, w.shape, s, p, d))):
        scale = float(np.abs(ref).max()) or 1.0
        err = float(np.abs(got - ref).max()) / scale
        # the test suite's criterion (tests/conftest.py assert_close): |got - ref| <= rtol * max|ref| + rtol * |ref|
        ok = got.shape == ref.shape and np.isfinite(err) and bool(np.all(np.abs(got - ref) <= rtol * scale + rtol * np.abs(ref)))
        if not ok:
            bad += 1
            print(f"BAD {name} err {err:.2e} {np.dtype(dtype).name} x{x.shape} w{w.shape} s{s} p{p} d{d}", flush=True)
print(f"FUZZ {'OK' if bad == 0 else 'FAILED'}: {ran} of {cases} cases ran, {bad} mismatches, seed {seed}, {time.time() - t0:.1f} s")
sys.exit(1 if bad else 0)

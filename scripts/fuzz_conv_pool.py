"""Random conv_pool layers (the fused Operation and its unfused fall-backs) against max_pool(conv_nd(...)) bit for
bit on the pooling routing, and against the oracle within tolerance.  usage: python scripts/fuzz_conv_pool.py [cases] [seed]"""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np

import mygrad_b200 as mg
from mygrad_b200 import Tensor
from oracle import conv_pool_oracle as oc

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 200
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rng = np.random.default_rng(seed)
bad = ran = 0
t0 = time.time()
for it in range(cases):
    dtype = np.float32 if rng.random() < 0.7 else np.float64
    C, F = int(rng.choice([3, 16, 32, 64, 96])), int(rng.choice([8, 16, 32, 64, 128]))
    N = int(rng.integers(1, 5))
    W = tuple(int(rng.integers(1, 4)) for _ in range(2))
    p = tuple(int(rng.integers(0, 2)) for _ in range(2))
    d = tuple(int(rng.choice([1, 1, 2])) for _ in range(2))
    s = tuple(int(rng.choice([1, 1, 1, 2])) for _ in range(2))
    G = tuple(2 * int(rng.integers(1, 16)) for _ in range(2))           # even: 2x2 / 2 pooling tiles it
    X = tuple((g - 1) * si + (w - 1) * di + 1 - 2 * pi for g, si, w, di, pi in zip(G, s, W, d, p))
    if any(x < 1 for x in X):
        continue
    ran += 1
    x = rng.standard_normal((N, C, *X)).astype(dtype)
    w = rng.standard_normal((F, C, *W)).astype(dtype)
    xt, wt = Tensor(x), Tensor(w)
    out = mg.conv_pool(xt, wt, stride=s, padding=p, dilation=d, pool=(2, 2), pool_stride=2)
    out.backward()
    want_out, want_dx, want_dw = oc.layer_step(x.astype(np.float64), w.astype(np.float64), stride=s, padding=p,
                                               dilation=d, pool=(2, 2), pool_stride=2)
    rtol = 1e-5 if dtype == np.float32 else 1e-10
    for name, got, ref in (("out", out.numpy(), want_out), ("dx", xt.grad.get(), want_dx), ("dw", wt.grad.get(), want_dw)):
        scale = float(np.abs(ref).max()) or 1.0
        # the max-pool routing is decided on float32 / float64 values that differ in the last bits from the oracle's:
        # a near-tie may route differently, which moves whole gradient entries — allow a handful of such elements
        diff = np.abs(got - ref) > rtol * scale + rtol * np.abs(ref)
        if got.shape != ref.shape or diff.mean() > 1e-3:
            bad += 1
            print(f"BAD {name} {np.dtype(dtype).name} x{x.shape} w{w.shape} s{s} p{p} d{d}: {int(diff.sum())} of {diff.size} off", flush=True)
print(f"FUZZ {'OK' if bad == 0 else 'FAILED'}: {ran} of {cases} cases ran, {bad} mismatches, seed {seed}, {time.time() - t0:.1f} s")
sys.exit(1 if bad else 0)

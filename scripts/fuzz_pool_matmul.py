"""Random max_pool geometries (bit-exact against the oracle, forward and backward, ties included) and random
skinny / general matmuls (classifier heads) with gradients.  usage: python scripts/fuzz_pool_matmul.py [cases] [seed]"""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np

import mygrad_b200 as mg
from mygrad_b200 import Tensor
from mygrad_b200.nnet.layers import max_pool
from oracle import conv_pool_oracle as oc

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 300
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rng = np.random.default_rng(seed)
bad = 0
t0 = time.time()
for it in range(cases):
    nd = int(rng.integers(1, 4))
    dtype = np.float32 if rng.random() < 0.5 else np.float64
    exact = rng.random() < 0.7                       # exact tilings (stride == pool) take the window-centric kernels
    pool = tuple(int(rng.choice([1, 2, 2, 2, 3])) for _ in range(nd))
    stride = pool if exact else tuple(int(rng.integers(1, 4)) for _ in range(nd))
    hi = {1: 300, 2: 70, 3: 20}[nd]
    G = tuple(int(rng.integers(1, hi)) for _ in range(nd))
    X = tuple((g - 1) * s + p for g, s, p in zip(G, stride, pool))
    lead = tuple(int(rng.integers(1, 5)) for _ in range(int(rng.integers(0, 3))))
    if int(np.prod(lead + X)) > 3e6:
        continue
    x = (rng.integers(-2, 3, size=(*lead, *X)) if rng.random() < 0.4 else rng.standard_normal((*lead, *X))).astype(dtype)
    xt = Tensor(x)
    y = max_pool(xt, pool, stride)
    g = rng.standard_normal(y.shape).astype(dtype)
    y.backward(g)
    ok = np.array_equal(y.numpy(), oc.max_pool_fwd(x, pool, stride)) and \
        np.array_equal(xt.grad.get(), oc.max_pool_bwd(g, x, pool, stride).astype(dtype))
    if not ok:
        bad += 1
        print(f"BAD pool {np.dtype(dtype).name} x{x.shape} pool{pool} stride{stride}", flush=True)
for it in range(cases // 3):
    dtype = np.float32 if rng.random() < 0.6 else np.float64
    n, d, m = int(rng.integers(1, 300)), int(rng.integers(1, 3000)), int(rng.choice([1, 2, 3, 5, 8, 10, 16, 17, 31, 32, 33, 64]))
    a, b = rng.standard_normal((n, d)).astype(dtype), rng.standard_normal((d, m)).astype(dtype)
    at, bt = Tensor(a), Tensor(b)
    c = mg.matmul(at, bt)
    g = rng.standard_normal((n, m)).astype(dtype)
    c.backward(g)
    a64, b64, g64 = a.astype(np.float64), b.astype(np.float64), g.astype(np.float64)
    rtol = 1e-5 if dtype == np.float32 else 1e-10
    for name, got, ref in (("c", c.numpy(), a64 @ b64), ("da", at.grad.get(), g64 @ b64.T), ("db", bt.grad.get(), a64.T @ g64)):
        scale = float(np.abs(ref).max()) or 1.0
        if got.shape != ref.shape or not bool(np.all(np.abs(got - ref) <= rtol * scale + rtol * np.abs(ref))):
            bad += 1
            print(f"BAD matmul {name} {np.dtype(dtype).name} n={n} d={d} m={m} err {np.abs(got - ref).max() / scale:.2e}", flush=True)
print(f"FUZZ {'OK' if bad == 0 else 'FAILED'}: {cases} pool + {cases // 3} matmul cases, {bad} mismatches, seed {seed}, {time.time() - t0:.1f} s")
sys.exit(1 if bad else 0)

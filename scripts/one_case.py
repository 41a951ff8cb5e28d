"""one conv_nd case against the oracle under several environment settings (debugging aid)
usage: python scripts/one_case.py "VAR=a,VAR2=b" "VAR=c" ...   (the case is edited below)"""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import mygrad_b200 as mg
from mygrad_b200 import Tensor
from mygrad_b200.nnet.layers import conv_nd
from oracle import conv_pool_oracle as oc

xs, ws, s, p, d, dtype = (4, 96, 2, 13, 19), (160, 96, 3, 3, 3), (1, 1, 1), (2, 0, 1), (1, 2, 2), np.float32
rng = np.random.default_rng(0)
x = rng.standard_normal(xs).astype(dtype); w = rng.standard_normal(ws).astype(dtype)
x64, w64 = x.astype(np.float64), w.astype(np.float64)
want = oc.conv_fwd(x64, w64, s, p, d)
g = rng.standard_normal(want.shape).astype(dtype); g64 = g.astype(np.float64)
refs = {"fwd": want, "dx": oc.conv_dx(g64, x.shape, w64, s, p, d), "dw": oc.conv_dw(g64, x64, w.shape, s, p, d)}
for setting in sys.argv[1:] or [""]:
    kv = dict(t.split("=") for t in setting.split(",") if t)
    os.environ.update(kv)
    xt, wt = Tensor(x), Tensor(w)
    y = conv_nd(xt, wt, stride=s, padding=p, dilation=d)
    y.backward(g)
    got = {"fwd": y.numpy(), "dx": xt.grad.get(), "dw": wt.grad.get()}
    print(setting or "(default)", {k: f"{np.abs(got[k] - refs[k]).max() / np.abs(refs[k]).max():.2e}" for k in refs}, flush=True)
    for k in kv: os.environ.pop(k, None)

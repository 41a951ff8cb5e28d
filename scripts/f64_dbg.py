import os, sys
sys.path.insert(0, '/root/repo')
import numpy as np
import mygrad_b200 as mg
from mygrad_b200.nnet.layers import conv_nd
from oracle import conv_pool_oracle as oc
rng = np.random.default_rng(0)
xs, ws = (2, 64, 16, 16), (128, 64, 3, 3)
xh = rng.standard_normal(xs); wh = rng.standard_normal(ws)
want = oc.conv_fwd(xh, wh, 1, 1, 1)
for env in (sys.argv[1:] or [""]):
    for kv in env.split(","):
        if kv: k, v = kv.split("="); os.environ[k] = v
    try:
        y = conv_nd(mg.Tensor(xh, constant=True), mg.Tensor(wh, constant=True), stride=1, padding=1)
        got = y.numpy()
        print(env, "max err", np.abs(got - want).max() / np.abs(want).max(), flush=True)
    except Exception as e:
        print(env, "FAILED", str(e)[-200:], flush=True)
        break

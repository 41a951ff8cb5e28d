"""debug helper: raw mgb_conv_wgrad call, dumps the result and the workspace"""
import ctypes as C, sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mygrad_b200 as mg
from mygrad_b200._lib import ConvDesc, lib, check

N, Cc, F, X, W = 2, 32, 32, (8, 8), (1, 1)
d = ConvDesc(); d.dtype, d.nd, d.N, d.C, d.F = 0, 2, N, Cc, F
for i in range(2):
    d.X[i], d.W[i], d.stride[i], d.pad[i], d.dil[i] = X[i], W[i], 1, 0, 1
rng = np.random.default_rng(0)
mode = sys.argv[1] if len(sys.argv) > 1 else "ones"
if mode == "ones":
    x = np.ones((N, Cc, *X), np.float32); dy = np.ones((N, F, *X), np.float32)
elif mode == "ramp":
    x = np.ones((N, Cc, *X), np.float32) * (np.arange(Cc, dtype=np.float32)[None, :, None, None] + 1)
    dy = np.ones((N, F, *X), np.float32) * (np.arange(F, dtype=np.float32)[None, :, None, None] * 100)
else:
    x = rng.standard_normal((N, Cc, *X)).astype(np.float32); dy = rng.standard_normal((N, F, *X)).astype(np.float32)
xd, dyd = mg.asdevice(x), mg.asdevice(dy)
dw = mg.full((F, Cc, *W), -7.0, np.float32)
n = C.c_size_t(0)
check(lib.mgb_conv_workspace_bytes(C.byref(d), 2, C.byref(n)))
ws = mg.full(((n.value + 3) // 4,), -3.0, np.float32)
check(lib.mgb_conv_wgrad(C.byref(d), C.c_void_p(dyd.ptr), C.c_void_p(xd.ptr), C.c_void_p(dw.ptr), C.c_void_p(ws.ptr), n, None))
mg.synchronize()
got = dw.get()
want = np.einsum("nfhw,nchw->fc", dy.astype(np.float64), x.astype(np.float64)).reshape(F, Cc, 1, 1)
print("ws bytes", n.value)
print("got[:4,:6]\n", got[:4, :6, 0, 0])
print("want[:4,:6]\n", want[:4, :6, 0, 0])
w = ws.get()
tail = w[-(F * Cc) - 64:]
print("workspace tail (partial region) unique head:", np.unique(tail)[:10], "count -3:", int((w == -3.0).sum()), "of", w.size)
print("max abs err", np.abs(got - want).max())

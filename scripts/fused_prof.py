"""MMA-thread cycle breakdown (MGB_HALO_PROF=1) of the C2 forward: plain epilogue vs pooled epilogue"""
import ctypes as C, os, sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import mygrad_b200 as mg
from mygrad_b200._lib import ConvDesc, PoolDesc, check, lib

N = int(os.environ.get("PROF_N", 256))
cd, pd = ConvDesc(), PoolDesc()
cd.dtype, cd.nd, cd.N, cd.C, cd.F = 0, 2, N, 64, 128
for i in range(2):
    cd.X[i], cd.W[i], cd.stride[i], cd.pad[i], cd.dil[i] = 56, 3, 1, 1, 1
    pd.X[i], pd.P[i], pd.stride[i] = 56, 2, 2
pd.dtype, pd.nd, pd.mode, pd.lead = 0, 2, 0, N * 128
x = mg.asdevice(np.random.default_rng(0).standard_normal((N, 64, 56, 56), dtype=np.float32))
w = mg.asdevice(np.random.default_rng(1).standard_normal((128, 64, 3, 3), dtype=np.float32))
y = mg.empty((N, 128, 56, 56)); p = mg.empty((N, 128, 28, 28)); am = mg.empty((N * 128 * 28 * 28 // 4,))
n = C.c_size_t(0)
check(lib.mgb_conv_workspace_bytes(C.byref(cd), 0, C.byref(n)))
ws = mg.empty((n.value // 8 + 1,), np.float64)
check(lib.mgb_conv_planes_bytes(C.byref(cd), 0, C.byref(n)))
xp = mg.empty((n.value // 8 + 1,), np.float64)
check(lib.mgb_conv_make_planes(C.byref(cd), 0, C.c_void_p(x.ptr), C.c_void_p(xp.ptr), None))
P = lambda a: C.c_void_p(a.ptr)
for rep in range(2):
    print("--- plain epilogue", file=sys.stderr, flush=True)
    check(lib.mgb_conv_fwd_p(C.byref(cd), P(x), P(w), P(y), P(xp), P(ws), ws.nbytes, None))
    mg.synchronize()
    print("--- pooled epilogue", file=sys.stderr, flush=True)
    check(lib.mgb_conv_pool_fwd(C.byref(cd), C.byref(pd), P(x), P(w), P(p), P(am), P(xp), P(ws), ws.nbytes, None))
    mg.synchronize()
# dX of the same layer (N tile 64)
dy = mg.asdevice(np.random.default_rng(2).standard_normal((N, 128, 56, 56), dtype=np.float32))
dx = mg.empty((N, 64, 56, 56))
check(lib.mgb_conv_workspace_bytes(C.byref(cd), 1, C.byref(n)))
ws1 = mg.empty((n.value // 8 + 1,), np.float64)
check(lib.mgb_conv_planes_bytes(C.byref(cd), 1, C.byref(n)))
dyp = mg.empty((n.value // 8 + 1,), np.float64)
check(lib.mgb_conv_make_planes(C.byref(cd), 1, C.c_void_p(dy.ptr), C.c_void_p(dyp.ptr), None))
for rep in range(int(os.environ.get("PROF_REPS", 1))):
    print("--- dgrad", file=sys.stderr, flush=True)
    check(lib.mgb_conv_dgrad_p(C.byref(cd), P(dy), P(w), P(dx), P(dyp), P(ws1), ws1.nbytes, None))
    mg.synchronize()
# dW of the same layer
dw = mg.empty((128, 64, 3, 3))
check(lib.mgb_conv_workspace_bytes(C.byref(cd), 2, C.byref(n)))
ws2 = mg.empty((n.value // 8 + 1,), np.float64)
for rep in range(3):
    print("--- wgrad", file=sys.stderr, flush=True)
    check(lib.mgb_conv_wgrad_p(C.byref(cd), P(dy), P(x), P(dw), P(dyp), P(xp), P(ws2), ws2.nbytes, None))
    mg.synchronize()

"""Row-halo wgrad (conv_wgrad2.cuh) against the NumPy oracle and against wgrad_tma_kernel
(MGB_WGRAD_ROWS=0), on shapes that exercise box shapes, c-block groups, dilation, strides and 3-D maps;
then C2 timing of both.  usage: python scripts/wgrad_rows_check.py [--time]"""
import ctypes as C, os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import mygrad_b200 as mg
from mygrad_b200._lib import check, lib, ConvDesc
from oracle import conv_pool_oracle as oc

stream = C.c_void_p(mg.current_stream())
P = lambda a: C.c_void_p(a.ptr)


def desc(xs, ws, stride, pad, dil):
    nd = len(xs) - 2
    cd = ConvDesc()
    cd.dtype, cd.nd = 0, nd
    cd.N, cd.C, cd.F = xs[0], xs[1], ws[0]
    for i in range(nd):
        cd.X[i], cd.W[i] = xs[2 + i], ws[2 + i]
        cd.stride[i], cd.pad[i], cd.dil[i] = stride[i], pad[i], dil[i]
    return cd


def run_wgrad(cd, dy, x, ws_shape):
    n = C.c_size_t(0)
    check(lib.mgb_conv_workspace_bytes(C.byref(cd), 2, C.byref(n)))
    ws = mg.empty((max(8, n.value) + 7) // 8, np.float64)
    dw = mg.empty(ws_shape, np.float32)
    fn = lambda: check(lib.mgb_conv_wgrad(C.byref(cd), P(dy), P(x), P(dw), P(ws), n.value, stream))
    fn()
    mg.synchronize()
    return dw, fn


CASES = [  # xs, ws, stride, pad, dil
    ((4, 64, 20, 20), (128, 64, 3, 3), (1, 1), (1, 1), (1, 1)),     # C2-like: N tile 192 = 3 row taps x 2 c-blocks
    ((3, 32, 17, 19), (64, 32, 3, 3), (1, 1), (1, 1), (1, 1)),      # one c-block, ragged boxes, F = 64
    ((2, 128, 12, 12), (256, 128, 2, 2), (1, 1), (0, 0), (1, 1)),   # 4 c-blocks in one CTA, 2 filter tiles, odd halo rows
    ((2, 256, 9, 9), (96, 256, 3, 3), (1, 2), (1, 1), (1, 1)),      # cblocks = 8 in groups of 2, F = 96, stride 2 on the last axis
    ((2, 256, 9, 9), (96, 256, 3, 3), (2, 2), (1, 1), (1, 1)),      # stride 2 on the middle axis: not applicable, per-tap kernel
    ((5, 16, 30, 30), (32, 16, 5, 5), (1, 1), (2, 2), (2, 2)),      # C padded to 32, dilation 2: block stride = 2 rows
    ((3, 64, 9, 70), (128, 64, 3, 3), (1, 1), (1, 1), (1, 1)),      # wide and short: 1 x 32 or 2 x 16 boxes
    ((2, 32, 6, 10, 10), (64, 32, 3, 3, 3), (1, 1, 1), (0, 0, 0), (1, 1, 1)),  # C4-like 3-D
    ((2, 64, 6, 9, 12), (128, 64, 2, 3, 3), (2, 1, 1), (0, 1, 1), (1, 1, 1)),  # 3-D, stride 2 on the outer axis
    ((2, 64, 40), (128, 64, 5), (1,), (2,), (1,)),                  # 1-D: one row tap
]

ok = True
rng = np.random.default_rng(5)
for xs, ws, st, pd, dl in CASES:
    xh = rng.standard_normal(xs).astype(np.float32)
    G = oc.conv_out_shape(xs[2:], ws[2:], st, pd, dl)
    dyh = rng.standard_normal((xs[0], ws[0], *G)).astype(np.float32)
    want = oc.conv_dw(dyh.astype(np.float64), xh.astype(np.float64), ws, st, pd, dl)
    cd = desc(xs, ws, st, pd, dl)
    x, dy = mg.asdevice(xh), mg.asdevice(dyh)
    res = {}
    for mode in ("1", "0"):
        os.environ["MGB_WGRAD_ROWS"] = mode
        dw, _ = run_wgrad(cd, dy, x, ws)
        res[mode] = dw.get()
    os.environ.pop("MGB_WGRAD_ROWS")
    scale = np.abs(want).max()
    e1 = np.abs(res["1"] - want).max() / scale
    e0 = np.abs(res["0"] - want).max() / scale
    good = e1 < 1e-5
    ok &= good
    print(f"{'ok ' if good else 'BAD'} x{xs} w{ws} s{st} p{pd} d{dl}: rows err {e1:.2e}, per-tap err {e0:.2e}", flush=True)

if "--time" in sys.argv:
    xs, ws = (256, 64, 56, 56), (128, 64, 3, 3)
    cd = desc(xs, ws, (1, 1), (1, 1), (1, 1))
    x = mg.asdevice(rng.standard_normal(xs).astype(np.float32))
    dy = mg.asdevice(rng.standard_normal((256, 128, 56, 56)).astype(np.float32))
    for rep in range(2):
        for mode in ("0", "1"):
            os.environ["MGB_WGRAD_ROWS"] = mode
            dw, fn = run_wgrad(cd, dy, x, ws)
            a, b = C.c_void_p(), C.c_void_p()
            check(lib.mgb_event_create(C.byref(a))); check(lib.mgb_event_create(C.byref(b)))
            fn(); fn()
            check(lib.mgb_event_record(a, stream))
            for _ in range(10):
                fn()
            check(lib.mgb_event_record(b, stream))
            ms = C.c_float(0)
            check(lib.mgb_event_synchronize(b)); check(lib.mgb_event_elapsed_ms(a, b, C.byref(ms)))
            print(f"C2 wgrad incl. both plane preps, MGB_WGRAD_ROWS={mode}: {ms.value / 10:.4f} ms", flush=True)
    os.environ["MGB_WGRAD_PROF"] = "1"
    for mode in ("0", "1"):
        os.environ["MGB_WGRAD_ROWS"] = mode
        run_wgrad(cd, dy, x, ws)
print("WGRAD_ROWS_CHECK", "OK" if ok else "FAILED")
sys.exit(0 if ok else 1)

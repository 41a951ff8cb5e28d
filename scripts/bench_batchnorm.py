"""HBM-roofline check of the batchnorm kernels on a C2-sized activation (N,128,56,56).
usage: python scripts/bench_batchnorm.py [f32|f64] [N]   -> one JSON line"""
import ctypes as C, json, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
import mygrad_b200 as mg
from mygrad_b200._lib import check, lib

dt = np.float64 if (len(sys.argv) > 1 and sys.argv[1] == "f64") else np.float32
n = int(sys.argv[2]) if len(sys.argv) > 2 else 256
c, s = 128, 56 * 56
code = 0 if dt == np.float32 else 1
rng = np.random.default_rng(0)
x = mg.asdevice(rng.standard_normal((n, c, s)).astype(dt))
g = mg.asdevice(rng.standard_normal((n, c, s)).astype(dt))
gamma = mg.asdevice(np.ones(c, dt)); beta = mg.asdevice(np.zeros(c, dt))
y, dx = mg.empty((n, c, s), dt), mg.empty((n, c, s), dt)
moments, sums = mg.empty((3, c), np.float64), mg.empty((2, c), np.float64)
nb = C.c_size_t(0); check(lib.mgb_batchnorm_workspace_bytes(n, c, C.byref(nb)))
ws = mg.empty(((nb.value + 7) // 8,), np.float64)
st = C.c_void_p(mg.current_stream()); P = lambda a: C.c_void_p(a.ptr)
eps = C.c_double(1e-5)
calls = {
    "stats": (lambda: check(lib.mgb_batchnorm_stats(code, n, c, s, P(x), P(moments), P(ws), nb, st)), 1),
    "fwd": (lambda: check(lib.mgb_batchnorm_fwd(code, n, c, s, P(x), P(moments), P(gamma), P(beta), eps, P(y), st)), 2),
    "bwd_reduce": (lambda: check(lib.mgb_batchnorm_bwd_reduce(code, n, c, s, P(g), P(x), P(moments), eps, P(sums), P(ws), nb, st)), 2),
    "bwd_dx": (lambda: check(lib.mgb_batchnorm_bwd_dx(code, n, c, s, P(g), P(x), P(moments), P(sums), P(gamma), eps, P(dx), st)), 3),
}
peak = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()).get("hbm_gbs", 6650.0) if (ROOT / "MEASURED_PEAKS.json").exists() else 6650.0
out = {"workload": f"batchnorm x=({n},{c},56,56) {np.dtype(dt).name}", "hbm_peak_gbs": peak, "kernels": {}}
for name, (fn, passes) in calls.items():
    fn(); fn()
    a, b = C.c_void_p(), C.c_void_p()
    check(lib.mgb_event_create(C.byref(a))); check(lib.mgb_event_create(C.byref(b)))
    check(lib.mgb_event_record(a, st))
    for _ in range(10):
        fn()
    check(lib.mgb_event_record(b, st)); check(lib.mgb_event_synchronize(b))
    ms = C.c_float(0); check(lib.mgb_event_elapsed_ms(a, b, C.byref(ms)))
    t = ms.value / 10
    nbytes = passes * x.nbytes                      # algorithmic: tensors of x's size read or written once each
    out["kernels"][name] = {"ms": round(t, 4), "bytes": nbytes, "gbs": round(nbytes / t / 1e6, 1),
                            "frac_of_hbm_peak": round(nbytes / t / 1e6 / peak, 3)}
print(json.dumps(out))

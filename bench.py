#!/usr/bin/env python
"""bench.py — conv_nd + max_pool forward+backward throughput on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload C2|C2f64|C1|C3|C4]

One "step" = one pass of the hot path over one synthetic batch through the public API:
``out = max_pool(conv_nd(x, w, ...), pool, 2); out.backward()`` (conv fwd, pool fwd, pool
bwd, conv dX, conv dW; plus the dW all-reduce when N > 1).  Prints ONE JSON line.

* ``value``  — whole-job samples/s with x and w already resident in HBM.
* ``e2e``    — same metric with HOST buffers: pinned x, w copied H2D and out, x.grad, w.grad
               copied D2H inside the timed region, every step.
* ``roofline`` — the slowest kernel of the step, algorithmic FLOPs (or bytes) over its
               CUDA-event duration, against the measured peak.
* ``cpu_baseline`` — the reference's NumPy path on the host cores (bounded sample).

``--impl reference`` times the unmodified reference (baseline/_ref, else the oracle port)
on the host cores for the same workload.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

if "reference" in sys.argv:
    # torchrun exports OMP_NUM_THREADS=1 to every rank: the reference arm (rank 0 only) must see all
    # host cores, so the BLAS thread caps are lifted BEFORE NumPy / OpenBLAS load
    for _k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS", "GOTO_NUM_THREADS"):
        os.environ.pop(_k, None)

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "conv_nd+max_pool fwd+bwd samples/sec"

WORKLOADS = {
    # BASELINE.json configs (C3 as written is rejected by the reference; 65x65 is the nearest valid shape)
    "C1": dict(x=(32, 3, 32, 32), w=(16, 3, 3, 3), stride=1, padding=1, dilation=1, pool=(2, 2), dtype="f64"),
    "C2": dict(x=(256, 64, 56, 56), w=(128, 64, 3, 3), stride=1, padding=1, dilation=1, pool=(2, 2), dtype="f32"),
    "C2f64": dict(x=(256, 64, 56, 56), w=(128, 64, 3, 3), stride=1, padding=1, dilation=1, pool=(2, 2), dtype="f64"),
    "C3": dict(x=(128, 128, 65, 65), w=(256, 128, 5, 5), stride=2, padding=0, dilation=2, pool=None, dtype="f64"),
    # config 4 is a strong-scaling config: batch 32 sharded over 8 GPUs = 4 samples per GPU ("C4shard")
    "C4": dict(x=(32, 32, 32, 64, 64), w=(64, 32, 3, 3, 3), stride=1, padding=0, dilation=1, pool=(2, 2, 2), dtype="f32"),
    "C4shard": dict(x=(4, 32, 32, 64, 64), w=(64, 32, 3, 3, 3), stride=1, padding=0, dilation=1, pool=(2, 2, 2), dtype="f32"),
}
# config 5: small CNN, full Tensor.backward (SURVEY.md §8d concretisation)
C5 = dict(batch=2048, dtype="f32")
NP_DTYPE = {"f32": np.float32, "f64": np.float64}


def workload_name(name: str, cfg: dict) -> str:
    x, w = cfg["x"], cfg["w"]
    pool = f"+max_pool{'x'.join(map(str, cfg['pool']))}" if cfg["pool"] else ""
    return (f"{name}: conv_nd x={x} w={w} stride={cfg['stride']} pad={cfg['padding']} "
            f"dil={cfg['dilation']}{pool} {cfg['dtype']}")


def geometry(cfg: dict):
    nd = len(cfg["x"]) - 2
    tup = lambda v: (v,) * nd if isinstance(v, int) else tuple(v)
    s, p, d = tup(cfg["stride"]), tup(cfg["padding"]), tup(cfg["dilation"])
    G = tuple((X + 2 * pi - ((W - 1) * di + 1)) // si + 1
              for X, W, si, pi, di in zip(cfg["x"][2:], cfg["w"][2:], s, p, d))
    N, Cc, F = cfg["x"][0], cfg["x"][1], cfg["w"][0]
    flops = 2.0 * N * F * np.prod(G) * Cc * np.prod(cfg["w"][2:])   # per GEMM
    esz = 4 if cfg["dtype"] == "f32" else 8
    y_elems = N * F * int(np.prod(G))
    p_elems = y_elems // int(np.prod(cfg["pool"])) if cfg["pool"] else 0
    return dict(G=G, flops=float(flops), esz=esz, y_bytes=y_elems * esz, p_bytes=p_elems * esz,
                x_bytes=int(np.prod(cfg["x"])) * esz, w_bytes=int(np.prod(cfg["w"])) * esz)


def make_inputs(cfg: dict):
    dt = NP_DTYPE[cfg["dtype"]]
    x = np.random.default_rng(0).standard_normal(cfg["x"]).astype(dt)
    w = np.random.default_rng(1).standard_normal(cfg["w"]).astype(dt)
    return x, w


# ---- clocks ---------------------------------------------------------------------------

class ClockSampler:
    """samples nvidia-smi while the timed region runs (B200_PROFILING.md clocks line)"""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [t.strip() for t in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        # under load = the upper half of the samples (idle gaps between launches pull the tail down)
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "samples": len(sm),
                "reasons": sorted(reasons)}


# ---- reference arm ------------------------------------------------------------------------

def load_reference():
    """the unmodified reference (pure Python) if it travelled with the repo, else the oracle port"""
    ref = ROOT / "baseline" / "_ref"
    if (ref / "mygrad").exists():
        sys.path.insert(0, str(ref))
        try:
            import mygrad as ref_mg
            from mygrad.nnet.layers import conv_nd as r_conv, max_pool as r_pool

            def step(x, w, cfg):
                xt, wt = ref_mg.Tensor(x), ref_mg.Tensor(w)
                y = r_conv(xt, wt, stride=cfg["stride"], padding=cfg["padding"], dilation=cfg["dilation"])
                o = r_pool(y, cfg["pool"], 2) if cfg["pool"] else y
                o.backward()
                return o.data, xt.grad, wt.grad

            return "reference", step
        except Exception:
            pass
    from oracle import conv_pool_oracle as oc

    def step(x, w, cfg):
        return oc.layer_step(x, w, stride=cfg["stride"], padding=cfg["padding"], dilation=cfg["dilation"],
                             pool=cfg["pool"], pool_stride=2)

    return "port", step


def host_threads() -> int:
    try:
        from threadpoolctl import threadpool_info

        n = [i.get("num_threads", 1) for i in threadpool_info() if i.get("user_api") == "blas"]
        if n:
            return int(max(n))
    except Exception:
        pass
    return os.cpu_count() or 1


def cpu_sample(cfg: dict, budget_s: float, steps: int = 1, warmup: int = 0):
    """times the reference on a batch-slice of the workload sized to ~budget_s in total"""
    kind, step = load_reference()
    x, w = make_inputs(dict(cfg, x=(min(2, cfg["x"][0]), *cfg["x"][1:])))
    t0 = time.perf_counter()
    step(x, w, cfg)
    t_probe = (time.perf_counter() - t0) / x.shape[0]        # s per sample, cold
    runs = max(1, steps + warmup)
    b = int(max(1, min(cfg["x"][0], budget_s / runs / max(t_probe, 1e-9))))
    b = 1 << (b.bit_length() - 1)                             # power of two <= b
    x, w = make_inputs(dict(cfg, x=(b, *cfg["x"][1:])))
    for _ in range(warmup):
        step(x, w, cfg)
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        step(x, w, cfg)
        times.append(time.perf_counter() - t0)
    return kind, b, times


def all_host_threads() -> int:
    """lets the BLAS behind NumPy use every host core again (torchrun pins OMP_NUM_THREADS=1)"""
    n = os.cpu_count() or 1
    try:
        from threadpoolctl import threadpool_limits

        threadpool_limits(limits=n, user_api="blas")
    except Exception:
        pass
    return host_threads()


def run_reference(args, cfg, name):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    cores = all_host_threads()
    if args.scaling == "strong":      # the reference has no ranks: whole job = the full batch either way
        pass
    kind, b, times = cpu_sample(cfg, budget_s=150.0, steps=args.steps, warmup=args.warmup)
    ms = 1e3 * float(np.mean(times))
    v = b / (ms * 1e-3)
    sample = f"batch {b} of {cfg['x'][0]} (same layer), mean of {args.steps} steps after {args.warmup} warm-up"
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": cfg["dtype"], "data": "synthetic",
        "config": {"workload": workload_name(name, cfg), "host": "NumPy/OpenBLAS on the box's CPU cores"},
        "cpu_baseline": {"value": v, "unit": "samples/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---- our arm ---------------------------------------------------------------------------------

def load_peaks():
    mp_file = ROOT / "MEASURED_PEAKS.json"
    measured = json.loads(mp_file.read_text()) if mp_file.exists() else {}
    src = "measured (MEASURED_PEAKS.json)" if measured else "fallback (B200_PROFILING.md)"
    return {
        "hbm_gbs": measured.get("hbm_gbs", 6650.0),
        "bf16_tflops_burst": measured.get("bf16_tflops", 1590.0),
        "bf16_tflops_sustained": measured.get("bf16_tflops_sustained", 1400.0),
        "source": src,
    }


def algorithmic_floor_ms(cfg, geo, tensor_peak_tflops, hbm_gbs):
    """composite roofline floor of one step: three conv GEMMs on the tensor pipe plus
    the pooling bytes (recompute-argmax form: fwd = y + p, bwd = 2y + p) at the HBM rate"""
    ms = 3.0 * geo["flops"] / (tensor_peak_tflops * 1e12) * 1e3
    if cfg["pool"]:
        ms += (3.0 * geo["y_bytes"] + 2.0 * geo["p_bytes"]) / (hbm_gbs * 1e9) * 1e3
    return ms


class Harness:
    """everything one rank needs to time the path: the library, the data-parallel group, events"""

    def __init__(self):
        import mygrad_b200 as mg
        from mygrad_b200._lib import check, lib
        from mygrad_b200.dist import DataParallel

        self.mg, self.lib, self.check = mg, lib, check
        self.rank = int(os.environ.get("RANK", 0))
        self.world = int(os.environ.get("WORLD_SIZE", 1))
        self.local = int(os.environ.get("LOCAL_RANK", 0))
        self.dist = None
        if self.world > 1:
            import torch
            import torch.distributed as dist

            torch.cuda.set_device(self.local)
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
            self.dist = dist
        self.dp = DataParallel(self.rank, self.world, self.local).init()
        self.stream = C.c_void_p(mg.current_stream())
        self.peaks = load_peaks()

    def close(self):
        self.barrier()
        self.dp.close()
        if self.dist is not None:
            self.dist.destroy_process_group()

    def barrier(self):
        self.mg.synchronize()
        if self.dist is not None:
            self.dist.barrier()
        self.mg.synchronize()

    def events(self):
        a, b = C.c_void_p(), C.c_void_p()
        self.check(self.lib.mgb_event_create(C.byref(a))); self.check(self.lib.mgb_event_create(C.byref(b)))
        return a, b

    def elapsed(self, a, b) -> float:
        ms = C.c_float(0)
        self.check(self.lib.mgb_event_synchronize(b))
        self.check(self.lib.mgb_event_elapsed_ms(a, b, C.byref(ms)))
        return float(ms.value)

    def max_over_ranks(self, v: float) -> float:
        if self.dist is None:
            return v
        import torch

        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def launches(self) -> int:
        n = C.c_uint64(0)
        self.lib.mgb_launch_count(C.byref(n))
        return int(n.value)

    # ---- one step through the public API ----------------------------------------------------------
    def make_step(self, cfg, fused: bool):
        """out = max_pool(conv_nd(x, w)); out.backward() — or the same layer through the opt-in fused
        Operation ``conv_pool`` (identical results; y is never written to HBM)"""
        mg, dp, world = self.mg, self.dp, self.world
        from mygrad_b200 import Tensor
        from mygrad_b200.nnet.layers import conv_nd, max_pool

        conv_kw = dict(stride=cfg["stride"], padding=cfg["padding"], dilation=cfg["dilation"])
        conv_pool = getattr(mg.nnet.layers, "conv_pool", None) if fused else None

        def step(xd, wd):
            xt, wt = Tensor(xd, copy=False), Tensor(wd, copy=False)
            y = None
            if conv_pool is not None and cfg["pool"]:
                o = conv_pool(xt, wt, pool=cfg["pool"], pool_stride=2, **conv_kw)
            else:
                y = conv_nd(xt, wt, **conv_kw)
                o = max_pool(y, cfg["pool"], 2) if cfg["pool"] else y
            with dp.fused_wgrad():      # dW: split-K reduction + cross-rank sum in one kernel
                o.backward()
            if world > 1:
                dp.allreduce([wt.grad])   # grads already summed by the fused kernel are skipped (dist.py)
            return o, xt, wt, y

        return step

    def time_steps(self, step, x_dev, w_dev, steps: int, warmup: int):
        """W warm-up steps, then K steps between CUDA events on the launching stream; max over ranks"""
        mg, lib, check = self.mg, self.lib, self.check
        for _ in range(warmup):
            step(x_dev, w_dev)
        sampler = ClockSampler(self.local)
        sampler.start()                       # spawns nvidia-smi: before the barrier, it takes tens of ms
        self.barrier()
        e0, e1 = self.events()
        # the host barrier leaves the ranks' STREAMS milliseconds apart; a one-element all-reduce queued right
        # before the start event lines them up on the device, so no rank's clock runs while it waits for a
        # late starter (the timed region itself is coupled by the dW exchange of every step)
        align = mg.zeros((1,), np.float32)
        self.dp.allreduce([align])
        n0 = self.launches()
        check(lib.mgb_event_record(e0, self.stream))
        for _ in range(steps):
            step(x_dev, w_dev)
        check(lib.mgb_event_record(e1, self.stream))
        total_ms = self.elapsed(e0, e1)
        self.barrier()
        n1 = self.launches()
        clocks = sampler.stop()
        if os.environ.get("MGB_BENCH_RANKTIMES"):
            print(f"[rank {self.rank}] {total_ms / steps:.3f} ms/step on its own stream", file=sys.stderr, flush=True)
        return self.max_over_ranks(total_ms) / steps, n1 - n0, clocks

    def time_graph(self, step, x_dev, w_dev, steps: int, warmup: int):
        """the step recorded as a CUDA graph: K replays between events on the graph's stream; outputs compared
        bit for bit with one eager step"""
        mg, lib, check = self.mg, self.lib, self.check
        o, xt, wt, _ = step(x_dev, w_dev)
        eager = (o.numpy(), xt.grad.get(), wt.grad.get())
        g = mg.StepGraph(lambda: step(x_dev, w_dev), warmup=max(2, warmup))
        gs = C.c_void_p(g.stream)
        for _ in range(warmup):
            g.replay()
        g.synchronize()
        e0, e1 = self.events()
        check(lib.mgb_event_record(e0, gs))
        for _ in range(steps):
            g.replay()
        check(lib.mgb_event_record(e1, gs))
        total_ms = self.elapsed(e0, e1)
        o, xt, wt, _ = g.result
        same = all(np.array_equal(a, b) for a, b in zip((o.numpy(), xt.grad.get(), wt.grad.get()), eager))
        nodes = g.kernel_nodes
        g.close()
        return {"ms_per_step": total_ms / steps, "kernel_nodes": nodes, "gpu_launches": nodes * steps,
                "bit_identical_to_eager": bool(same),
                "api": "StepGraph(step).replay() — one cudaGraphLaunch per step, same kernels and buffers"}

    # ---- per-phase timing straight through the C ABI (preallocated outputs: only kernels are timed) ----
    def time_phases(self, cfg, geo, x_dev, w_dev):
        mg, lib, check, stream = self.mg, self.lib, self.check, self.stream
        from mygrad_b200._lib import ConvDesc, PoolDesc

        dt = NP_DTYPE[cfg["dtype"]]
        nd = len(cfg["x"]) - 2
        tup = lambda v: (v,) * nd if isinstance(v, int) else tuple(v)
        cd = ConvDesc()
        cd.dtype, cd.nd = (0 if cfg["dtype"] == "f32" else 1), nd
        cd.N, cd.C, cd.F = cfg["x"][0], cfg["x"][1], cfg["w"][0]
        for i in range(nd):
            cd.X[i], cd.W[i] = cfg["x"][2 + i], cfg["w"][2 + i]
            cd.stride[i], cd.pad[i], cd.dil[i] = tup(cfg["stride"])[i], tup(cfg["padding"])[i], tup(cfg["dilation"])[i]
        y = mg.empty((cfg["x"][0], cfg["w"][0], *geo["G"]), dt)
        # upstream gradients: seed-2 normal data, as BASELINE.md §2 specifies for the per-phase timings
        rng2 = np.random.default_rng(2)
        gy = mg.asdevice(rng2.standard_normal(y.shape, dtype=np.float32).astype(dt, copy=False))
        dx, dw = mg.empty(cfg["x"], dt), mg.empty(cfg["w"], dt)
        wss = []
        for which in range(3):
            n = C.c_size_t(0)
            check(lib.mgb_conv_workspace_bytes(C.byref(cd), which, C.byref(n)))
            wss.append((mg.empty((max(8, n.value) + 7) // 8, np.float64), n.value))
        P = lambda a: C.c_void_p(a.ptr)
        PN = lambda a: C.c_void_p(a.ptr if a is not None else 0)
        phases = {}

        def timed(fn, reps=5):
            fn()
            a, b = self.events()
            check(lib.mgb_event_record(a, stream))
            for _ in range(reps):
                fn()
            check(lib.mgb_event_record(b, stream))
            return self.elapsed(a, b) / reps

        # the step builds the channels-last operand planes of the tensor-core path ONCE per tensor
        # (x planes: fwd + dW, dy planes: dX + dW; ConvND keeps them between forward and backward), so
        # the phases time the same entry points the step calls: plane prep on its own (HBM-bound), then
        # the *_p kernels on prebuilt planes.  float64 has no planes (bytes = 0, NULL).
        def planes_for(tensor, src):
            n = C.c_size_t(0)
            check(lib.mgb_conv_planes_bytes(C.byref(cd), tensor, C.byref(n)))
            if n.value == 0:
                return None, None
            buf = mg.empty(((n.value + 7) // 8,), np.float64)
            ms = timed(lambda: check(lib.mgb_conv_make_planes(C.byref(cd), tensor, P(src), P(buf), stream)))
            return buf, ms

        x_planes, ms_px = planes_for(0, x_dev)
        dy_planes, ms_pdy = planes_for(1, gy)
        if ms_px is not None:
            phases["plane_prep_x"] = {"ms": ms_px, "bytes": 3 * geo["x_bytes"]}        # read x, write hi + lo
            phases["plane_prep_dy"] = {"ms": ms_pdy, "bytes": 3 * geo["y_bytes"]}
        phases["conv_fwd"] = {"ms": timed(lambda: check(lib.mgb_conv_fwd_p(
            C.byref(cd), P(x_dev), P(w_dev), P(y), PN(x_planes), P(wss[0][0]), wss[0][1], stream))),
            "flops": geo["flops"]}
        if cfg["pool"]:
            npool = len(cfg["pool"])
            pd = PoolDesc()
            pd.dtype, pd.nd, pd.mode = cd.dtype, npool, 0
            pd.lead = int(np.prod(y.shape[: y.ndim - npool]))
            for i in range(npool):
                pd.X[i], pd.P[i], pd.stride[i] = y.shape[y.ndim - npool + i], cfg["pool"][i], 2
            pooled = mg.empty((*y.shape[: y.ndim - npool],
                               *[(X - p_) // 2 + 1 for X, p_ in zip(y.shape[y.ndim - npool:], cfg["pool"])]), dt)
            gp = mg.asdevice(rng2.standard_normal(pooled.shape, dtype=np.float32).astype(dt, copy=False))
            gy2 = mg.empty(y.shape, dt)
            phases["pool_fwd"] = {"ms": timed(lambda: check(lib.mgb_pool_fwd(C.byref(pd), P(y), P(pooled), stream))),
                                  "bytes": geo["y_bytes"] + geo["p_bytes"]}
            phases["pool_bwd"] = {"ms": timed(lambda: check(lib.mgb_pool_bwd(C.byref(pd), P(gp), P(y), P(gy2), stream))),
                                  "bytes": 2 * geo["y_bytes"] + geo["p_bytes"]}
            del gy2
            # the opt-in fused Operation (conv_pool): pooled epilogue + dy written straight as planes
            yes = C.c_int(0)
            check(lib.mgb_conv_pool_supported(C.byref(cd), C.byref(pd), C.byref(yes)))
            if yes.value and dy_planes is not None:
                # what the drop-in pair runs instead of pool_bwd + plane_prep_dy: dy and its planes in one pass
                dyp1, gy3 = mg.empty((dy_planes.size,), np.float64), mg.empty(y.shape, dt)
                phases["pool_bwd_dy_planes"] = {"ms": timed(lambda: check(lib.mgb_pool_bwd_dy_planes(
                    C.byref(cd), C.byref(pd), P(gp), P(y), P(gy3), P(dyp1), stream))),
                    "bytes": 4 * geo["y_bytes"] + geo["p_bytes"],
                    "note": "replaces pool_bwd + plane_prep_dy in the step (y + dp in; dy + hi/lo planes out)"}
                phases["pool_bwd"]["two_pass_only"] = True
                phases["plane_prep_dy"]["two_pass_only"] = True
                del dyp1, gy3
                am = mg.empty(((pooled.size + 3) // 4,), np.float32)
                phases["fused_conv_pool_fwd"] = {"ms": timed(lambda: check(lib.mgb_conv_pool_fwd(
                    C.byref(cd), C.byref(pd), P(x_dev), P(w_dev), P(pooled), P(am), PN(x_planes), P(wss[0][0]),
                    wss[0][1], stream))), "flops": geo["flops"], "fused_only": True}
                dyp2 = mg.empty((dy_planes.size,), np.float64)
                phases["fused_pool_bwd_planes"] = {"ms": timed(lambda: check(lib.mgb_pool_bwd_planes(
                    C.byref(cd), C.byref(pd), P(gp), P(am), P(dyp2), stream))),
                    "bytes": geo["p_bytes"] + geo["p_bytes"] // 4 + 2 * geo["y_bytes"], "fused_only": True}
                del dyp2
        phases["conv_dgrad"] = {"ms": timed(lambda: check(lib.mgb_conv_dgrad_p(
            C.byref(cd), P(gy), P(w_dev), P(dx), PN(dy_planes), P(wss[1][0]), wss[1][1], stream))),
            "flops": geo["flops"]}
        phases["conv_wgrad"] = {"ms": timed(lambda: check(lib.mgb_conv_wgrad_p(
            C.byref(cd), P(gy), P(x_dev), P(dw), PN(dy_planes), PN(x_planes), P(wss[2][0]), wss[2][1], stream))),
            "flops": geo["flops"]}
        for v in phases.values():
            if "flops" in v:
                v["tflops"] = v["flops"] / (v["ms"] * 1e-3) / 1e12
            else:
                v["gbs"] = v["bytes"] / (v["ms"] * 1e-3) / 1e9
        return phases

    def probe_peaks(self):
        tf = C.c_double(0)
        out = {}
        self.check(self.lib.mgb_probe_mma_peak(0, 4096, C.byref(tf), self.stream)); out["fp64_dmma_tflops"] = tf.value
        self.check(self.lib.mgb_probe_mma_peak(1, 4096, C.byref(tf), self.stream)); out["tf32_mma_sync_tflops"] = tf.value
        out["hbm_gbs"] = self.peaks["hbm_gbs"]
        out["bf16_tflops_burst"] = self.peaks["bf16_tflops_burst"]
        out["bf16_tflops_sustained"] = self.peaks["bf16_tflops_sustained"]
        return out

    def roofline(self, cfg, geo, phases, peaks, ms_per_step, workload):
        """the dominant kernel against its measured ceiling, and the whole step against the composite floor.
        The phases are isolated launches (5 back to back, ~3 ms): their denominator is the BURST figure.
        3xTF32 useful peak = bf16 / 2 (TF32 runs at half the bf16 rate) / 3 (three MMAs per fp32 product)."""
        top = max((k for k in phases if not phases[k].get("fused_only")), key=lambda k: phases[k]["ms"])
        ph = phases[top]
        src = self.peaks["source"]
        if cfg["dtype"] == "f64":
            t_burst = t_sust = peaks["fp64_dmma_tflops"]
            basis = "FP64 DMMA register-resident loop measured in this run (mgb_probe_mma_peak)"
        else:
            t_burst, t_sust = peaks["bf16_tflops_burst"] / 6.0, peaks["bf16_tflops_sustained"] / 6.0
            basis = (f"{src} bf16 burst {peaks['bf16_tflops_burst']} TF/s / 2 / 3 (3xTF32); "
                     f"sustained basis {peaks['bf16_tflops_sustained']} / 6 = {t_sust:.1f} TF/s is frac_vs_sustained")
        if "flops" in ph:
            achieved = ph["tflops"]
            roof = {"bound": "tensor", "kernel": top, "achieved": achieved, "peak": t_burst, "unit": "TFLOP/s",
                    "frac": achieved / t_burst, "frac_vs_sustained": achieved / t_sust, "traffic": None,
                    "peak_basis": basis}
        else:
            achieved = ph["gbs"]
            roof = {"bound": "hbm", "kernel": top, "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": achieved / peaks["hbm_gbs"], "traffic": None, "peak_basis": src}
        for tf in ("r02b_traffic.json", "r02_traffic.json", "r01_traffic.json"):     # DRAM bytes of that kernel from the committed ncu capture
            tfile = ROOT / "profiles" / tf
            if tfile.exists():
                t = json.loads(tfile.read_text()).get(workload, {}).get(top)
                if t is not None:
                    roof["traffic"] = t
                    roof["traffic_from"] = f"profiles/{tf}"
                    break
        roof["per_kernel_frac"] = {k: (v["tflops"] / t_burst if "flops" in v else v["gbs"] / peaks["hbm_gbs"])
                                   for k, v in phases.items()}
        floor_b = algorithmic_floor_ms(cfg, geo, t_burst, peaks["hbm_gbs"])
        floor_s = algorithmic_floor_ms(cfg, geo, t_sust, peaks["hbm_gbs"])
        roof["step_floor_ms"] = floor_b
        roof["step_frac"] = floor_b / ms_per_step
        roof["step_frac_vs_sustained"] = floor_s / ms_per_step
        roof["step_conv_tflops"] = 3.0 * geo["flops"] / (ms_per_step * 1e-3) / 1e12
        roof["step_basis"] = ("floor = 3 conv GEMMs at the tensor peak + pooling bytes (fwd y+p, bwd 2y+p) at the "
                              "measured HBM rate; step_frac = floor / measured ms_per_step")
        return roof

    # ---- parity of this very configuration, outside the timed region ---------------------------
    def parity(self, cfg, fused: bool):
        """m samples per rank of the workload's layer, every rank on its own shard through the SAME step
        function that was timed (fused wgrad + cross-rank sum included).  Checked against the oracle
        (oracle/conv_pool_oracle.py, pinned to the reference by tests/golden):
          out            vs conv_fwd + max_pool_fwd                      (tolerance)
          pool routing   y.grad == max_pool_bwd(ones, the GPU's own y)    (bit-exact)
          dX             vs conv_dx(dy routed by the GPU)                 (tolerance)
          dW (summed)    vs conv_dw over ALL ranks' shards, and bit-identical on every rank
        The oracle is fed the dy the GPU routed because a float32 near-tie inside a pooling window may
        resolve differently on y values that differ in the last bits.  The fused conv_pool Operation must
        reproduce the unfused path bit for bit."""
        mg, world, rank = self.mg, self.world, self.rank
        from oracle import conv_pool_oracle as oc

        geo1 = geometry(dict(cfg, x=(1, *cfg["x"][1:])))
        m = 2 if 3 * geo1["flops"] < 12e9 else 1
        total = m * world
        dt = NP_DTYPE[cfg["dtype"]]
        xs = np.random.default_rng(7).standard_normal((total, *cfg["x"][1:])).astype(dt)
        ws = np.random.default_rng(1).standard_normal(cfg["w"]).astype(dt)
        lo, hi = rank * m, (rank + 1) * m
        xd, wd = mg.asdevice(xs[lo:hi]), mg.asdevice(ws)
        o, xt, wt, y = self.make_step(cfg, False)(xd, wd)
        got = (o.numpy(), xt.grad.get(), wt.grad.get())
        kw = (cfg["stride"], cfg["padding"], cfg["dilation"])
        errs, exact = {}, {}
        tol = 1e-5 if cfg["dtype"] == "f32" else 1e-10

        def err(g, w_):
            scale = float(np.max(np.abs(w_))) or 1.0
            return float(np.max(np.abs(g.astype(np.float64) - w_.astype(np.float64)))) / scale

        if fused:
            fo, fxt, fwt, _ = self.make_step(cfg, True)(xd, wd)
            fgot = (fo.numpy(), fxt.grad.get(), fwt.grad.get())
            exact = {k: bool(np.array_equal(a, b_)) for k, a, b_ in zip(("out", "dx", "dw"), fgot, got)}
            errs = {k: err(a, b_) for k, a, b_ in zip(("out", "dx", "dw"), fgot, got)}
            ok_local = all(e <= 2 * tol for e in errs.values())
            dw_for_hash = fgot[2]
        else:
            y_h, dy_h = y.numpy(), y.grad.get()
            want_y = oc.conv_fwd(xs[lo:hi], ws, *kw)
            if cfg["pool"]:
                errs["out"] = err(got[0], oc.max_pool_fwd(want_y, cfg["pool"], 2))
                exact["pool_fwd_on_gpu_y"] = bool(np.array_equal(got[0], oc.max_pool_fwd(y_h, cfg["pool"], 2)))
                exact["pool_routing"] = bool(np.array_equal(
                    dy_h, oc.max_pool_bwd(np.ones(o.shape, dt), y_h, cfg["pool"], 2).astype(dt)))
            else:
                errs["out"] = err(got[0], want_y)
            errs["dx"] = err(got[1], oc.conv_dx(dy_h, xs[lo:hi].shape, ws, *kw, dtype=dt))
            ok_local = all(e <= 2 * tol for e in errs.values()) and all(exact.values())
            dw_for_hash = got[2]
        same, oks, dys = True, [ok_local], None
        if not fused:
            dys = [dy_h]
        if self.dist is not None:
            import hashlib

            box = [None] * world
            self.dist.all_gather_object(box, (hashlib.sha1(dw_for_hash.tobytes()).hexdigest(), ok_local,
                                              None if fused else dy_h))
            same = all(b[0] == box[0][0] for b in box)
            oks = [b[1] for b in box]
            dys = None if fused else [b[2] for b in box]
        res = None
        if rank == 0:
            if not fused:
                errs["dw"] = err(got[2], oc.conv_dw(np.concatenate(dys), xs, ws.shape, *kw).astype(dt))
            ok = bool(same and all(oks) and all(e <= 2 * tol for e in errs.values()))
            res = {"ok": ok, "checker": "oracle (NumPy restatement, pinned to the reference by tests/golden)"
                   if not fused else "the unfused path (bit_exact: same kernels -> same bits)",
                   "samples": total, "max_norm_err": errs, "bit_exact": exact,
                   "tolerance": f"max|got-want| <= 2*{tol:g}*max|want|",
                   "dw_bitwise_equal_across_ranks": bool(same), "ranks_ok": [bool(k) for k in oks]}
        return res


def measure_workload(h: "Harness", args, cfg, name, steps, warmup, with_fused=True):
    """device-resident timing of one workload (+ per-phase kernels and rooflines on rank 0)"""
    mg = h.mg
    geo = geometry(cfg)
    xh, wh = make_inputs(cfg)
    x_dev, w_dev = mg.asdevice(xh), mg.asdevice(wh)
    del xh, wh
    batch = cfg["x"][0]
    scale_n = h.world if args.scaling == "weak" else 1      # strong: cfg already holds this rank's shard
    rec = {}
    ms, launches, clocks = h.time_steps(h.make_step(cfg, fused=False), x_dev, w_dev, steps, warmup)
    rec.update(ms_per_step=ms, gpu_launches=launches, clocks=clocks)
    fused_ok = with_fused and cfg["pool"] and getattr(mg.nnet.layers, "conv_pool", None) is not None
    if fused_ok:
        fms, fl, _ = h.time_steps(h.make_step(cfg, fused=True), x_dev, w_dev, steps, warmup)
        rec["fused"] = {"ms_per_step": fms, "gpu_launches": fl,
                        "api": "conv_pool(x, w, pool=...) — opt-in fused Operation, same results; y stays on chip"}
    # launch-bound steps (config 1): the same step recorded once as a CUDA graph (mygrad_b200.StepGraph) and replayed
    # with one launch per step.  Single GPU only: the peer-memory collective's epoch counters are kernel arguments.
    if h.world == 1 and name == "C1" and getattr(mg, "StepGraph", None) is not None:
        try:
            rec["graph"] = h.time_graph(h.make_step(cfg, fused=False), x_dev, w_dev, steps, warmup)
        except Exception as e:      # reported, never fatal: the eager numbers above stand
            rec["graph"] = {"error": str(e)[:200]}
    rec["parity"] = h.parity(cfg, fused=False)
    if fused_ok:
        pf = h.parity(cfg, fused=True)
        if h.rank == 0:
            rec["fused"]["parity"] = pf
    if h.rank == 0:
        rec["phases"] = h.time_phases(cfg, geo, x_dev, w_dev)
    return rec, geo, batch, (x_dev, w_dev)


def run_ours(args, cfg, name):
    h = Harness()
    mg, lib, check, dp = h.mg, h.lib, h.check, h.dp
    rank, world = h.rank, h.world
    full_batch = cfg["x"][0]
    if args.scaling == "strong":     # total work fixed: this rank takes its slice of the batch
        from mygrad_b200.dist import shard_bounds

        lo, hi = shard_bounds(full_batch, rank, world)
        cfg = dict(cfg, x=(hi - lo, *cfg["x"][1:]))
    dt = NP_DTYPE[cfg["dtype"]]
    conv_kw = dict(stride=cfg["stride"], padding=cfg["padding"], dilation=cfg["dilation"])
    rec, geo, batch, (x_dev, w_dev) = measure_workload(h, args, cfg, name, args.steps, args.warmup)
    ms_per_step = rec["ms_per_step"]
    job_batch = full_batch if args.scaling == "strong" else world * batch
    value = job_batch / (ms_per_step * 1e-3)

    # ---- end to end with host buffers ------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        # the host-buffer API: pinned NumPy in, pinned NumPy out, batch chunks pipelined over
        # H2D / compute / D2H streams (mygrad_b200/host.py)
        xh, wh = make_inputs(cfg)
        px, pw = mg.pinned_empty(cfg["x"], dt), mg.pinned_empty(cfg["w"], dt)
        px[...] = xh; pw[...] = wh
        del xh, wh
        pipe = mg.HostLayerPipeline(pool=cfg["pool"], pool_stride=2, chunks=args.e2e_chunks, **conv_kw)

        def e2e_step():
            o, dx_h, dw_h = pipe.step(px, pw)
            if world > 1:   # dW of the host API is per rank; sum it like the device path does
                g = mg.asdevice(dw_h)
                dp.allreduce([g])
                g.get(dw_h)
            return o, dx_h, dw_h

        for _ in range(2):
            res = e2e_step()
        h.barrier()
        ke = max(3, min(args.steps, 10))
        t0 = time.perf_counter()
        for _ in range(ke):
            res = e2e_step()
        h.barrier()
        e2e_ms = h.max_over_ranks((time.perf_counter() - t0) * 1e3 / ke)
        e2e = {"value": job_batch / (e2e_ms * 1e-3), "unit": "samples/s", "ms_per_step": e2e_ms,
               "h2d_bytes_per_step": int(px.nbytes + pw.nbytes),
               "d2h_bytes_per_step": int(sum(r.nbytes for r in res)),
               "api": f"HostLayerPipeline.step (chunks={args.e2e_chunks}; H2D/compute/D2H overlapped)"}
        pipe.close()
        # what bounds this number: the pinned-memory PCIe rate of the box, measured here on EVERY rank at
        # once (one direction alone and both at once); the step moves d2h_bytes out while h2d_bytes come in
        nb = 256 << 20
        hb_in, hb_out = mg.pinned_empty((nb // 4,), np.float32), mg.pinned_empty((nb // 4,), np.float32)
        db_in, db_out = mg.empty((nb // 4,), np.float32), mg.empty((nb // 4,), np.float32)
        s1, s2 = C.c_void_p(), C.c_void_p()
        check(lib.mgb_stream_create(C.byref(s1))); check(lib.mgb_stream_create(C.byref(s2)))

        def copies(h2d: bool, d2h: bool, reps: int = 4) -> float:
            def go():
                if h2d:
                    check(lib.mgb_memcpy_h2d(C.c_void_p(db_in.ptr), C.c_void_p(hb_in.ctypes.data), C.c_size_t(nb), s1))
                if d2h:
                    check(lib.mgb_memcpy_d2h(C.c_void_p(hb_out.ctypes.data), C.c_void_p(db_out.ptr), C.c_size_t(nb), s2))
            go()
            check(lib.mgb_stream_synchronize(s1)); check(lib.mgb_stream_synchronize(s2))
            h.barrier()                                   # all ranks copy at the same time
            t0 = time.perf_counter()
            for _ in range(reps):
                go()
            check(lib.mgb_stream_synchronize(s1)); check(lib.mgb_stream_synchronize(s2))
            return nb * reps / (time.perf_counter() - t0) / 1e9

        pcie = {"h2d_gbs": copies(True, False), "d2h_gbs": copies(False, True), "both_each_gbs": copies(True, True)}
        lib.mgb_stream_destroy(s1); lib.mgb_stream_destroy(s2)
        if h.dist is not None:   # every rank's rate with all ranks copying at once (the host side is shared)
            box = [None] * world
            h.dist.all_gather_object(box, pcie)
            pcie = {k: min(b[k] for b in box) for k in pcie}
            pcie["per_rank_both_each_gbs"] = [round(b["both_each_gbs"], 1) for b in box]
            pcie["note"] = "minimum over ranks, all ranks copying concurrently"
        e2e["pcie"] = pcie
        e2e["d2h_gbs_achieved"] = e2e["d2h_bytes_per_step"] / (e2e_ms * 1e-3) / 1e9
        e2e["bound"] = "PCIe: frac = achieved D2H rate / measured D2H rate with both directions busy (all ranks at once)"
        e2e["frac_of_pcie"] = e2e["d2h_gbs_achieved"] / pcie["both_each_gbs"]
        del px, pw, hb_in, hb_out, db_in, db_out

    peaks = h.probe_peaks() if rank == 0 else {}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        kind, b, times = cpu_sample(cfg, budget_s=20.0)
        cpu = {"value": b / times[0], "unit": "samples/s", "cores": host_threads(), "kind": kind,
               "sample": f"batch {b} of {batch} (same layer), one timed step after a 2-sample probe"}

    # ---- the north-star headline (FP64 roofline on the C2 shape) rides along on the default line ----
    c2_f64 = None
    if name == "C2" and not args.no_f64:
        del x_dev, w_dev
        mg.empty_cache()
        cfg64 = dict(WORKLOADS["C2f64"])
        if args.scaling == "strong":
            cfg64["x"] = cfg["x"]
        k64 = max(3, args.steps // 4)
        r64, geo64, b64, _ = measure_workload(h, args, cfg64, "C2f64", k64, 3, with_fused=False)
        if rank == 0:
            roof64 = h.roofline(cfg64, geo64, r64["phases"], peaks, r64["ms_per_step"], "C2f64")
            jb = full_batch if args.scaling == "strong" else world * b64
            c2_f64 = {"workload": workload_name("C2f64", cfg64), "value": jb / (r64["ms_per_step"] * 1e-3),
                      "unit": "samples/s", "ms_per_step": r64["ms_per_step"], "steps": k64, "dtype": "f64",
                      "roofline": roof64, "phases": r64["phases"], "parity": r64["parity"], "clocks": r64["clocks"],
                      "gpu_launches": r64["gpu_launches"],
                      "north_star": ">= 60% of the FP64-tensor roofline over the whole step (roofline.step_frac)"}

    collective = "peer" if (dp.fuses(int(np.prod(cfg["w"])) * geo["esz"])
                            and os.environ.get("MGB_DP_FUSED", "1") != "0") else "nccl"
    if rank == 0:
        roof = h.roofline(cfg, geo, rec["phases"], peaks, ms_per_step, args.workload)
        line = {
            "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": cfg["dtype"], "data": "synthetic",
            "config": {"workload": workload_name(name, dict(cfg, x=(full_batch, *cfg["x"][1:])) if args.scaling == "strong" else cfg),
                       "per_gpu_batch": batch, "job_batch": job_batch,
                       "api": "max_pool(conv_nd(x, w, ...), pool, 2).backward() (drop-in, unfused)",
                       "parallelism": (f"dp{world} (batch shards; dW summed by "
                                       + ("the fused split-K + peer-memory kernel over NVLink" if collective == "peer"
                                          else "ncclAllReduce") + ")") if world > 1 else "single GPU",
                       "l2": "inputs larger than L2 (x+y+dy > 126 MB), no flush between steps"
                             if geo["x_bytes"] + 2 * geo["y_bytes"] > 126e6 else "working set fits L2; not flushed"},
            "clocks": rec["clocks"], "e2e": e2e, "gpu_launches": rec["gpu_launches"],
            "parity_ok": bool(rec["parity"] and rec["parity"]["ok"]), "parity": rec["parity"],
            "roofline": roof, "cpu_baseline": cpu, "phases": rec["phases"], "peaks": peaks,
        }
        if "fused" in rec:
            f = rec["fused"]
            f["value"] = job_batch / (f["ms_per_step"] * 1e-3)
            f["step_frac"] = roof["step_floor_ms"] / f["ms_per_step"]
            line["fused"] = f
        if "graph" in rec:
            gr = rec["graph"]
            if "ms_per_step" in gr:
                gr["value"] = job_batch / (gr["ms_per_step"] * 1e-3)
            line["graph"] = gr
        if c2_f64 is not None:
            line["c2_f64"] = c2_f64
        print(json.dumps(line), flush=True)
    h.close()


def c5_params(dt):
    r = np.random.default_rng(1)
    return dict(
        w1=(r.standard_normal((32, 3, 3, 3)) * 0.1).astype(dt), b1=np.zeros((1, 32, 1, 1), dt),
        w2=(r.standard_normal((64, 32, 3, 3)) * 0.05).astype(dt), b2=np.zeros((1, 64, 1, 1), dt),
        wd=(r.standard_normal((4096, 10)) * 0.01).astype(dt), bd=np.zeros((10,), dt))


def run_c5(args):
    """BASELINE.json config 5: 2x(conv_nd 3x3 p1 + bias + relu + max_pool 2) + dense 4096->10 +
    softmax cross-entropy, batch 2048 per GPU, full Tensor.backward; parameter grads are summed
    over ranks with one all-reduce per tensor on the communication stream."""
    import mygrad_b200 as mg
    from mygrad_b200 import Tensor
    from mygrad_b200._lib import check, lib
    from mygrad_b200.dist import DataParallel, scale_for_mean_loss
    from mygrad_b200.nnet import conv_nd, max_pool, relu, softmax_crossentropy

    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist

        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dp = DataParallel(rank, world, local).init()
    dt = np.float32
    B = C5["batch"]
    xh = np.random.default_rng(100 + rank).standard_normal((B, 3, 32, 32)).astype(dt)
    labels = np.random.default_rng(200 + rank).integers(0, 10, B)
    x_dev = mg.asdevice(xh)
    params = {k: mg.asdevice(v) for k, v in c5_params(dt).items()}

    def step():
        p = {k: Tensor(v, copy=False) for k, v in params.items()}
        x = Tensor(x_dev, copy=False, constant=True)
        h = max_pool(relu(conv_nd(x, p["w1"], stride=1, padding=1) + p["b1"]), (2, 2), 2)
        h = max_pool(relu(conv_nd(h, p["w2"], stride=1, padding=1) + p["b2"]), (2, 2), 2)
        loss = softmax_crossentropy(h.reshape(B, -1) @ p["wd"] + p["bd"], labels)
        with dp.fused_wgrad() as fused:
            loss.backward()
        if world > 1:   # the loss averages over the local batch: rescale so the sum is the global mean
            dp.allreduce([t.grad for t in p.values()], scale=scale_for_mean_loss(B, B * world))
        return loss, p

    def barrier():
        mg.synchronize()
        if dist is not None:
            dist.barrier()
        mg.synchronize()

    stream = C.c_void_p(mg.current_stream())
    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    n0, n1 = C.c_uint64(0), C.c_uint64(0)
    e0, e1 = C.c_void_p(), C.c_void_p()
    check(lib.mgb_event_create(C.byref(e0))); check(lib.mgb_event_create(C.byref(e1)))
    align = mg.zeros((1,), np.float32)      # line the ranks' streams up on the device (see run_ours)
    dp.allreduce([align])
    lib.mgb_launch_count(C.byref(n0))
    check(lib.mgb_event_record(e0, stream))
    for _ in range(args.steps):
        loss, _ = step()
    check(lib.mgb_event_record(e1, stream))
    check(lib.mgb_event_synchronize(e1))
    ms = C.c_float(0)
    check(lib.mgb_event_elapsed_ms(e0, e1, C.byref(ms)))
    barrier()
    lib.mgb_launch_count(C.byref(n1))
    clocks = sampler.stop()
    total_ms = float(ms.value)
    if dist is not None:
        import torch

        t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    # end to end: host batch + labels in, loss out, every step
    px = mg.pinned_empty(xh.shape, dt)
    px[...] = xh
    t0 = time.perf_counter()
    ke = max(3, min(args.steps, 10))
    for _ in range(ke):
        x_dev.set_from_host(px, sync=False)
        loss, _ = step()
        lv = loss.item()
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / ke
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = c5_reference(xh[:256], labels[:256], dt)
    if rank == 0:
        line = {
            "metric": "small CNN (config 5) fwd+bwd samples/sec", "value": world * B / (ms_per_step * 1e-3),
            "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "C5: x=(2048,3,32,32) -> conv3x3(32)+b+relu+pool2 -> conv3x3(64)+b+relu+pool2 -> "
                                   "dense 4096x10 -> softmax_crossentropy; Tensor.backward; x constant",
                       "per_gpu_batch": B,
                       "parallelism": (f"dp{world} (parameter grads: " + ("peer-memory kernels over NVLink"
                                       if dp.peer_slot_bytes > 0 else "ncclAllReduce") + ")") if world > 1 else "single GPU"},
            "clocks": clocks, "gpu_launches": int(n1.value - n0.value), "loss": lv,
            "e2e": {"value": world * B / (e2e_ms * 1e-3), "unit": "samples/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": int(px.nbytes + labels.nbytes), "d2h_bytes_per_step": 4},
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    barrier()
    dp.close()
    if dist is not None:
        dist.destroy_process_group()


def c5_reference(x, labels, dt):
    """the reference (or nothing, if it did not travel) on a 256-sample slice of config 5"""
    ref = ROOT / "baseline" / "_ref"
    if not (ref / "mygrad").exists():
        return None
    sys.path.insert(0, str(ref))
    import mygrad as rm
    from mygrad.nnet.activations import relu as r_relu
    from mygrad.nnet.layers import conv_nd as r_conv, max_pool as r_pool
    from mygrad.nnet.losses import softmax_crossentropy as r_xent

    p = {k: rm.Tensor(v) for k, v in c5_params(dt).items()}

    def step():
        xt = rm.Tensor(x, constant=True)
        h = r_pool(r_relu(r_conv(xt, p["w1"], stride=1, padding=1) + p["b1"]), (2, 2), 2)
        h = r_pool(r_relu(r_conv(h, p["w2"], stride=1, padding=1) + p["b2"]), (2, 2), 2)
        loss = r_xent(h.reshape(x.shape[0], -1) @ p["wd"] + p["bd"], labels)
        loss.backward()
        return float(loss.data)

    step()
    t0 = time.perf_counter()
    step()
    dtm = time.perf_counter() - t0
    return {"value": x.shape[0] / dtm, "unit": "samples/s", "cores": host_threads(), "kind": "reference",
            "sample": f"batch {x.shape[0]} of 2048, one timed step after one warm-up"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C2", choices=sorted(WORKLOADS) + ["C5"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-chunks", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-f64", action="store_true", help="skip the C2-shape float64 sub-record of the default line")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: every GPU runs the workload's batch; strong: the batch is sharded over the GPUs")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.workload == "C5":
        if args.impl == "reference":
            if int(os.environ.get("RANK", 0)) == 0:
                dtp = np.float32
                r = np.random.default_rng(100)
                cpu = c5_reference(r.standard_normal((256, 3, 32, 32)).astype(dtp), r.integers(0, 10, 256), dtp)
                print(json.dumps({"impl": "reference", "metric": "small CNN (config 5) fwd+bwd samples/sec",
                                  "value": cpu and cpu["value"], "unit": "samples/s", "n_gpus": args.gpus,
                                  "cpu_baseline": cpu}), flush=True)
            return
        return run_c5(args)
    cfg = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, cfg, args.workload)
    else:
        run_ours(args, cfg, args.workload)


if __name__ == "__main__":
    main()

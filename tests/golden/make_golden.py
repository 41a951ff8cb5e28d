"""Generates the golden vectors in tests/golden/*.npz by running the LIVE reference
(rsokl/MyGrad, imported from /root/reference/src or baseline/_ref) on seeded inputs.

Run in the build container only (the GPU box has no reference checkout):

    python tests/golden/make_golden.py

Every case stores its inputs, its parameters (JSON) and the reference's outputs
(forward value, x.grad, w.grad as left by Tensor.backward).  The fixtures pin the
behaviours the reference's own tests do not: pool tie-breaking, NaN/inf routing,
overlapping-window accumulation, fp32 / mixed-dtype convolution, strided+dilated dX,
zero-size inputs, and grad accumulation / un-broadcast semantics of Operation.backward.
"""
from __future__ import annotations

import json
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
for cand in (Path("/root/reference/src"), HERE.parent.parent / "baseline" / "_ref"):
    if (cand / "mygrad").exists():
        sys.path.insert(0, str(cand))
        break
import mygrad as mg  # noqa: E402
from mygrad.nnet.layers import conv_nd, max_pool  # noqa: E402
from mygrad.nnet.layers.utils import sliding_window_view  # noqa: E402


def rng_arr(seed, shape, dtype):
    return np.random.default_rng(seed).standard_normal(shape).astype(dtype)


# ---- convolution: value + both gradients for a random upstream grad ------------------
CONV_CASES = [
    # name, x_shape, w_shape, stride, padding, dilation, x_dtype, w_dtype
    ("c1d_basic", (2, 3, 9), (4, 3, 3), 1, 0, 1, "f8", "f8"),
    ("c1d_s2_p1", (3, 2, 10), (5, 2, 4), 2, 1, 1, "f8", "f8"),
    ("c1d_d2", (2, 2, 11), (3, 2, 3), 1, 2, 2, "f8", "f8"),
    ("c2d_mixed_stride", (2, 3, 7, 8), (4, 3, 3, 2), (1, 2), (1, 0), 1, "f8", "f8"),
    ("c2d_s2_p1", (3, 2, 9, 9), (5, 2, 3, 3), 2, 1, 1, "f8", "f8"),
    ("c2d_s2_d2", (2, 2, 13, 13), (3, 2, 3, 3), 2, 0, 2, "f8", "f8"),
    ("c2d_s3_d2_p2", (2, 3, 11, 11), (4, 3, 2, 3), (3, 2), (2, 1), (2, 1), "f8", "f8"),
    ("c2d_1x1", (2, 5, 6, 6), (7, 5, 1, 1), 1, 0, 1, "f8", "f8"),
    ("c2d_f32", (3, 4, 10, 10), (6, 4, 3, 3), 1, 1, 1, "f4", "f4"),
    ("c2d_f32_s2", (2, 8, 12, 12), (16, 8, 4, 4), 2, 1, 1, "f4", "f4"),
    ("c2d_x32_w64", (2, 3, 8, 8), (4, 3, 3, 3), 1, 1, 1, "f4", "f8"),
    ("c2d_x64_w32", (2, 3, 8, 8), (4, 3, 3, 3), 1, 1, 1, "f8", "f4"),
    ("c3d_basic", (2, 2, 5, 6, 7), (3, 2, 2, 3, 2), 1, (0, 1, 0), 1, "f8", "f8"),
    ("c3d_s2_f32", (2, 3, 8, 8, 8), (4, 3, 2, 2, 2), 2, 0, 1, "f4", "f4"),
    ("c3d_d2", (1, 2, 9, 7, 8), (2, 2, 2, 2, 3), 1, 1, (2, 1, 2), "f8", "f8"),
    ("c2d_wide_channels", (2, 40, 6, 6), (70, 40, 3, 3), 1, 1, 1, "f8", "f8"),
    ("c2d_tall_batch", (9, 2, 17, 19), (3, 2, 5, 3), 1, (2, 1), 1, "f8", "f8"),
    ("config1_small_batch", (4, 3, 32, 32), (16, 3, 3, 3), 1, 1, 1, "f8", "f8"),
]


def run_conv_cases():
    out, manifest = {}, {}
    for i, (name, xs, ws, s, p, d, xdt, wdt) in enumerate(CONV_CASES):
        x = rng_arr(100 + i, xs, xdt)
        w = rng_arr(200 + i, ws, wdt)
        xt, wt = mg.Tensor(x), mg.Tensor(w)
        y = conv_nd(xt, wt, stride=s, padding=p, dilation=d)
        g = rng_arr(300 + i, y.shape, y.dtype)
        y.backward(g)
        out[f"{name}/x"], out[f"{name}/w"], out[f"{name}/g"] = x, w, g
        out[f"{name}/y"], out[f"{name}/dx"], out[f"{name}/dw"] = y.data, xt.grad, wt.grad
        manifest[name] = {"stride": s, "padding": p, "dilation": d}
    # zero-size spatial input that only exists through padding (test_conv.py:247-273)
    for name, nd, pad in (("zero_size_1d", 1, 2), ("zero_size_2d", 2, (1, 3))):
        pads = (pad,) * nd if isinstance(pad, int) else pad
        x = np.zeros((1, 1) + (0,) * nd)
        w = rng_arr(7, (1, 1) + tuple(2 * q for q in pads), "f8")
        xt, wt = mg.Tensor(x), mg.Tensor(w)
        y = conv_nd(xt, wt, stride=1, padding=pad)
        y.backward()
        out[f"{name}/x"], out[f"{name}/w"], out[f"{name}/g"] = x, w, np.ones(y.shape)
        out[f"{name}/y"], out[f"{name}/dx"], out[f"{name}/dw"] = y.data, xt.grad, wt.grad
        manifest[name] = {"stride": 1, "padding": pad, "dilation": 1}
    np.savez_compressed(HERE / "conv_cases.npz", manifest=json.dumps(manifest), **out)
    return len(manifest)


# ---- pooling: value + gradient, including ties / NaN / inf / overlaps ----------------
def pool_inputs():
    r = np.random.default_rng(7)
    cases = []
    # (name, x, pool, stride)
    cases.append(("p1d_overlap", r.standard_normal((2, 3, 11)).astype("f4"), (3,), 2))
    cases.append(("p2d_tiled", r.standard_normal((2, 3, 8, 6)).astype("f8"), (2, 2), 2))
    cases.append(("p2d_rect_overlap", r.standard_normal((2, 2, 9, 10)).astype("f4"), (3, 4), (2, 3)))
    cases.append(("p2d_gaps", r.standard_normal((2, 2, 11, 11)).astype("f8"), (2, 2), 3))
    cases.append(("p3d_tiled", r.standard_normal((2, 2, 4, 6, 8)).astype("f4"), (2, 2, 2), 2))
    cases.append(("p3d_mixed", r.standard_normal((3, 5, 7, 9)).astype("f8"), (2, 3, 3), (1, 2, 3)))
    cases.append(("p_no_lead", r.standard_normal((6, 8)).astype("f8"), (2, 2), 2))
    cases.append(("p_trailing_only", r.standard_normal((3, 4, 5, 12)).astype("f4"), (4,), 4))
    # ties: small integer alphabet, overlapping and tiled windows
    ties = r.integers(0, 3, size=(2, 2, 8, 8)).astype("f4")
    cases.append(("p2d_ties_tiled", ties, (2, 2), 2))
    cases.append(("p2d_ties_overlap", ties, (3, 3), 1))
    cases.append(("p2d_all_equal", np.ones((1, 1, 4, 4), "f8"), (2, 2), 2))
    # NaN / inf routing
    special = r.standard_normal((2, 1, 6, 6)).astype("f8")
    special[0, 0, 0, 1] = np.nan
    special[0, 0, 1, 0] = np.nan
    special[0, 0, 2, 3] = np.inf
    special[0, 0, 4, 4] = -np.inf
    special[1, 0, 3, 3] = np.nan
    special[1, 0, 5, 0] = np.inf
    special[1, 0, 5, 1] = np.inf
    cases.append(("p2d_special_tiled", special, (2, 2), 2))
    cases.append(("p2d_special_overlap", special, (2, 2), 1))
    cases.append(("p2d_special_f32", special.astype("f4"), (3, 3), 3))
    # C1-shaped pool input at reduced batch
    cases.append(("config1_pool", r.standard_normal((4, 16, 32, 32)).astype("f8"), (2, 2), 2))
    return cases


def ref_mean_pool(xt, pool, st):
    """mean-pool composed ENTIRELY of the reference's own autodiff ops: strided basic-index slices of a
    Tensor (one per window offset; tensor_base.py __getitem__), added and divided by the window size.
    Its ``backward`` is therefore the reference's gradient of mean-pooling."""
    import itertools

    nd = len(pool)
    G = [(X - P) // s + 1 for X, P, s in zip(xt.shape[xt.ndim - nd:], pool, st)]
    acc = None
    for offs in itertools.product(*[range(p) for p in pool]):
        idx = (Ellipsis,) + tuple(slice(o, o + s * (g - 1) + 1, s) for o, s, g in zip(offs, st, G))
        term = xt[idx]
        acc = term if acc is None else acc + term
    return acc / float(np.prod(pool))


def run_pool_cases():
    out, manifest = {}, {}
    for i, (name, x, pool, stride) in enumerate(pool_inputs()):
        xt = mg.Tensor(x)
        y = max_pool(xt, pool, stride)
        g = np.random.default_rng(400 + i).standard_normal(y.shape).astype(x.dtype)
        y.backward(g)
        # mean-pool has no reference function: the composed definition
        nd = len(pool)
        st = (stride,) * nd if isinstance(stride, int) else stride
        win = sliding_window_view(x, pool, st)                       # (G..., lead..., P...)
        mean = win.mean(axis=tuple(range(-nd, 0)))
        axes = tuple(range(mean.ndim))
        lead_nd = x.ndim - nd
        mean = np.ascontiguousarray(mean.transpose(axes[nd:] + axes[:nd])) if lead_nd else mean
        out[f"{name}/x"], out[f"{name}/g"] = x, g
        out[f"{name}/y"], out[f"{name}/dx"], out[f"{name}/mean"] = y.data, xt.grad, mean
        if np.isfinite(x).all():      # gradient of mean-pooling from the reference's autodiff
            xm = mg.Tensor(x)
            ym = ref_mean_pool(xm, pool, st)
            assert ym.shape == y.shape
            np.testing.assert_allclose(ym.data, mean, rtol=1e-5 if x.dtype == np.float32 else 1e-12, atol=1e-6)
            ym.backward(g)
            out[f"{name}/mean_dx"] = xm.grad
        manifest[name] = {"pool": pool, "stride": stride}
    np.savez_compressed(HERE / "pool_cases.npz", manifest=json.dumps(manifest), **out)
    return len(manifest)


# ---- whole layer through the reference's Tensor graph ---------------------------------
LAYER_CASES = [
    # name, x_shape, w_shape, stride, padding, dilation, pool, pool_stride, dtype
    ("layer_c1_like", (4, 3, 16, 16), (8, 3, 3, 3), 1, 1, 1, (2, 2), 2, "f8"),
    ("layer_c2_like_f32", (2, 8, 12, 12), (16, 8, 3, 3), 1, 1, 1, (2, 2), 2, "f4"),
    ("layer_c3_like", (2, 4, 13, 13), (6, 4, 3, 3), 2, 0, 2, None, None, "f8"),
    ("layer_c4_like_f32", (2, 3, 6, 8, 8), (4, 3, 3, 3, 3), 1, 0, 1, (2, 2, 2), 2, "f4"),
    ("layer_overlap_pool", (2, 3, 9, 9), (4, 3, 3, 3), 1, 0, 1, (3, 3), 2, "f8"),
    # >= 16 channels: the tcgen05 kernels, and the geometry of the fused conv_pool Operation
    ("layer_tc_f32", (3, 16, 12, 12), (32, 16, 3, 3), 1, 1, 1, (2, 2), 2, "f4"),
    ("layer_tc_dilated_f32", (2, 32, 10, 14), (16, 32, 3, 3), 1, 2, 2, (2, 2), 2, "f4"),
]


def run_layer_cases():
    out, manifest = {}, {}
    for i, (name, xs, ws, s, p, d, pool, ps, dt) in enumerate(LAYER_CASES):
        x, w = rng_arr(500 + i, xs, dt), rng_arr(600 + i, ws, dt)
        xt, wt = mg.Tensor(x), mg.Tensor(w)
        y = conv_nd(xt, wt, stride=s, padding=p, dilation=d)
        o = max_pool(y, pool, ps) if pool else y
        o.backward()  # all-ones seed, like bench.py's step
        out[f"{name}/x"], out[f"{name}/w"] = x, w
        out[f"{name}/out"], out[f"{name}/conv"], out[f"{name}/dconv"] = o.data, y.data, y.grad
        out[f"{name}/dx"], out[f"{name}/dw"] = xt.grad, wt.grad
        manifest[name] = {"stride": s, "padding": p, "dilation": d, "pool": pool, "pool_stride": ps}
    np.savez_compressed(HERE / "layer_cases.npz", manifest=json.dumps(manifest), **out)
    return len(manifest)


# ---- einsum through the reference's EinSum (linalg/ops.py:79-243) ------------------------------
EINSUM_CASES = [
    # name, subscripts, operand shapes, dtype, same_tensor (operand indices that are ONE tensor)
    ("matmul", "ij,jk->ik", [(4, 5), (5, 3)], "f8", None),
    ("matmul_implicit", "ij,jk", [(3, 4), (4, 6)], "f8", None),
    ("transpose_implicit", "ba", [(2, 3)], "f8", None),
    ("trace", "ii", [(5, 5)], "f8", None),
    ("diag", "ii->i", [(4, 4)], "f8", None),
    ("docstring_ijk_k", "ijk,k->ji", [(2, 3, 4), (4,)], "f8", None),
    ("sum_without_partner", "ijk,jk->k", [(3, 2, 4), (2, 4)], "f8", None),
    ("full_sum", "ij->", [(3, 4)], "f8", None),
    ("inner", "i,i", [(7,), (7,)], "f8", None),
    ("outer", "i,j->ij", [(3,), (4,)], "f8", None),
    ("three_operands", "ij,jk,kl->il", [(2, 3), (3, 4), (4, 5)], "f8", None),
    ("tensor_product", "ijk,jil->kl", [(3, 4, 5), (4, 3, 2)], "f8", None),
    ("trace_repeat_kept", "iji->ij", [(3, 4, 3)], "f8", None),
    ("double_trace", "ijkji,k->jk", [(2, 3, 4, 3, 2), (4,)], "f8", None),
    ("ellipsis_matmul", "...ij,...jk->...ik", [(5, 1, 2, 3), (4, 3, 2)], "f8", None),
    ("ellipsis_trace", "i...i", [(3, 2, 3)], "f8", None),
    ("ellipsis_diag", "...ii->...i", [(2, 3, 3)], "f8", None),
    ("ellipsis_middle", "ij...,jk...->ik...", [(2, 3, 4), (3, 5, 4)], "f8", None),
    ("same_tensor_sq_norm", "...,...->", [(3, 4), (3, 4)], "f8", (0, 1)),
    ("same_tensor_two_labels", "ij,ji->", [(3, 3), (3, 3)], "f8", (0, 1)),
    ("same_tensor_matmul", "ij,jk->ik", [(3, 3), (3, 3)], "f8", (0, 1)),
    ("scaled", ",ij->ij", [(), (2, 3)], "f8", None),
    ("f32_matmul", "ij,jk->ik", [(8, 16), (16, 4)], "f4", None),
    ("f32_batched", "bij,bjk->bik", [(3, 4, 5), (3, 5, 2)], "f4", None),
    ("upper_case_labels", "Ab,bC->AC", [(2, 3), (3, 4)], "f8", None),
]


def run_einsum_cases():
    out, manifest = {}, {}
    for i, (name, subs, shapes, dt, same) in enumerate(EINSUM_CASES):
        arrs = [rng_arr(900 + 10 * i + k, s, dt) for k, s in enumerate(shapes)]
        ts = [mg.Tensor(a) for a in arrs]
        if same:
            for k in same[1:]:
                ts[k] = ts[same[0]]
                arrs[k] = arrs[same[0]]
        y = mg.einsum(subs, *ts)
        g = np.random.default_rng(950 + i).standard_normal(y.shape).astype(dt)
        y.backward(g)
        out[f"{name}/y"], out[f"{name}/g"] = np.asarray(y.data), g
        for k, (a, t) in enumerate(zip(arrs, ts)):
            out[f"{name}/op{k}"], out[f"{name}/grad{k}"] = a, t.grad
        manifest[name] = {"subscripts": subs, "nops": len(shapes), "same": same}
    np.savez_compressed(HERE / "einsum_cases.npz", manifest=json.dumps(manifest), **out)
    return len(manifest)


# ---- Operation.backward semantics: accumulation, un-broadcast, dtype casting -------------
def run_graph_cases():
    out = {}
    r = np.random.default_rng(11)
    # (1) one tensor consumed by two convs and a pool: grads accumulate (+=)
    x = r.standard_normal((2, 2, 8, 8))
    w1, w2 = r.standard_normal((3, 2, 3, 3)), r.standard_normal((3, 2, 3, 3))
    xt, w1t, w2t = mg.Tensor(x), mg.Tensor(w1), mg.Tensor(w2)
    total = conv_nd(xt, w1t, stride=1, padding=1) + conv_nd(xt, w2t, stride=1, padding=1)
    o = max_pool(total, (2, 2), 2)
    o.backward()
    out.update({"acc/x": x, "acc/w1": w1, "acc/w2": w2, "acc/out": o.data, "acc/dx": xt.grad,
                "acc/dw1": w1t.grad, "acc/dw2": w2t.grad, "acc/dtotal": total.grad})
    # (2) broadcast bias: un-broadcast reduction (N,F,H,W) -> (1,F,1,1) and -> (F,1,1)
    b4, b3 = r.standard_normal((1, 3, 1, 1)), r.standard_normal((3, 1, 1))
    xt, w1t, b4t, b3t = mg.Tensor(x), mg.Tensor(w1), mg.Tensor(b4), mg.Tensor(b3)
    y = conv_nd(xt, w1t, stride=1, padding=1) + b4t + b3t
    g = r.standard_normal(y.shape)
    y.backward(g)
    out.update({"bias/b4": b4, "bias/b3": b3, "bias/g": g, "bias/y": y.data,
                "bias/db4": b4t.grad, "bias/db3": b3t.grad, "bias/dx": xt.grad})
    # (3) reduce_broadcast shape algebra on bare arrays
    from mygrad._utils import reduce_broadcast
    shapes = [((4, 3, 5, 6), (3, 1, 6)), ((4, 3, 5, 6), (1, 1, 1, 1)), ((4, 3, 5, 6), ()),
              ((7, 5), (5,)), ((7, 5), (7, 1)), ((2, 3, 4, 5, 6), (3, 1, 5, 1)), ((6, 1, 4), (6, 1, 4)),
              ((3, 4, 5), (4, 5)), ((64, 33, 7, 9), (1, 33, 1, 1))]
    for i, (gs, vs) in enumerate(shapes):
        a = r.standard_normal(gs)
        out[f"rb{i}/g"], out[f"rb{i}/out"] = a, np.asarray(reduce_broadcast(a, vs))
        out[f"rb{i}/vshape"] = np.array(vs, dtype=np.int64)
    out["rb_count"] = np.array(len(shapes))
    # (4) fp32 pool grad: float64 scatter buffer cast back to float32 by Operation.backward
    xf = r.standard_normal((2, 2, 7, 7)).astype("f4")
    xt = mg.Tensor(xf)
    o = max_pool(xt, (3, 3), 2)
    gf = r.standard_normal(o.shape).astype("f4")
    o.backward(gf)
    out.update({"pool32/x": xf, "pool32/g": gf, "pool32/dx": xt.grad})
    np.savez_compressed(HERE / "graph_cases.npz", **out)
    return 4


# ---- config-5-like small CNN through the reference (scaled down) -----------
def run_cnn_case():
    from mygrad.nnet.activations import relu
    from mygrad.nnet.losses import softmax_crossentropy

    r = np.random.default_rng(21)
    out = {}
    for tag, dt in (("f32", "f4"), ("f64", "f8")):
        x = r.standard_normal((8, 3, 16, 16)).astype(dt)
        labels = r.integers(0, 10, 8)
        w1 = (r.standard_normal((16, 3, 3, 3)) * 0.1).astype(dt)
        b1 = (r.standard_normal((1, 16, 1, 1)) * 0.1).astype(dt)
        w2 = (r.standard_normal((32, 16, 3, 3)) * 0.05).astype(dt)
        b2 = (r.standard_normal((1, 32, 1, 1)) * 0.1).astype(dt)
        wd = (r.standard_normal((32 * 4 * 4, 10)) * 0.01).astype(dt)
        bd = (r.standard_normal((10,)) * 0.1).astype(dt)
        params = {k: mg.Tensor(v) for k, v in dict(w1=w1, b1=b1, w2=w2, b2=b2, wd=wd, bd=bd).items()}
        xt = mg.Tensor(x)
        h = max_pool(relu(conv_nd(xt, params["w1"], stride=1, padding=1) + params["b1"]), (2, 2), 2)
        h = max_pool(relu(conv_nd(h, params["w2"], stride=1, padding=1) + params["b2"]), (2, 2), 2)
        logits = h.reshape(8, -1) @ params["wd"] + params["bd"]
        loss = softmax_crossentropy(logits, labels)
        loss.backward()
        out[f"{tag}/x"], out[f"{tag}/labels"] = x, labels
        out[f"{tag}/loss"], out[f"{tag}/logits"], out[f"{tag}/dx"] = loss.data, logits.data, xt.grad
        for k, t in params.items():
            out[f"{tag}/{k}"], out[f"{tag}/d{k}"] = t.data, t.grad
    np.savez_compressed(HERE / "cnn_case.npz", **out)
    return 2




# ---- batchnorm (src/mygrad/nnet/layers/batchnorm.py): value, saved moments, all three grads ---------
BN_CASES = [
    # name, x_shape, dtype, gamma?, beta?, eps
    ("bn_2d_affine", (6, 4), "f8", True, True, 1e-8),
    ("bn_docstring", (3, 1), "f8", False, False, 0.0),
    ("bn_3d_gamma_only", (4, 3, 5), "f8", True, False, 1e-5),
    ("bn_4d_affine", (5, 6, 7, 8), "f8", True, True, 1e-5),
    ("bn_4d_plain", (3, 4, 6, 6), "f8", False, False, 1e-3),
    ("bn_4d_beta_only", (3, 4, 5, 6), "f8", False, True, 1e-6),
    ("bn_4d_f32", (8, 16, 12, 12), "f4", True, True, 1e-5),
    ("bn_5d_f32", (2, 3, 4, 5, 6), "f4", True, True, 1e-5),
    ("bn_offset_mean", (16, 3, 9, 9), "f8", True, True, 1e-8),     # mean >> std: variance must not cancel
    ("bn_odd_run", (3, 5, 7, 3), "f4", True, True, 1e-5),          # run length 21: scalar (unvectorised) path
]


def run_bn_cases():
    from mygrad.nnet.layers import batchnorm

    out, manifest = {}, {}
    for i, (name, xs, dt, has_g, has_b, eps) in enumerate(BN_CASES):
        x = rng_arr(700 + i, xs, dt)
        if name == "bn_docstring":
            x = np.array([1.0, 4.0, 1.0]).reshape(3, 1)          # batchnorm.py:152-157
        if name == "bn_offset_mean":
            x = x * 1e-3 + 1e3
        c = xs[1]
        gamma = mg.Tensor(rng_arr(800 + i, (c,), dt) + 1) if has_g else None
        beta = mg.Tensor(rng_arr(900 + i, (c,), dt)) if has_b else None
        xt = mg.Tensor(x)
        y = batchnorm(xt, gamma=gamma, beta=beta, eps=eps)
        op = y.creator
        g = rng_arr(1000 + i, y.shape, dt)
        y.backward(g)
        out[f"{name}/x"], out[f"{name}/g"], out[f"{name}/y"], out[f"{name}/dx"] = x, g, y.data, xt.grad
        out[f"{name}/mean"], out[f"{name}/var"] = op.mean, op.var
        if has_g:
            out[f"{name}/gamma"], out[f"{name}/dgamma"] = gamma.data, gamma.grad
        if has_b:
            out[f"{name}/beta"], out[f"{name}/dbeta"] = beta.data, beta.grad
        manifest[name] = {"eps": eps}
    np.savez_compressed(HERE / "batchnorm_cases.npz", manifest=json.dumps(manifest), **out)
    return len(manifest)


# ---- on-disk format and the public window view ---------------------------------
def run_io_and_window_cases():
    # a file written by the reference's own mg.save (src/mygrad/_io.py:54-63): data + grad keys
    x = mg.arange(10.0)
    (x * x).backward()
    mg.save(HERE / "saved_by_reference.npz", x)
    y = mg.Tensor(rng_arr(41, (3, 4), "f4"))
    mg.save(HERE / "saved_by_reference_nograd.npz", y)
    # sliding_window_view values (utils.py:7-228) incl. the docstring example (utils.py:60-97)
    out, manifest = {}, {}
    cases = [
        ("doc_example", np.arange(36.0).reshape(6, 6), (3, 2), (2, 2), None),
        ("w1d_step2", rng_arr(42, (3, 11), "f8"), (4,), 2, None),
        ("w2d_dilated", rng_arr(43, (2, 3, 9, 10), "f4"), (2, 3), (2, 1), (3, 2)),
        ("w3d", rng_arr(44, (2, 5, 6, 7), "f8"), (2, 3, 2), (1, 2, 3), None),
        ("w3d_dil_int", rng_arr(45, (7, 8, 9), "f4"), (2, 2, 2), 2, 2),
        ("no_lead", rng_arr(46, (8, 9), "f8"), (3, 3), 3, None),
    ]
    for name, arr, win, step, dil in cases:
        out[f"{name}/x"] = arr
        out[f"{name}/view"] = np.ascontiguousarray(sliding_window_view(arr, win, step, dil))
        manifest[name] = {"window_shape": win, "step": step, "dilation": dil}
    np.savez_compressed(HERE / "window_cases.npz", manifest=json.dumps(manifest), **out)
    return len(manifest)


if __name__ == "__main__" and "--einsum-only" in sys.argv:
    print("einsum cases:", run_einsum_cases())
    sys.exit(0)

if __name__ == "__main__":
    print("reference:", mg.__file__)
    print("io / window cases:", run_io_and_window_cases())
    print("batchnorm cases:", run_bn_cases())
    print("conv cases:", run_conv_cases())
    print("pool cases:", run_pool_cases())
    print("layer cases:", run_layer_cases())
    print("graph cases:", run_graph_cases())
    print("cnn cases:", run_cnn_case())
    print("einsum cases:", run_einsum_cases())
    for f in sorted(HERE.glob("*.npz")):
        print(f.name, f.stat().st_size, "bytes")

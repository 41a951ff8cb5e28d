"""StepGraph: a recorded step replays bit-identically, follows in-place input updates, and refuses steps whose
allocations do not repeat."""
import numpy as np
import pytest

import mygrad_b200 as mg
from mygrad_b200 import StepGraph, Tensor
from mygrad_b200.nnet.layers import conv_nd, max_pool
from oracle import conv_pool_oracle as oc

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype,shapes", [(np.float64, ((8, 3, 16, 16), (16, 3, 3, 3))),
                                          (np.float32, ((4, 32, 12, 12), (32, 32, 3, 3)))])
def test_replay_matches_eager_and_follows_input_updates(dtype, shapes):
    xs, ws = shapes
    rng = np.random.default_rng(3)
    xh = rng.standard_normal(xs).astype(dtype)
    wh = rng.standard_normal(ws).astype(dtype)
    x_dev, w_dev = mg.asdevice(xh), mg.asdevice(wh)

    def step():
        x, w = Tensor(x_dev, copy=False), Tensor(w_dev, copy=False)
        out = max_pool(conv_nd(x, w, stride=1, padding=1), (2, 2), 2)
        out.backward()
        return out, x, w

    out, x, w = step()
    eager = (out.numpy(), x.grad.get(), w.grad.get())
    g = StepGraph(step)
    assert g.kernel_nodes >= 5
    for _ in range(3):
        out, x, w = g.replay()
    g.synchronize()
    for got, want in zip((out.numpy(), x.grad.get(), w.grad.get()), eager):
        np.testing.assert_array_equal(got, want)          # same kernels, same buffers: bit-identical
    # new input, same buffers: the replay computes the new step
    xh2 = rng.standard_normal(xs).astype(dtype)
    x_dev.set_from_host(xh2)
    mg.synchronize()
    out, x, w = g.replay()
    g.synchronize()
    want_out, want_dx, want_dw = oc.layer_step(xh2.astype(np.float64), wh.astype(np.float64), stride=1, padding=1,
                                               dilation=1, pool=(2, 2), pool_stride=2)
    rtol = 1e-10 if dtype == np.float64 else 1e-5
    for got, want in ((out.numpy(), want_out), (x.grad.get(), want_dx), (w.grad.get(), want_dw)):
        np.testing.assert_allclose(got, want, rtol=rtol, atol=rtol * float(np.abs(want).max()))
    g.close()


def test_capture_refuses_a_step_whose_allocations_change():
    sizes = iter([64, 64, 4096])

    def step():
        return mg.zeros((next(sizes),), np.float32)

    with pytest.raises(RuntimeError, match="StepGraph"):
        StepGraph(step, warmup=2)
    # the library is usable afterwards
    a = mg.zeros((8,), np.float32)
    assert float(a.get().sum()) == 0.0

"""einsum against golden vectors written by the LIVE reference's ``mygrad.einsum``
(tests/golden/make_golden.py: value and every operand's gradient for a random upstream gradient) — matrix
products, traces and diagonals, ellipsis broadcasting, labels summed without a partner, one tensor fed
several times — plus NumPy's own einsum on extra forward cases and the error conventions."""
import numpy as np
import pytest

import mygrad_b200 as mg
from mygrad_b200 import Tensor, einsum

from conftest import GoldenFile, assert_close

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", GoldenFile("einsum_cases.npz").names())
def test_einsum_vs_reference_golden(name):
    gf = GoldenFile("einsum_cases.npz")
    c, m = gf.case(name), gf.manifest[name]
    ops = [Tensor(c[f"op{k}"]) for k in range(m["nops"])]
    if m["same"]:
        for k in m["same"][1:]:
            ops[k] = ops[m["same"][0]]
    y = einsum(m["subscripts"], *ops)
    dt = c["y"].dtype
    assert y.shape == c["y"].shape and y.dtype == dt
    assert_close(y.numpy(), c["y"], dt, err_msg=name)
    y.backward(c["g"])
    for k, t in enumerate(ops):
        want = c[f"grad{k}"]
        assert t.grad.shape == want.shape and t.grad.dtype == want.dtype, (name, k)
        assert_close(t.grad.get(), want, dt, scale=max(1.0, float(np.abs(want).max())), err_msg=f"{name} grad{k}")


FWD_CASES = [
    ("ij,jk->ik", [(64, 300), (300, 48)]),          # long sums, many outputs: one thread per output
    ("ij->", [(200, 300)]),                          # one output, long sum: one CTA per output
    ("bij,bjk->bik", [(4, 16, 32), (4, 32, 8)]),
    ("ii->", [(700, 700)]),
    ("i,i->", [(100000,), (100000,)]),
    ("ijkl->lkji", [(3, 4, 5, 6)]),
    ("...i,...i->...", [(6, 1, 40), (5, 40)]),
    ("ab,cd->abcd", [(2, 3), (4, 5)]),
]


@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("subs,shapes", FWD_CASES, ids=[c[0] for c in FWD_CASES])
def test_einsum_forward_vs_numpy(subs, shapes, dt):
    rng = np.random.default_rng(len(subs))
    arrs = [rng.standard_normal(s).astype(dt) for s in shapes]
    want = np.einsum(subs, *[a.astype(np.float64) for a in arrs])
    got = einsum(subs, *arrs)
    assert got.dtype == dt and got.constant          # NumPy operands are constants
    scale = float(np.abs(np.einsum(subs, *[np.abs(a.astype(np.float64)) for a in arrs])).max())
    assert_close(got.numpy(), want.astype(dt), dt, scale=scale, err_msg=subs)


def test_einsum_sublist_form_mixed_dtypes_and_constants():
    rng = np.random.default_rng(0)
    a, b = rng.standard_normal((2, 3)).astype(np.float32), rng.standard_normal((3, 4))
    ta, tb = Tensor(a), Tensor(b, constant=True)
    y = einsum(ta, [0, 1], tb, [1, 2], [0, 2])
    assert y.dtype == np.float64 and not y.constant
    np.testing.assert_allclose(y.numpy(), a.astype(np.float64) @ b, rtol=1e-12)
    y.backward()
    assert tb.grad is None and ta.grad.dtype == np.float32
    np.testing.assert_allclose(ta.grad.get(), (np.ones((2, 4)) @ b.T).astype(np.float32), rtol=1e-6)
    assert einsum("ij,jk", tb, Tensor(b.T.copy(), constant=True)).constant
    # an ellipsis the explicit output omits is summed away, as in the reference ("...,...->", funcs.py docstring)
    np.testing.assert_allclose(einsum("...i->i", np.ones((2, 3))).numpy(), [2.0, 2.0, 2.0])
    # implicit output sorts labels; the ellipsis goes first
    x = rng.standard_normal((2, 3, 4))
    np.testing.assert_allclose(einsum("...ab->...ba", x).numpy(), np.einsum("...ab->...ba", x))
    np.testing.assert_allclose(einsum("b...a", x).numpy(), np.einsum("b...a", x))


@pytest.mark.parametrize("args", [
    ("ij->ii", np.ones((2, 2))),
    ("ij,jk->il", np.ones((2, 2)), np.ones((2, 2))),
    ("ij,jk", np.ones((2, 2))),
    ("ijk", np.ones((2, 2))),
    ("ij,jk->ik", np.ones((2, 3)), np.ones((4, 2))),
    ("i.j", np.ones((2, 3))),
    ("i->>i", np.ones(3)),
    ("i$", np.ones(3)),
])
def test_einsum_errors_match_numpy(args):
    with pytest.raises(ValueError):
        np.einsum(*args)
    with pytest.raises(ValueError):
        einsum(*args)

"""GPU parity of compute_cross_distances on the tensor cores (engine 1 of yb_cross_distances_l2:
split-precision FP16 operands, both norms folded into the contraction; yael/nn.c:100-129).

Tolerance (BASELINE.json north_star): L2 distances within 1e-5 relative of the reference's CPU
implementation.  The reference's own value is fl32(|a|^2 + |b|^2) - 2 <a,b> in FP32, i.e. it carries
an absolute rounding error of a few ulp of |a|^2 + |b|^2 itself, so the bound used here is
1e-5 * dist + 4 ulp(|a|^2 + |b|^2): relative for ordinary pairs, absolute at the scale of the
reference's own cancellation error for near-duplicates."""
import numpy as np
import pytest

import yael_b200

pytestmark = pytest.mark.gpu


@pytest.fixture
def cross_tensor():
    L = yael_b200.lib()
    L.yb_set_cross_engine(1)
    yield L
    L.yb_set_cross_engine(-1)


def _check(got, a, b, want):
    a64, b64 = a.astype(np.float64), b.astype(np.float64)
    na2, nb2 = (a64 * a64).sum(1), (b64 * b64).sum(1)
    exact = nb2[:, None] + na2[None, :] - 2.0 * b64 @ a64.T
    scale = nb2[:, None] + na2[None, :]
    tol = 1e-5 * np.abs(exact) + 4 * 2.0 ** -23 * scale
    assert np.isfinite(got).all()
    assert (np.abs(got - exact) <= tol).all(), float((np.abs(got - exact) / tol).max())
    # and against the oracle's FP32 value (which has the same class of error against exact)
    assert (np.abs(got - want) <= 2 * tol).all()


@pytest.mark.parametrize("na,nb,d", [(300, 5000, 128), (1000, 3000, 64), (257, 1030, 100), (512, 2048, 960),
                                     (130, 700, 20), (400, 4000, 129)])
def test_cross_distances_tensor_engine_uniform(yn, ob, cross_tensor, na, nb, d):
    r = np.random.RandomState(na + nb + d)
    a = r.rand(na, d).astype(np.float32)
    b = r.rand(nb, d).astype(np.float32)
    got = yn.cross_distances(a, b)
    assert cross_tensor.yb_last_cross_engine() == 1
    _check(got, a, b, ob.orc_cross(a, b, ob.DOT_F32_SEQ))


def test_cross_distances_tensor_engine_sift_like(yn, ob, cross_tensor):
    r = np.random.RandomState(3)
    a = np.minimum(255, r.gamma(1.2, 25.0, (300, 128))).astype(np.int32).astype(np.float32)
    b = np.minimum(255, r.gamma(1.2, 25.0, (6000, 128))).astype(np.int32).astype(np.float32)
    got = yn.cross_distances(a, b)
    assert cross_tensor.yb_last_cross_engine() == 1
    _check(got, a, b, ob.orc_cross(a, b, ob.DOT_F32_SEQ))


def test_cross_distances_tensor_engine_scales_and_offsets(yn, ob, cross_tensor):
    # a large common offset (centring removes it) and a tiny scale (the power-of-two scale handles it)
    r = np.random.RandomState(4)
    for scale, shift in ((1e-4, 0.0), (300.0, 1000.0), (1.0, -7.5)):
        a = (r.randn(260, 48) * scale + shift).astype(np.float32)
        b = (r.randn(2000, 48) * scale + shift).astype(np.float32)
        got = yn.cross_distances(a, b)
        assert cross_tensor.yb_last_cross_engine() == 1
        a64, b64 = a.astype(np.float64), b.astype(np.float64)
        exact = ((b64[:, None, :] - a64[None, :, :]) ** 2).sum(2)
        # relative to the distance itself: the centred operands do not cancel like the reference's formula
        assert (np.abs(got - exact) <= 1e-5 * exact + 1e-30).all()


def test_cross_distances_out_of_fp16_range_falls_back(yn, ob, cross_tensor):
    r = np.random.RandomState(5)
    a = r.rand(300, 32).astype(np.float32)
    b = r.rand(40000, 32).astype(np.float32)
    b[-1, 3] = 3e6        # outside the sampled rows: the scale does not cover it
    got = yn.cross_distances(a, b)
    assert cross_tensor.yb_last_cross_engine() == 0
    assert np.array_equal(got, ob.orc_cross(a, b, ob.DOT_F32_SEQ))


def test_cross_distances_small_or_strided_stay_exact(yn, ob):
    # automatic engine: small problems keep the bit-exact FP32 engine
    L = yael_b200.lib()
    r = np.random.RandomState(6)
    a = r.rand(100, 32).astype(np.float32)
    b = r.rand(500, 32).astype(np.float32)
    got = yn.cross_distances(a, b)
    assert L.yb_last_cross_engine() == 0
    assert np.array_equal(got, ob.orc_cross(a, b, ob.DOT_F32_SEQ))

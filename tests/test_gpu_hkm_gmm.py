"""GPU parity of the remaining consumers of SURVEY.md 8(f)-N4 through the drop-in C API:
hkm_quantize / hkm_learn (yael/hkm.c) and the GMM E-step gmm_compute_p (yael/gmm.c:211-367),
against the compiled-reference golden (tests/golden/hkm_gmm.npz) and the oracle.

Bars: leaves bit-identical (integer result of exact k = 1 searches with the reference's
arithmetic); posteriors within 1e-5 absolute of the reference (floating point: the reference's two
sgemm calls have no defined summation order) -- and in fact bit-identical to the oracle's
sequential-FMA order, which is asserted too."""
import ctypes as C
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_hkm_quantize_matches_reference_golden(yn):
    g = np.load(os.path.join(GOLD, "hkm_gmm.npz"))
    levels = [g["hkm_level%d" % l] for l in range(3)]
    assert np.array_equal(yn.hkm_quantize(levels, 5, g["hkm_query"]), g["hkm_quantize_query"])
    assert np.array_equal(yn.hkm_quantize(levels, 5, g["hkm_points"]), g["hkm_quantize_points"])


@pytest.mark.parametrize("n,d,bf,nlevel", [(20000, 128, 10, 3), (3000, 7, 2, 6), (5000, 200, 40, 2), (257, 3, 33, 1)])
def test_hkm_quantize_matches_oracle(yn, ob, n, d, bf, nlevel):
    r = np.random.RandomState(n + d + bf)
    levels = [r.rand(bf ** (l + 1), d).astype(np.float32) for l in range(nlevel)]
    v = r.rand(n, d).astype(np.float32)
    # exact ties between children (duplicated rows): the lowest child wins, as in nn_single_full
    levels[0][bf - 1] = levels[0][0]
    got = yn.hkm_quantize(levels, bf, v)
    assert np.array_equal(got, ob.orc_hkm_quantize(levels, bf, v))
    assert got.min() >= 0 and got.max() < bf ** nlevel


def _tree_data(r, bf, nlevel, d, per_leaf):
    """points around bf^nlevel leaf centres that are nested: siblings are close, cousins far"""
    centres = np.zeros((1, d))
    scale = 1.0
    for l in range(nlevel):
        step = r.randn(centres.shape[0], bf, d) * scale
        centres = (centres[:, None, :] + step).reshape(-1, d)
        scale *= 0.15
    pts = centres[:, None, :] + r.randn(centres.shape[0], per_leaf, d) * scale * 0.2
    v = pts.reshape(-1, d).astype(np.float32)
    return v[r.permutation(len(v))]


def test_hkm_learn_tree_is_consistent(yn, ob):
    r = np.random.RandomState(8)
    bf, nlevel, d = 4, 2, 16
    v = _tree_data(r, bf, nlevel, d, 150)
    levels, assign = yn.hkm_learn(v, nlevel, bf, niter=12)
    assert [x.shape for x in levels] == [(bf, d), (bf * bf, d)]
    assert assign.min() >= 0 and assign.max() < bf ** nlevel
    # every table row is the mean of the points the learning assigned to that node
    # (kmeans.c:278-288: the centroids returned are the means under the returned assignment)
    for l in range(nlevel):
        node = assign // bf ** (nlevel - 1 - l)
        for c in range(bf ** (l + 1)):
            pts = v[node == c]
            assert len(pts) > 0
            np.testing.assert_allclose(levels[l][c], pts.astype(np.float64).mean(0), rtol=0, atol=2e-5)
    # the quantiser agrees with the oracle on the learned tree
    assert np.array_equal(yn.hkm_quantize(levels, bf, v), ob.orc_hkm_quantize(levels, bf, v))


def test_hkm_learn_matches_compiled_reference(yn, ob):
    # both implementations draw their k-means seeds from lrand48 (kmeans.c:379-380, seed 0 at
    # hkm.c:88): with the generator reset before each run the trees must coincide on contracting data
    if not ob.have_ref():
        pytest.skip("oracle/_ref not built")
    r = np.random.RandomState(21)
    bf, nlevel, d = 3, 3, 12
    v = _tree_data(r, bf, nlevel, d, 60)
    libc = C.CDLL(None)
    libc.srand48.argtypes = [C.c_long]
    libc.srand48(77)
    want_levels, want_assign = ob.ref_hkm_learn(v, nlevel, bf, niter=10)
    libc.srand48(77)
    levels, assign = yn.hkm_learn(v, nlevel, bf, niter=10)
    assert np.array_equal(assign, want_assign)
    for a, b in zip(levels, want_levels):
        np.testing.assert_allclose(a, b, rtol=0, atol=1e-5)


def test_gmm_posteriors_match_reference_golden(yn):
    g = np.load(os.path.join(GOLD, "hkm_gmm.npz"))
    mix = (g["gmm_w"], g["gmm_mu"], g["gmm_sigma"])
    for flags, name in ((1, "gmm_p_w"), (0, "gmm_p_now")):
        p = yn.gmm_compute_p(mix, g["gmm_v"], flags)
        np.testing.assert_allclose(p, g[name], rtol=0, atol=1e-5)   # the bar (floating point)
        assert np.array_equal(p, g[name])                           # what this build achieves


@pytest.mark.parametrize("n,k,d", [(20000, 256, 64), (1000, 1000, 128), (4097, 65, 33), (3, 2, 1)])
def test_gmm_posteriors_match_oracle(yn, ob, n, k, d):
    r = np.random.RandomState(n + k + d)
    mu = r.rand(k, d).astype(np.float32)
    sigma = (0.05 + 0.2 * r.rand(k, d)).astype(np.float32)
    w = r.rand(k).astype(np.float32)
    w /= w.sum()
    v = r.rand(n, d).astype(np.float32)
    for flags in (1, 0):
        p = yn.gmm_compute_p((w, mu, sigma), v, flags)
        want = ob.orc_gmm_compute_p(w, mu, sigma, v, flags)
        np.testing.assert_allclose(p, want, rtol=0, atol=1e-5)
        np.testing.assert_allclose(p.sum(1), 1.0, atol=1e-4)
        assert ((p == 0) == (want == 0)).all()    # the 2^-24 cut-off of softmax_ref (gmm.c:265,281)
        # exp() is the one step whose last bit may differ between libm and the device
        assert np.abs(p.view(np.int32).astype(np.int64) - want.view(np.int32).astype(np.int64)).max() <= 2


def test_gmm_posteriors_device_pointers(yn, ob):
    import yael_b200
    from devmem import DevArray
    r = np.random.RandomState(4)
    k, d, n = 40, 20, 700
    mu, sigma = r.rand(k, d).astype(np.float32), (0.1 + r.rand(k, d)).astype(np.float32)
    w = np.full(k, 1.0 / k, np.float32)
    v = r.rand(n, d).astype(np.float32)
    from yael_b200 import _lib
    L = yael_b200.lib()
    f = lambda a: a.ctypes.data_as(C.POINTER(C.c_float))
    g = _lib.GmmT(d, k, f(w), f(mu), f(sigma))
    dv, dp = DevArray(v), DevArray(shape=(n, k), dtype=np.float32)
    L.gmm_compute_p_thread(n, C.cast(dv.ptr, C.POINTER(C.c_float)), C.byref(g),
                           C.cast(dp.ptr, C.POINTER(C.c_float)), 1, 4)
    got = dp.get()
    dv.free()
    dp.free()
    assert np.array_equal(got, yn.gmm_compute_p((w, mu, sigma), v, 1))


def test_hkm_quantize_without_levels_is_one_leaf():
    # yael/hkm.c:144-162 with nlevel = 0: the loop body never runs, every point lands in leaf 0
    import yael_b200
    from yael_b200 import _lib
    L = yael_b200.lib()
    v = np.random.RandomState(1).rand(100, 8).astype(np.float32)
    h = _lib.HkmT(0, 4, 1, 8, None)
    idx = np.full(100, -7, np.int32)
    L.hkm_quantize(C.byref(h), 100, v.ctypes.data_as(C.POINTER(C.c_float)), idx.ctypes.data_as(C.POINTER(C.c_int)))
    assert (idx == 0).all()

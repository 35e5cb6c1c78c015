"""GPU parity of the exact FP32 path (engine 0) against the oracle, through the C ABI.

Bar: bit-exact for everything this path computes -- the distance kernels reproduce the
reference's rounding sequence (sequential FP32 FMA dot product, float / double norms), and
selection is defined as (value, id) order, which the oracle's canonical variants restate.
"""
import numpy as np
import pytest

import yael_b200
from devmem import DevArray

pytestmark = pytest.mark.gpu


def rs(seed):
    return np.random.RandomState(seed)


@pytest.fixture(autouse=True)
def _exact_engine():
    import yael_b200
    yael_b200.lib().yb_set_knn_engine(0)
    yield
    yael_b200.lib().yb_set_knn_engine(-1)


@pytest.mark.parametrize("na,nb,d", [(300, 200, 64), (65, 129, 128), (1, 1, 1), (130, 7, 3),
                                     (257, 255, 96), (64, 64, 37)])
def test_cross_distances_bit_exact(yn, ob, na, nb, d):
    r = rs(na * 1000 + nb)
    a = r.rand(na, d).astype(np.float32)
    b = r.rand(nb, d).astype(np.float32)
    got = yn.cross_distances(a, b)
    want = ob.orc_cross(a, b, ob.DOT_F32_SEQ)
    assert got.shape == want.shape
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


@pytest.mark.parametrize("dtype", [1, 2, 3, 4, 5, 6, 16])
@pytest.mark.parametrize("d", [8, 13])
def test_cross_distances_alt(yn, ob, dtype, d):
    r = rs(dtype * 10 + d)
    a = r.rand(40, d).astype(np.float32)
    b = r.rand(33, d).astype(np.float32)
    got = yn.cross_distances(a, b, dtype)
    if ob.have_ref():
        L = ob.ref()
        import ctypes as C
        want = np.empty((33, 40), np.float32)
        L.compute_cross_distances_alt_nonpacked = L.compute_cross_distances_alt_nonpacked
        L.compute_cross_distances_alt_nonpacked.argtypes = [C.c_int] * 4 + [
            C.POINTER(C.c_float), C.c_int, C.POINTER(C.c_float), C.c_int, C.POINTER(C.c_float), C.c_int]
        L.compute_cross_distances_alt_nonpacked(dtype, d, 40, 33, ob.fp(a), d, ob.fp(b), d, ob.fp(want), 40)
        if dtype == 16:  # sgemm: accumulation order of the BLAS edge kernels is not pinned
            np.testing.assert_allclose(got, want, rtol=1e-6)
        else:
            assert np.array_equal(got, want)


@pytest.mark.parametrize("nq,nb,d,k", [(100, 5000, 128, 1), (100, 5000, 128, 10), (33, 777, 96, 100),
                                        (5, 100, 16, 100), (257, 3000, 20, 7), (10, 70000, 32, 50)])
def test_knn_exact_engine(yn, ob, nq, nb, d, k):
    r = rs(nq + nb + k)
    b = r.rand(nb, d).astype(np.float32)
    q = r.rand(nq, d).astype(np.float32)
    idx, dis = yn.knn(q, b, k)
    widx, wdis = ob.orc_knn(b, q, k, ob.DOT_F32_SEQ, canonical=True)
    assert np.array_equal(dis.view(np.uint32), wdis.view(np.uint32))
    assert np.array_equal(idx, widx)


def test_knn_integer_ties_and_padding(yn, ob):
    r = rs(5)
    b = r.randint(0, 3, (500, 8)).astype(np.float32)
    q = r.randint(0, 3, (40, 8)).astype(np.float32)
    for k in (1, 3, 60):
        idx, dis = yn.knn(q, b, k)
        widx, wdis = ob.orc_knn(b, q, k, canonical=True)
        assert np.array_equal(idx, widx) and np.array_equal(dis, wdis)
        # the reference (heap tie order) returns the same distance multiset
        hidx, hdis = ob.orc_knn(b, q, k)
        assert np.array_equal(dis, hdis)
    # NaN rows are never selected, short lists padded with -1 / all-ones bits (nn.c:515-518)
    b2 = b.copy()
    b2[1::2] = np.nan
    idx, dis = yn.knn(q, b2, 300)
    widx, wdis = ob.orc_knn(b2, q, 300)
    assert (idx[:, 250:] == -1).all() and (widx[:, 250:] == -1).all()
    assert (dis[:, 250:].view(np.uint32) == 0xFFFFFFFF).all()
    assert np.array_equal(dis[:, :250], wdis[:, :250])


def test_knn_weights(yn, ob):
    r = rs(6)
    b = r.rand(2000, 24).astype(np.float32)
    q = r.rand(50, 24).astype(np.float32)
    w = (0.5 + r.rand(2000)).astype(np.float32)
    idx, dis = yn.knn_weighted(q, b, w, 5)
    widx, wdis = ob.orc_knn(b, q, 5, weights=w)
    assert np.array_equal(idx, widx) and np.array_equal(dis, wdis)


@pytest.mark.parametrize("n,k", [(100000, 1), (100000, 7), (100000, 100), (1000, 100), (50, 50),
                                 (30000, 5000)])
def test_k_min(yn, ob, n, k):
    r = rs(n + k)
    v = r.rand(3, n).astype(np.float32)
    v[1] = r.randint(0, 50, n)  # heavy ties
    got = yn.kmin(v, k)
    for i in range(3):
        want = ob.orc_k_min(v[i], k, canonical=True)
        assert np.array_equal(got[i], want)
    gotmax = yn.kmax(v, k)
    for i in range(3):
        want = ob.orc_k_min(-v[i], k, canonical=True)
        assert np.array_equal(gotmax[i], want)


def test_k_min_is_prefix_of_full_sort(yn):
    # the reference's own check, test/matlab/test_kmin.m:20
    r = rs(11)
    v = r.rand(4, 20000).astype(np.float32)
    got = yn.kmin(v, 100)
    for i in range(4):
        assert np.array_equal(got[i], np.argsort(v[i], kind="stable")[:100])


@pytest.mark.parametrize("nc", [4, 8, 16, 24, 5, 64])
def test_compute_hamming(yn, ob, nc):
    r = rs(nc)
    a = r.randint(0, 256, (70, nc)).astype(np.uint8)
    b = r.randint(0, 256, (50, nc)).astype(np.uint8)
    assert np.array_equal(yn.hamming_distances(a, b), ob.orc_compute_hamming(a, b))


@pytest.mark.parametrize("nq,nb,nc,k", [(300, 20000, 8, 100), (10, 5000, 16, 7), (700, 3000, 4, 1),
                                         (5, 100, 8, 100), (64, 40000, 32, 33), (3, 1000, 5, 10)])
def test_nn_hamming_bit_exact(yn, ob, nq, nb, nc, k):
    r = rs(nq + nb)
    b = r.randint(0, 256, (nb, nc)).astype(np.uint8)
    q = r.randint(0, 256, (nq, nc)).astype(np.uint8)
    b[::17] = q[0]  # planted duplicates: distance-0 ties resolved by id
    idx, dis = yn.knn_hamming(q, b, k)
    widx, wdis = ob.orc_nn_hamming(b, q, k)
    assert np.array_equal(dis, wdis)
    assert np.array_equal(idx, widx)


def test_match_hamming_self_consistency(yn, ob):
    # test/matlab/test_hamming.m:5-22: thresholded output == filtered full matrix
    r = rs(3)
    a = r.randint(0, 256, (80, 8)).astype(np.uint8)
    b = r.randint(0, 256, (1000, 8)).astype(np.uint8)
    full = yn.hamming_distances(a, b)  # [nb][na]
    for ht in (0, 20, 26, 64):
        pairs, scores = yn.match_hamming(a, b, ht)
        qi, bj = np.nonzero(full.T <= ht)  # query-major, base ascending
        assert np.array_equal(pairs[:, 0], qi) and np.array_equal(pairs[:, 1], bj)
        assert np.array_equal(scores, full.T[qi, bj])


def test_crossmatch_hamming_matches_oracle(yn, ob):
    # crossmatch_hamming* (yael/hamming.c:310-395, 751-829): pairs i < j of one set, (i, j) order;
    # sizes straddle the 256-row scan blocks, thresholds from "nothing" to "everything"
    import ctypes as C
    r = rs(5)
    for n, nc in ((1, 8), (2, 8), (257, 8), (1500, 16), (700, 4), (513, 24)):
        db = r.randint(0, 256, (n, nc)).astype(np.uint8)
        if n > 10:
            db[::7] = db[1]
        for ht in (-1, 0, nc * 3, nc * 8):
            pairs, scores = yn.crossmatch_hamming(db, ht)
            m = C.c_size_t(0)
            ob.oracle().orc_crossmatch_hamming_count(ob.u8p(db), n, ht, nc, C.byref(m))
            assert len(scores) == m.value
            if ht == nc * 8:
                assert m.value == n * (n - 1) // 2
            widx = np.empty((m.value, 2), np.int32)
            wham = np.empty(m.value, np.uint16)
            if m.value:
                ob.oracle().orc_crossmatch_hamming_prealloc(ob.u8p(db), n, ht, nc, ob.ip(widx), ob.u16p(wham))
            assert np.array_equal(pairs, widx) and np.array_equal(scores, wham)


def test_kmeans_step_teacher_forced(yn, ob):
    r = rs(1234)
    v = r.rand(20000, 32).astype(np.float32)
    c0 = v[r.permutation(20000)[:64]].copy()
    cent, qerr, dis, assign, nassign = yn.kmeans(v, 64, niter=1, verbose=False, init=c0, output="all")
    q, wc, wa, wd, wn = ob.orc_kmeans_step(v, c0)
    assert np.array_equal(assign, wa) and np.array_equal(nassign, wn)
    assert np.array_equal(dis.view(np.uint32), wd.view(np.uint32))
    np.testing.assert_allclose(cent, wc, rtol=0, atol=1e-6)
    assert abs(qerr - q / 20000) < 1e-5 * qerr


def test_kmeans_full_matches_oracle_small(yn, ob):
    r = rs(99)
    v = r.rand(5000, 16).astype(np.float32)
    import os
    os.environ["YAEL_B200_EXACT_UPDATE"] = "1"
    try:
        got = yn.kmeans(v, 32, niter=12, seed=777, verbose=False, output="all", nt=4)
    finally:
        del os.environ["YAEL_B200_EXACT_UPDATE"]
    want = ob.orc_kmeans(v, 32, 12, ob.KMEANS_QUIET | 4, 777)
    # exact-order update + bit-exact assignment => the whole trajectory is identical
    assert np.array_equal(got[0], want[1])
    assert np.array_equal(got[3], want[3]) and np.array_equal(got[4], want[4])
    assert got[1] == pytest.approx(want[0], rel=1e-6)


def test_kmeans_empty_cluster_split(yn, ob):
    r = rs(5)
    v = np.repeat(r.rand(30, 8).astype(np.float32), 50, axis=0)
    import os
    os.environ["YAEL_B200_EXACT_UPDATE"] = "1"
    try:
        got = yn.kmeans(v, 40, niter=10, seed=5, verbose=False, output="all")
    finally:
        del os.environ["YAEL_B200_EXACT_UPDATE"]
    want = ob.orc_kmeans(v, 40, 10, ob.KMEANS_QUIET | 1, 5)
    assert np.array_equal(got[4], want[4])
    assert np.array_equal(got[0], want[1])


@pytest.mark.parametrize("n,k,d,skew", [(300000, 2048, 128, 0), (200000, 2048, 96, 5000), (150000, 1024, 200, 3000),
                                         (120000, 1024, 30, 0), (400000, 4096, 64, 70000), (5000, 1024, 16, 0), (600000, 8192, 128, 0)])
def test_kmeans_accumulate_short_segments_bit_exact(n, k, d, skew):
    """Centroid update with many centroids (yb_kmeans_accumulate's short-segment path: one-pass
    scatter + per-segment id sort): sums must equal the reference's strict point-order FP32 sums
    (yael/kmeans.c:278-283) bit for bit -- also when one cluster is much longer than the in-warp
    sort (id-range passes) or longer than the path's limit (device-side fall-back to the general
    radix-sort path)."""
    L = yael_b200.lib()
    r = rs(n + k + d)
    v = r.rand(n, d).astype(np.float32)
    assign = r.randint(0, k, n).astype(np.int32)
    if skew:
        assign[r.permutation(n)[:skew]] = 7
    assign[assign == 11] = 12                       # an empty cluster
    dis = r.rand(n).astype(np.float32)
    dv, da, dd = DevArray(v), DevArray(assign), DevArray(dis)
    ds = DevArray(shape=(k, d), dtype=np.float32)
    dn = DevArray(shape=(k,), dtype=np.int32)
    dq = DevArray(shape=(1,), dtype=np.float64)
    rc = L.yb_kmeans_accumulate(d, n, k, dv.ptr, da.ptr, dd.ptr, ds.ptr, dn.ptr, dq.ptr, 1, None)  # exact_order
    assert rc == 0, L.yb_last_error()
    L.yb_sync(None)
    want = np.zeros((k, d), np.float32)
    np.add.at(want, assign, v)                      # unbuffered, in index order, float32 adds
    assert np.array_equal(dn.get(), np.bincount(assign, minlength=k))
    assert np.array_equal(ds.get(), want)
    assert dq.get()[0] == pytest.approx(float(dis.astype(np.float64).sum()), rel=1e-12)
    for a in (dv, da, dd, ds, dn, dq):
        a.free()


@pytest.mark.parametrize("n,k,d,skew", [(600000, 8192, 128, 40000), (600000, 8192, 64, 0)])
def test_kmeans_accumulate_default_mode_splits_only_long_clusters(n, k, d, skew):
    """Default (non-exact) mode with many centroids: clusters of at most 2048 points are summed in the
    reference's strict point order (bit-identical); a longer cluster -- the first iteration after a
    random-point initialisation has one of 12 860 points at BASELINE configs[3] -- is walked in
    2048-row pieces by the general path, so its sums carry the rounding of that grouping (checked
    against float64 at the scale of FP32 accumulation error)."""
    L = yael_b200.lib()
    r = rs(n + k + d + skew)
    v = r.rand(n, d).astype(np.float32)
    assign = r.randint(0, k, n).astype(np.int32)
    if skew:
        assign[r.permutation(n)[:skew]] = 7
    dis = r.rand(n).astype(np.float32)
    dv, da, dd = DevArray(v), DevArray(assign), DevArray(dis)
    ds = DevArray(shape=(k, d), dtype=np.float32)
    dn = DevArray(shape=(k,), dtype=np.int32)
    dq = DevArray(shape=(1,), dtype=np.float64)
    rc = L.yb_kmeans_accumulate(d, n, k, dv.ptr, da.ptr, dd.ptr, ds.ptr, dn.ptr, dq.ptr, 0, None)
    assert rc == 0, L.yb_last_error()
    L.yb_sync(None)
    got, cnt = ds.get(), dn.get()
    want = np.zeros((k, d), np.float32)
    np.add.at(want, assign, v)
    assert np.array_equal(cnt, np.bincount(assign, minlength=k))
    small = cnt <= 2048
    assert np.array_equal(got[small], want[small])
    if skew:
        assert (~small).sum() == 1 and cnt[7] > 2048
        exact = v[assign == 7].astype(np.float64).sum(0)
        # strict-order FP32 and pieced FP32 both sit within the accumulation error of the exact sums
        tol = cnt[7] * 2.0 ** -24 * np.abs(exact) * 4
        assert (np.abs(got[7] - exact) <= tol).all()
        assert (np.abs(want[7] - exact) <= tol).all()
    for a in (dv, da, dd, ds, dn, dq):
        a.free()

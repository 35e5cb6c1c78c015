"""GPU: the product against the committed golden vectors of the unmodified reference, the
shard-merge kernels (the exchange step of the multi-GPU paths, emulated on one device), the
drop-in wrappers, and size-independent properties at larger shapes."""
import ctypes as C
import os

import numpy as np
import pytest

import yael_b200
from devmem import DevArray

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def gold(name):
    return np.load(os.path.join(GOLD, name + ".npz"))


def test_reference_own_test_vector(yn):
    g = gold("ynumpy_knn")
    idx, dis = yn.knn(g["query"], g["base"], 2)
    assert np.array_equal(idx, g["idx"]) and np.array_equal(dis, g["dis"])


@pytest.mark.parametrize("engine", [0, 1])
def test_knn_uniform_golden(yn, engine):
    g = gold("knn_uniform_seed1234")
    r = np.random.RandomState(1234)
    b = r.random_sample((4000, 128)).astype(np.float32)
    q = r.random_sample((64, 128)).astype(np.float32)
    L = yael_b200.lib()
    L.yb_set_knn_engine(engine)
    try:
        for k in (1, 10, 100):
            idx, dis = yn.knn(q, b, k)
            # north_star: distances within 1e-5 relative, ids identical (no ties in this data)
            np.testing.assert_allclose(dis, g["dis%d" % k], rtol=1e-5)
            assert np.array_equal(idx, g["idx%d" % k])
    finally:
        L.yb_set_knn_engine(-1)


def test_knn_semantics_golden(yn):
    g = gold("knn_semantics")
    bt = np.zeros((6, 1), np.float32)
    bt[:, 0] = [1, 1, 1, 1, .5, 1]
    qt = np.zeros((1, 1), np.float32)
    idx, dis = yn.knn(qt, bt, 3)
    assert np.array_equal(dis, g["tie_dis"])          # same distances as the reference
    assert np.array_equal(idx, [[4, 0, 1]])           # ties by id (the reference: heap slot)
    bn = np.array([[0.], [np.nan], [2.]], np.float32)
    idx, dis = yn.knn(qt, bn, 3)
    assert np.array_equal(idx, g["nan_idx"])
    assert np.array_equal(dis.view(np.uint32), g["nan_dis"].view(np.uint32))
    idx, dis = yn.knn(np.zeros((2, 2), np.float32), np.ones((5, 2), np.float32), 1)
    assert np.array_equal(idx, g["k1_tie_idx"]) and np.array_equal(dis, g["k1_tie_dis"])


def test_cross_distances_golden(yn):
    g = gold("cross_distances")
    np.testing.assert_allclose(yn.cross_distances(g["a"], g["b"]), g["dist"], rtol=1e-5, atol=1e-6)


def test_kmeans_golden_exact_update(yn, monkeypatch):
    g = gold("kmeans_seed777")
    monkeypatch.setenv("YAEL_B200_EXACT_UPDATE", "1")
    L = yael_b200.lib()
    L.yb_set_knn_engine(0)
    try:
        v = np.random.RandomState(1234).random_sample((20000, 32)).astype(np.float32)
        for name, init in (("random", "random"), ("pp", "kmeans++")):
            cent, qerr, dis, assign, nassign = yn.kmeans(v, 64, niter=15, seed=777, redo=2, nt=4,
                                                         verbose=False, init=init, output="all")
            # identical trajectory: same assignments and sizes; centroids to the last ulps
            assert np.array_equal(nassign, g[name + "_nassign"])
            assert np.array_equal(assign, g[name + "_assign"])
            np.testing.assert_allclose(cent, g[name + "_cent"], rtol=0, atol=1e-6)
            assert qerr == pytest.approx(float(g[name + "_qerr"]), rel=1e-6)
    finally:
        L.yb_set_knn_engine(-1)


def test_kmeans_golden_default_engines(yn):
    # default configuration (tensor-core assignment, parallel update): north_star tolerance
    g = gold("kmeans_seed777")
    v = np.random.RandomState(1234).random_sample((20000, 32)).astype(np.float32)
    init = v[np.random.RandomState(3).permutation(20000)[:64]].copy()
    from oracle import bindings as ob
    cent, qerr, dis, assign, nassign = yn.kmeans(v, 64, niter=1, verbose=False, init=init, output="all")
    q, wc, wa, wd, wn = ob.orc_kmeans_step(v, init)
    mism = assign != wa
    assert np.all(np.abs(dis[mism] - wd[mism]) <= 1e-5 * wd[mism])
    np.testing.assert_allclose(cent, wc, atol=1e-4)


def test_kmeans_empty_split_golden(yn, monkeypatch):
    g = gold("kmeans_empty_split")
    monkeypatch.setenv("YAEL_B200_EXACT_UPDATE", "1")
    cent, qerr, dis, assign, nassign = yn.kmeans(g["v"], 40, niter=10, seed=5, verbose=False, output="all")
    assert np.array_equal(nassign, g["nassign"])
    np.testing.assert_allclose(cent, g["cent"], rtol=0, atol=1e-6)


def test_kmin_golden_values(yn):
    g = gold("kmin")
    val = g["val"]
    for k, key in ((7, "k7"), (100, "k100"), (1, "k1")):
        got = yn.kmin(val[None, :], k)[0]
        assert np.array_equal(val[got], val[g[key]])  # same values (tie order is ours: by index)
    got = yn.kmin(val[None, :1000].copy(), 100)[0]
    assert np.array_equal(got, g["small100"])


def test_hamming_golden(yn):
    g = gold("hamming")
    for nc in (4, 8, 16, 24, 5):
        assert np.array_equal(yn.hamming_distances(g["a%d" % nc], g["b%d" % nc]), g["dis%d" % nc])
    pairs, scores = yn.match_hamming(g["a8"], g["b8"], 28)
    assert np.array_equal(pairs, g["match_ht28_idx"]) and np.array_equal(scores, g["match_ht28_ham"])


def test_crossmatch_hamming_golden(yn):
    g = gold("hamming_crossmatch")
    for nc in (4, 8, 16, 5):
        pairs, scores = yn.crossmatch_hamming(g["db%d" % nc], int(g["ht%d" % nc]))
        assert np.array_equal(pairs, g["idx%d" % nc]) and np.array_equal(scores, g["ham%d" % nc])


@pytest.mark.parametrize("t", [1, 2, 3, 4, 5, 6, 16])
def test_knn_alt_types_and_weights_golden(yn, t):
    """knn_full with every distance type of compute_cross_distances_alt (yael/nn.c:280-350), with
    and without the per-base weights (yael/nn.c:497-500), k = 1 (nn_single_full) and k = 5, against
    the compiled reference (scripts/make_golden.py section 10)."""
    g = gold("knn_alt_weighted")
    b, q, w = g["base"], g["query"], g["weights"]
    for k in (1, 5):
        for wname, ww in (("", None), ("w", w)):
            if ww is None:
                idx, dis = yn.knn(q, b, k, distance_type=t)
            else:
                idx, dis = yn.knn_weighted(q, b, ww, k, distance_type=t)
            widx, wdis = g["idx_t%d_k%d%s" % (t, k, wname)], g["dis_t%d_k%d%s" % (t, k, wname)]
            assert np.array_equal(idx, widx), (t, k, wname)
            if t == 16:   # sgemm in the reference: summation order inside BLAS edge tiles
                np.testing.assert_allclose(dis, wdis, rtol=2e-6)
            else:
                assert np.array_equal(dis, wdis), (t, k, wname)


def test_kmeans_l1_is_refused_loudly(yn):
    v = np.random.RandomState(0).random_sample((100, 4)).astype(np.float32)
    with pytest.raises(NotImplementedError):
        yn.kmeans(v, 4, distance_type=1, verbose=False)


def test_knn_merge_equals_unsharded(yn):
    # the multi-GPU exchange step on one device: G shard results -> merged == single search
    L = yael_b200.lib()
    r = np.random.RandomState(3)
    b = r.randint(0, 6, (4000, 16)).astype(np.float32)  # ties across shards
    q = r.randint(0, 6, (50, 16)).astype(np.float32)
    k, G = 20, 4
    L.yb_set_knn_engine(0)
    try:
        widx, wdis = yn.knn(q, b, k)
        parts_i, parts_d = [], []
        for gI in range(G):
            lo, hi = 4000 * gI // G, 4000 * (gI + 1) // G
            i_, d_ = yn.knn(q, np.ascontiguousarray(b[lo:hi]), k)
            parts_i.append(i_ + lo)
            parts_d.append(d_)
    finally:
        L.yb_set_knn_engine(-1)
    gi, gd = DevArray(np.stack(parts_i)), DevArray(np.stack(parts_d))
    oi, od = DevArray(shape=(50, k), dtype=np.int32), DevArray(shape=(50, k), dtype=np.float32)
    assert L.yb_knn_merge(50, k, G, gi.ptr, gd.ptr, oi.ptr, od.ptr, None) == 0
    L.yb_sync(None)
    assert np.array_equal(oi.get(), widx) and np.array_equal(od.get(), wdis)


def test_knn_merge_strided_padding_and_unsorted_input():
    # the merge ranks sorted lists (binary searches) and falls back to a sort for unsorted ones;
    # both must give the k best by (distance, id), padding (-1, NaN bits) last; the strided entry
    # reads ids and distances of a shard from one [2][nq][k] block (one collective)
    L = yael_b200.lib()
    r = np.random.RandomState(12)
    nq, k = 37, 25
    for G in (2, 5, 8):
        dis = np.sort(r.randint(0, 60, (G, nq, k)).astype(np.float32), axis=2)   # heavy ties
        ids = np.empty((G, nq, k), np.int32)
        for g in range(G):
            for q in range(nq):
                ids[g, q] = g * 1000 + np.arange(k)      # ascending ids inside equal distances
        # short lists: the tail of some lists is padding
        dis[1, ::3, k - 7:] = np.float32(np.nan)
        ids[1, ::3, k - 7:] = -1
        keys = (dis.astype(np.float64) * 1e6 + ids).transpose(1, 0, 2).reshape(nq, G * k)
        keys[np.isnan(keys)] = np.inf
        order = np.argsort(keys, axis=1, kind="stable")[:, :k]
        want_i = ids.transpose(1, 0, 2).reshape(nq, G * k)[np.arange(nq)[:, None], order]
        want_d = dis.transpose(1, 0, 2).reshape(nq, G * k)[np.arange(nq)[:, None], order]
        for variant in ("plain", "strided", "unsorted"):
            di, dd = ids.copy(), dis.copy()
            if variant == "unsorted":   # reverse one list: still the same multiset
                di[0] = di[0, :, ::-1]
                dd[0] = dd[0, :, ::-1]
            oi, od = DevArray(shape=(nq, k), dtype=np.int32), DevArray(shape=(nq, k), dtype=np.float32)
            if variant == "strided":
                blk = np.stack([di, dd.view(np.int32)], axis=1)   # [G][2][nq][k]
                gb = DevArray(blk)
                rc = L.yb_knn_merge_strided(nq, k, G, gb.ptr, gb.ptr + 4 * nq * k, 2 * nq * k,
                                            oi.ptr, od.ptr, None)
            else:
                gi, gd = DevArray(di), DevArray(dd)
                rc = L.yb_knn_merge(nq, k, G, gi.ptr, gd.ptr, oi.ptr, od.ptr, None)
            assert rc == 0, L.yb_last_error()
            L.yb_sync(None)
            got_i, got_d = oi.get(), od.get()
            assert np.array_equal(got_i, want_i), (G, variant)
            assert np.array_equal(got_d[want_i >= 0], want_d[want_i >= 0])
            assert (got_i[want_i < 0] == -1).all()


def test_hamming_merge_bit_identical_for_any_shard_count(yn):
    L = yael_b200.lib()
    r = np.random.RandomState(4)
    b = r.randint(0, 256, (6000, 8)).astype(np.uint8)
    q = r.randint(0, 256, (40, 8)).astype(np.uint8)
    k = 30
    widx, wdis = yn.knn_hamming(q, b, k)
    for G in (2, 3, 8):
        parts_i, parts_d = [], []
        for gI in range(G):
            lo, hi = 6000 * gI // G, 6000 * (gI + 1) // G
            i_, d_ = yn.knn_hamming(q, np.ascontiguousarray(b[lo:hi]), k)
            parts_i.append(i_ + lo)
            parts_d.append(d_)
        gi, gd = DevArray(np.stack(parts_i)), DevArray(np.stack(parts_d))
        oi, od = DevArray(shape=(40, k), dtype=np.int32), DevArray(shape=(40, k), dtype=np.uint16)
        assert L.yb_nn_hamming_merge(40, k, G, gi.ptr, gd.ptr, oi.ptr, od.ptr, None) == 0
        L.yb_sync(None)
        assert np.array_equal(oi.get(), widx) and np.array_equal(od.get(), wdis)


def test_drop_in_wrappers(yn, ob):
    L = yael_b200.lib()
    f, i = C.POINTER(C.c_float), C.POINTER(C.c_int)
    r = np.random.RandomState(5)
    b = r.random_sample((500, 16)).astype(np.float32)
    q = r.random_sample((40, 16)).astype(np.float32)
    # nn(): returns the sum of distances as double (yael/nn.c:608-621)
    a = np.empty(40, np.int32)
    tot = L.nn(40, 500, 16, b.ctypes.data_as(f), q.ctypes.data_as(f), a.ctypes.data_as(i))
    widx, wdis = ob.orc_knn(b, q, 1)
    assert np.array_equal(a, widx[:, 0]) and tot == pytest.approx(float(wdis.astype(np.float64).sum()), rel=1e-6)
    # knn_reorder_shortlist (yael/nn.c:528-580)
    idx = np.ascontiguousarray(np.tile(np.arange(12, dtype=np.int32)[::-1], (40, 1)))
    idx[3, 5:] = -1
    widx = idx.copy()
    wd = np.zeros((40, 12), np.float32)
    ob.oracle().orc_knn_reorder_shortlist(40, 500, 16, 12, ob.fp(b), ob.fp(q), ob.ip(widx), ob.fp(wd), 0)
    dis = yn.knn_reorder_shortlist(q, b, idx)
    assert np.array_equal(idx, widx)
    np.testing.assert_allclose(dis[3, :5], wd[3, :5], rtol=1e-6)
    np.testing.assert_allclose(dis[0], wd[0], rtol=1e-6)
    # compute_distances_1 (yael/nn.c:132-162)
    out = np.empty(500, np.float32)
    L.compute_distances_1(16, 500, q.ctypes.data_as(f), b.ctypes.data_as(f), out.ctypes.data_as(f))
    w = np.empty(500, np.float32)
    ob.oracle().orc_distances_1(16, 500, ob.fp(q), ob.fp(b), 16, ob.fp(w), 0)
    assert np.array_equal(out, w)
    # non-packed cross distances: gaps in the output belong to the caller
    outp = np.full((40, 600), -7.0, np.float32)
    L.compute_cross_distances_nonpacked(16, 500, 40, b.ctypes.data_as(f), 16, q.ctypes.data_as(f), 16,
                                        outp.ctypes.data_as(f), 600)
    assert np.array_equal(outp[:, :500], ob.orc_cross(b, q)) and (outp[:, 500:] == -7.0).all()


def test_drop_in_thin_wrappers(ob):
    """knn / knn_thread / nn_thread / compute_cross_distances_thread (yael/nn.c:624-632, 704-726,
    777-792): thin calls into knn_full / compute_cross_distances; knn and knn_thread return a
    malloc'd block the caller frees."""
    L = yael_b200.lib()
    libc = C.CDLL(None)
    libc.free.argtypes = [C.c_void_p]
    f, i = C.POINTER(C.c_float), C.POINTER(C.c_int)
    r = np.random.RandomState(15)
    b = r.random_sample((700, 24)).astype(np.float32)
    q = r.random_sample((33, 24)).astype(np.float32)
    widx, wdis = ob.orc_knn(b, q, 6, canonical=True)
    for name, extra in (("knn", ()), ("knn_thread", (3,))):
        vw = np.empty((33, 6), np.int32)
        p = getattr(L, name)(33, 700, 24, 6, b.ctypes.data_as(f), q.ctypes.data_as(f), vw.ctypes.data_as(i), *extra)
        got = np.ctypeslib.as_array(p, shape=(33, 6)).copy()
        libc.free(C.cast(p, C.c_void_p))
        assert np.array_equal(vw, widx) and np.array_equal(got, wdis), name
    w1i, w1d = ob.orc_knn(b, q, 1)
    vw = np.empty(33, np.int32)
    tot = L.nn_thread(33, 700, 24, b.ctypes.data_as(f), q.ctypes.data_as(f), vw.ctypes.data_as(i), 5)
    assert np.array_equal(vw, w1i[:, 0])
    assert tot == pytest.approx(float(w1d.astype(np.float64).sum()), rel=1e-6)
    out = np.empty((33, 700), np.float32)
    L.compute_cross_distances_thread(24, 700, 33, b.ctypes.data_as(f), q.ctypes.data_as(f),
                                     out.ctypes.data_as(f), 4)
    assert np.array_equal(out, ob.orc_cross(b, q))


def test_device_pointers_accepted_by_drop_in_api(ob):
    # the drop-in functions take CUDA device pointers as well as host pointers
    L = yael_b200.lib()
    r = np.random.RandomState(6)
    b = r.random_sample((300, 8)).astype(np.float32)
    q = r.random_sample((20, 8)).astype(np.float32)
    db, dq = DevArray(b), DevArray(q)
    di, dd = DevArray(shape=(20, 3), dtype=np.int32), DevArray(shape=(20, 3), dtype=np.float32)
    fn = L.knn_full
    fn.argtypes = [C.c_int] * 5 + [C.c_void_p] * 5
    fn(2, 20, 300, 8, 3, db.ptr, dq.ptr, None, di.ptr, dd.ptr)
    widx, wdis = ob.orc_knn(b, q, 3, canonical=True)
    assert np.array_equal(di.get(), widx) and np.array_equal(dd.get(), wdis)
    L.knn_full.argtypes = [C.c_int] * 5 + [C.POINTER(C.c_float)] * 3 + [C.POINTER(C.c_int), C.POINTER(C.c_float)]


def test_large_knn_properties(yn):
    # BASELINE-shaped slice: results independent of the engine, sorted, and consistent with a
    # brute-force check on a few queries
    r = np.random.RandomState(11)
    b = r.random_sample((200000, 128)).astype(np.float32)
    q = r.random_sample((2000, 128)).astype(np.float32)
    idx, dis = yn.knn(q, b, 100)
    assert yael_b200.lib().yb_last_knn_engine() == 1
    assert (np.diff(dis, axis=1) >= 0).all()
    assert all(len(set(row)) == 100 for row in idx[:50])
    for j in (0, 7, 1999):
        d64 = ((b.astype(np.float64) - q[j].astype(np.float64)) ** 2).sum(1)
        want = np.argsort(d64, kind="stable")[:100]
        assert np.array_equal(idx[j], want)
        np.testing.assert_allclose(dis[j], d64[want], rtol=1e-5)

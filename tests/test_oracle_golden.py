"""CPU: the oracle (our C restatement) against the committed golden vectors that were produced
by the unmodified reference (scripts/make_golden.py), and -- when oracle/_ref is present --
against the compiled reference itself on fresh seeded inputs.  This is what pins the oracle."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def gold(name):
    return np.load(os.path.join(GOLD, name + ".npz"))


def test_reference_own_test_vector(ob):
    # test/py/test_ynumpy.py:13-26 data; closed form 4*(j + 1/4 - i)^2
    g = gold("ynumpy_knn")
    idx, dis = ob.orc_knn(g["base"], g["query"], 2, nt=1)
    assert np.array_equal(idx, g["idx"]) and np.array_equal(dis, g["dis"])
    assert np.array_equal(g["idx"], [[0, 1], [1, 2], [2, 3]])
    np.testing.assert_allclose(g["dis"], [[.25, 2.25]] * 3)


def test_knn_uniform_golden(ob):
    g = gold("knn_uniform_seed1234")
    r = np.random.RandomState(1234)
    b = r.random_sample((4000, 128)).astype(np.float32)
    q = r.random_sample((64, 128)).astype(np.float32)
    for k in (1, 10, 100):
        idx, dis = ob.orc_knn(b, q, k, ob.DOT_F32_SEQ, nt=4)
        assert np.array_equal(idx, g["idx%d" % k])
        # the sequential-FMA dot reproduces the sgemm build bit for bit except in BLAS edge tiles
        np.testing.assert_allclose(dis, g["dis%d" % k], rtol=2e-7)
        idx64, dis64 = ob.orc_knn(b, q, k, ob.DOT_F64, nt=4)
        np.testing.assert_allclose(dis64, g["dis%d" % k], rtol=1e-5)


def test_knn_semantics_golden(ob):
    g = gold("knn_semantics")
    bt = np.zeros((6, 1), np.float32)
    bt[:, 0] = [1, 1, 1, 1, .5, 1]
    qt = np.zeros((1, 1), np.float32)
    idx, dis = ob.orc_knn(bt, qt, 3, nt=1)
    assert np.array_equal(idx, g["tie_idx"]) and np.array_equal(idx, [[4, 2, 1]])  # heap-slot ties
    bn = np.array([[0.], [np.nan], [2.]], np.float32)
    idx, dis = ob.orc_knn(bn, qt, 3, nt=1)
    assert np.array_equal(idx, g["nan_idx"]) and np.array_equal(idx, [[0, 2, -1]])
    assert np.array_equal(dis.view(np.uint32), g["nan_dis"].view(np.uint32))
    idx, dis = ob.orc_knn(np.ones((5, 2), np.float32), np.zeros((2, 2), np.float32), 1, nt=1)
    assert np.array_equal(idx, g["k1_tie_idx"]) and (idx == 0).all()  # k=1: lowest id on ties
    # canonical order differs from the heap order only inside ties
    cidx, cdis = ob.orc_knn(bt, qt, 3, canonical=True)
    assert np.array_equal(cidx, [[4, 0, 1]]) and np.array_equal(cdis, [[.25, 1., 1.]])


def test_cross_distances_golden(ob):
    g = gold("cross_distances")
    got = ob.orc_cross(g["a"], g["b"], ob.DOT_F32_SEQ)
    np.testing.assert_allclose(got, g["dist"], rtol=3e-7, atol=1e-6)
    assert got.shape == (29, 37)


def test_kmeans_golden(ob):
    g = gold("kmeans_seed777")
    v = np.random.RandomState(1234).random_sample((20000, 32)).astype(np.float32)
    for name, flags in (("random", ob.KMEANS_QUIET | 4), ("pp", ob.KMEANS_QUIET | ob.KMEANS_INIT_BERKELEY | 4)):
        qe, cent, dis, assign, nassign = ob.orc_kmeans(v, 64, 15, flags, 777, redo=2)
        assert np.array_equal(nassign, g[name + "_nassign"])
        assert np.array_equal(assign, g[name + "_assign"])
        assert np.array_equal(cent, g[name + "_cent"])
        assert np.float32(qe) == g[name + "_qerr"]


def test_kmeans_empty_split_golden(ob):
    g = gold("kmeans_empty_split")
    qe, cent, dis, assign, nassign = ob.orc_kmeans(g["v"], 40, 10, ob.KMEANS_QUIET | 1, 5)
    assert np.array_equal(nassign, g["nassign"]) and np.array_equal(cent, g["cent"])


def test_rng_golden(ob):
    g = gold("rng")
    import ctypes as C
    p = ob.oracle().orc_random_perm_r(1000, 4242)
    perm = np.ctypeslib.as_array(p, shape=(1000,)).copy()
    assert np.array_equal(perm, g["perm_n1000_seed4242"])
    x = np.empty(257, np.float32)
    ob.oracle().orc_fvec_randn_r(ob.fp(x), 257, 99)
    assert np.array_equal(x, g["randn_n257_seed99"])


def test_kmin_golden(ob):
    g = gold("kmin")
    val = g["val"]
    assert np.array_equal(ob.orc_k_min(val, 7), g["k7"])
    assert np.array_equal(ob.orc_k_min(val, 100), g["k100"])
    assert np.array_equal(ob.orc_k_min(val, 1), g["k1"])
    assert np.array_equal(ob.orc_k_min(val[:1000], 100), g["small100"])
    vt = np.random.RandomState(4).randint(0, 50, 100000).astype(np.float32)
    assert np.array_equal(ob.orc_k_min(vt, 1), g["ties_k1"])
    # test/matlab/test_kmin.m:20: k-min == prefix of the full sort
    assert np.array_equal(ob.orc_k_min(val, 100, canonical=True), np.argsort(val, kind="stable")[:100])


def test_hamming_golden(ob):
    g = gold("hamming")
    for nc in (4, 8, 16, 24, 5):
        a, b = g["a%d" % nc], g["b%d" % nc]
        assert np.array_equal(ob.orc_compute_hamming(a, b), g["dis%d" % nc])
        # independent check: popcount via unpackbits
        want = (np.unpackbits(a[None, :, :] ^ b[:, None, :], axis=2).sum(2)).astype(np.uint16)
        assert np.array_equal(g["dis%d" % nc], want)
    import ctypes as C
    a, b = g["a8"], g["b8"]
    n = C.c_size_t(0)
    ob.oracle().orc_match_hamming_count(ob.u8p(a), ob.u8p(b), 23, 31, 28, 8, C.byref(n))
    assert n.value == len(g["match_ht28_ham"])
    idx = np.empty((n.value, 2), np.int32)
    ham = np.empty(n.value, np.uint16)
    ob.oracle().orc_match_hamming_thres_prealloc(ob.u8p(a), ob.u8p(b), 23, 31, 28, 8, ob.ip(idx), ob.u16p(ham))
    assert np.array_equal(idx, g["match_ht28_idx"]) and np.array_equal(ham, g["match_ht28_ham"])


def test_crossmatch_hamming_golden(ob):
    # crossmatch_hamming_count / _prealloc of the compiled reference (yael/hamming.c:368-395, 793-829)
    import ctypes as C
    g = gold("hamming_crossmatch")
    for nc in (4, 8, 16, 5):
        db, ht = g["db%d" % nc], int(g["ht%d" % nc])
        n = C.c_size_t(0)
        ob.oracle().orc_crossmatch_hamming_count(ob.u8p(db), len(db), ht, nc, C.byref(n))
        assert n.value == len(g["ham%d" % nc])
        idx = np.empty((n.value, 2), np.int32)
        ham = np.empty(n.value, np.uint16)
        m = ob.oracle().orc_crossmatch_hamming_prealloc(ob.u8p(db), len(db), ht, nc, ob.ip(idx), ob.u16p(ham))
        assert m == n.value
        assert np.array_equal(idx, g["idx%d" % nc]) and np.array_equal(ham, g["ham%d" % nc])
        # independent check: upper triangle of the unpackbits distance matrix
        full = np.unpackbits(db[:, None, :] ^ db[None, :, :], axis=2).sum(2)
        i, j = np.nonzero(np.triu(full <= ht, 1))
        assert np.array_equal(idx[:, 0], i) and np.array_equal(idx[:, 1], j)


def test_nn_hamming_oracle_is_stable_select(ob):
    r = np.random.RandomState(9)
    b = r.randint(0, 256, (3000, 8)).astype(np.uint8)
    q = r.randint(0, 256, (20, 8)).astype(np.uint8)
    idx, dis = ob.orc_nn_hamming(b, q, 50)
    full = ob.orc_compute_hamming(b, q)  # [nq][nb]
    for j in range(20):
        order = np.lexsort((np.arange(3000), full[j]))[:50]
        assert np.array_equal(idx[j], order) and np.array_equal(dis[j], full[j][order])


@pytest.mark.skipif(not os.path.exists(os.path.join(os.path.dirname(GOLD), "..", "oracle", "_ref", "libyael_ref.so")),
                    reason="oracle/_ref not built")
def test_oracle_matches_compiled_reference_fresh_inputs(ob):
    r = np.random.RandomState(77)
    b = r.randint(0, 4, (3000, 16)).astype(np.float32)  # exact arithmetic: tie order must match too
    q = r.randint(0, 4, (200, 16)).astype(np.float32)
    for k in (1, 5, 50):
        i0, d0 = ob.ref_knn(b, q, k, nt=3)
        i1, d1 = ob.orc_knn(b, q, k, nt=3)
        assert np.array_equal(i0, i1) and np.array_equal(d0, d1)
    v = r.random_sample((5000, 16)).astype(np.float32)
    r0 = ob.ref_kmeans(v, 32, 10, ob.KMEANS_QUIET | 2, 4321)
    r1 = ob.orc_kmeans(v, 32, 10, ob.KMEANS_QUIET | 2, 4321)
    assert np.array_equal(r0[1], r1[1]) and np.array_equal(r0[4], r1[4])


def _golden_subsets(g):
    ends = g["subset_ends"]
    idx = g["subset_indexes"]
    return [idx[(ends[i - 1] if i else 0):ends[i]].tolist() for i in range(len(ends))]


def test_vlad_bof_golden(ob):
    # yael/vlad.c:10-139, golden from the compiled reference (scripts/make_golden.py, fixture 11)
    g = gold("vlad_bof")
    c, v, subs = g["centroids"], g["v"], _golden_subsets(g)
    assert np.array_equal(ob.orc_vlad(c, v), g["vlad"])
    assert np.array_equal(ob.orc_vlad(c, v, weights=g["weights"]), g["vlad_weighted"])
    assert np.array_equal(ob.orc_vlad(c, v, subsets=subs), g["vlad_subsets"])
    assert np.array_equal(ob.orc_bof(c, v), g["bof"])
    assert np.array_equal(ob.orc_bof(c, v, ma=3), g["bof_ma3"])
    assert np.array_equal(ob.orc_bof(c, v, subsets=subs), g["bof_subsets"])
    assert g["bof"].sum() == len(v) and g["bof_ma3"].sum() == 3 * len(v)


def test_hkm_quantize_and_gmm_posteriors_golden(ob):
    # yael/hkm.c:144-162 and yael/gmm.c:211-367, golden from the compiled reference
    # (scripts/make_golden.py, fixture 12): leaves identical; posteriors identical with the
    # sequential-FMA dot order and within 1e-4 with the float64 order (BLAS order is unspecified)
    g = gold("hkm_gmm")
    levels = [g["hkm_level%d" % l] for l in range(3)]
    assert np.array_equal(ob.orc_hkm_quantize(levels, 5, g["hkm_query"]), g["hkm_quantize_query"])
    assert np.array_equal(ob.orc_hkm_quantize(levels, 5, g["hkm_points"]), g["hkm_quantize_points"])
    for flags, name in ((1, "gmm_p_w"), (0, "gmm_p_now")):
        p = ob.orc_gmm_compute_p(g["gmm_w"], g["gmm_mu"], g["gmm_sigma"], g["gmm_v"], flags)
        assert np.array_equal(p, g[name])
        p64 = ob.orc_gmm_compute_p(g["gmm_w"], g["gmm_mu"], g["gmm_sigma"], g["gmm_v"], flags, ob.DOT_F64)
        np.testing.assert_allclose(p64, g[name], atol=1e-4)
        np.testing.assert_allclose(g[name].sum(1), 1.0, atol=1e-5)
    assert np.array_equal(g["gmm_p_w"], g["gmm_p_w_nt3"])   # independent of the thread count
    if ob.have_ref():
        assert np.array_equal(ob.ref_hkm_quantize(levels, 5, g["hkm_query"]), g["hkm_quantize_query"])
        assert np.array_equal(ob.ref_gmm_compute_p(g["gmm_w"], g["gmm_mu"], g["gmm_sigma"], g["gmm_v"], 1),
                              g["gmm_p_w"])

"""GPU parity of the tcgen05 TF32 engine: raw tensor-core scores against the error model the
certificate relies on, and the full shortlist -> re-rank pipeline against the oracle.

Tolerances (BASELINE.json north_star): L2 distances within 1e-5 relative of the reference's
CPU implementation, ids identical except for ties inside that tolerance.  The re-rank uses the
reference's arithmetic, so in practice the distances are bit-identical."""
import numpy as np
import pytest

import yael_b200
from devmem import DevArray, check_knn

pytestmark = pytest.mark.gpu


def rs(seed):
    return np.random.RandomState(seed)


@pytest.mark.parametrize("nq,nb,d", [(128, 256, 128), (128, 256, 32), (130, 700, 128), (77, 1000, 96),
                                     (300, 5000, 64), (5, 300, 100), (256, 2048, 8)])
def test_tf32_scores_within_error_model(nq, nb, d):
    L = yael_b200.lib()
    r = rs(nq + nb + d)
    base = r.rand(nb, d).astype(np.float32)
    query = r.rand(nq, d).astype(np.float32)
    db, dq = DevArray(base), DevArray(query)
    out = DevArray(shape=(nq, nb), dtype=np.float32)
    rc = L.yb_debug_tf32_scores(nq, nb, d, db.ptr, dq.ptr, out.ptr, None)
    assert rc == 0, L.yb_last_error()
    L.yb_sync(None)
    got = out.get()
    b64, q64 = base.astype(np.float64), query.astype(np.float64)
    exact = (b64 * b64).sum(1)[None, :] - 2.0 * q64 @ b64.T
    err = np.abs(got - exact)
    bound = (1.05 / 256.0) * np.linalg.norm(q64, axis=1)[:, None] * np.linalg.norm(b64, axis=1).max()
    assert np.isfinite(got).all()
    assert (err <= bound + 1e-5).all(), "max err %g vs bound %g" % (err.max(), bound.min())
    # and it really is a TF32-class result, not an FP32 one or garbage
    assert err.max() < 0.05 * np.abs(exact).max()
    for a in (db, dq, out):
        a.free()


@pytest.mark.parametrize("nq,nb,d,scale", [(128, 256, 128, 1.0), (130, 700, 128, 1e-6), (77, 1000, 96, 3e4),
                                           (300, 5000, 64, 1.0), (5, 300, 100, 255.0), (256, 2048, 8, 1.0),
                                           (64, 1500, 20, 1e-3)])
def test_f16_scores_within_error_model(nq, nb, d, scale):
    # FP16 operands (kind::f16), power-of-two scale chosen on the device: rounding to nearest keeps
    # 11 significant bits per operand, so the score is within 2 * 2^-10 |q||b| of exact
    L = yael_b200.lib()
    r = rs(nq + nb + d + 1)
    base = ((r.rand(nb, d) - 0.3) * scale).astype(np.float32)
    query = ((r.rand(nq, d) - 0.3) * scale).astype(np.float32)
    base[7, : d // 2] = 0.0          # zeros and tiny values (sub-normal after conversion)
    base[9] *= np.float32(1e-7)
    db, dq = DevArray(base), DevArray(query)
    out = DevArray(shape=(nq, nb), dtype=np.float32)
    rc = L.yb_debug_f16_scores(nq, nb, d, db.ptr, dq.ptr, out.ptr, None)
    assert rc == 0, L.yb_last_error()
    L.yb_sync(None)
    got = out.get()
    b64, q64 = base.astype(np.float64), query.astype(np.float64)
    exact = (b64 * b64).sum(1)[None, :] - 2.0 * q64 @ b64.T
    err = np.abs(got - exact)
    qn, bn = np.linalg.norm(q64, axis=1)[:, None], np.linalg.norm(b64, axis=1)
    bound = (1.05 / 512.0) * qn * bn.max() + 1e-6 * np.abs(exact) + 1e-30
    assert np.isfinite(got).all()
    assert (err <= bound).all(), "max err ratio %g" % (err / bound).max()
    assert err.max() > 0  # it is a reduced-precision result, not an FP32 one
    for a in (db, dq, out):
        a.free()


def test_knn_f16_out_of_range_value_falls_back_to_tf32(yn, ob):
    # the FP16 scale comes from a SAMPLE of the rows (8x head room); an un-sampled outlier beyond it
    # raises the overflow flag and the pass is repeated on TF32 operands -- same exact result
    L = yael_b200.lib()
    L.yb_set_knn_engine(1)
    try:
        r = rs(99)
        nb, nq, d, k = 100000, 64, 64, 10
        b = r.rand(nb, d).astype(np.float32)
        q = r.rand(nq, d).astype(np.float32)
        idx, dis = yn.knn(q, b, k)
        assert L.yb_last_knn_engine() == 1 and L.yb_last_knn_operands() == 2
        widx, wdis = ob.orc_knn(b, q, k, ob.DOT_F32_SEQ, canonical=True)
        check_knn(idx, dis, widx, wdis, b, q)
        b[500] = 1e3  # rows 256..780 are not in the sample: 1e3 >> 8 x the sampled maximum
        idx, dis = yn.knn(q, b, k)
        assert L.yb_last_knn_engine() == 1 and L.yb_last_knn_operands() == 0
        widx, wdis = ob.orc_knn(b, q, k, ob.DOT_F32_SEQ, canonical=True)
        check_knn(idx, dis, widx, wdis, b, q)
    finally:
        L.yb_set_knn_engine(-1)


def test_knn_too_tight_thresholds_are_retried_and_still_exact(yn, ob, monkeypatch):
    # YAEL_B200_J2 forces admission thresholds far too tight: many queries come up short of
    # candidates, fail their certificate, and go through the tensor retry (4x looser) and, if need
    # be, the exact engine -- the result must not change
    L = yael_b200.lib()
    L.yb_set_knn_engine(1)
    try:
        r = rs(21)
        nb, nq, d, k = 120000, 400, 64, 50
        b = r.rand(nb, d).astype(np.float32)
        q = r.rand(nq, d).astype(np.float32)
        widx, wdis = ob.orc_knn(b, q, k, ob.DOT_F32_SEQ, canonical=True)
        for j2 in ("2", "5"):
            monkeypatch.setenv("YAEL_B200_J2", j2)
            idx, dis = yn.knn(q, b, k)
            assert L.yb_last_knn_engine() == 1
            assert L.yb_last_knn_uncertified() > 0      # the scenario really happened
            check_knn(idx, dis, widx, wdis, b, q)
    finally:
        L.yb_set_knn_engine(-1)


@pytest.mark.parametrize("operands", ["tf32", "f16"])
def test_knn_both_operand_kinds_match_oracle(yn, ob, operands, monkeypatch):
    monkeypatch.setenv("YAEL_B200_OPERANDS", operands)
    L = yael_b200.lib()
    L.yb_set_knn_engine(1)
    try:
        r = rs(5)
        for nb, nq, d, k in ((60000, 300, 128, 100), (50000, 1000, 100, 1), (40000, 100, 24, 7)):
            b = (r.rand(nb, d) * 200).astype(np.float32)   # SIFT-like magnitudes
            q = (r.rand(nq, d) * 200).astype(np.float32)
            idx, dis = yn.knn(q, b, k)
            assert L.yb_last_knn_engine() == 1
            assert L.yb_last_knn_operands() == (0 if operands == "tf32" else 2)
            widx, wdis = ob.orc_knn(b, q, k, ob.DOT_F32_SEQ, canonical=True)
            check_knn(idx, dis, widx, wdis, b, q)
    finally:
        L.yb_set_knn_engine(-1)


@pytest.fixture
def tf32_engine():
    L = yael_b200.lib()
    L.yb_set_knn_engine(1)
    yield L
    L.yb_set_knn_engine(-1)


@pytest.mark.parametrize("nq,nb,d,k", [(256, 20000, 128, 10), (100, 5000, 128, 100), (1000, 30000, 96, 1),
                                        (130, 3000, 64, 5), (64, 100000, 32, 50), (700, 2500, 128, 100)])
def test_knn_tf32_engine_matches_oracle(yn, ob, tf32_engine, nq, nb, d, k):
    r = rs(nq * 7 + nb + k)
    b = r.rand(nb, d).astype(np.float32)
    q = r.rand(nq, d).astype(np.float32)
    idx, dis = yn.knn(q, b, k)
    assert tf32_engine.yb_last_knn_engine() == 1
    widx, wdis = ob.orc_knn(b, q, k, ob.DOT_F32_SEQ, canonical=True)
    check_knn(idx, dis, widx, wdis, b, q)
    # uniform data never needs the exact fallback in bulk
    assert tf32_engine.yb_last_knn_uncertified() <= nq // 20


def test_knn_tf32_sift_like_integers(yn, ob, tf32_engine):
    # integer coordinates 0..255 are exact in TF32: the tensor pass is exact, ties abound
    r = rs(42)
    b = np.minimum(255, r.gamma(1.2, 25.0, (20000, 128))).astype(np.int32).astype(np.float32)
    q = np.minimum(255, r.gamma(1.2, 25.0, (200, 128))).astype(np.int32).astype(np.float32)
    idx, dis = yn.knn(q, b, 100)
    widx, wdis = ob.orc_knn(b, q, 100, canonical=True)
    assert np.array_equal(dis, wdis)
    assert np.array_equal(idx, widx)


def test_knn_tf32_adversarial_near_duplicates(yn, ob, tf32_engine):
    # clusters of near-duplicates with large norms: the certificate must route the hard
    # queries to the exact engine instead of returning a wrong neighbour
    r = rs(7)
    centers = (r.rand(50, 64) * 100).astype(np.float32)
    b = (centers[r.randint(0, 50, 20000)] + r.randn(20000, 64) * 1e-3).astype(np.float32)
    q = (centers[r.randint(0, 50, 300)] + r.randn(300, 64) * 1e-3).astype(np.float32)
    idx, dis = yn.knn(q, b, 10)
    widx, wdis = ob.orc_knn(b, q, 10, canonical=True)
    assert np.array_equal(dis, wdis)
    assert np.array_equal(idx, widx)


def test_knn_tf32_nan_and_padding(yn, ob, tf32_engine):
    r = rs(8)
    b = r.rand(3000, 32).astype(np.float32)
    b[5::7] = np.nan
    q = r.rand(150, 32).astype(np.float32)
    idx, dis = yn.knn(q, b, 20)
    widx, wdis = ob.orc_knn(b, q, 20, canonical=True)
    check_knn(idx, dis, widx, wdis, b, q)
    assert not np.isin(idx, np.arange(5, 3000, 7)).any()


def test_kmeans_assignment_tf32(yn, ob, tf32_engine):
    r = rs(1234)
    v = r.rand(50000, 128).astype(np.float32)
    c0 = v[r.permutation(50000)[:256]].copy()
    cent, qerr, dis, assign, nassign = yn.kmeans(v, 256, niter=1, verbose=False, init=c0, output="all")
    q, wc, wa, wd, wn = ob.orc_kmeans_step(v, c0)
    mism = assign != wa
    # any disagreement must be a tie inside 1e-5 relative
    assert np.all(np.abs(dis[mism] - wd[mism]) <= 1e-5 * wd[mism])
    assert mism.mean() < 1e-3
    np.testing.assert_allclose(dis, wd, rtol=1e-5)
    np.testing.assert_allclose(cent, wc, atol=1e-4)


# ---------------------------------------------------------------- database fed from host memory
def _resident_knn(L, b, q, k):
    db, dq = DevArray(b), DevArray(q)
    oi = DevArray(shape=(q.shape[0], k), dtype=np.int32)
    od = DevArray(shape=(q.shape[0], k), dtype=np.float32)
    rc = L.yb_knn_l2(q.shape[0], b.shape[0], b.shape[1], k, db.ptr, dq.ptr, None, oi.ptr, od.ptr, 0, None)
    assert rc == 0, L.yb_last_error()
    L.yb_sync(None)
    out = oi.get(), od.get()
    for a in (db, dq, oi, od):
        a.free()
    return out


@pytest.mark.parametrize("nb,d,k,chunks", [(200003, 128, 10, None), (262144 + 4096 + 300, 96, 100, 3),
                                           (196608 + 100, 128, 7, 8), (400000, 64, 32, 2)])
def test_knn_host_database_streamed_equals_resident(yn, tf32_engine, monkeypatch, nb, d, k, chunks):
    """knn_full() on a host-resident database overlaps the transfer with the scan (sample tiles
    first, chunked tensor passes, yb_knn_l2_hostbase); the result must be the resident path's,
    bit for bit, ragged tails and any chunk count included."""
    if chunks:
        monkeypatch.setenv("YAEL_B200_H2D_CHUNKS", str(chunks))
    r = rs(nb + d + k)
    b = r.rand(nb, d).astype(np.float32)
    q = r.rand(300, d).astype(np.float32)
    idx, dis = yn.knn(q, b, k)          # host arrays -> streamed path
    assert tf32_engine.yb_last_knn_engine() == 1
    widx, wdis = _resident_knn(tf32_engine, b, q, k)
    assert np.array_equal(dis, wdis)
    assert np.array_equal(idx, widx)
    monkeypatch.setenv("YAEL_B200_NO_STREAMED_H2D", "1")
    idx2, dis2 = yn.knn(q, b, k)        # copy-then-scan
    assert np.array_equal(dis2, wdis) and np.array_equal(idx2, widx)


def test_knn_host_database_streamed_sorted_rows(yn, ob, tf32_engine):
    """Rows sorted by cluster (position correlates with content): thresholds come from sample
    tiles spread over the whole database, so nothing degrades; checked against the oracle."""
    r = rs(99)
    nb, d, k = 230000, 128, 20
    centers = r.rand(64, d).astype(np.float32) * 4
    lab = np.sort(r.randint(0, 64, nb))
    b = (centers[lab] + r.randn(nb, d) * 0.05).astype(np.float32)
    q = (centers[r.randint(0, 64, 200)] + r.randn(200, d) * 0.05).astype(np.float32)
    idx, dis = yn.knn(q, b, k)
    n_streamed = tf32_engine.yb_last_knn_uncertified()
    widx, wdis = ob.orc_knn(b, q, k, canonical=True)
    check_knn(idx, dis, widx, wdis, b, q)
    ridx, rdis = _resident_knn(tf32_engine, b, q, k)
    assert np.array_equal(dis, rdis) and np.array_equal(idx, ridx)
    # tight clusters defeat the TF32 certificate for many queries in BOTH paths (they are then
    # answered by the exact engine); feeding from the host must not make that worse
    assert n_streamed <= tf32_engine.yb_last_knn_uncertified() + 5


# ---------------------------------------------------------------- queries at the centring vector
def _thin_shell(r, nb, d, center, radius, ulps):
    """Rows on a thin shell around `center`: |b - center|^2 = radius^2 (1 + j 2^-22), j < ulps, so
    their exact distances to the centre are a few FP32 ulps apart."""
    u = r.randn(nb, d)
    u /= np.linalg.norm(u, axis=1)[:, None]
    scale = radius * np.sqrt(1.0 + r.randint(0, ulps, nb) * 2.0 ** -22)
    return (center[None, :] + u * scale[:, None]).astype(np.float32)


@pytest.mark.parametrize("k", [1, 10])
def test_knn_queries_at_and_near_the_column_mean(yn, ob, tf32_engine, k):
    """The Cauchy-Schwarz part of the tensor-score error bound is proportional to |q - mu| and
    vanishes for a query AT the centring vector, while the FP32 roundings of |b - mu|^2 and of the
    accumulator do not: the certificate / the k = 1 margin carry an absolute term for them.  Rows on
    a thin shell around the mean (distances 1-4 ulp apart) must therefore fail the certificate and be
    answered by the exact engine -- never a wrong id."""
    r = rs(31 + k)
    nb, d = 30000, 64                       # <= 32768 rows: the device's mean is over all of them
    center = (r.rand(d) * 3 + 1).astype(np.float64)
    b = _thin_shell(r, nb, d, center, 2.0, 5)
    mu = b.astype(np.float64).mean(0)
    q = np.stack([mu, mu + 1e-6 * r.randn(d), mu * (1 + 1e-7), mu + 1e-4 * r.randn(d),
                  mu + 1e-2 * r.randn(d)] + [b[i] * 0.5 + mu * 0.5 for i in range(11)]).astype(np.float32)
    idx, dis = yn.knn(q, b, k)
    assert tf32_engine.yb_last_knn_engine() == 1
    widx, wdis = ob.orc_knn(b, q, k, canonical=True)
    assert np.array_equal(dis, wdis)
    assert np.array_equal(idx, widx)
    # the near-mean queries cannot be certified on the tensor path: they went to the exact engine
    assert tf32_engine.yb_last_knn_uncertified() >= 3


def test_knn_query_at_the_mean_with_separated_rows(yn, ob, tf32_engine):
    # same query positions, but rows at well separated radii: the tensor path certifies them
    r = rs(77)
    nb, d, k = 30000, 64, 10
    center = (r.rand(d) * 3 + 1).astype(np.float64)
    u = r.randn(nb, d)
    u /= np.linalg.norm(u, axis=1)[:, None]
    b = (center[None, :] + u * (0.5 + 2.0 * r.rand(nb))[:, None]).astype(np.float32)
    mu = b.astype(np.float64).mean(0)
    q = np.stack([mu, mu + 1e-6 * r.randn(d), mu + 1e-3 * r.randn(d)]).astype(np.float32)
    idx, dis = yn.knn(q, b, k)
    widx, wdis = ob.orc_knn(b, q, k, canonical=True)
    assert np.array_equal(dis, wdis) and np.array_equal(idx, widx)


def test_kmeans_points_at_the_centroid_mean(yn, ob, tf32_engine):
    """k-means assignment (k = 1 margin mode): points at / near the mean of the centroids, centroids
    on a thin shell around it.  The margin must not collapse with |q - mu|."""
    r = rs(5)
    k, d, n = 512, 32, 40000
    center = (r.rand(d) * 2).astype(np.float64)
    c0 = _thin_shell(r, k, d, center, 1.5, 5)
    mu = c0.astype(np.float64).mean(0)
    v = np.concatenate([np.tile(mu, (2000, 1)), mu + 1e-6 * r.randn(2000, d), mu + 1e-3 * r.randn(6000, d),
                        c0[r.randint(0, k, n - 10000)] + 0.05 * r.randn(n - 10000, d)]).astype(np.float32)
    cent, qerr, dis, assign, nassign = yn.kmeans(v, k, niter=1, verbose=False, init=c0, output="all")
    _, wc, wa, wd, wn = ob.orc_kmeans_step(v, c0)
    assert np.array_equal(dis, wd)
    assert np.array_equal(assign, wa)
    assert np.array_equal(nassign, wn)


# ---------------------------------------------------------------- d beyond the resident query tile
@pytest.mark.parametrize("nq,nb,d,k", [(300, 20000, 256, 10), (256, 30000, 384, 100), (400, 8000, 960, 20),
                                        (1000, 20000, 200, 1), (512, 10000, 960, 1), (260, 5000, 1000, 7)])
def test_knn_tensor_engine_large_d_matches_oracle(yn, ob, tf32_engine, nq, nb, d, k):
    # d > 240: the query tile no longer fits in shared memory next to the database ring; the 2-SM kernel
    # streams the query chunks through the ring with the database chunks (k_knn_2sm<.., STREAM>).
    # GIST1M has d = 960; the reference has no limit on d (yael/nn.c:451-525).
    r = rs(nq + nb + d + k)
    b = r.rand(nb, d).astype(np.float32)
    q = r.rand(nq, d).astype(np.float32)
    idx, dis = yn.knn(q, b, k)
    assert tf32_engine.yb_last_knn_engine() == 1, "d = %d fell back to the exact engine" % d
    widx, wdis = ob.orc_knn(b, q, k, ob.DOT_F32_SEQ, canonical=True)
    check_knn(idx, dis, widx, wdis, b, q)
    assert tf32_engine.yb_last_knn_uncertified() <= nq // 10


@pytest.mark.parametrize("d,k", [(128, 100), (96, 10), (64, 1), (128, 1)])
def test_knn_streamed_query_chunks_equal_resident_tile(yn, tf32_engine, monkeypatch, d, k):
    # the streamed variant forced at a d the resident layout handles too: identical results
    r = rs(d + k)
    b = r.rand(40000, d).astype(np.float32)
    q = r.rand(520, d).astype(np.float32)
    i0, d0 = yn.knn(q, b, k)
    monkeypatch.setenv("YAEL_B200_STREAM", "1")
    i1, d1 = yn.knn(q, b, k)
    assert tf32_engine.yb_last_knn_engine() == 1
    assert np.array_equal(i0, i1) and np.array_equal(d0, d1)


def test_kmeans_assignment_large_d(yn, ob, tf32_engine):
    r = rs(99)
    v = r.rand(20000, 320).astype(np.float32)
    c0 = v[r.permutation(20000)[:300]].copy()
    cent, qerr, dis, assign, nassign = yn.kmeans(v, 300, niter=1, verbose=False, init=c0, output="all")
    q, wc, wa, wd, wn = ob.orc_kmeans_step(v, c0)
    mism = assign != wa
    assert np.all(np.abs(dis[mism] - wd[mism]) <= 1e-5 * wd[mism])
    assert mism.mean() < 1e-3
    np.testing.assert_allclose(dis, wd, rtol=1e-5)
    np.testing.assert_allclose(cent, wc, atol=1e-4)


# ---------------------------------------------------------------- k = 1 re-rank: lane per (query, candidate) pair
@pytest.mark.parametrize("nq,nb,d", [(1000, 30000, 100), (333, 9000, 36), (2049, 70000, 128), (65, 4000, 8),
                                     (500, 20000, 124)])
def test_knn_k1_lane_rerank_matches_oracle_and_warp_kernel(yn, ob, tf32_engine, monkeypatch, nq, nb, d):
    # k_rerank_k1_lanes (d <= 128, d % 4 == 0): rows shorter than 128 coordinates, a ragged last
    # warp of queries, and the same answer as the warp-per-query kernel (YAEL_B200_K1_LANES=0)
    r = rs(nq + nb + d)
    b = r.rand(nb, d).astype(np.float32)
    q = r.rand(nq, d).astype(np.float32)
    b[7] = b[3]                      # exact duplicate rows: the lowest id wins (nn.c:404-440)
    q[0] = b[3]
    idx, dis = yn.knn(q, b, 1)
    assert tf32_engine.yb_last_knn_engine() == 1
    widx, wdis = ob.orc_knn(b, q, 1, ob.DOT_F32_SEQ, canonical=True)
    check_knn(idx, dis, widx, wdis, b, q)
    assert idx[0, 0] == 3
    import subprocess, sys, os, json
    code = ("import numpy as np, json, sys; sys.path.insert(0, %r); import yael_b200; from yael_b200 import ynumpy as yn;"
            "yael_b200.lib().yb_set_knn_engine(1); r = np.random.RandomState(%d);"
            "b = r.rand(%d, %d).astype(np.float32); q = r.rand(%d, %d).astype(np.float32); b[7] = b[3]; q[0] = b[3];"
            "i, d = yn.knn(q, b, 1); print(json.dumps([i.ravel().tolist(), d.view(np.int32).ravel().tolist()]))"
            % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), (nq + nb + d) % (2 ** 31), nb, d, nq, d))
    env = dict(os.environ, YAEL_B200_K1_LANES="0")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    oi, od = json.loads(out.stdout.strip().splitlines()[-1])
    assert np.array_equal(np.array(oi, np.int32), idx.ravel())
    assert np.array_equal(np.array(od, np.int32), dis.view(np.int32).ravel())


def test_knn_k1_many_near_ties_overflow_goes_to_exact_engine(yn, ob, tf32_engine):
    # hundreds of rows within the margin of the best score: more than RL_CAP (16) candidates survive the
    # score filter, the query is flagged and redone by the exact engine -- still the reference's answer
    r = rs(99)
    d, nb = 64, 20000
    centre = r.rand(d).astype(np.float32)
    b = r.rand(nb, d).astype(np.float32)
    b[:600] = centre + (1e-6 * r.randn(600, d)).astype(np.float32)
    q = np.tile(centre, (40, 1)).astype(np.float32) + (1e-6 * r.randn(40, d)).astype(np.float32)
    idx, dis = yn.knn(q, b, 1)
    widx, wdis = ob.orc_knn(b, q, 1, ob.DOT_F32_SEQ, canonical=True)
    check_knn(idx, dis, widx, wdis, b, q)

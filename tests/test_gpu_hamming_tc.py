"""GPU parity of the tensor-core Hamming engine (yb_hamming_tc.cu): nn_hamming as an exact
E4M3 (+-1) contraction on tcgen05, reusing the fused threshold / top-k epilogue of the
distance GEMM kernel.

Bar: BIT-EXACT ids and uint16 distances, (distance, id) order -- against the oracle
(orc_nn_hamming = compute_hamming, yael/hamming.c:177-219, + stable selection) at sizes it
finishes in seconds, and at BASELINE.json's full size (10M x 64 bit, 10k queries, k = 100) against
the popcount engine, which the oracle pins at the small sizes."""
import ctypes as C

import numpy as np
import pytest

import yael_b200
from devmem import DevArray

pytestmark = pytest.mark.gpu


def rs(seed):
    return np.random.RandomState(seed)


@pytest.fixture
def tc_engine():
    L = yael_b200.lib()
    L.yb_set_hamming_engine(1)
    yield L
    L.yb_set_hamming_engine(-1)


@pytest.mark.parametrize("nq,nb,nc", [(128, 256, 8), (130, 700, 8), (77, 1000, 16), (300, 5000, 4),
                                      (5, 300, 32), (256, 2048, 64), (200, 3000, 24), (33, 999, 5)])
def test_e4m3_scores_are_four_times_hamming(nq, nb, nc):
    # the raw tensor-core pass: every score must be EXACTLY 4 * popcount(q xor b)
    L = yael_b200.lib()
    r = rs(nq + nb + nc)
    base = r.randint(0, 256, (nb, nc)).astype(np.uint8)
    query = r.randint(0, 256, (nq, nc)).astype(np.uint8)
    base[::13] = query[0]          # distance 0
    base[1::13] = ~query[min(1, nq - 1)]  # distance = all bits
    db, dq = DevArray(base), DevArray(query)
    out = DevArray(shape=(nq, nb), dtype=np.float32)
    rc = L.yb_debug_hamming_tc_scores(nq, nb, nc, db.ptr, dq.ptr, out.ptr, None)
    assert rc == 0, L.yb_last_error()
    L.yb_sync(None)
    got = out.get()
    want = 4.0 * np.unpackbits(query[:, None, :] ^ base[None, :, :], axis=2).sum(2)
    assert np.array_equal(got, want.astype(np.float32))
    for a in (db, dq, out):
        a.free()


@pytest.mark.parametrize("nq,nb,nc,k", [(300, 40000, 8, 100), (130, 70000, 16, 10), (64, 33000, 4, 50),
                                         (257, 50000, 8, 1), (40, 36000, 32, 33), (700, 3000, 8, 20),
                                         (100, 66000, 5, 64), (10, 100000, 64, 100)])
def test_nn_hamming_tensor_engine_bit_exact(yn, ob, tc_engine, nq, nb, nc, k):
    r = rs(nq * 3 + nb + nc)
    b = r.randint(0, 256, (nb, nc)).astype(np.uint8)
    q = r.randint(0, 256, (nq, nc)).astype(np.uint8)
    b[::17] = q[0]       # planted exact duplicates: distance-0 ties resolved by id
    b[5::1001] = q[-1]
    b[7::1001, 0] ^= 1   # and near-duplicates
    idx, dis = yn.knn_hamming(q, b, k)
    assert tc_engine.yb_last_hamming_engine() == 1
    widx, wdis = ob.orc_nn_hamming(b, q, k)
    assert np.array_equal(dis, wdis)
    assert np.array_equal(idx, widx)


def test_tensor_engine_survives_adversarial_row_order(yn, ob, tc_engine):
    # rows sorted by distance to query 0: the sampled thresholds are useless for it and lists
    # overflow; the answer must still be exact (list compaction, certificate, scan fallback)
    r = rs(77)
    nb, nq, k = 60000, 200, 100
    b = r.randint(0, 256, (nb, 8)).astype(np.uint8)
    q = r.randint(0, 256, (nq, 8)).astype(np.uint8)
    d0 = np.unpackbits(b ^ q[0], axis=1).sum(1)
    b = np.ascontiguousarray(b[np.argsort(-d0, kind="stable")])  # best rows come last
    idx, dis = yn.knn_hamming(q, b, k)
    widx, wdis = ob.orc_nn_hamming(b, q, k)
    assert np.array_equal(dis, wdis) and np.array_equal(idx, widx)
    # all-identical database: every distance ties, ids decide
    b2 = np.repeat(q[:1], 40000, axis=0)
    idx, dis = yn.knn_hamming(q[:50], b2, 64)
    widx, wdis = ob.orc_nn_hamming(b2, q[:50], 64)
    assert np.array_equal(dis, wdis) and np.array_equal(idx, widx)


def test_engine_selection_and_offsets(yn, ob):
    L = yael_b200.lib()
    r = rs(5)
    b = r.randint(0, 256, (2000, 8)).astype(np.uint8)
    q = r.randint(0, 256, (20, 8)).astype(np.uint8)
    yn.knn_hamming(q, b, 5)
    assert L.yb_last_hamming_engine() == 0  # small problems stay on the popcount scan
    # device-level call with an id offset (what a database shard passes), both engines
    nb, nq, k = 50000, 300, 30
    b = r.randint(0, 256, (nb, 8)).astype(np.uint8)
    q = r.randint(0, 256, (nq, 8)).astype(np.uint8)
    db, dq = DevArray(b), DevArray(q)
    res = []
    for eng in (0, 1):
        L.yb_set_hamming_engine(eng)
        oi = DevArray(shape=(nq, k), dtype=np.int32)
        od = DevArray(shape=(nq, k), dtype=np.uint16)
        rc = L.yb_nn_hamming(nq, nb, 8, k, db.ptr, dq.ptr, oi.ptr, od.ptr, 1000000, None)
        assert rc == 0, L.yb_last_error()
        L.yb_sync(None)
        assert L.yb_last_hamming_engine() == eng
        res.append((oi.get(), od.get()))
        oi.free()
        od.free()
    L.yb_set_hamming_engine(-1)
    widx, wdis = ob.orc_nn_hamming(b, q, k)
    for gi, gd in res:
        assert np.array_equal(gi, widx + 1000000) and np.array_equal(gd, wdis)
    db.free()
    dq.free()


def test_baseline_shape_tensor_equals_popcount_engine():
    # BASELINE configs[2]: 10M x 64-bit codes, 10k queries, k = 100 -- the two engines are
    # independent implementations (POPC on CUDA cores vs E4M3 MMAs) and must agree bit for bit
    L = yael_b200.lib()
    r = rs(1236)
    nb, nq, k = 10_000_000, 10_000, 100
    b = r.randint(0, 2 ** 63, nb, dtype=np.int64).view(np.uint8).reshape(nb, 8)
    q = r.randint(0, 2 ** 63, nq, dtype=np.int64).view(np.uint8).reshape(nq, 8)
    b[::100003] = q[3]  # planted duplicates
    db, dq = DevArray(b), DevArray(q)
    out = []
    for eng in (0, 1):
        L.yb_set_hamming_engine(eng)
        oi = DevArray(shape=(nq, k), dtype=np.int32)
        od = DevArray(shape=(nq, k), dtype=np.uint16)
        rc = L.yb_nn_hamming(nq, nb, 8, k, db.ptr, dq.ptr, oi.ptr, od.ptr, 0, None)
        assert rc == 0, L.yb_last_error()
        L.yb_sync(None)
        assert L.yb_last_hamming_engine() == eng
        out.append((oi.get(), od.get(), L.yb_last_hamming_fallbacks()))
        oi.free()
        od.free()
    L.yb_set_hamming_engine(-1)
    db.free()
    dq.free()
    assert np.array_equal(out[0][1], out[1][1])
    assert np.array_equal(out[0][0], out[1][0])
    # sanity of the result itself: sorted by (distance, id), distances are the true popcounts
    gi, gd = out[1][0], out[1][1]
    assert (np.diff(gd.astype(np.int32), axis=1) >= 0).all()
    same = np.diff(gd.astype(np.int32), axis=1) == 0
    assert (np.diff(gi, axis=1)[same] > 0).all()
    for qq in (0, 3, 9999):
        true = np.unpackbits(b[gi[qq]] ^ q[qq], axis=1).sum(1)
        assert np.array_equal(true, gd[qq])
    assert gd[3, 0] == 0
    # on iid codes the certificate holds for (nearly) every query
    assert out[1][2] <= nq // 20, "fallbacks: %d" % out[1][2]


@pytest.mark.parametrize("nq,nb,nc,slots", [(128, 768, 8, 3), (130, 2000, 8, 3), (77, 1001, 4, 3),
                                            (64, 1000, 16, 2), (200, 5000, 8, 2), (50, 999, 5, 3)])
def test_packed_accumulators_are_exact(nq, nb, nc, slots):
    # `slots` consecutive database rows share one FP32 accumulator at scales 1, 2^8, 2^16: the
    # tensor core must deliver dot_0 + 2^8 dot_1 + 2^16 dot_2 EXACTLY (integers below 2^23)
    L = yael_b200.lib()
    r = rs(nq + nb + nc + slots)
    base = r.randint(0, 256, (nb, nc)).astype(np.uint8)
    query = r.randint(0, 256, (nq, nc)).astype(np.uint8)
    base[::7] = query[0]                     # ham 0 -> dot = +bits in every slot position
    base[3::11] = ~query[min(2, nq - 1)]     # ham = bits -> dot = -bits
    ncomb = (nb + slots - 1) // slots
    db, dq = DevArray(base), DevArray(query)
    out = DevArray(shape=(nq, ncomb), dtype=np.float32)
    rc = L.yb_debug_hamming_tc_packed(nq, nb, nc, slots, db.ptr, dq.ptr, out.ptr, None)
    assert rc == 0, L.yb_last_error()
    L.yb_sync(None)
    got = out.get().astype(np.float64)
    W = 1 if nc <= 8 else 2
    bits = 64 * W   # codes are zero-padded to W words: padded bits are equal on both sides
    ham = np.unpackbits(query[:, None, :] ^ base[None, :, :], axis=2).sum(2).astype(np.int64)
    dot = bits - 2 * ham                                   # [nq][nb]
    pad = ncomb * slots - nb
    dot = np.concatenate([dot, np.zeros((nq, pad), np.int64)], axis=1).reshape(nq, ncomb, slots)
    want = -2.0 * sum(dot[:, :, i] * float(256 ** i) for i in range(slots))
    assert np.array_equal(got, want)
    for a in (db, dq, out):
        a.free()


@pytest.mark.parametrize("slots", ["1", "2", "3"])
def test_tensor_engine_every_packing_is_bit_exact(yn, ob, tc_engine, monkeypatch, slots):
    # 64-bit codes can pack up to three database rows per accumulator (default from 8 M rows on);
    # YAEL_B200_HAM_SLOTS forces a packing.
    # nb not a multiple of the packing, k = 128 (the largest k with three slots), heavy ties.
    monkeypatch.setenv("YAEL_B200_HAM_SLOTS", slots)
    r = rs(int(slots) + 40)
    for nb, nq, nc, k in ((50001, 300, 8, 100), (40003, 64, 8, 128), (35000, 130, 4, 10), (70001, 50, 16, 40)):
        b = r.randint(0, 256, (nb, nc)).astype(np.uint8)
        q = r.randint(0, 256, (nq, nc)).astype(np.uint8)
        b[::17] = q[0]
        b[-1] = q[1]            # the very last row (a partially filled accumulator) is a hit
        b[3::5003] = ~q[2]      # distance = every bit
        idx, dis = yn.knn_hamming(q, b, k)
        assert tc_engine.yb_last_hamming_engine() == 1
        widx, wdis = ob.orc_nn_hamming(b, q, k)
        assert np.array_equal(dis, wdis)
        assert np.array_equal(idx, widx)

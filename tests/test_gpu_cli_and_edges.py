"""GPU: (1) the reference's own CLIs (progs/kmeans.c, progs/knn.c), compiled UNMODIFIED by
`make -C oracle progs` and linked against libyael_b200.so, run BASELINE configs[0] ("progs/kmeans
on synthetic fvecs: n=100k, d=128, k=256, niter=20, fixed seed") and a k-NN job from .fvecs files;
results are compared with the oracle.  (2) edge cases of the hot path: empty inputs, k = nb,
dimensions that are not a multiple of 4, k too large for the tensor engine, duplicates."""
import os
import subprocess

import numpy as np
import pytest

import yael_b200
from devmem import check_knn

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROGS = os.path.join(ROOT, "oracle", "_ref", "progs")


def write_fvecs(path, m):
    n, d = m.shape
    out = np.empty((n, d + 1), np.float32)
    out[:, 0] = np.array([d], np.int32).view(np.float32)[0]
    out[:, 1:] = m
    out.tofile(path)


def read_vecs(path, dtype):
    raw = np.fromfile(path, dtype=np.int32)
    d = int(raw[0])
    return raw.reshape(-1, d + 1)[:, 1:].view(dtype).copy()


needs_progs = pytest.mark.skipif(not os.path.exists(os.path.join(PROGS, "kmeans")),
                                 reason="reference CLIs not built (make -C oracle progs)")


@needs_progs
def test_reference_kmeans_cli_baseline_config0(tmp_path, ob):
    # BASELINE.json configs[0]
    v = np.random.RandomState(1234).random_sample((100000, 128)).astype(np.float32)
    write_fvecs(str(tmp_path / "v.fvecs"), v)
    env = dict(os.environ, YAEL_B200_EXACT_UPDATE="1")
    p = subprocess.run([os.path.join(PROGS, "kmeans"), "-i", str(tmp_path / "v.fvecs"), "-k", "256",
                        "-niter", "20", "-seed", "1234", "-nt", "8", "-o", str(tmp_path / "c.fvecs")],
                       capture_output=True, text=True, env=env, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    cent = read_vecs(str(tmp_path / "c.fvecs"), np.float32)
    assert cent.shape == (256, 128)
    # progs/kmeans.c:152: kmeans(d, n, k, niter, v, nt, seed, nredo, centroids, NULL, NULL, nassign)
    qerr, wcent, _, _, wn = ob.orc_kmeans(v, 256, 20, 8 | ob.KMEANS_QUIET, 1234)
    # the assignment is exact and the update runs in the reference's order: same trajectory
    np.testing.assert_allclose(cent, wcent, rtol=0, atol=1e-6)
    assert " -> " in p.stdout and "Total number of iterations: 20" in p.stdout


@needs_progs
def test_reference_knn_cli(tmp_path, ob):
    r = np.random.RandomState(7)
    b = r.random_sample((20000, 64)).astype(np.float32)
    q = r.random_sample((300, 64)).astype(np.float32)
    write_fvecs(str(tmp_path / "b.fvecs"), b)
    write_fvecs(str(tmp_path / "q.fvecs"), q)
    p = subprocess.run([os.path.join(PROGS, "knn"), "-b", str(tmp_path / "b.fvecs"), "-q",
                        str(tmp_path / "q.fvecs"), "-k", "10", "-onn", str(tmp_path / "nn.ivecs"),
                        "-odis", str(tmp_path / "dis.fvecs"), "-silent"],
                       capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    idx = read_vecs(str(tmp_path / "nn.ivecs"), np.int32)
    dis = read_vecs(str(tmp_path / "dis.fvecs"), np.float32)
    widx, wdis = ob.orc_knn(b, q, 10, canonical=True)
    assert np.array_equal(idx, widx)
    # progs/knn.c:226 re-orders with compute_distances_1 (both norms in double): same order here
    np.testing.assert_allclose(dis, wdis, rtol=1e-5)


def test_empty_and_degenerate_inputs(yn, ob):
    b = np.random.RandomState(1).random_sample((50, 8)).astype(np.float32)
    idx, dis = yn.knn(np.zeros((0, 8), np.float32), b, 3)
    assert idx.shape == (0, 3) and dis.shape == (0, 3)
    # k == nb: every row is returned, sorted
    q = b[:5].copy()
    idx, dis = yn.knn(q, b, 50)
    assert all(sorted(row) == list(range(50)) for row in idx.tolist())
    assert (np.diff(dis, axis=1) >= 0).all() and (idx[:, 0] == np.arange(5)).all()
    # one base row, one query, d = 1
    i1, d1 = yn.knn(np.array([[2.0]], np.float32), np.array([[5.0]], np.float32), 1)
    assert i1.tolist() == [[0]] and d1.tolist() == [[9.0]]
    # all rows identical: ties resolved by id
    same = np.ones((40, 4), np.float32)
    idx, dis = yn.knn(np.zeros((3, 4), np.float32), same, 7)
    assert (idx == np.arange(7)).all() and (dis == 4.0).all()


@pytest.mark.parametrize("d", [1, 3, 5, 30, 100, 127])
def test_tensor_engine_any_dimension(yn, ob, d):
    # the tensor pass runs on pitch-padded copies, so d need not be a multiple of 4
    L = yael_b200.lib()
    r = np.random.RandomState(d)
    b = r.random_sample((6000, d)).astype(np.float32)
    q = r.random_sample((200, d)).astype(np.float32)
    L.yb_set_knn_engine(1)
    try:
        idx, dis = yn.knn(q, b, 10)
        assert L.yb_last_knn_engine() == 1
    finally:
        L.yb_set_knn_engine(-1)
    widx, wdis = ob.orc_knn(b, q, 10, canonical=True)
    check_knn(idx, dis, widx, wdis, b, q)


def test_large_k_and_high_dimension_use_exact_engine(yn, ob):
    L = yael_b200.lib()
    r = np.random.RandomState(2)
    b = r.random_sample((3000, 200)).astype(np.float32)   # d > 128
    q = r.random_sample((30, 200)).astype(np.float32)
    idx, dis = yn.knn(q, b, 5)
    assert L.yb_last_knn_engine() == 0
    widx, wdis = ob.orc_knn(b, q, 5, canonical=True)
    assert np.array_equal(idx, widx) and np.array_equal(dis, wdis)
    b2 = r.random_sample((20000, 32)).astype(np.float32)
    q2 = r.random_sample((100, 32)).astype(np.float32)
    idx, dis = yn.knn(q2, b2, 1000)                          # k' would not fit the tensor engine
    assert L.yb_last_knn_engine() == 0
    widx, wdis = ob.orc_knn(b2, q2, 1000, canonical=True)
    assert np.array_equal(idx, widx) and np.array_equal(dis, wdis)


def test_kmeans_sift_like_integer_data(yn, ob, monkeypatch):
    # integer coordinates: arithmetic is exact everywhere, so the whole run must match the oracle
    monkeypatch.setenv("YAEL_B200_EXACT_UPDATE", "1")
    r = np.random.RandomState(3)
    v = np.minimum(255, r.gamma(1.2, 25.0, (30000, 64))).astype(np.int32).astype(np.float32)
    cent, qerr, dis, assign, nassign = yn.kmeans(v, 128, niter=8, seed=99, verbose=False, output="all", nt=4)
    w = ob.orc_kmeans(v, 128, 8, ob.KMEANS_QUIET | 4, 99)
    assert np.array_equal(nassign, w[4]) and np.array_equal(assign, w[3])
    np.testing.assert_allclose(cent, w[1], rtol=0, atol=1e-4)


def test_hamming_edge_cases(yn, ob):
    r = np.random.RandomState(4)
    b = r.randint(0, 256, (130, 8)).astype(np.uint8)
    q = r.randint(0, 256, (3, 8)).astype(np.uint8)
    idx, dis = yn.knn_hamming(q, b, 130)       # k == nb
    widx, wdis = ob.orc_nn_hamming(b, q, 130)
    assert np.array_equal(idx, widx) and np.array_equal(dis, wdis)
    z = np.zeros((500, 8), np.uint8)            # everything ties at distance 0
    idx, dis = yn.knn_hamming(z[:2], z, 20)
    assert (idx == np.arange(20)).all() and (dis == 0).all()
    pairs, scores = yn.match_hamming(q, b, -1)   # nothing matches
    assert len(scores) == 0

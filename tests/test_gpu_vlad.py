"""GPU parity of the consumers of the k = 1 search: vlad_compute[_weighted /
_subsets] and bof_compute[_ma / _subsets] (yael/vlad.c:10-139) through the drop-in C API, against
the compiled-reference golden and the oracle.  The assignment is the library's k-NN (exact distances,
(distance, id) order), the aggregation follows the reference's summation order: bit-identical."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _subsets(g):
    ends, idx = g["subset_ends"], g["subset_indexes"]
    return [idx[(ends[i - 1] if i else 0):ends[i]].tolist() for i in range(len(ends))]


def test_vlad_bof_match_reference_golden(yn):
    g = np.load(os.path.join(GOLD, "vlad_bof.npz"))
    c, v, subs = g["centroids"], g["v"], _subsets(g)
    assert np.array_equal(yn.vlad(c, v), g["vlad"])
    assert np.array_equal(yn.vlad(c, v, weights=g["weights"]), g["vlad_weighted"])
    assert np.array_equal(yn.vlad(c, v, subsets=subs), g["vlad_subsets"])
    assert np.array_equal(yn.bof(c, v), g["bof"])
    assert np.array_equal(yn.bof(c, v, ma=3), g["bof_ma3"])
    assert np.array_equal(yn.bof(c, v, subsets=subs), g["bof_subsets"])


@pytest.mark.parametrize("n,k,d", [(20000, 256, 128), (5000, 100, 30), (70000, 1024, 64), (300, 7, 5)])
def test_vlad_bof_match_oracle(yn, ob, n, k, d):
    # SIFT-like sizes: the k = 1 search runs on the tensor engine for the larger shapes
    r = np.random.RandomState(n + k + d)
    c = r.rand(k, d).astype(np.float32)
    v = r.rand(n, d).astype(np.float32)
    w = (0.5 + r.rand(n)).astype(np.float32)
    subs = [r.permutation(n)[: n // 3].tolist(), list(range(n // 2, n)), []]
    assert np.array_equal(yn.vlad(c, v), ob.orc_vlad(c, v))
    assert np.array_equal(yn.vlad(c, v, weights=w), ob.orc_vlad(c, v, weights=w))
    assert np.array_equal(yn.vlad(c, v, subsets=subs), ob.orc_vlad(c, v, subsets=subs))
    assert np.array_equal(yn.bof(c, v), ob.orc_bof(c, v))
    assert np.array_equal(yn.bof(c, v, ma=4), ob.orc_bof(c, v, ma=4))
    assert np.array_equal(yn.bof(c, v, subsets=subs), ob.orc_bof(c, v, subsets=subs))


def test_vlad_of_kmeans_centroids_is_small_but_exact(yn, ob):
    # residual sums against k-means centroids nearly cancel: the summation ORDER decides the bits
    r = np.random.RandomState(3)
    v = r.rand(30000, 32).astype(np.float32)
    c = yn.kmeans(v, 64, niter=8, verbose=False, seed=5)
    got, want = yn.vlad(c, v), ob.orc_vlad(c, v)
    assert np.array_equal(got, want)
    assert np.abs(want).max() < 0.1 * len(v) / 64   # far below the plain per-centroid sums (~235)

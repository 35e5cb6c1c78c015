"""Generate tests/golden/*.npz from the UNMODIFIED reference compiled by oracle/Makefile
(oracle/_ref/libyael_ref.so: /root/reference sources + SciPy's OpenBLAS 0.3.x, gcc -O3 -msse4).

Run in the build container (needs /root/reference); the fixtures are committed so that the GPU
box and later rounds never need the reference tree.  Inputs are regenerated in the tests from the
same seeds where that is cheap; otherwise they are stored."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import bindings as ob  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)
assert ob.have_ref(), "build oracle/_ref first (make -C oracle)"
L = ob.ref()


ONLY = set(sys.argv[1:])  # fixture names to (re)write; none given = all


def save(name, **kw):
    if ONLY and name not in ONLY:
        return
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **kw)
    print(name, {k: getattr(v, "shape", v) for k, v in kw.items()})


# 1. the only deterministic k-NN input of the reference's own tests (test/py/test_ynumpy.py:13-26)
base = np.array([range(i, i + 4) for i in range(5)], dtype=np.float32)
quer = np.array([[x + 0.25 for x in range(i, i + 4)] for i in range(3)], dtype=np.float32)
idx, dis = ob.ref_knn(base, quer, 2, nt=1)
save("ynumpy_knn", base=base, query=quer, idx=idx, dis=dis)

# 2. k-NN, seeded uniform data (twin of BASELINE config 2), k in {1, 10, 100}
r = np.random.RandomState(1234)
b = r.random_sample((4000, 128)).astype(np.float32)
q = r.random_sample((64, 128)).astype(np.float32)
out = {}
for k in (1, 10, 100):
    i_, d_ = ob.ref_knn(b, q, k, nt=4)
    out["idx%d" % k], out["dis%d" % k] = i_, d_
save("knn_uniform_seed1234", **out)

# 3. tie / NaN semantics
bt = np.zeros((6, 1), np.float32)
bt[:, 0] = [1, 1, 1, 1, .5, 1]
qt = np.zeros((1, 1), np.float32)
it, dt = ob.ref_knn(bt, qt, 3, nt=1)
bn = np.array([[0.], [np.nan], [2.]], np.float32)
inn, dnn = ob.ref_knn(bn, qt, 3, nt=1)
i1, d1 = ob.ref_knn(np.ones((5, 2), np.float32), np.zeros((2, 2), np.float32), 1, nt=1)
save("knn_semantics", tie_idx=it, tie_dis=dt, nan_idx=inn, nan_dis=dnn, k1_tie_idx=i1, k1_tie_dis=d1)

# 4. cross distances incl. non-packed leading dimensions
a = r.random_sample((37, 24)).astype(np.float32)
bb = r.random_sample((29, 24)).astype(np.float32)
save("cross_distances", a=a, b=bb, dist=ob.ref_cross(a, bb))

# 5. k-means: config-1 twin (n=20000, d=32, k=64), random / k-means++ / normalised, and the
#    empty-cluster split on duplicated points
v = np.random.RandomState(1234).random_sample((20000, 32)).astype(np.float32)
out = {}
for name, flags in (("random", ob.KMEANS_QUIET | 4), ("pp", ob.KMEANS_QUIET | ob.KMEANS_INIT_BERKELEY | 4)):
    qe, cent, dis_, assign, nassign = ob.ref_kmeans(v, 64, 15, flags, 777, redo=2)
    out[name + "_qerr"], out[name + "_cent"], out[name + "_nassign"] = np.float32(qe), cent, nassign
    out[name + "_assign"] = assign
save("kmeans_seed777", **out)
v2 = np.repeat(np.random.RandomState(5).random_sample((30, 8)).astype(np.float32), 50, axis=0)
qe, cent, dis_, assign, nassign = ob.ref_kmeans(v2, 40, 10, ob.KMEANS_QUIET | 1, 5)
save("kmeans_empty_split", v=v2, qerr=np.float32(qe), cent=cent, nassign=nassign)

# 6. RNG sequences the k-means driver depends on (glibc rand_r)
perm = L.ivec_new_random_perm_r(1000, 4242)
perm = np.ctypeslib.as_array(perm, shape=(1000,)).copy()
g = np.empty(257, np.float32)
L.fvec_randn_r(ob.fp(g), 257, 99)
save("rng", perm_n1000_seed4242=perm, randn_n257_seed99=g)

# 7. k-min: heap regime (n > 20k), quickselect regime, argmin, ties
val = np.random.RandomState(3).random_sample(50000).astype(np.float32)
vt = np.random.RandomState(4).randint(0, 50, 100000).astype(np.float32)
save("kmin", val=val, k7=ob.ref_k_min(val, 7), k100=ob.ref_k_min(val, 100), k1=ob.ref_k_min(val, 1),
     small100=ob.ref_k_min(val[:1000], 100), ties=vt[:0], ties_k1=ob.ref_k_min(vt, 1))

# 8. Hamming: full matrix for 4 / 8 / 16 / 24 / 5 byte codes; threshold matches
hout = {}
rh = np.random.RandomState(6)
for nc in (4, 8, 16, 24, 5):
    ha = rh.randint(0, 256, (23, nc)).astype(np.uint8)
    hb = rh.randint(0, 256, (31, nc)).astype(np.uint8)
    hout["a%d" % nc], hout["b%d" % nc], hout["dis%d" % nc] = ha, hb, ob.ref_compute_hamming(ha, hb)
ha, hb = hout["a8"], hout["b8"]
n = C.c_size_t(0)
L.match_hamming_count(ob.u8p(ha), ob.u8p(hb), 23, 31, 28, 8, C.byref(n))
midx = np.empty((n.value, 2), np.int32)
mham = np.empty(n.value, np.uint16)
L.match_hamming_thres_prealloc(ob.u8p(ha), ob.u8p(hb), 23, 31, 28, 8, ob.ip(midx), ob.u16p(mham))
hout["match_ht28_idx"], hout["match_ht28_ham"] = midx, mham
save("hamming", **hout)

# 9. Hamming cross-matching inside one set (crossmatch_hamming_count / _prealloc,
#    yael/hamming.c:368-395, 793-829) for 4 / 8 / 16 / 5 byte codes, planted near-duplicates
cout = {}
rc = np.random.RandomState(8)
for nc, ht in ((4, 10), (8, 24), (16, 52), (5, 14)):
    db = rc.randint(0, 256, (300, nc)).astype(np.uint8)
    db[::37] = db[3]            # exact duplicates (distance 0)
    db[5::41, 0] ^= 0x11        # and near-duplicates of whatever sits there
    n = C.c_size_t(0)
    L.crossmatch_hamming_count(ob.u8p(db), 300, ht, nc, C.byref(n))
    cidx = np.empty((n.value, 2), np.int32)
    cham = np.empty(n.value, np.uint16)
    m = L.crossmatch_hamming_prealloc(ob.u8p(db), 300, ht, nc, ob.ip(cidx), ob.u16p(cham))
    assert m == n.value
    cout["db%d" % nc], cout["ht%d" % nc] = db, np.int32(ht)
    cout["idx%d" % nc], cout["ham%d" % nc] = cidx, cham
save("hamming_crossmatch", **cout)

# 10. knn_full with the other distance types and per-base weights (yael/nn.c:280-350, 497-500),
#     k = 1 (nn_single_full, yael/nn.c:383-446) and k = 5; positive data so chi2 is well defined
ra = np.random.RandomState(10)
ab = (ra.random_sample((300, 12)) + 0.05).astype(np.float32)
aq = (ra.random_sample((40, 12)) + 0.05).astype(np.float32)
aw = (0.5 + 1.5 * ra.random_sample(300)).astype(np.float32)
aout = {"base": ab, "query": aq, "weights": aw}
for t in (1, 2, 3, 4, 5, 6, 16):
    for k in (1, 5):
        for wname, w in (("", None), ("w", aw)):
            i_ = np.empty((40, k), np.int32)
            d_ = np.empty((40, k), np.float32)
            L.knn_full_thread(t, 40, 300, 12, k, ob.fp(ab), ob.fp(aq), ob.fp(w) if w is not None else None,
                              ob.ip(i_), ob.fp(d_), 2)
            aout["idx_t%d_k%d%s" % (t, k, wname)], aout["dis_t%d_k%d%s" % (t, k, wname)] = i_, d_
save("knn_alt_weighted", **aout)

# 11. consumers of the k = 1 search (yael/vlad.c:10-139): VLAD (plain, weighted, subsets) and bag of
#     features (plain, multiple assignment, subsets); SIFT-like codebook size, ragged subsets
rv = np.random.RandomState(11)
vc = rv.random_sample((64, 32)).astype(np.float32)
vv = rv.random_sample((3000, 32)).astype(np.float32)
vw = (0.25 + rv.random_sample(3000)).astype(np.float32)
subs = [rv.permutation(3000)[:700].tolist(), list(range(1000, 1900)), [], [5, 5, 5, 2999, 0]]
sidx, sends = ob._subsets(subs)
save("vlad_bof", centroids=vc, v=vv, weights=vw, subset_indexes=sidx, subset_ends=sends,
     vlad=ob.ref_vlad(vc, vv), vlad_weighted=ob.ref_vlad(vc, vv, weights=vw),
     vlad_subsets=ob.ref_vlad(vc, vv, subsets=subs), bof=ob.ref_bof(vc, vv), bof_ma3=ob.ref_bof(vc, vv, ma=3),
     bof_subsets=ob.ref_bof(vc, vv, subsets=subs))

# 12. further consumers: hkm_quantize (yael/hkm.c:144-162) over a tree learned
#     by the reference's hkm_learn, and the GMM E-step gmm_compute_p (yael/gmm.c:305-367) with and
#     without the mixture weights
rh = np.random.RandomState(12)
hv = rh.random_sample((4000, 16)).astype(np.float32)
hlevels, hassign = ob.ref_hkm_learn(hv, 3, 5, niter=8)
hq = rh.random_sample((1500, 16)).astype(np.float32)
gk, gd = 32, 24
gmu = rh.random_sample((gk, gd)).astype(np.float32)
gsigma = (0.02 + 0.1 * rh.random_sample((gk, gd))).astype(np.float32)
gw = rh.random_sample(gk).astype(np.float32)
gw /= gw.sum()
gv = rh.random_sample((600, gd)).astype(np.float32)
save("hkm_gmm", hkm_level0=hlevels[0], hkm_level1=hlevels[1], hkm_level2=hlevels[2], hkm_learn_assign=hassign,
     hkm_points=hv, hkm_query=hq, hkm_quantize_query=ob.ref_hkm_quantize(hlevels, 5, hq),
     hkm_quantize_points=ob.ref_hkm_quantize(hlevels, 5, hv),
     gmm_w=gw, gmm_mu=gmu, gmm_sigma=gsigma, gmm_v=gv,
     gmm_p_w=ob.ref_gmm_compute_p(gw, gmu, gsigma, gv, 1), gmm_p_now=ob.ref_gmm_compute_p(gw, gmu, gsigma, gv, 0),
     gmm_p_w_nt3=ob.ref_gmm_compute_p(gw, gmu, gsigma, gv, 1, nt=3))

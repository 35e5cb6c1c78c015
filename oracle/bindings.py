"""ctypes bindings for the CHECKERS: oracle/liboracle.so (our C restatement) and
oracle/_ref/libyael_ref.so (the unmodified reference compiled by oracle/Makefile).

TEST INFRASTRUCTURE ONLY.  Imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs; the product package (yael_b200) never imports it.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "liboracle.so")
REF_SO = os.path.join(HERE, "_ref", "libyael_ref.so")

DOT_F32_SEQ, DOT_F64 = 0, 1

KMEANS_QUIET = 0x10000
KMEANS_INIT_BERKELEY = 0x20000
KMEANS_NORMALIZE_CENTS = 0x40000
KMEANS_INIT_RANDOM = 0x80000
KMEANS_INIT_USER = 0x100000

_f = C.POINTER(C.c_float)
_i = C.POINTER(C.c_int)
_u8 = C.POINTER(C.c_uint8)
_u16 = C.POINTER(C.c_uint16)


def build(quiet=True):
    """Compile liboracle.so (and _ref when the reference tree is present)."""
    out = subprocess.run(["make", "-C", HERE], capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + out.stdout + out.stderr)
    if not quiet:
        print(out.stdout)
    # the reference's own CLIs (progs/knn.c, progs/kmeans.c), unmodified, linked against the
    # PRODUCT library: the drop-in check of tests/test_gpu_cli_and_edges.py.  Needs the reference
    # tree (this container) and the built product; the binaries travel to the GPU box.
    ref_root = os.environ.get("YAEL_REF", "/root/reference")
    product = os.path.join(os.path.dirname(HERE), "yael_b200", "libyael_b200.so")
    if os.path.exists(os.path.join(ref_root, "progs", "knn.c")) and os.path.exists(product):
        out = subprocess.run(["make", "-C", HERE, "progs"], capture_output=True, text=True)
        if out.returncode != 0:
            raise RuntimeError("oracle progs build failed:\n" + out.stdout + out.stderr)


def fp(a):
    return a.ctypes.data_as(_f) if a is not None else None


def ip(a):
    return a.ctypes.data_as(_i) if a is not None else None


def u8p(a):
    return a.ctypes.data_as(_u8) if a is not None else None


def u16p(a):
    return a.ctypes.data_as(_u16) if a is not None else None


def f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class HkmT(C.Structure):
    """hkm_t (yael/hkm.h:11-17)"""
    _fields_ = [("nlevel", C.c_int), ("bf", C.c_int), ("k", C.c_int), ("d", C.c_int),
                ("centroids", C.POINTER(_f))]


class GmmT(C.Structure):
    """gmm_t (yael/gmm.h:20-26)"""
    _fields_ = [("d", C.c_int), ("k", C.c_int), ("w", _f), ("mu", _f), ("sigma", _f)]


_oracle = None
_ref = None


def oracle():
    global _oracle
    if _oracle is None:
        if not os.path.exists(ORACLE_SO):
            build()
        L = C.CDLL(ORACLE_SO)
        L.orc_cross_distances.argtypes = [C.c_int] * 3 + [_f, _f, _f, C.c_int]
        L.orc_cross_distances_nonpacked.argtypes = [C.c_int] * 3 + [_f, C.c_int, _f, C.c_int, _f, C.c_int, C.c_int]
        L.orc_distances_1.argtypes = [C.c_int, C.c_int, _f, _f, C.c_int, _f, C.c_int]
        L.orc_knn_full.argtypes = [C.c_int] * 4 + [_f, _f, _f, _i, _f, C.c_int, C.c_int]
        L.orc_knn_canonical.argtypes = [C.c_int] * 4 + [_f, _f, _i, _f, C.c_int, C.c_int]
        L.orc_knn_reorder_shortlist.argtypes = [C.c_int] * 4 + [_f, _f, _i, _f, C.c_int]
        L.orc_fvec_k_min.argtypes = [_f, C.c_int, _i, C.c_int]
        L.orc_fvecs_k_min.argtypes = [_f, C.c_long, C.c_long, _i, C.c_int]
        L.orc_fvec_k_min_canonical.argtypes = [_f, C.c_int, _i, C.c_int]
        L.orc_kmeans.argtypes = [C.c_int] * 4 + [_f, C.c_int, C.c_long, C.c_int, _f, _f, _i, _i, C.c_int]
        L.orc_kmeans.restype = C.c_float
        L.orc_kmeans_step.argtypes = [C.c_int] * 3 + [_f, _f, _f, _i, _f, _i, C.c_int, C.c_int]
        L.orc_kmeans_step.restype = C.c_double
        L.orc_kmeans_reassign_empty.argtypes = [C.c_int] * 3 + [_f, _i, _i, C.c_uint]
        L.orc_kmeans_reassign_empty.restype = C.c_int
        L.orc_random_perm_r.argtypes = [C.c_int, C.c_uint]
        L.orc_random_perm_r.restype = C.POINTER(C.c_int)
        L.orc_fvec_randn_r.argtypes = [_f, C.c_long, C.c_uint]
        L.orc_hamming.argtypes = [_u8, _u8, C.c_int]
        L.orc_hamming.restype = C.c_uint16
        L.orc_compute_hamming.argtypes = [_u16, _u8, _u8, C.c_int, C.c_int, C.c_int]
        L.orc_nn_hamming.argtypes = [C.c_int] * 4 + [_u8, _u8, _i, _u16, C.c_int]
        L.orc_match_hamming_count.argtypes = [_u8, _u8, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_size_t)]
        L.orc_match_hamming_thres_prealloc.argtypes = [_u8, _u8, C.c_int, C.c_int, C.c_int, C.c_int, _i, _u16]
        L.orc_match_hamming_thres_prealloc.restype = C.c_size_t
        L.orc_crossmatch_hamming_count.argtypes = [_u8, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_size_t)]
        L.orc_crossmatch_hamming_prealloc.argtypes = [_u8, C.c_long, C.c_int, C.c_int, _i, _u16]
        L.orc_crossmatch_hamming_prealloc.restype = C.c_size_t
        L.orc_hkm_quantize.argtypes = [C.c_int] * 3 + [_f, C.c_int, _f, _i, C.c_int]
        L.orc_gmm_compute_p.argtypes = [C.c_int] * 3 + [_f, _f, _f, _f, _f, C.c_int, C.c_int]
        _oracle = L
    return _oracle


def have_ref():
    return os.path.exists(REF_SO)


def ref():
    """The compiled, unmodified reference (prototypes: yael/nn.h:41-214, kmeans.h:41-44,
    sorting.h:19-40, hamming.h:24-50, vector.h)."""
    global _ref
    if _ref is None:
        os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")  # yael threads itself (README:161-162)
        L = C.CDLL(REF_SO)
        L.knn_full.argtypes = [C.c_int] * 5 + [_f, _f, _f, _i, _f]
        L.knn_full_thread.argtypes = [C.c_int] * 5 + [_f, _f, _f, _i, _f, C.c_int]
        L.knn_reorder_shortlist.argtypes = [C.c_int] * 4 + [_f, _f, _i, _f]
        L.compute_cross_distances.argtypes = [C.c_int] * 3 + [_f, _f, _f]
        L.compute_cross_distances_nonpacked.argtypes = [C.c_int] * 3 + [_f, C.c_int, _f, C.c_int, _f, C.c_int]
        L.compute_distances_1.argtypes = [C.c_int, C.c_int, _f, _f, _f]
        L.kmeans.argtypes = [C.c_int] * 4 + [_f, C.c_int, C.c_long, C.c_int, _f, _f, _i, _i]
        L.kmeans.restype = C.c_float
        L.fvec_k_min.argtypes = [_f, C.c_int, _i, C.c_int]
        L.fvecs_k_min.argtypes = [_f, C.c_long, C.c_long, _i, C.c_int]
        L.compute_hamming.argtypes = [_u16, _u8, _u8, C.c_int, C.c_int, C.c_int]
        L.match_hamming_count.argtypes = [_u8, _u8, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_size_t)]
        L.match_hamming_thres_prealloc.argtypes = [_u8, _u8, C.c_int, C.c_int, C.c_int, C.c_int, _i, _u16]
        L.match_hamming_thres_prealloc.restype = C.c_size_t
        L.crossmatch_hamming_count.argtypes = [_u8, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_size_t)]
        L.crossmatch_hamming_prealloc.argtypes = [_u8, C.c_long, C.c_int, C.c_int, _i, _u16]
        L.crossmatch_hamming_prealloc.restype = C.c_size_t
        L.ivec_new_random_perm_r.argtypes = [C.c_int, C.c_uint]
        L.ivec_new_random_perm_r.restype = C.POINTER(C.c_int)
        L.fvec_randn_r.argtypes = [_f, C.c_long, C.c_uint]
        L.count_cpu.restype = C.c_int
        L.hkm_quantize.argtypes = [C.POINTER(HkmT), C.c_int, _f, _i]
        L.hkm_learn.argtypes = [C.c_int] * 4 + [_f, C.c_int, C.c_int, C.c_int, C.POINTER(_i)]
        L.hkm_learn.restype = C.POINTER(HkmT)
        L.hkm_delete.argtypes = [C.POINTER(HkmT)]
        L.gmm_compute_p.argtypes = [C.c_int, _f, C.POINTER(GmmT), _f, C.c_int]
        L.gmm_compute_p_thread.argtypes = [C.c_int, _f, C.POINTER(GmmT), _f, C.c_int, C.c_int]
        _ref = L
    return _ref


# ---------------------------------------------------------------- numpy conveniences


def orc_knn(base, query, k, dot_mode=DOT_F32_SEQ, nt=8, canonical=False, weights=None):
    base, query = f32(base), f32(query)
    nq, d = query.shape
    nb = base.shape[0]
    idx = np.empty((nq, k), np.int32)
    dis = np.empty((nq, k), np.float32)
    if canonical:
        oracle().orc_knn_canonical(nq, nb, d, k, fp(base), fp(query), ip(idx), fp(dis), dot_mode, nt)
    else:
        w = f32(weights) if weights is not None else None
        oracle().orc_knn_full(nq, nb, d, k, fp(base), fp(query), fp(w), ip(idx), fp(dis), dot_mode, nt)
    return idx, dis


def ref_knn(base, query, k, nt=8, weights=None):
    base, query = f32(base), f32(query)
    nq, d = query.shape
    nb = base.shape[0]
    idx = np.empty((nq, k), np.int32)
    dis = np.empty((nq, k), np.float32)
    w = f32(weights) if weights is not None else None
    ref().knn_full_thread(2, nq, nb, d, k, fp(base), fp(query), fp(w), ip(idx), fp(dis), nt)
    return idx, dis


def orc_cross(a, b, dot_mode=DOT_F32_SEQ):
    a, b = f32(a), f32(b)
    out = np.empty((b.shape[0], a.shape[0]), np.float32)
    oracle().orc_cross_distances(a.shape[1], a.shape[0], b.shape[0], fp(a), fp(b), fp(out), dot_mode)
    return out


def ref_cross(a, b):
    a, b = f32(a), f32(b)
    out = np.empty((b.shape[0], a.shape[0]), np.float32)
    ref().compute_cross_distances(a.shape[1], a.shape[0], b.shape[0], fp(a), fp(b), fp(out))
    return out


def _kmeans_call(fn, v, k, niter, flags, seed, redo, init, extra):
    v = f32(v)
    n, d = v.shape
    cent = np.zeros((k, d), np.float32)
    if init is not None:
        cent[:] = init
    dis = np.empty(n, np.float32)
    assign = np.empty(n, np.int32)
    nassign = np.empty(k, np.int32)
    q = fn(d, n, k, niter, fp(v), flags, seed, redo, fp(cent), fp(dis), ip(assign), ip(nassign), *extra)
    return q, cent, dis, assign, nassign


def orc_kmeans(v, k, niter, flags, seed, redo=1, init=None, dot_mode=DOT_F32_SEQ):
    return _kmeans_call(oracle().orc_kmeans, v, k, niter, flags, seed, redo, init, (dot_mode,))


def ref_kmeans(v, k, niter, flags, seed, redo=1, init=None):
    return _kmeans_call(ref().kmeans, v, k, niter, flags, seed, redo, init, ())


def orc_kmeans_step(v, cent, dot_mode=DOT_F32_SEQ, nt=8):
    v, cent = f32(v), f32(cent)
    n, d = v.shape
    k = cent.shape[0]
    out = np.empty_like(cent)
    assign = np.empty(n, np.int32)
    dis = np.empty(n, np.float32)
    nassign = np.empty(k, np.int32)
    q = oracle().orc_kmeans_step(d, n, k, fp(v), fp(cent), fp(out), ip(assign), fp(dis), ip(nassign), dot_mode, nt)
    return q, out, assign, dis, nassign


def orc_nn_hamming(base, query, k, nt=8):
    base = np.ascontiguousarray(base, np.uint8)
    query = np.ascontiguousarray(query, np.uint8)
    nq, nc = query.shape
    idx = np.empty((nq, k), np.int32)
    dis = np.empty((nq, k), np.uint16)
    oracle().orc_nn_hamming(nq, base.shape[0], nc, k, u8p(base), u8p(query), ip(idx), u16p(dis), nt)
    return idx, dis


def orc_compute_hamming(a, b):
    a = np.ascontiguousarray(a, np.uint8)
    b = np.ascontiguousarray(b, np.uint8)
    out = np.empty((b.shape[0], a.shape[0]), np.uint16)
    oracle().orc_compute_hamming(u16p(out), u8p(a), u8p(b), a.shape[0], b.shape[0], a.shape[1])
    return out


def ref_compute_hamming(a, b):
    a = np.ascontiguousarray(a, np.uint8)
    b = np.ascontiguousarray(b, np.uint8)
    out = np.empty((b.shape[0], a.shape[0]), np.uint16)
    ref().compute_hamming(u16p(out), u8p(a), u8p(b), a.shape[0], b.shape[0], a.shape[1])
    return out


def orc_k_min(val, k, canonical=False):
    val = f32(val)
    idx = np.empty(k, np.int32)
    fn = oracle().orc_fvec_k_min_canonical if canonical else oracle().orc_fvec_k_min
    fn(fp(val), val.shape[0], ip(idx), k)
    return idx


def ref_k_min(val, k):
    val = f32(val)
    idx = np.empty(k, np.int32)
    ref().fvec_k_min(fp(val), val.shape[0], ip(idx), k)
    return idx


def ref_nn_hamming_blocked(base, query, k, block=1 << 20, threads=None):
    """Hamming k-NN the way SURVEY.md 8(c)-4 defines the oracle for the NEW nn_hamming: the
    reference's compute_hamming (yael/hamming.c:177-219; the compiled reference when present, else
    the restatement) over database blocks -- one block per host thread at a time -- followed by a
    stable (distance, id) selection.  Returns (idx[nq][k], dis[nq][k], kind)."""
    from concurrent.futures import ThreadPoolExecutor
    base = np.ascontiguousarray(base, np.uint8)
    query = np.ascontiguousarray(query, np.uint8)
    nq, nc = query.shape
    nb = base.shape[0]
    assert k <= nb
    use_ref = have_ref()
    fn = ref().compute_hamming if use_ref else oracle().orc_compute_hamming
    threads = threads or len(os.sched_getaffinity(0))
    starts = list(range(0, nb, block))
    # upper bound on every query's k-th distance from the first block (exact selection there)
    b0 = base[:min(nb, max(block, k))]
    d0 = np.empty((b0.shape[0], nq), np.uint16)
    fn(u16p(d0), u8p(query), u8p(b0), nq, b0.shape[0], nc)
    bound = np.partition(d0, k - 1, axis=0)[k - 1]          # [nq]

    def scan(s):
        blk = base[s:s + block]
        d = np.empty((blk.shape[0], nq), np.uint16)
        fn(u16p(d), u8p(query), u8p(blk), nq, blk.shape[0], nc)   # dis[j * nq + i] = ham(q_i, b_j)
        j, i = np.nonzero(d <= bound[None, :])
        return i.astype(np.int32), (j + s).astype(np.int64), d[j, i]

    with ThreadPoolExecutor(threads) as ex:
        parts = list(ex.map(scan, starts))
    qi = np.concatenate([p[0] for p in parts])
    bj = np.concatenate([p[1] for p in parts])
    dv = np.concatenate([p[2] for p in parts])
    order = np.lexsort((bj, dv, qi))                       # by query, then (distance, id)
    qi, bj, dv = qi[order], bj[order], dv[order]
    first = np.searchsorted(qi, np.arange(nq))
    idx = np.empty((nq, k), np.int32)
    dis = np.empty((nq, k), np.uint16)
    for q in range(nq):
        idx[q] = bj[first[q]:first[q] + k]
        dis[q] = dv[first[q]:first[q] + k]
    return idx, dis, ("reference" if use_ref else "port")


# ---- consumers of the k = 1 search: VLAD / bag of features (yael/vlad.c)
def _subsets(subsets):
    idx = np.ascontiguousarray(np.concatenate([np.asarray(s, np.int32) for s in subsets]) if subsets
                               else np.zeros(0, np.int32), np.int32)
    ends = np.ascontiguousarray(np.cumsum([len(s) for s in subsets]), np.int32)
    return idx, ends


def _vlad_family(L, prefix, dot, cent, v, weights=None, subsets=None, ma=None, bof=False):
    """One calling convention for the oracle (orc_*, extra dot_mode argument) and the compiled
    reference (vlad_* / bof_*)."""
    cent, v = f32(cent), f32(v)
    k, d = cent.shape
    n = v.shape[0]
    extra = [] if dot is None else [dot]
    if subsets is not None:
        idx, ends = _subsets(subsets)
        out = np.zeros((len(subsets), k) if bof else (len(subsets), k, d), np.float32)
        fn = getattr(L, prefix + ("bof_compute_subsets" if bof else "vlad_compute_subsets"))
        fn.argtypes = [C.c_int, C.c_int, _f, C.c_int, _f, C.c_int, _i, _i, _f] + [C.c_int] * len(extra)
        fn.restype = None
        fn(k, d, fp(cent), n, fp(v), len(subsets), ip(idx), ip(ends), fp(out), *extra)
        return out
    if bof:
        out = np.zeros(k, np.int32)
        if prefix:   # oracle: one entry point for ma >= 1
            fn = L.orc_bof_compute_ma
            fn.argtypes = [C.c_int, C.c_int, _f, C.c_int, _f, _i, C.c_int, C.c_int]
            fn.restype = None
            fn(k, d, fp(cent), n, fp(v), ip(out), ma or 1, dot)
        elif ma:
            fn = L.bof_compute_ma
            fn.argtypes = [C.c_int, C.c_int, _f, C.c_int, _f, _i, C.c_int, C.c_float, C.c_int]
            fn.restype = None
            fn(k, d, fp(cent), n, fp(v), ip(out), ma, 0.0, 1)
        else:
            fn = L.bof_compute
            fn.argtypes = [C.c_int, C.c_int, _f, C.c_int, _f, _i]
            fn.restype = None
            fn(k, d, fp(cent), n, fp(v), ip(out))
        return out
    out = np.zeros((k, d), np.float32)
    if prefix:
        fn = L.orc_vlad_compute
        fn.argtypes = [C.c_int, C.c_int, _f, C.c_int, _f, _f, _f, C.c_int]
        fn.restype = None
        fn(k, d, fp(cent), n, fp(v), fp(f32(weights)) if weights is not None else None, fp(out), dot)
    elif weights is not None:
        fn = L.vlad_compute_weighted
        fn.argtypes = [C.c_int, C.c_int, _f, C.c_int, _f, _f, _f]
        fn.restype = None
        fn(k, d, fp(cent), n, fp(v), fp(f32(weights)), fp(out))
    else:
        fn = L.vlad_compute
        fn.argtypes = [C.c_int, C.c_int, _f, C.c_int, _f, _f]
        fn.restype = None
        fn(k, d, fp(cent), n, fp(v), fp(out))
    return out


def orc_vlad(cent, v, weights=None, subsets=None, dot_mode=DOT_F32_SEQ):
    return _vlad_family(oracle(), "orc_", dot_mode, cent, v, weights=weights, subsets=subsets)


def orc_bof(cent, v, ma=None, subsets=None, dot_mode=DOT_F32_SEQ):
    return _vlad_family(oracle(), "orc_", dot_mode, cent, v, subsets=subsets, ma=ma, bof=True)


def ref_vlad(cent, v, weights=None, subsets=None):
    return _vlad_family(ref(), "", None, cent, v, weights=weights, subsets=subsets)


def ref_bof(cent, v, ma=None, subsets=None):
    return _vlad_family(ref(), "", None, cent, v, subsets=subsets, ma=ma, bof=True)


# ---- hierarchical k-means quantiser (yael/hkm.c:144-162) and GMM E-step (yael/gmm.c:305-367)
def make_hkm(levels, bf):
    """hkm_t over numpy tables: levels[l] is [bf^(l+1)][d]; returns (struct, keep-alive list)."""
    levels = [f32(x) for x in levels]
    d = levels[0].shape[1]
    for l, x in enumerate(levels):
        assert x.shape == (bf ** (l + 1), d)
    ptrs = (_f * len(levels))(*[fp(x) for x in levels])
    h = HkmT(len(levels), bf, bf ** len(levels), d, C.cast(ptrs, C.POINTER(_f)))
    return h, (levels, ptrs)


def ref_hkm_learn(v, nlevel, bf, niter=10):
    """hkm_learn of the compiled reference (yael/hkm.c:35-118; kmeans seeds come from lrand48)."""
    v = f32(v)
    n, d = v.shape
    out = _i()
    h = ref().hkm_learn(n, d, nlevel, bf, fp(v), niter, 1, 0, C.byref(out))
    levels = [np.ctypeslib.as_array(h.contents.centroids[l], shape=(bf ** (l + 1), d)).copy()
              for l in range(nlevel)]
    assign = np.ctypeslib.as_array(out, shape=(n,)).copy()
    ref().hkm_delete(h)
    return levels, assign


def ref_hkm_quantize(levels, bf, v):
    v = f32(v)
    h, keep = make_hkm(levels, bf)
    idx = np.empty(v.shape[0], np.int32)
    ref().hkm_quantize(C.byref(h), v.shape[0], fp(v), ip(idx))
    return idx


def orc_hkm_quantize(levels, bf, v, dot_mode=DOT_F32_SEQ):
    v = f32(v)
    cat = np.ascontiguousarray(np.concatenate([f32(x).reshape(-1) for x in levels]), np.float32)
    idx = np.empty(v.shape[0], np.int32)
    oracle().orc_hkm_quantize(len(levels), bf, v.shape[1], fp(cat), v.shape[0], fp(v), ip(idx), dot_mode)
    return idx


def make_gmm(w, mu, sigma):
    w, mu, sigma = f32(w), f32(mu), f32(sigma)
    k, d = mu.shape
    assert sigma.shape == (k, d) and w.shape == (k,)
    return GmmT(d, k, fp(w), fp(mu), fp(sigma)), (w, mu, sigma)


def ref_gmm_compute_p(w, mu, sigma, v, flags=1, nt=1):
    v = f32(v)
    g, keep = make_gmm(w, mu, sigma)
    p = np.empty((v.shape[0], g.k), np.float32)
    if nt > 1:
        ref().gmm_compute_p_thread(v.shape[0], fp(v), C.byref(g), fp(p), flags, nt)
    else:
        ref().gmm_compute_p(v.shape[0], fp(v), C.byref(g), fp(p), flags)
    return p


def orc_gmm_compute_p(w, mu, sigma, v, flags=1, dot_mode=DOT_F32_SEQ):
    v, w, mu, sigma = f32(v), f32(w), f32(mu), f32(sigma)
    k, d = mu.shape
    p = np.empty((v.shape[0], k), np.float32)
    oracle().orc_gmm_compute_p(v.shape[0], d, k, fp(w), fp(mu), fp(sigma), fp(v), fp(p), flags, dot_mode)
    return p

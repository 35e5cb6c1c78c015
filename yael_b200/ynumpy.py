"""numpy front-end mirroring the reference's yael/ynumpy.py call shapes
(knn :46-67, cross_distances :69-84, kmeans :87-130, kmin/kmax :372-390), bound to
libyael_b200.so through the reference's own C prototypes.  Same argument names and order,
same return values, same error behaviour (kmeans raises RuntimeError when the C call
returns a negative qerr)."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import lib

KMEANS_QUIET = 0x10000
KMEANS_INIT_BERKELEY = 0x20000
KMEANS_NORMALIZE_CENTS = 0x40000
KMEANS_INIT_RANDOM = 0x80000
KMEANS_INIT_USER = 0x100000
KMEANS_L1 = 0x200000
KMEANS_CHI2 = 0x400000

_f = C.POINTER(C.c_float)
_i = C.POINTER(C.c_int)
_u8 = C.POINTER(C.c_uint8)
_u16 = C.POINTER(C.c_uint16)


def _check_row_float32(a):
    # yael/ynumpy.py:24-27
    if a.dtype != np.float32:
        raise TypeError("expected float32 matrix, got %s" % a.dtype)
    if not a.flags.c_contiguous:
        raise TypeError("expected C order matrix")


def _check_row_uint8(a):
    if a.dtype != np.uint8:
        raise TypeError("expected uint8 matrix, got %s" % a.dtype)
    if not a.flags.c_contiguous:
        raise TypeError("expected C order matrix")


def _fp(a):
    return a.ctypes.data_as(_f)


def _ip(a):
    return a.ctypes.data_as(_i)


def knn(queries, base, nnn=1, distance_type=2, nt=1):
    """yael/ynumpy.py:46-67 -> knn_full_thread (yael/nn.c:679-699)."""
    _check_row_float32(base)
    _check_row_float32(queries)
    n, d = base.shape
    nq, d2 = queries.shape
    assert d == d2, "base and queries must have same nb of rows (got %d != %d) " % (d, d2)
    assert nnn <= n
    _lib.require_gpu()
    idx = np.empty((nq, nnn), dtype=np.int32)
    dis = np.empty((nq, nnn), dtype=np.float32)
    lib().knn_full_thread(distance_type, nq, n, d, nnn, _fp(base), _fp(queries), None,
                          _ip(idx), _fp(dis), nt)
    return idx, dis


def knn_weighted(queries, base, weights, nnn=1, nt=1, distance_type=2):
    """knn_full_thread with the per-base-vector weights of yael/nn.c:497-500 (any distance type)."""
    _check_row_float32(base)
    _check_row_float32(queries)
    weights = np.ascontiguousarray(weights, dtype=np.float32)
    n, d = base.shape
    nq = queries.shape[0]
    _lib.require_gpu()
    idx = np.empty((nq, nnn), dtype=np.int32)
    dis = np.empty((nq, nnn), dtype=np.float32)
    lib().knn_full_thread(distance_type, nq, n, d, nnn, _fp(base), _fp(queries), _fp(weights),
                          _ip(idx), _fp(dis), nt)
    return idx, dis


def knn_reorder_shortlist(queries, base, idx):
    """yael/nn.c:528-580; idx is re-ordered in place, distances returned."""
    _check_row_float32(base)
    _check_row_float32(queries)
    assert idx.dtype == np.int32 and idx.flags.c_contiguous
    nq, k = idx.shape
    _lib.require_gpu()
    dis = np.zeros((nq, k), dtype=np.float32)
    lib().knn_reorder_shortlist(nq, base.shape[0], base.shape[1], k, _fp(base), _fp(queries),
                                _ip(idx), _fp(dis))
    return dis


def cross_distances(a, b, distance_type=12):
    """yael/ynumpy.py:69-84 -> compute_cross_distances_alt_nonpacked (yael/nn.c:280-350)."""
    _check_row_float32(a)
    na, d = a.shape
    _check_row_float32(b)
    nb, d2 = b.shape
    assert d2 == d
    _lib.require_gpu()
    dis = np.empty((nb, na), dtype=np.float32)
    lib().compute_cross_distances_alt_nonpacked(distance_type, d, na, nb, _fp(a), d, _fp(b), d,
                                                _fp(dis), na)
    return dis


def kmeans(v, k, distance_type=2, nt=1, niter=30, seed=0, redo=1, verbose=True,
           normalize=False, init='random', output='centroids'):
    """yael/ynumpy.py:87-130 -> kmeans (yael/kmeans.c:332-447).  init may also be an array of
    k initial centroids (KMEANS_INIT_USER, yael/kmeans.h:15)."""
    _check_row_float32(v)
    n, d = v.shape
    _lib.require_gpu()
    centroids = np.zeros((k, d), dtype=np.float32)
    dis = np.empty(n, dtype=np.float32)
    assign = np.empty(n, dtype=np.int32)
    nassign = np.empty(k, dtype=np.int32)
    flags = nt
    if not verbose:
        flags |= KMEANS_QUIET
    if distance_type == 2:
        pass
    elif distance_type in (1, 3):
        # KMEANS_L1 / KMEANS_CHI2 (yael/kmeans.h:16-17): medians / Newton steps per coordinate, no
        # contraction -- outside the B200 hot path; say so instead of reporting a failed clustering
        raise NotImplementedError("yael_b200.kmeans: distance_type %d (KMEANS_%s) is not provided; "
                                  "only L2 k-means (distance_type=2) is"
                                  % (distance_type, "L1" if distance_type == 1 else "CHI2"))
    else:
        raise ValueError("kmeans: unknown distance_type %r" % (distance_type,))
    if isinstance(init, np.ndarray):
        assert init.shape == (k, d)
        centroids[:] = init
        flags |= KMEANS_INIT_USER
    elif init == 'random':
        flags |= KMEANS_INIT_RANDOM
    elif init == 'kmeans++':
        flags |= KMEANS_INIT_BERKELEY
    if normalize:
        flags |= KMEANS_NORMALIZE_CENTS
    qerr = lib().kmeans(d, n, k, niter, _fp(v), flags, seed, redo, _fp(centroids), _fp(dis),
                        _ip(assign), _ip(nassign))
    if qerr < 0:
        raise RuntimeError("kmeans: clustering failed. Is dataset diverse enough?")
    if output == 'centroids':
        return centroids
    return centroids, qerr, dis, assign, nassign


def kmin(v, k):
    """yael/ynumpy.py:372-380: indices of the k smallest values of each line."""
    _check_row_float32(v)
    n, d = v.shape
    assert k <= d
    _lib.require_gpu()
    idx = np.empty((n, k), dtype='int32')
    lib().fvecs_k_min(_fp(v), d, n, _ip(idx), k)
    return idx


def kmax(v, k):
    """yael/ynumpy.py:382-390."""
    _check_row_float32(v)
    n, d = v.shape
    assert k <= d
    _lib.require_gpu()
    idx = np.empty((n, k), dtype='int32')
    lib().fvecs_k_max(_fp(v), d, n, _ip(idx), k)
    return idx


def hamming_distances(a, b):
    """compute_hamming (yael/hamming.c:177-219): uint16 matrix dis[j, i] = ham(a_i, b_j)."""
    _check_row_uint8(a)
    _check_row_uint8(b)
    assert a.shape[1] == b.shape[1]
    _lib.require_gpu()
    dis = np.empty((b.shape[0], a.shape[0]), dtype=np.uint16)
    lib().compute_hamming(dis.ctypes.data_as(_u16), a.ctypes.data_as(_u8), b.ctypes.data_as(_u8),
                          a.shape[0], b.shape[0], a.shape[1])
    return dis


def knn_hamming(queries, base, nnn=1):
    """NEW nn_hamming: (idx, dis) of the nnn closest base codes, ordered by (distance, id)."""
    _check_row_uint8(base)
    _check_row_uint8(queries)
    assert base.shape[1] == queries.shape[1]
    assert nnn <= base.shape[0]
    _lib.require_gpu()
    nq = queries.shape[0]
    idx = np.empty((nq, nnn), dtype=np.int32)
    dis = np.empty((nq, nnn), dtype=np.uint16)
    lib().nn_hamming(nq, base.shape[0], base.shape[1], nnn, base.ctypes.data_as(_u8),
                     queries.ctypes.data_as(_u8), _ip(idx), dis.ctypes.data_as(_u16))
    return idx, dis


def match_hamming(a, b, ht):
    """match_hamming_count + match_hamming_thres_prealloc (yael/hamming.c:283-300, 704-748):
    returns (pairs[n, 2] as (qid, bid), scores[n])."""
    _check_row_uint8(a)
    _check_row_uint8(b)
    _lib.require_gpu()
    n = C.c_size_t(0)
    lib().match_hamming_count(a.ctypes.data_as(_u8), b.ctypes.data_as(_u8), a.shape[0],
                              b.shape[0], ht, a.shape[1], C.byref(n))
    pairs = np.empty((n.value, 2), dtype=np.int32)
    scores = np.empty(n.value, dtype=np.uint16)
    if n.value:
        lib().match_hamming_thres_prealloc(a.ctypes.data_as(_u8), b.ctypes.data_as(_u8),
                                           a.shape[0], b.shape[0], ht, a.shape[1], _ip(pairs),
                                           scores.ctypes.data_as(_u16))
    return pairs, scores


def crossmatch_hamming(a, ht):
    """crossmatch_hamming_count + crossmatch_hamming_prealloc (yael/hamming.c:368-395, 793-829):
    all pairs i < j of ONE code set within ht; returns (pairs[n, 2] as (i, j), scores[n])."""
    _check_row_uint8(a)
    _lib.require_gpu()
    n = C.c_size_t(0)
    lib().crossmatch_hamming_count(a.ctypes.data_as(_u8), a.shape[0], ht, a.shape[1], C.byref(n))
    pairs = np.empty((n.value, 2), dtype=np.int32)
    scores = np.empty(n.value, dtype=np.uint16)
    if n.value:
        lib().crossmatch_hamming_prealloc(a.ctypes.data_as(_u8), a.shape[0], ht, a.shape[1],
                                          _ip(pairs), scores.ctypes.data_as(_u16))
    return pairs, scores


# ---- consumers of the k = 1 search (include/yael/vlad.h; not in the reference's ynumpy: same
# ---- call shapes as the C functions, arrays in / arrays out)
def _subset_arrays(subsets):
    idx = np.ascontiguousarray(np.concatenate([np.asarray(s, np.int32) for s in subsets]) if len(subsets)
                               else np.zeros(0, np.int32), np.int32)
    ends = np.ascontiguousarray(np.cumsum([len(s) for s in subsets]), np.int32)
    return idx, ends


def vlad(centroids, v, weights=None, subsets=None):
    """vlad_compute / vlad_compute_weighted / vlad_compute_subsets (yael/vlad.c:10-79)."""
    _check_row_float32(centroids)
    _check_row_float32(v)
    k, d = centroids.shape
    n = v.shape[0]
    _lib.require_gpu()
    if subsets is not None:
        idx, ends = _subset_arrays(subsets)
        desc = np.zeros((len(subsets), k, d), np.float32)
        lib().vlad_compute_subsets(k, d, _fp(centroids), n, _fp(v), len(subsets), _ip(idx), _ip(ends), _fp(desc))
        return desc
    desc = np.zeros((k, d), np.float32)
    if weights is not None:
        w = np.ascontiguousarray(weights, np.float32)
        lib().vlad_compute_weighted(k, d, _fp(centroids), n, _fp(v), _fp(w), _fp(desc))
    else:
        lib().vlad_compute(k, d, _fp(centroids), n, _fp(v), _fp(desc))
    return desc


def bof(centroids, v, ma=None, subsets=None):
    """bof_compute / bof_compute_ma / bof_compute_subsets (yael/vlad.c:82-139)."""
    _check_row_float32(centroids)
    _check_row_float32(v)
    k, d = centroids.shape
    n = v.shape[0]
    _lib.require_gpu()
    if subsets is not None:
        idx, ends = _subset_arrays(subsets)
        desc = np.zeros((len(subsets), k), np.float32)
        lib().bof_compute_subsets(k, d, _fp(centroids), n, _fp(v), len(subsets), _ip(idx), _ip(ends), _fp(desc))
        return desc
    desc = np.zeros(k, np.int32)
    if ma:
        lib().bof_compute_ma(k, d, _fp(centroids), n, _fp(v), _ip(desc), int(ma), 0.0, 1)
    else:
        lib().bof_compute(k, d, _fp(centroids), n, _fp(v), _ip(desc))
    return desc


# ---- hierarchical k-means quantiser (include/yael/hkm.h, yael/hkm.c)
def _hkm_struct(levels, bf):
    levels = [np.ascontiguousarray(x, np.float32) for x in levels]
    d = levels[0].shape[1]
    for l, x in enumerate(levels):
        if x.shape != (bf ** (l + 1), d):
            raise ValueError("level %d: expected a %d x %d table, got %s" % (l, bf ** (l + 1), d, x.shape))
    ptrs = (_f * len(levels))(*[_fp(x) for x in levels])
    h = _lib.HkmT(len(levels), bf, bf ** len(levels), d, C.cast(ptrs, C.POINTER(_f)))
    return h, (levels, ptrs)


def hkm_learn(v, nlevel, bf, niter=20, nt=1, verbose=False):
    """hkm_learn (yael/hkm.c:35-118): returns (levels, assign): levels[l] is the [bf^(l+1)][d]
    table of level l, assign the leaf of every learning point."""
    _check_row_float32(v)
    _lib.require_gpu()
    n, d = v.shape
    out = _i()
    h = lib().hkm_learn(n, d, nlevel, bf, _fp(v), niter, nt, int(verbose), C.byref(out))
    levels = [np.ctypeslib.as_array(h.contents.centroids[l], shape=(bf ** (l + 1), d)).copy()
              for l in range(nlevel)]
    assign = np.ctypeslib.as_array(out, shape=(n,)).copy()
    lib().hkm_delete(h)
    libc = C.CDLL(None)
    libc.free.argtypes = [C.c_void_p]
    libc.free(C.cast(out, C.c_void_p))  # malloc'd by the library (hkm.c:112-115), freed by the caller
    return levels, assign


def hkm_quantize(levels, bf, v):
    """hkm_quantize (yael/hkm.c:144-162): the leaf of every row of v."""
    _check_row_float32(v)
    _lib.require_gpu()
    h, keep = _hkm_struct(levels, bf)
    if v.shape[1] != h.d:
        raise ValueError("points and tree have different dimensions")
    idx = np.empty(v.shape[0], np.int32)
    lib().hkm_quantize(C.byref(h), v.shape[0], _fp(v), _ip(idx))
    return idx


# ---- GMM E-step (include/yael/gmm.h, yael/gmm.c:305-367); gmm_npy = (w, mu, sigma) as in the
# ---- reference's ynumpy (yael/ynumpy.py:268-300)
GMM_FLAGS_W = 1


def gmm_compute_p(gmm_npy, v, flags=GMM_FLAGS_W):
    """gmm_compute_p: p[n][k], the posterior of every mixture component for every row of v."""
    w, mu, sigma = [np.ascontiguousarray(x, np.float32) for x in gmm_npy]
    _check_row_float32(v)
    _lib.require_gpu()
    k, d = mu.shape
    if sigma.shape != (k, d) or w.shape != (k,) or v.shape[1] != d:
        raise ValueError("inconsistent mixture / point shapes")
    g = _lib.GmmT(d, k, _fp(w), _fp(mu), _fp(sigma))
    p = np.empty((v.shape[0], k), np.float32)
    lib().gmm_compute_p(v.shape[0], _fp(v), C.byref(g), _fp(p), flags)
    return p

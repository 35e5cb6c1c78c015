"""CPU: the C-ABI library loads without a GPU, exports every symbol the headers declare, the
host-side pieces (RNG, heap, sort helpers, file format) behave like the reference's, and the
product fails LOUDLY -- never falls back -- when no device is present."""
import ctypes as C
import glob
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def declared_symbols():
    names = set()
    pat = re.compile(r"^\s*(?:[A-Za-z_][\w\s\*]*?)\b([A-Za-z_]\w*)\s*\([^;{]*\)\s*;", re.M)
    for h in glob.glob(os.path.join(ROOT, "include", "**", "*.h"), recursive=True):
        src = open(h).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        src = re.sub(r"typedef\s+struct\s*\{.*?\}\s*\w+\s*;", "", src, flags=re.S)
        for m in pat.finditer(src):
            n = m.group(1)
            if n not in ("defined", "sizeof"):
                names.add(n)
    return names


def test_library_loads_and_exports_every_declared_symbol():
    import yael_b200
    L = yael_b200.lib()  # attaches prototypes of the binding tables: fails on a missing symbol
    raw = C.CDLL(yael_b200.LIB_PATH)
    missing = [n for n in sorted(declared_symbols()) if not hasattr(raw, n)]
    assert not missing, "declared in include/ but not exported: %s" % missing
    assert L.yb_version().startswith(b"yael_b200")
    # the drop-in layer covers the reference's hot-path API
    for n in ("knn_full", "knn_full_thread", "nn", "nn_thread", "knn", "knn_thread",
              "knn_reorder_shortlist", "compute_cross_distances", "compute_cross_distances_nonpacked",
              "compute_cross_distances_thread", "compute_distances_1", "kmeans", "fvec_k_min",
              "fvecs_k_min", "fvec_k_max", "compute_hamming", "nn_hamming", "hamming",
              "match_hamming_count", "match_hamming_thres", "fbinheap_addn_label_range"):
        assert hasattr(raw, n), n


def test_no_cpu_fallback_without_gpu():
    import yael_b200
    L = yael_b200.lib()
    if L.yb_device_count() > 0:
        pytest.skip("a GPU is present")
    from yael_b200 import ynumpy
    b = np.zeros((10, 4), np.float32)
    with pytest.raises(yael_b200.YaelB200Error):
        ynumpy.knn(b, b, 1)
    # the raw C call aborts with a message instead of computing on the host
    code = ("import numpy as np, ctypes as C, yael_b200; L=yael_b200.lib();"
            "b=np.zeros((10,4),np.float32); i=np.zeros((10,1),np.int32); d=np.zeros((10,1),np.float32);"
            "f=C.POINTER(C.c_float); L.knn_full(2,10,10,4,1,b.ctypes.data_as(f),b.ctypes.data_as(f),None,"
            "i.ctypes.data_as(C.POINTER(C.c_int)),d.ctypes.data_as(f)); print('COMPUTED')")
    p = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True)
    assert p.returncode != 0 and "COMPUTED" not in p.stdout
    assert "yael_b200" in p.stderr


def test_host_rng_matches_reference_sequences():
    import yael_b200
    L = yael_b200.lib()
    g = np.load(os.path.join(GOLD, "rng.npz"))
    p = L.ivec_new_random_perm_r(1000, 4242)
    perm = np.ctypeslib.as_array(p, shape=(1000,)).copy()
    assert np.array_equal(perm, g["perm_n1000_seed4242"])
    x = np.empty(257, np.float32)
    L.fvec_randn_r(x.ctypes.data_as(C.POINTER(C.c_float)), 257, 99)
    assert np.array_equal(x, g["randn_n257_seed99"])


def test_host_binheap_matches_oracle_heap(ob):
    import yael_b200
    L = yael_b200.lib()
    r = np.random.RandomState(1)
    v = r.randint(0, 30, 5000).astype(np.float32)  # many ties: exercises the heap-slot order
    v[::97] = np.nan
    k = 25
    h = L.fbinheap_new(k)
    f, i = C.POINTER(C.c_float), C.POINTER(C.c_int)
    for s in range(0, 5000, 256):  # as knn_full feeds it (yael/nn.c:504-507)
        blk = np.ascontiguousarray(v[s:s + 256])
        L.fbinheap_addn_label_range(h, len(blk), s, blk.ctypes.data_as(f))
    lab = np.empty(k, np.int32)
    val = np.empty(k, np.float32)
    L.fbinheap_sort(h, lab.ctypes.data_as(i), val.ctypes.data_as(f))
    L.fbinheap_delete(h)
    O = ob.oracle()
    oh = O.orc_heap_new(k) if hasattr(O, "orc_heap_new") else None
    O.orc_heap_new.restype = C.c_void_p
    O.orc_heap_addn_range.argtypes = [C.c_void_p, C.c_int, C.c_int, f]
    O.orc_heap_sorted.argtypes = [C.c_void_p, i, f]
    O.orc_heap_free.argtypes = [C.c_void_p]
    oh = O.orc_heap_new(k)
    for s in range(0, 5000, 256):
        blk = np.ascontiguousarray(v[s:s + 256])
        O.orc_heap_addn_range(oh, len(blk), s, blk.ctypes.data_as(f))
    olab = np.empty(k, np.int32)
    oval = np.empty(k, np.float32)
    O.orc_heap_sorted(oh, olab.ctypes.data_as(i), oval.ctypes.data_as(f))
    O.orc_heap_free(oh)
    assert np.array_equal(lab, olab) and np.array_equal(val, oval)
    assert L.fbinheap_sizeof(100) == 824


def test_host_sort_helpers_and_file_format(tmp_path):
    import yael_b200
    L = yael_b200.lib()
    f, i = C.POINTER(C.c_float), C.POINTER(C.c_int)
    t = np.array([3, 1, 2, 1, 0.5], np.float32)
    perm = np.empty(5, np.int32)
    L.fvec_sort_index(t.ctypes.data_as(f), 5, perm.ctypes.data_as(i))
    assert perm.tolist() == [4, 1, 3, 2, 0]
    assert L.fvec_arg_min(t.ctypes.data_as(f), 5) == 4
    # .fvecs round trip: [int32 d][d floats] per vector (doc/file_format.rst:4-20)
    m = np.arange(12, dtype=np.float32).reshape(3, 4)
    path = str(tmp_path / "x.fvecs").encode()
    assert L.fvecs_write(path, 4, 3, m.ctypes.data_as(f)) == 3
    raw = np.fromfile(path.decode(), dtype=np.int32)
    assert raw.size == 3 * 5 and (raw[::5] == 4).all()
    d, n = C.c_int(), C.c_int()
    assert L.fvecs_fsize(path, C.byref(d), C.byref(n)) == 3 * 20
    assert (d.value, n.value) == (4, 3)
    back = np.empty((3, 4), np.float32)
    assert L.fvecs_read(path, 4, 3, back.ctypes.data_as(f)) == 3
    assert np.array_equal(back, m)
    assert L.count_cpu() >= 1


def test_python_frontend_argument_checks():
    from yael_b200 import ynumpy
    with pytest.raises(TypeError):
        ynumpy._check_row_float32(np.zeros((2, 2), np.float64))
    with pytest.raises(TypeError):
        ynumpy._check_row_float32(np.zeros((4, 4), np.float32)[:, ::2])


def test_host_only_entry_points_hamming_and_recompute_exact_dists(ob):
    """hamming() (yael/hamming.c:66-78) and knn_recompute_exact_dists() (yael/nn.c:583-600) stay on
    the host (a scalar popcount; pointer chasing over a partially loaded base): compare with the
    oracle / the compiled reference and with numpy."""
    import yael_b200
    L = yael_b200.lib()
    u8 = C.POINTER(C.c_uint8)
    r = np.random.RandomState(3)
    for nc in (4, 8, 16, 5, 24):
        a = r.randint(0, 256, nc).astype(np.uint8)
        b = r.randint(0, 256, nc).astype(np.uint8)
        want = int(np.unpackbits(a ^ b).sum())
        assert L.hamming(a.ctypes.data_as(u8), b.ctypes.data_as(u8), nc) == want
        assert ob.oracle().orc_hamming(ob.u8p(a), ob.u8p(b), nc) == want
    # knn_recompute_exact_dists: base rows [label0, label0 + nb) are loaded; per query, entries
    # kp[q].. of the (label-sorted) shortlist are recomputed until a label falls outside
    f, i = C.POINTER(C.c_float), C.POINTER(C.c_int)
    nq, nb, d, k, label0 = 7, 50, 12, 9, 100
    b = r.random_sample((nb, d)).astype(np.float32)
    v = r.random_sample((nq, d)).astype(np.float32)
    idx = np.sort(r.randint(label0, label0 + 2 * nb, (nq, k)), axis=1).astype(np.int32)
    kp0 = np.array([0, 2, 0, 1, 0, 0, 3], np.int32)

    def run(fn):
        kp = kp0.copy()
        dis = np.full((nq, k), -1.0, np.float32)
        fn(nq, nb, d, k, b.ctypes.data_as(f), v.ctypes.data_as(f), label0, kp.ctypes.data_as(i),
           idx.ctypes.data_as(i), dis.ctypes.data_as(f))
        return kp, dis

    kp, dis = run(L.knn_recompute_exact_dists)
    for q in range(nq):
        j = int(kp0[q])
        while j < k and idx[q, j] - label0 < nb:
            row = b[idx[q, j] - label0].astype(np.float64)
            assert dis[q, j] == np.float32(((row - v[q].astype(np.float64)) ** 2).sum()) or \
                abs(dis[q, j] - ((row - v[q]) ** 2).sum()) < 1e-5
            j += 1
        assert kp[q] == j and (dis[q, j:] == -1.0).all() and (dis[q, :kp0[q]] == -1.0).all()
    if ob.have_ref():
        R = ob.ref()
        R.knn_recompute_exact_dists.argtypes = [C.c_int] * 4 + [f, f, C.c_int, i, i, f]
        R.knn_recompute_exact_dists.restype = None
        rkp, rdis = run(R.knn_recompute_exact_dists)
        assert np.array_equal(kp, rkp) and np.array_equal(dis, rdis)


def test_hkm_and_gmm_files_round_trip(tmp_path):
    # hkm_write / hkm_read (yael/hkm.c:181-232) and gmm_write / gmm_read (yael/gmm.c:810-836): host
    # code, no device needed; the byte layout is the reference's (checked against a hand-built file)
    import yael_b200
    from yael_b200 import _lib, ynumpy
    L = yael_b200.lib()
    r = np.random.RandomState(5)
    bf, d = 3, 4
    levels = [r.rand(bf ** (l + 1), d).astype(np.float32) for l in range(2)]
    h, keep = ynumpy._hkm_struct(levels, bf)
    path = str(tmp_path / "tree.hkm").encode()
    L.hkm_write(path, C.byref(h))
    raw = open(path, "rb").read()
    want = np.array([2, bf, d], np.int32).tobytes()
    for x in levels:
        want += np.array([x.size], np.int32).tobytes() + x.tobytes()
    assert raw == want
    h2 = L.hkm_read(path)
    assert (h2.contents.nlevel, h2.contents.bf, h2.contents.k, h2.contents.d) == (2, bf, bf * bf, d)
    for l, x in enumerate(levels):
        got = np.ctypeslib.as_array(h2.contents.centroids[l], shape=x.shape)
        assert np.array_equal(got, x)
    node = np.ctypeslib.as_array(L.hkm_get_centroids(h2, 1, 2), shape=(bf, d))
    assert np.array_equal(node, levels[1][2 * bf:3 * bf])
    L.hkm_delete(h2)

    libc = C.CDLL(None)
    libc.fopen.restype = C.c_void_p
    libc.fopen.argtypes = [C.c_char_p, C.c_char_p]
    libc.fclose.argtypes = [C.c_void_p]
    k = 5
    w, mu, sg = r.rand(k).astype(np.float32), r.rand(k, d).astype(np.float32), r.rand(k, d).astype(np.float32)
    fp_ = lambda a: a.ctypes.data_as(C.POINTER(C.c_float))
    g = _lib.GmmT(d, k, fp_(w), fp_(mu), fp_(sg))
    gpath = str(tmp_path / "mix.gmm").encode()
    f = libc.fopen(gpath, b"w")
    L.gmm_write(C.byref(g), f)
    libc.fclose(f)
    assert open(gpath, "rb").read() == np.array([d, k], np.int32).tobytes() + w.tobytes() + mu.tobytes() + sg.tobytes()
    f = libc.fopen(gpath, b"r")
    g2 = L.gmm_read(f)
    libc.fclose(f)
    assert (g2.contents.d, g2.contents.k) == (d, k)
    assert np.array_equal(np.ctypeslib.as_array(g2.contents.sigma, shape=(k, d)), sg)
    L.gmm_delete(g2)

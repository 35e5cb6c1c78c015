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

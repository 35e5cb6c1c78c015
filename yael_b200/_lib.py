"""ctypes loader for libyael_b200.so (the product library).

The library must exist: there is no Python or CPU fallback.  `build()` compiles it in-tree
with nvcc for sm_100a (works without a GPU); loading works without a GPU too, but every
compute call needs one.
"""
import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("YAEL_B200_LIB") or os.path.join(HERE, "libyael_b200.so")  # override: A/B runs
CSRC = os.path.join(HERE, "csrc")

_f = C.POINTER(C.c_float)
_i = C.POINTER(C.c_int)
_u8 = C.POINTER(C.c_uint8)
_u16 = C.POINTER(C.c_uint16)
_vp = C.c_void_p


def build(verbose=False):
    """Compile every CUDA kernel for sm_100a and link libyael_b200.so in-tree."""
    out = subprocess.run(["make", "-C", CSRC, "-j8"], capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("yael_b200 build failed:\n" + out.stdout[-4000:] + out.stderr[-4000:])
    if verbose:
        print(out.stdout[-2000:])
    return LIB_PATH


class YaelB200Error(RuntimeError):
    pass


class HkmT(C.Structure):
    """hkm_t (include/yael/hkm.h, reference yael/hkm.h:11-17)"""
    _fields_ = [("nlevel", C.c_int), ("bf", C.c_int), ("k", C.c_int), ("d", C.c_int),
                ("centroids", C.POINTER(C.POINTER(C.c_float)))]


class GmmT(C.Structure):
    """gmm_t (include/yael/gmm.h, reference yael/gmm.h:20-26)"""
    _fields_ = [("d", C.c_int), ("k", C.c_int), ("w", C.POINTER(C.c_float)),
                ("mu", C.POINTER(C.c_float)), ("sigma", C.POINTER(C.c_float))]


# (name, restype, argtypes) of the drop-in layer: exactly the reference's prototypes
DROPIN = {
    # include/yael/nn.h  (reference yael/nn.h:41-214)
    "knn_full": (None, [C.c_int] * 5 + [_f, _f, _f, _i, _f]),
    "knn_full_thread": (None, [C.c_int] * 5 + [_f, _f, _f, _i, _f, C.c_int]),
    "nn": (C.c_double, [C.c_int] * 3 + [_f, _f, _i]),
    "nn_thread": (C.c_double, [C.c_int] * 3 + [_f, _f, _i, C.c_int]),
    "knn": (_f, [C.c_int] * 4 + [_f, _f, _i]),
    "knn_thread": (_f, [C.c_int] * 4 + [_f, _f, _i, C.c_int]),
    "knn_reorder_shortlist": (None, [C.c_int] * 4 + [_f, _f, _i, _f]),
    "knn_recompute_exact_dists": (None, [C.c_int] * 4 + [_f, _f, C.c_int, _i, _i, _f]),
    "compute_cross_distances": (None, [C.c_int] * 3 + [_f, _f, _f]),
    "compute_cross_distances_nonpacked": (None, [C.c_int] * 3 + [_f, C.c_int, _f, C.c_int, _f, C.c_int]),
    "compute_cross_distances_thread": (None, [C.c_int] * 3 + [_f, _f, _f, C.c_int]),
    "compute_cross_distances_alt": (None, [C.c_int] * 4 + [_f, _f, _f]),
    "compute_cross_distances_alt_nonpacked": (None, [C.c_int] * 4 + [_f, C.c_int, _f, C.c_int, _f, C.c_int]),
    "compute_cross_distances_alt_thread": (None, [C.c_int] * 4 + [_f, _f, _f, C.c_int]),
    "compute_distances_1": (None, [C.c_int, C.c_int, _f, _f, _f]),
    "compute_distances_1_nonpacked": (None, [C.c_int, C.c_int, _f, _f, C.c_int, _f]),
    "compute_distances_1_thread": (None, [C.c_int, C.c_int, _f, _f, _f, C.c_int]),
    "compute_distances_1_nonpacked_thread": (None, [C.c_int, C.c_int, _f, _f, C.c_int, _f, C.c_int]),
    # include/yael/kmeans.h  (reference yael/kmeans.h:41-66)
    "kmeans": (C.c_float, [C.c_int] * 4 + [_f, C.c_int, C.c_long, C.c_int, _f, _f, _i, _i]),
    "clustering_kmeans": (_f, [C.c_int, C.c_int, _f, C.c_int, C.c_int, C.c_double]),
    "clustering_kmeans_assign": (_f, [C.c_int, C.c_int, _f, C.c_int, C.c_int, C.c_double, C.POINTER(_i)]),
    "clustering_kmeans_assign_with_score": (
        _f, [C.c_int, C.c_int, _f, C.c_int, C.c_int, C.c_double, C.c_int, C.POINTER(C.c_double), C.POINTER(_i)]),
    # include/yael/sorting.h  (reference yael/sorting.h:19-40)
    "fvec_k_min": (None, [_f, C.c_int, _i, C.c_int]),
    "fvec_k_max": (None, [_f, C.c_int, _i, C.c_int]),
    "fvecs_k_min": (None, [_f, C.c_long, C.c_long, _i, C.c_int]),
    "fvecs_k_max": (None, [_f, C.c_long, C.c_long, _i, C.c_int]),
    "fvec_sort_index": (None, [_f, C.c_int, _i]),
    "fvec_arg_min": (C.c_int, [_f, C.c_long]),
    # include/yael/hamming.h  (reference yael/hamming.h:24-66) + NEW nn_hamming
    "hamming": (C.c_uint16, [_u8, _u8, C.c_int]),
    "compute_hamming": (None, [_u16, _u8, _u8, C.c_int, C.c_int, C.c_int]),
    "nn_hamming": (None, [C.c_int] * 4 + [_u8, _u8, _i, _u16]),
    "match_hamming_count": (None, [_u8, _u8, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_size_t)]),
    "match_hamming_thres": (None, [_u8, _u8, C.c_int, C.c_int, C.c_int, C.c_int, C.c_size_t,
                                   C.POINTER(_vp), C.POINTER(C.c_size_t)]),
    "match_hamming_thres_prealloc": (C.c_size_t, [_u8, _u8, C.c_int, C.c_int, C.c_int, C.c_int, _i, _u16]),
    "crossmatch_hamming_count": (None, [_u8, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_size_t)]),
    "crossmatch_hamming": (None, [_u8, C.c_long, C.c_int, C.c_int, C.c_long, C.POINTER(_vp),
                                  C.POINTER(C.c_size_t)]),
    "crossmatch_hamming_prealloc": (C.c_size_t, [_u8, C.c_long, C.c_int, C.c_int, _i, _u16]),
    "compute_hamming_thread": (None, [_u16, _u8, _u8, C.c_int, C.c_int, C.c_int]),
    # include/yael/vlad.h  (reference yael/vlad.h:9-49)
    "vlad_compute": (None, [C.c_int, C.c_int, _f, C.c_int, _f, _f]),
    "vlad_compute_weighted": (None, [C.c_int, C.c_int, _f, C.c_int, _f, _f, _f]),
    "vlad_compute_subsets": (None, [C.c_int, C.c_int, _f, C.c_int, _f, C.c_int, _i, _i, _f]),
    "bof_compute": (None, [C.c_int, C.c_int, _f, C.c_int, _f, _i]),
    "bof_compute_ma": (None, [C.c_int, C.c_int, _f, C.c_int, _f, _i, C.c_int, C.c_float, C.c_int]),
    "bof_compute_subsets": (None, [C.c_int, C.c_int, _f, C.c_int, _f, C.c_int, _i, _i, _f]),
    # include/yael/hkm.h  (reference yael/hkm.h:21-40)
    "hkm_learn": (C.POINTER(HkmT), [C.c_int] * 4 + [_f, C.c_int, C.c_int, C.c_int, C.POINTER(_i)]),
    "hkm_delete": (None, [C.POINTER(HkmT)]),
    "hkm_quantize": (None, [C.POINTER(HkmT), C.c_int, _f, _i]),
    "hkm_write": (None, [C.c_char_p, C.POINTER(HkmT)]),
    "hkm_read": (C.POINTER(HkmT), [C.c_char_p]),
    "hkm_get_centroids": (_f, [C.POINTER(HkmT), C.c_int, C.c_int]),
    # include/yael/gmm.h  (reference yael/gmm.h:63-66,121-131)
    "gmm_compute_p": (None, [C.c_int, _f, C.POINTER(GmmT), _f, C.c_int]),
    "gmm_compute_p_thread": (None, [C.c_int, _f, C.POINTER(GmmT), _f, C.c_int, C.c_int]),
    "gmm_delete": (None, [C.POINTER(GmmT)]),
    "gmm_write": (None, [C.POINTER(GmmT), _vp]),
    "gmm_read": (C.POINTER(GmmT), [_vp]),
    # include/yael/binheap.h  (reference yael/binheap.h:31-87)
    "fbinheap_new": (_vp, [C.c_int]),
    "fbinheap_sizeof": (C.c_size_t, [C.c_int]),
    "fbinheap_init": (None, [_vp, C.c_int]),
    "fbinheap_delete": (None, [_vp]),
    "fbinheap_reset": (None, [_vp]),
    "fbinheap_add": (None, [_vp, C.c_int, C.c_float]),
    "fbinheap_pop": (None, [_vp]),
    "fbinheap_addn": (None, [_vp, C.c_int, _i, _f]),
    "fbinheap_addn_label_range": (None, [_vp, C.c_int, C.c_int, _f]),
    "fbinheap_sort_labels": (None, [_vp, _i]),
    "fbinheap_sort_values": (None, [_vp, _f]),
    "fbinheap_sort": (None, [_vp, _i, _f]),
    # include/yael/vector.h (subset) and machinedeps.h
    "fvec_new": (_f, [C.c_long]),
    "ivec_new": (_i, [C.c_long]),
    "fvec_randn_r": (None, [_f, C.c_long, C.c_uint]),
    "fvec_rand_r": (None, [_f, C.c_long, C.c_uint]),
    "ivec_new_random_perm_r": (_i, [C.c_int, C.c_uint]),
    "ivec_new_random_idx_r": (_i, [C.c_int, C.c_int, C.c_uint]),
    "fvec_sum": (C.c_double, [_f, C.c_long]),
    "fvec_norm": (C.c_double, [_f, C.c_long, C.c_double]),
    "fvec_normalize": (C.c_double, [_f, C.c_long, C.c_double]),
    "fvec_purge_nans": (C.c_long, [_f, C.c_long, C.c_float]),
    "fvecs_fsize": (C.c_long, [C.c_char_p, _i, _i]),
    "ivecs_fsize": (C.c_long, [C.c_char_p, _i, _i]),
    "bvecs_fsize": (C.c_long, [C.c_char_p, _i, _i]),
    "fvecs_read": (C.c_int, [C.c_char_p, C.c_int, C.c_int, _f]),
    "fvecs_new_read": (C.c_int, [C.c_char_p, _i, C.POINTER(_f)]),
    "ivecs_new_read": (C.c_int, [C.c_char_p, _i, C.POINTER(_i)]),
    "fvecs_write": (C.c_int, [C.c_char_p, C.c_int, C.c_int, _f]),
    "ivecs_write": (C.c_int, [C.c_char_p, C.c_int, C.c_int, _i]),
    "count_cpu": (C.c_int, []),
    "getmillisecs": (C.c_double, []),
}

# the device-level layer (include/yael_b200.h); pointers are raw addresses (c_void_p) so
# torch tensors' data_ptr() can be passed straight through
DEVICE = {
    "yb_version": (C.c_char_p, []),
    "yb_last_error": (C.c_char_p, []),
    "yb_device_count": (C.c_int, []),
    "yb_set_device": (C.c_int, [C.c_int]),
    "yb_sync": (C.c_int, [_vp]),
    "yb_launch_count": (C.c_long, [C.c_int]),
    "yb_last_knn_engine": (C.c_int, []),
    "yb_last_knn_operands": (C.c_int, []),
    "yb_last_knn_uncertified": (C.c_long, []),
    "yb_set_knn_engine": (None, [C.c_int]),
    "yb_prof_enable": (None, [C.c_int]),
    "yb_prof_ms": (C.c_double, [C.c_int, C.POINTER(C.c_long), C.c_int]),
    "yb_malloc": (_vp, [C.c_size_t]),
    "yb_free": (None, [_vp]),
    "yb_is_device_ptr": (C.c_int, [_vp]),
    "yb_gather_rows": (C.c_int, [_vp, _vp, C.c_int, C.c_int, _vp, _vp]),
    "yb_h2d": (C.c_int, [_vp, _vp, C.c_size_t, _vp]),
    "yb_d2h": (C.c_int, [_vp, _vp, C.c_size_t, _vp]),
    "yb_release_scratch": (None, []),
    "yb_cross_distances_l2": (C.c_int, [C.c_int] * 3 + [_vp, C.c_int, _vp, C.c_int, _vp, C.c_int, _vp]),
    "yb_set_cross_engine": (None, [C.c_int]),
    "yb_last_cross_engine": (C.c_int, []),
    "yb_distances_1": (C.c_int, [C.c_int, C.c_int, _vp, _vp, C.c_int, _vp, _vp]),
    "yb_cross_distances_alt": (C.c_int, [C.c_int] * 4 + [_vp, C.c_int, _vp, C.c_int, _vp, C.c_int, _vp]),
    "yb_knn_l2": (C.c_int, [C.c_int] * 4 + [_vp, _vp, _vp, _vp, _vp, C.c_int, _vp]),
    "yb_knn_alt": (C.c_int, [C.c_int] * 5 + [_vp, _vp, _vp, _vp, _vp, _vp]),
    "yb_knn_l2_hostbase": (C.c_int, [C.c_int] * 4 + [_vp] * 5 + [C.c_int, _vp]),
    "yb_knn_merge": (C.c_int, [C.c_int] * 3 + [_vp, _vp, _vp, _vp, _vp]),
    "yb_knn_merge_strided": (C.c_int, [C.c_int] * 3 + [_vp, _vp, C.c_long, _vp, _vp, _vp]),
    "yb_knn_reorder_shortlist": (C.c_int, [C.c_int] * 4 + [_vp, _vp, _vp, _vp, _vp]),
    "yb_k_min_rows": (C.c_int, [_vp, C.c_long, C.c_long, C.c_long, C.c_int, C.c_int, _vp, _vp, _vp]),
    "yb_kmeans_accumulate": (C.c_int, [C.c_int] * 3 + [_vp, _vp, _vp, _vp, _vp, _vp, C.c_int, _vp]),
    "yb_kmeans_scale": (C.c_int, [C.c_int, C.c_int, _vp, _vp, _vp, C.c_int, _vp]),
    "yb_kmeans_dev": (C.c_float, [C.c_int] * 4 + [_vp, C.c_int, C.c_long, C.c_int, _f, _f, _i, _i, _vp, _vp]),
    "yb_debug_tf32_clocks": (C.c_int, [_vp, C.c_int]),
    "yb_debug_tf32_scores": (C.c_int, [C.c_int] * 3 + [_vp, _vp, _vp, _vp]),
    "yb_debug_f16_scores": (C.c_int, [C.c_int] * 3 + [_vp, _vp, _vp, _vp]),
    "yb_debug_popc_pairs_per_s": (C.c_double, [_vp]),
    "yb_compute_hamming": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int, C.c_int, _vp]),
    "yb_nn_hamming": (C.c_int, [C.c_int] * 4 + [_vp, _vp, _vp, _vp, C.c_int, _vp]),
    "yb_nn_hamming_merge": (C.c_int, [C.c_int] * 3 + [_vp, _vp, _vp, _vp, _vp]),
    "yb_set_hamming_engine": (None, [C.c_int]),
    "yb_last_hamming_engine": (C.c_int, []),
    "yb_last_hamming_fallbacks": (C.c_long, []),
    "yb_debug_hamming_tc_scores": (C.c_int, [C.c_int] * 3 + [_vp, _vp, _vp, _vp]),
    "yb_debug_hamming_tc_packed": (C.c_int, [C.c_int] * 4 + [_vp, _vp, _vp, _vp]),
    "yb_vlad_accumulate": (C.c_int, [C.c_int, C.c_int, _vp, C.c_long, _vp, _vp, _vp, _vp, _vp, _vp]),
    "yb_bof_accumulate": (C.c_int, [C.c_int, C.c_long, _vp, _vp, C.c_long, _vp, _vp, _vp]),
    "yb_debug_store_pattern_gbs": (C.c_double, [_vp, C.c_long, C.c_int, C.c_int, C.c_int, _vp]),
    "yb_hkm_quantize": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(_vp), C.c_long, _vp, _vp, _vp]),
    "yb_gmm_posteriors": (C.c_int, [C.c_long, C.c_int, C.c_int] + [_vp] * 9),
    "yb_match_hamming_count": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp]),
    "yb_match_hamming_thres": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, _vp]),
    "yb_crossmatch_hamming_count": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, _vp, _vp]),
    "yb_crossmatch_hamming": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, _vp]),
    # sharded hot path: NCCL exchange inside the library (yb_comm.cu)
    "yb_comm_available": (C.c_int, []),
    "yb_comm_unique_id": (C.c_int, [_vp]),
    "yb_comm_create": (_vp, [_vp, C.c_int, C.c_int]),
    "yb_comm_create_all": (C.c_int, [C.c_int, _i, C.POINTER(_vp)]),
    "yb_comm_destroy": (None, [_vp]),
    "yb_comm_p2p": (C.c_int, [_vp]),
    "yb_comm_rank": (C.c_int, [_vp]),
    "yb_comm_world": (C.c_int, [_vp]),
    "yb_comm_allreduce_f32": (C.c_int, [_vp, _vp, C.c_long, _vp]),
    "yb_comm_allgather": (C.c_int, [_vp, _vp, _vp, C.c_long, _vp]),
    "yb_knn_l2_sharded": (C.c_int, [_vp] + [C.c_int] * 4 + [_vp, _vp, C.c_int, _vp, _vp, _vp]),
    "yb_knn_l2_sharded_hostbase": (C.c_int, [_vp] + [C.c_int] * 4 + [_vp, _vp, _vp, C.c_int, _vp, _vp, _vp]),
    "yb_nn_hamming_sharded": (C.c_int, [_vp] + [C.c_int] * 4 + [_vp, _vp, C.c_int, _vp, _vp, _vp]),
    "yb_kmeans_sharded": (C.c_float, [_vp, C.c_int, C.c_int, C.c_long, C.c_int, C.c_int, _vp, C.c_int,
                                      C.c_long, _f, _f, _i, _i, _vp]),
    # the drop-in layer's in-process multi-GPU mode (yb_mgpu.cu)
    "yb_mgpu_set_devices": (C.c_int, [C.c_int, _i]),
    "yb_mgpu_device_count": (C.c_int, []),
    "yb_mgpu_last_used": (C.c_int, []),
    "yb_mgpu_knn_full": (C.c_int, [C.c_int] * 4 + [_f, _f, _i, _f]),
    "yb_mgpu_nn_hamming": (C.c_int, [C.c_int] * 4 + [_u8, _u8, _i, _u16]),
    "yb_mgpu_kmeans": (C.c_int, [C.c_int] * 4 + [_f, C.c_int, C.c_long, C.c_int, _f, _f, _i, _i, _f]),
}

_lib = None


def lib():
    """The loaded library, prototypes attached.  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise YaelB200Error(
                "libyael_b200.so is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or make -C yael_b200/csrc). There is no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for table in (DROPIN, DEVICE):
            for name, (res, args) in table.items():
                if name.startswith("yb_debug_") and os.environ.get("YAEL_B200_LIB") and not hasattr(L, name):
                    continue  # A/B runs against an older build (developer override only)
                fn = getattr(L, name)  # AttributeError here = header/library mismatch
                fn.restype = res
                fn.argtypes = args
        _lib = L
    return _lib


def check(rc, what):
    if rc != 0:
        raise YaelB200Error("%s failed (code %d): %s" % (what, rc, lib().yb_last_error().decode()))


def require_gpu():
    n = lib().yb_device_count()
    if n <= 0:
        raise YaelB200Error("no CUDA device: yael_b200 has no CPU fallback (%s)"
                            % lib().yb_last_error().decode())
    return n

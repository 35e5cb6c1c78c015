"""Tiny device-memory helper for tests that call the yb_ device-level C ABI directly, and the
k-NN comparison rule shared by the GPU parity tests."""
import ctypes as C

import numpy as np

import yael_b200


class DevArray:
    def __init__(self, arr=None, shape=None, dtype=None):
        L = yael_b200.lib()
        if arr is not None:
            arr = np.ascontiguousarray(arr)
            shape, dtype = arr.shape, arr.dtype
        self.shape, self.dtype = tuple(shape), np.dtype(dtype)
        self.nbytes = int(np.prod(self.shape)) * self.dtype.itemsize
        self.ptr = L.yb_malloc(max(self.nbytes, 1))
        if arr is not None and self.nbytes:
            rc = L.yb_h2d(self.ptr, arr.ctypes.data_as(C.c_void_p), self.nbytes, None)
            assert rc == 0, L.yb_last_error()
            L.yb_sync(None)

    def get(self):
        L = yael_b200.lib()
        out = np.empty(self.shape, self.dtype)
        if self.nbytes:
            rc = L.yb_d2h(out.ctypes.data_as(C.c_void_p), self.ptr, self.nbytes, None)
            assert rc == 0, L.yb_last_error()
            L.yb_sync(None)
        return out

    def free(self):
        if self.ptr:
            yael_b200.lib().yb_free(self.ptr)
            self.ptr = None


def check_knn(idx, dis, widx, wdis, b, q, rtol=1e-5):
    """North-star tolerance: distances within 1e-5 relative of the oracle's, ids identical except
    for ties inside that tolerance.  A differing id is accepted only if ITS OWN distance to the
    query, recomputed here in float64 from the rows, is within the tolerance of the oracle's
    distance at that rank (a wrong id carrying a right distance does not pass), and no id may
    appear twice in a result list."""
    valid = widx >= 0
    assert np.array_equal(valid, idx >= 0)
    np.testing.assert_allclose(dis[valid], wdis[valid], rtol=rtol, atol=1e-6)
    diff = (idx != widx) & valid
    if diff.any():
        b64 = b.astype(np.float64)
        for qi, j in zip(*np.nonzero(diff)):
            own = ((b64[idx[qi, j]] - q[qi].astype(np.float64)) ** 2).sum()
            tol = rtol * max(abs(float(wdis[qi, j])), 1e-6) + 2e-6 * float((q[qi].astype(np.float64) ** 2).sum())
            assert abs(own - float(wdis[qi, j])) <= tol, (qi, j, idx[qi, j], widx[qi, j], own, wdis[qi, j])
        for qi in np.unique(np.nonzero(diff)[0]):
            row = idx[qi][idx[qi] >= 0]
            assert len(set(row.tolist())) == len(row), "duplicate id in the result list of query %d" % qi

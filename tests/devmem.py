"""Tiny device-memory helper for tests that call the yb_ device-level C ABI directly."""
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

"""ctypes wrapper of oracle/libr8boracle.so -- the C restatement (test infrastructure only)."""
import ctypes as C
import os

import numpy as np

from oracle_util import PORT_LIB

_L = None


def _lib():
    global _L
    if _L is None:
        L = C.CDLL(PORT_LIB)
        L.r8bo_create.restype = C.c_void_p
        L.r8bo_create.argtypes = [C.c_double, C.c_double, C.c_int, C.c_double, C.c_double, C.c_int]
        for f in ("r8bo_clear", "r8bo_delete"):
            getattr(L, f).argtypes = [C.c_void_p]
            getattr(L, f).restype = None
        L.r8bo_max_out_len.argtypes = [C.c_void_p]
        L.r8bo_stage_count.argtypes = [C.c_void_p]
        L.r8bo_stage_kind.argtypes = [C.c_void_p, C.c_int]
        L.r8bo_in_len_before_out_pos.argtypes = [C.c_void_p, C.c_int]
        L.r8bo_stage_data.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.r8bo_process.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        _L = L
    return _L


class PortOracle:
    def __init__(self, extfft=0):
        self.extfft = int(extfft)
        self.name = "port(extfft=%d)" % self.extfft
        oracle = self

        class Resampler:
            def __init__(self, src, dst, max_in_len, tb=2.0, atten=180.15):
                L = _lib()
                self.h = L.r8bo_create(src, dst, max_in_len, tb, atten, oracle.extfft)
                self.max_out_len = L.r8bo_max_out_len(self.h)
                self._out = np.empty(self.max_out_len + 16)

            def __del__(self):
                if getattr(self, "h", None):
                    _lib().r8bo_delete(self.h)
                    self.h = None

            def process(self, x):
                x = np.ascontiguousarray(x, dtype=np.float64)
                n = _lib().r8bo_process(self.h, x.ctypes.data, len(x), self._out.ctypes.data, len(self._out))
                return self._out[:n].copy()

            def clear(self):
                _lib().r8bo_clear(self.h)

            def in_len_before_out_pos(self, p):
                return _lib().r8bo_in_len_before_out_pos(self.h, p)

            def stage_kinds(self):
                return [_lib().r8bo_stage_kind(self.h, i) for i in range(_lib().r8bo_stage_count(self.h))]

            def stage_data(self, i):
                n = _lib().r8bo_stage_data(self.h, i, None, 0)
                a = np.empty(n)
                _lib().r8bo_stage_data(self.h, i, a.ctypes.data, n)
                return a

        self.Resampler = Resampler

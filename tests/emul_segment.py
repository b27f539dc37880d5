"""TEST-ONLY stand-in for `whatshap_b200._lib.Segment`: the same five operations executed by the host
emulation harness (tests/emul/emul.cpp: the kernels' __host__ __device__ functions stepped serially).
Lets the multi-GPU pedigree scheme of `whatshap_b200.multigpu` be checked against the oracle in the
GPU-less container; never part of the product."""
import ctypes as C
import os
import subprocess

import numpy as np

from whatshap_b200._abi import CProblem, CSolution, FlatSolution, raise_for

EMUL_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "emul")
_U32P = C.POINTER(C.c_uint32)
_lib = None


def lib():
    global _lib
    if _lib is None:
        subprocess.run(["make", "-C", EMUL_DIR, "all"], check=True, capture_output=True)
        L = C.CDLL(os.path.join(EMUL_DIR, "libwhemul.so"))
        L.whemul_segment_create.argtypes = [C.POINTER(CProblem), C.c_int, C.POINTER(C.c_void_p), C.c_char_p, C.c_size_t]
        L.whemul_segment_transfer.argtypes = [C.c_void_p, _U32P, C.c_char_p, C.c_size_t]
        L.whemul_segment_sweep.argtypes = [C.c_void_p, _U32P, _U32P, C.c_char_p, C.c_size_t]
        L.whemul_segment_exits.argtypes = [C.c_void_p, C.c_int, _U32P, C.c_char_p, C.c_size_t]
        L.whemul_segment_finish.argtypes = [C.c_void_p, C.c_int, C.POINTER(CSolution), C.c_char_p, C.c_size_t]
        L.whemul_segment_destroy.argtypes = [C.c_void_p]
        L.whemul_segment_destroy.restype = None
        _lib = L
    return _lib


class EmulSegment:
    def __init__(self, prob, continues):
        self.prob, self.T, self._h = prob, 4 ** prob.n_trios, C.c_void_p()
        cp, err = prob.as_c(), C.create_string_buffer(512)
        raise_for(lib().whemul_segment_create(C.byref(cp), int(bool(continues)), C.byref(self._h), err, len(err)), err.value.decode())

    @staticmethod
    def _ptr(a):
        return a.ctypes.data_as(_U32P)

    def transfer(self):
        m, err = np.zeros((self.T, self.T), np.uint32), C.create_string_buffer(512)
        raise_for(lib().whemul_segment_transfer(self._h, self._ptr(m), err, len(err)), err.value.decode())
        return m

    def sweep(self, in_vec=None):
        out, err = np.zeros(self.T, np.uint32), C.create_string_buffer(512)
        vec = None if in_vec is None else np.ascontiguousarray(in_vec, np.uint32)
        raise_for(lib().whemul_segment_sweep(self._h, None if vec is None else self._ptr(vec), self._ptr(out), err, len(err)), err.value.decode())
        return out

    def exits(self, is_last):
        out, err = np.zeros(self.T, np.uint32), C.create_string_buffer(512)
        raise_for(lib().whemul_segment_exits(self._h, int(bool(is_last)), self._ptr(out), err, len(err)), err.value.decode())
        return out

    def finish(self, entry):
        sol = FlatSolution(self.prob.n_cols, self.prob.n_reads, self.prob.n_ind)
        cs, err = sol.as_c(), C.create_string_buffer(512)
        raise_for(lib().whemul_segment_finish(self._h, int(entry), C.byref(cs), err, len(err)), err.value.decode())
        sol.cost = int(cs.cost)
        return sol

    def close(self):
        if self._h:
            lib().whemul_segment_destroy(self._h)
            self._h = C.c_void_p()

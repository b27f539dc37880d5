"""TEST-ONLY: the genotyping kernels' per-cell code (whatshap_b200/csrc/gl_device.h) and host packer (gl_pack.cpp)
executed serially on the host by tests/emul/libwhemul.so — a stand-in for `whmec_genotype` where there is no GPU."""
import ctypes as C
import os
import subprocess

import numpy as np

from whatshap_b200._abi import CProblem, raise_for

EMUL_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "emul")
_lib = None


def lib():
    global _lib
    if _lib is None:
        subprocess.run(["make", "-C", EMUL_DIR, "all"], check=True, capture_output=True)
        _lib = C.CDLL(os.path.join(EMUL_DIR, "libwhemul.so"))
        _lib.whemul_genotype.argtypes = [C.POINTER(CProblem), C.POINTER(C.c_double), C.c_char_p, C.c_size_t]
        _lib.whemul_genotype_grouped.argtypes = [C.POINTER(CProblem), C.POINTER(C.c_double), C.c_uint64, C.POINTER(C.c_uint32), C.c_char_p, C.c_size_t]
    return _lib


def genotype(prob, device: int = 0, budget_doubles: int = 0):
    """budget_doubles > 0: the tables (chains of a single individual) are processed in groups whose backward tables and
    projection buffers fit that many doubles, each with its own launch schedule and buffers — what whmec_genotype does when
    they do not fit the device together.  The second return value reports the number of groups."""
    out = np.zeros((prob.n_ind, prob.n_cols, 3), np.float64)
    cp, err = prob.as_c(), C.create_string_buffer(512)
    groups = C.c_uint32(0)
    raise_for(lib().whemul_genotype_grouped(C.byref(cp), out.ctypes.data_as(C.POINTER(C.c_double)), budget_doubles, C.byref(groups), err, len(err)),
              err.value.decode() or "a single table exceeds the budget")
    return out, {"groups": int(groups.value)}

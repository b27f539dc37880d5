"""Host phases of whmec_solve for a single individual, timed without a GPU through the test-only emulation library
(tests/emul: whemul_time_host_product = packer without deltas + tile planner + output pass on an arbitrary path, heap
retention as in the library).  Usage:  python scripts/host_phase_timing.py [path/to/libwhemul.so] [workload ...]
WHMEC_HOST_THREADS selects the host threads.  Best of 10 calls per workload."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from whatshap_b200 import synth  # noqa: E402
from whatshap_b200._abi import CProblem  # noqa: E402

args = sys.argv[1:]
path = args.pop(0) if args and args[0].endswith(".so") else os.path.join(ROOT, "tests", "emul", "libwhemul.so")
lib = C.CDLL(path)
lib.whemul_time_host_product.argtypes = [C.POINTER(CProblem), C.POINTER(C.c_double)]
for name in args or ["cfg2", "cfg3", "cfg4"]:
    prob = synth.config(name)
    cp, out = prob.as_c(), (C.c_double * 3)()
    best = [1e9] * 3
    for _ in range(10):
        assert lib.whemul_time_host_product(C.byref(cp), out) == 0
        best = [min(a, b) for a, b in zip(best, out)]
    print("%-5s threads %-3s pack %6.2f ms  plan %6.2f ms  outputs %6.2f ms" % ((name, os.environ.get("WHMEC_HOST_THREADS", "all")) + tuple(best)), flush=True)

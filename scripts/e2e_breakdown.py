"""Where the end-to-end time of whmec_solve goes (GPU box).  For every workload: the library's own phase timers
(WHMEC_TIMING=1, stderr), the wall time of the C call alone and of the Python wrapper around it, for several host thread counts.
Usage: python scripts/e2e_breakdown.py [workload ...]   (subprocess per setting: the thread pool is sized at first use)"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import ctypes as C, os, sys, time
sys.path.insert(0, %r)
from whatshap_b200 import _lib, synth
from whatshap_b200._abi import CStats, FlatSolution
name = sys.argv[1]
prob = synth.config(name)
for _ in range(3):
    _lib.solve(prob)
best_py = best_c = 1e9
for it in range(6):
    t0 = time.perf_counter()
    sol = FlatSolution(prob.n_cols, prob.n_reads, prob.n_ind)
    cp, cs, st = prob.as_c(), sol.as_c(), CStats()
    err = C.create_string_buffer(512)
    t1 = time.perf_counter()
    rc = _lib.lib().whmec_solve(C.byref(cp), C.byref(cs), 0, C.byref(st), err, len(err))
    t2 = time.perf_counter()
    assert rc == 0, err.value
    best_py = min(best_py, t2 - t0)
    best_c = min(best_c, t2 - t1)
print("RESULT %%s threads=%%s  wrapper+call %%.2f ms  C call %%.2f ms" %% (name, os.environ.get("WHMEC_HOST_THREADS", "all"), best_py * 1e3, best_c * 1e3), flush=True)
''' % ROOT

if __name__ == "__main__":
    for name in sys.argv[1:] or ["cfg3", "cfg2", "cfg5"]:
        for threads, extra in ((None, {}), ("32", {}), ("16", {}), (None, {"WHMEC_PINNED_UPLOAD": "0"})) if name == "cfg3" else ((None, {}), (None, {"WHMEC_PINNED_UPLOAD": "0"})):
            env = dict(os.environ, WHMEC_TIMING="1", **extra)
            if threads:
                env["WHMEC_HOST_THREADS"] = threads
            r = subprocess.run([sys.executable, "-c", CHILD, name], env=env, capture_output=True, text=True, timeout=300)
            lines = [l for l in r.stderr.splitlines() if l.startswith("[whmec]")]
            print("==", name, "threads", threads or "all", extra, flush=True)
            for l in lines[-8:]:
                print("   ", l)
            print("   ", (r.stdout.strip().splitlines() or ["(no result) " + r.stderr[-300:]])[-1], flush=True)

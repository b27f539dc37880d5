"""B200 check of the group-wise whmec_solve (WHMEC_SOLVE_GROUPS, csrc/grouped.h): the result must equal the
ordinary single-plan solve bit for bit, and the end-to-end time from host arrays is printed for both.
    python scripts/gpu_grouped_check.py            (needs a GPU; not run in round 1: no GPU minutes were left)"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from whatshap_b200 import _lib, synth  # noqa: E402


def timed(prob, groups, repeats=5):
    if groups:
        os.environ["WHMEC_SOLVE_GROUPS"] = str(groups)
    else:
        os.environ.pop("WHMEC_SOLVE_GROUPS", None)
    best, sol = 1e9, None
    for _ in range(repeats):
        t = time.perf_counter()
        sol, stats = _lib.solve(prob)
        best = min(best, time.perf_counter() - t)
    return sol, best, stats


for name, n in (("cfg3", None), ("cfg2", None), ("cfg3", 5000)):
    prob = synth.config(name, n)
    whole, t_whole, _ = timed(prob, 0)
    for groups in (2, 4, 8):
        got, t, stats = timed(prob, groups)
        print(f"{name} n={prob.n_cols}: {groups} groups {t * 1e3:.1f} ms vs {t_whole * 1e3:.1f} ms, identical: {got.same_as(whole)}, "
              f"launches {stats['kernel_launches']}, device sweeps {stats['sweep_ms']:.1f} ms", flush=True)
        assert got.same_as(whole), got.diff(whole)

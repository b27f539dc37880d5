"""B200 check of the extremes the emulation test `test_extreme_values_and_pedigree_sizes` covers on the CPU: weights,
recombination costs and likelihoods that wrap u32 arithmetic (column kernel without the no-overflow shortcuts), T = 64 and
T = 256 pedigrees, 30 active reads.  CUDA result vs the CPU checker.   python scripts/gpu_extremes_check.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from oracle import checker  # noqa: E402
from test_emulation import _extreme_problems  # noqa: E402
from whatshap_b200 import _lib  # noqa: E402

ck = checker.best()
bad = 0
for label, prob in _extreme_problems():
    want = ck.solve(prob)
    got, stats = _lib.solve(prob)
    ok = got.same_as(want)
    bad += not ok
    print(("ok      " if ok else "MISMATCH"), label, "path", stats["path_kind"], "" if ok else got.diff(want), flush=True)
print("mismatches:", bad)
assert bad == 0

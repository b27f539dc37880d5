"""Parity fuzz of the multi-tile paths (coverage 15-18, irregular spans): tile kernel with global bits,
canonical and tile-major hand-offs, vs the compiled reference and vs the column kernel.
    python scripts/gpu_fuzz_highcov.py [seconds [min_coverage max_coverage]]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import checker  # noqa: E402
from whatshap_b200 import _lib, synth  # noqa: E402

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
cov_lo, cov_hi = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (15, 18)
ck = checker.best()
rng = np.random.default_rng(int(time.time()) & 0xFFFF)
t0 = time.time()
n = errs = multi = 0
while time.time() - t0 < budget:
    cov = int(rng.integers(cov_lo, cov_hi + 1))
    if n % 2:  # irregular spans
        prob = synth.random_problem(rng, int(rng.integers(12, 40)), cov, "single", distrust=bool(rng.integers(0, 3) == 0),
                                    conflict_free=True, max_phred=int(rng.integers(1, 40)), mean_len=float(rng.choice([10, 16, 24])),
                                    gap=float(rng.choice([0.0, 0.1])), burst=6)
    else:      # sliding windows: long runs of steady-state columns (the fast / thread-packed / packed 16-bit column code)
        length = int(rng.integers(cov + 6, 56))
        prob = synth.sliding_window(length, cov, block_len=int(rng.integers(cov + 4, length + 1)), seed=int(rng.integers(1 << 30)),
                                    gap=float(rng.random() * 0.12), max_phred=int(rng.choice([1, 2, 40, 90])))
        if rng.random() < 0.3:  # some homozygous sites
            prob.gt = prob.gt.copy()
            prob.gt[0, rng.random(prob.n_cols) < 0.1] = int(rng.integers(0, 3))
    os.environ.pop("WHMEC_FORCE_COLUMN_KERNEL", None)
    got, st = _lib.solve(prob)
    want = ck.solve(prob)
    if st["max_active"] > 15:
        multi += 1
    ok = got.same_as(want)
    if ok and n % 4 == 0:
        os.environ["WHMEC_FORCE_COLUMN_KERNEL"] = "1"
        col, _ = _lib.solve(prob)
        ok = col.same_as(want)
    if not ok:
        errs += 1
        print("MISMATCH cov", st["max_active"], "path", st["path_kind"], got.diff(want), flush=True)
    n += 1
print(f"fuzzed {n} problems ({multi} with more than 15 active reads) in {time.time() - t0:.0f} s, mismatches {errs}, checker {ck.kind}")
sys.exit(1 if errs else 0)

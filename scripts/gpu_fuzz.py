"""Long-running parity fuzz on a GPU box: CUDA (all kernel paths) vs the strongest CPU checker.
    python scripts/gpu_fuzz.py [seconds]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import checker  # noqa: E402
from whatshap_b200 import _lib, synth  # noqa: E402

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
ck = checker.best()
rng = np.random.default_rng(int(time.time()) & 0xFFFF)
peds = list(synth.PEDIGREES)
t0 = time.time()
n = errs = 0
kinds = {}
while time.time() - t0 < budget:
    ped = peds[n % len(peds)]
    single = ped == "single"
    maxcov = int(rng.integers(2, 14 if single else (7 if ped in ("quartet", "three_generations") else 10)))
    prob = synth.random_problem(rng, int(rng.integers(1, 50 if single else 24)), maxcov, ped, distrust=bool(rng.integers(0, 2)),
                                conflict_free=bool(rng.integers(0, 5)), max_phred=int(rng.integers(1, 40)),
                                mean_len=float(rng.choice([2, 4, 8, 14])), gap=float(rng.choice([0.0, 0.1, 0.3])))
    for env in ({}, {"WHMEC_FORCE_COLUMN_KERNEL": "1"}, {"WHMEC_PED_SEQUENTIAL": "1"}):
        if env and ((single and "WHMEC_PED_SEQUENTIAL" in env) or (not single and "WHMEC_FORCE_COLUMN_KERNEL" in env)):
            continue
        for k in ("WHMEC_FORCE_COLUMN_KERNEL", "WHMEC_PED_SEQUENTIAL"):
            os.environ.pop(k, None)
        os.environ.update(env)
        want = werr = got = gerr = None
        try:
            want = ck.solve(prob)
        except RuntimeError as e:
            werr = str(e)
        try:
            got, st = _lib.solve(prob)
            kinds[st["path_kind"]] = kinds.get(st["path_kind"], 0) + 1
        except RuntimeError as e:
            gerr = str(e)
        if werr != gerr or (want is not None and not got.same_as(want)):
            errs += 1
            print("MISMATCH", ped, env, werr, gerr, None if want is None or got is None else got.diff(want), flush=True)
    n += 1
print(f"fuzzed {n} problems in {time.time() - t0:.0f} s, mismatches {errs}, kernel paths used {kinds}, checker {ck.kind}")
sys.exit(1 if errs else 0)

"""B200 check of the fused per-chain pedigree sweep (WHMEC_PED_CHAIN=1, ped_chain_kernel in csrc/whmec.cu): results
must equal the batched sweep bit for bit (and the CPU checker on small problems); sweep times are printed for both.
    python scripts/gpu_ped_chain_check.py        (needs a GPU; written in round 1 after the GPU minutes ran out)"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import checker  # noqa: E402
from whatshap_b200 import _lib, synth  # noqa: E402


def solve(prob, chain):
    if chain:
        os.environ["WHMEC_PED_CHAIN"] = "1"
    else:
        os.environ.pop("WHMEC_PED_CHAIN", None)
    try:
        return _lib.solve(prob)
    finally:
        os.environ.pop("WHMEC_PED_CHAIN", None)


ck = checker.best()
rng = np.random.default_rng(17)
bad = done = 0
for it in range(200):
    ped = ["trio", "quartet", "three_generations", "trio_child_first"][it % 4]
    prob = synth.random_problem(rng, int(rng.integers(4, 60)), int(rng.integers(2, 7)), pedigree=ped, distrust=it % 3 == 0,
                                mean_len=float(rng.choice([1.5, 3.0, 6.0])))
    try:
        want = ck.solve(prob)
    except RuntimeError:
        continue
    got, stats = solve(prob, True)
    done += 1
    if not got.same_as(want):
        bad += 1
        print("MISMATCH", it, ped, got.diff(want), stats)
print(f"random pedigrees: {done} solved with the chain kernel, {bad} mismatches")
for name, n in (("cfg5", 4000), ("cfg5", None)):
    prob = synth.config(name, n)
    base, st0 = solve(prob, False)
    got, st1 = solve(prob, True)
    print(f"{name} n={prob.n_cols}: identical {got.same_as(base)}; sweep {st0['sweep_ms']:.2f} ms / {st0['kernel_launches']} launches (batched) "
          f"vs {st1['sweep_ms']:.2f} ms / {st1['kernel_launches']} launches (per-chain blocks)", flush=True)
    assert got.same_as(base), got.diff(base)
assert bad == 0

"""CPU fuzz of the genotyping DP: host packer + launch schedule + the kernels' per-cell code (tests/emul) against the compiled
reference GenotypeDPTable (oracle/_ref).  No GPU.  Authoring container only.
    python scripts/cpu_fuzz_genotype.py <seed> <seconds>"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import emul_genotype  # noqa: E402
from oracle import checker  # noqa: E402
from whatshap_b200 import synth  # noqa: E402

seed, budget = int(sys.argv[1]), float(sys.argv[2])
rng = np.random.default_rng(seed)
ref = checker.reference()
assert ref is not None, "needs the compiled reference (oracle/_ref)"
peds = ("single", "single", "two_unrelated", "trio", "trio_child_first", "quartet", "three_generations")
t0 = time.time()
n, worst = 0, 0.0
while time.time() - t0 < budget:
    ped = peds[n % len(peds)]
    cov = int(rng.integers(2, 11 if ped == "single" else (6 if ped in ("two_unrelated", "trio", "trio_child_first") else 5)))
    prob = synth.genotyping_problem(rng, int(rng.integers(2, 50)), cov, ped, prior=("uniform", "random", "sparse")[n % 3],
                                    max_phred=int(rng.choice([3, 40, 60, 300])), gap=float(rng.random() * 0.3),
                                    mean_len=float(rng.uniform(2, 10)), burst=int(rng.integers(2, 6)))
    budget_doubles = 0 if ped != "single" or n % 2 else int(rng.integers(200, 5000))
    try:
        got, _ = emul_genotype.genotype(prob, budget_doubles=budget_doubles)
    except Exception as e:  # a budget below one table is refused: retry without
        got, _ = emul_genotype.genotype(prob)
    want = ref.genotype(prob)
    d = np.abs(got - want)
    d = d[~(np.isnan(got) & np.isnan(want))]
    assert not np.isnan(d).any(), (seed, n, ped)
    worst = max(worst, float(d.max()) if d.size else 0.0)
    assert worst < 1e-9, (seed, n, ped, worst)
    n += 1
print("seed", seed, "problems", n, "max abs difference of the normalised likelihoods", worst)

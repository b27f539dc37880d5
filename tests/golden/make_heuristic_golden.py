"""Golden vectors of the reference's PedMecHeuristic (src/pedmecheuristic.cpp) for tests/test_heuristic.py: random pedigree problems
solved by the reference compiled in place (oracle/_ref: whref_heuristic).  Authoring container only (needs /root/reference).
    python tests/golden/make_heuristic_golden.py        ->  tests/golden/heuristic.npz
Problems on which the reference itself is undefined (distrusted genotypes with a zero mutation cost: empty phasing lists,
src/pedmecheuristic.cpp:505-530) are avoided by recombination costs >= 1 and at least two columns."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import checker  # noqa: E402
from whatshap_b200 import synth  # noqa: E402


def cases(count=60, seed0=20250999):
    shapes = list(synth.PEDIGREES)
    out = []
    k = 0
    while len(out) < count:
        rng = np.random.default_rng(seed0 + k)
        k += 1
        ped = shapes[int(rng.integers(len(shapes)))]
        prob = synth.random_problem(rng, int(rng.integers(2, 60)), int(rng.integers(2, 10)), pedigree=ped, distrust=bool(rng.integers(2)),
                                    max_phred=int(rng.choice([3, 10, 40])), mean_len=float(rng.choice([2.0, 4.0, 8.0])))
        if prob.n_reads == 0:
            continue
        prob.recombcost = np.maximum(prob.recombcost, 1).astype(np.uint32)
        out.append((prob, int(rng.choice([1, 2, 4, 16, 256]))))
    return out


PROBLEM_FIELDS = ("positions", "read_off", "ent_col", "ent_allele", "ent_phred", "read_ind", "recombcost", "trios", "gt")


def main():
    ref = checker.reference()
    assert ref is not None, "needs the compiled reference (oracle/_ref)"
    blob = {}
    todo = cases()
    for i, (prob, row_limit) in enumerate(todo):
        want = ref.heuristic(prob, row_limit, True)
        for f in PROBLEM_FIELDS:
            blob["%d.%s" % (i, f)] = getattr(prob, f)
        blob["%d.n_ind" % i] = np.array(prob.n_ind)
        blob["%d.distrust" % i] = np.array(int(prob.distrust))
        blob["%d.row_limit" % i] = np.array(row_limit)
        blob["%d.n_samples" % i] = np.array(want.n_samples)
        blob["%d.partition" % i] = want.partition
        blob["%d.transmission" % i] = want.transmission
        blob["%d.haplotypes" % i] = want.haplotypes[: want.n_samples]
        blob["%d.mutated" % i] = want.mutated[: want.n_samples]
    blob["n"] = np.array(len(todo))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "heuristic.npz"), **blob)
    print("wrote", len(todo), "cases")


if __name__ == "__main__":
    main()

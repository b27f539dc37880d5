"""CPU fuzz of the host heuristic solver `whmec_heuristic` (csrc/heuristic.cpp) against the reference's PedMecHeuristic compiled in
place (oracle/_ref, whref_heuristic).  Every problem runs in a forked child: on some inputs the REFERENCE itself crashes (e.g. empty
phasing lists with distrusted genotypes index out of bounds, src/pedmecheuristic.cpp:505-530); those are counted and skipped.
    python scripts/cpu_fuzz_heuristic.py <seed> <seconds> [no-mutations]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import checker  # noqa: E402
from whatshap_b200 import _lib, synth  # noqa: E402


def make(seed):
    rng = np.random.default_rng(seed)
    shapes = list(synth.PEDIGREES)
    ped = shapes[int(rng.integers(len(shapes)))]
    distrust = bool(rng.integers(2))
    prob = synth.random_problem(rng, int(rng.integers(2, 40)), int(rng.integers(2, 9)), pedigree=ped, distrust=distrust,
                                max_phred=int(rng.choice([3, 10, 40])), mean_len=float(rng.choice([2.0, 4.0, 8.0])))
    return prob, int(rng.choice([1, 2, 4, 16, 256])), ped, distrust


ALLOW = True


def main():
    global ALLOW
    seed0, secs = int(sys.argv[1]), float(sys.argv[2])
    ALLOW = not (len(sys.argv) > 3 and sys.argv[3] == "no-mutations")  # third argument: allow_mutations = False (NaN costs, Q2)
    ref = checker.reference()
    _lib.lib()
    t0, k, bad, ref_crash, ours_crash = time.time(), 0, 0, 0, 0
    while time.time() - t0 < secs:
        seed = seed0 * 1000000 + k
        k += 1
        prob, rl, ped, distrust = make(seed)
        if prob.n_reads == 0:
            continue
        pid = os.fork()
        if pid == 0:
            code = 0
            try:
                want = ref.heuristic(prob, rl, ALLOW)
            except Exception:
                os._exit(3)
            os.write(1, b"")  # (reference survived)
            try:
                got = _lib.heuristic(prob, rl, ALLOW)
                code = 0 if got.same_as(want) else 1
            except Exception:
                code = 2
            os._exit(code)
        _, status = os.waitpid(pid, 0)
        if os.WIFSIGNALED(status):
            # which side crashed?  run the product alone
            pid2 = os.fork()
            if pid2 == 0:
                try:
                    _lib.heuristic(prob, rl, ALLOW)
                except Exception:
                    pass
                os._exit(0)
            _, st2 = os.waitpid(pid2, 0)
            if os.WIFSIGNALED(st2):
                ours_crash += 1
                print("PRODUCT CRASH seed", seed, ped, distrust, rl, flush=True)
            else:
                ref_crash += 1  # the product survives alone: the reference went down
        elif os.WEXITSTATUS(status) == 1:
            bad += 1
            print("MISMATCH seed", seed, ped, distrust, rl, prob.n_cols, prob.n_reads, flush=True)
        elif os.WEXITSTATUS(status) == 2:
            bad += 1
            print("PRODUCT ERROR seed", seed, ped, distrust, rl, flush=True)
    print("seed", seed0, "problems", k, "mismatches", bad, "reference crashed on", ref_crash, "product crashed on", ours_crash)


if __name__ == "__main__":
    main()

"""The CPU checkers themselves: the plain-C restatement (oracle/mec_oracle.c) must reproduce the
golden vectors produced by the unmodified reference, and — where the compiled reference is
available (oracle/_ref) — agree with it on fresh seeded instances."""
import numpy as np
import pytest

import golden_io
from conftest import solve_or_error
from whatshap_b200 import synth


@pytest.mark.parametrize("group", golden_io.GROUPS)
def test_port_reproduces_reference_golden_vectors(port_checker, group):
    assert golden_io.check(port_checker.solve, group) > 0


def test_known_answer_costs_of_the_reference_tests():
    """Costs hard-coded in the reference's tests/test_pedigreephasing.py (see test_pedigreephasing.py here)."""
    import test_pedigreephasing as tp

    costs = {label.split("/", 1)[1]: sol.cost for label, _, sol, _ in golden_io.load("reference_cases") if label.startswith("pedigree/")}
    for name, case in tp.CASES.items():
        assert costs[name] == case["cost"], name


def test_port_agrees_with_compiled_reference_on_fresh_instances(port_checker):
    from oracle import checker as ck

    ref = ck.reference()
    if ref is None:
        pytest.skip("compiled reference (oracle/_ref) not present on this machine")
    rng = np.random.default_rng(4242)
    peds = list(synth.PEDIGREES)
    for it in range(250):
        ped = peds[it % len(peds)]
        maxcov = 6 if ped in ("quartet", "three_generations") else 9
        prob = synth.random_problem(rng, int(rng.integers(1, 14)), int(rng.integers(1, maxcov)), ped,
                                    distrust=bool(rng.integers(0, 2)), conflict_free=bool(rng.integers(0, 4)),
                                    max_phred=int(rng.integers(1, 8)))
        a, ea = solve_or_error(ref.solve, prob)
        b, eb = solve_or_error(port_checker.solve, prob)
        assert ea == eb, (it, ea, eb)
        if a is not None:
            assert a.same_as(b), (it, ped, a.diff(b))


def test_blocks_decompose_exactly(checker):
    """T = 1: cost adds and partitions / super-reads concatenate over DP-independent blocks
    (SURVEY.md §8(e)); this is what lets blocks shard across GPUs."""
    prob = synth.sliding_window(90, 7, block_len=30, seed=11)
    whole = checker.solve(prob)
    cost = 0
    for b0 in range(0, 90, 30):
        part = checker.solve(prob.slice_columns(b0, b0 + 30))
        cost += part.cost
        assert np.array_equal(part.path_index, whole.path_index[b0:b0 + 30])
        assert np.array_equal(part.sr_allele, whole.sr_allele[:, :, b0:b0 + 30])
        assert np.array_equal(part.sr_quality, whole.sr_quality[:, b0:b0 + 30])
    assert cost == whole.cost

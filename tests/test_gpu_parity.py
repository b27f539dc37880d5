"""GPU parity: the CUDA path, called through the C ABI, against the CPU checker (bit-exact)."""
import numpy as np
import pytest

from conftest import solve_or_error
from whatshap_b200 import synth

pytestmark = pytest.mark.gpu

PEDS = ["single", "single", "trio", "quartet", "two_unrelated", "trio_child_first", "three_generations"]


def assert_same(gpu, checker, prob, label=""):
    want, werr = solve_or_error(checker.solve, prob)
    got, gerr = solve_or_error(lambda p: gpu.solve(p)[0], prob)
    assert werr == gerr, (label, werr, gerr)
    if want is not None:
        assert got.same_as(want), (label, got.diff(want))


@pytest.mark.parametrize("seed", range(8))
def test_fuzz_irregular_instances(gpu, checker, seed):
    """Random spans, gaps, blanks, ties, trusted/distrusted genotypes, Mendelian conflicts."""
    rng = np.random.default_rng(1000 + seed)
    for it in range(60):
        ped = PEDS[it % len(PEDS)]
        maxcov = 6 if ped in ("quartet", "three_generations") else 9
        prob = synth.random_problem(
            rng,
            int(rng.integers(1, 18)),
            int(rng.integers(1, maxcov)),
            ped,
            distrust=bool(rng.integers(0, 2)),
            conflict_free=bool(rng.integers(0, 4)),
            max_phred=int(rng.integers(1, 8)),
        )
        assert_same(gpu, checker, prob, f"seed={seed} it={it} ped={ped}")


@pytest.mark.parametrize(
    "n,c,stride,block,gap",
    [(120, 10, 1, 50, 0.0), (90, 8, 2, 45, 0.1), (64, 12, 1, 32, 0.0), (40, 14, 1, 20, 0.05), (30, 16, 1, 30, 0.0)],
)
def test_sliding_window_blocks(gpu, checker, n, c, stride, block, gap):
    """Coverage-capped single-individual ReadSets; chain ends drop up to c reads at once."""
    prob = synth.sliding_window(n, c, stride=stride, block_len=block, seed=n * 31 + c, gap=gap)
    assert_same(gpu, checker, prob)


def test_trio_blocks(gpu, checker):
    prob = synth.trio(90, 3, block_len=30, seed=5)
    assert_same(gpu, checker, prob)
    prob = synth.trio(40, 4, block_len=40, seed=6)
    assert_same(gpu, checker, prob)


def test_empty(gpu, checker):
    prob = synth.sliding_window(0, 5)
    sol, _ = gpu.solve(prob)
    assert sol.cost == 0

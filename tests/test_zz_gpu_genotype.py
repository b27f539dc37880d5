"""GPU parity of whmec_genotype (forward-backward genotyping DP) against the reference-generated golden vectors and
the CPU checker, through the C ABI.  Tolerance as in tests/test_genotype.py: 1e-9 absolute on the normalised likelihoods."""
import numpy as np
import pytest

from oracle import checker
from whatshap_b200 import _lib, synth

pytestmark = pytest.mark.gpu

TOL = 1e-9


def close(a, b):
    return a.shape == b.shape and bool(np.all((np.abs(a - b) <= TOL) | (np.isnan(a) & np.isnan(b))))


def test_golden_vectors():
    from test_genotype import golden

    for label, prob, want, _ in golden():
        got, stats = _lib.genotype(prob)
        assert close(got, want), (label, float(np.nanmax(np.abs(got - want))))
        assert stats["path_kind"] == 4 and stats["kernel_launches"] > 0


def test_fuzz_against_checker():
    ck = checker.best()
    rng = np.random.default_rng(5)
    for it in range(80):
        ped = ("single", "trio", "quartet", "two_unrelated", "three_generations")[it % 5]
        prob = synth.genotyping_problem(rng, int(rng.integers(2, 40)), int(rng.integers(2, 10 if ped == "single" else 5)), ped,
                                        prior=("uniform", "random", "sparse")[it % 3])
        got, _ = _lib.genotype(prob)
        assert close(got, ck.genotype(prob)), (it, ped)


def test_wide_columns_and_many_chains():
    """Coverage 15 (the default cap of `whatshap genotype`): 32 768 cells per column; likelihoods sum to 1; a prefix of the
    chains equals the checker (the CPU reference needs ~1 ms per column at this width)."""
    prob = synth.genotyping_problem(np.random.default_rng(9), 400, 15, "single", prior="random", burst=8, mean_len=10.0)
    got, stats = _lib.genotype(prob)
    assert np.allclose(got.sum(axis=2), 1.0, atol=1e-9)
    assert close(got, checker.best().genotype(prob))

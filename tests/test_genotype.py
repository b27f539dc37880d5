"""The genotyping DP (the reference's GenotypeDPTable) without a GPU: the C restatement
(oracle/gl_oracle.c) against the reference-generated golden vectors and the live compiled reference, the kernels'
per-cell code + host packer (stepped on the host by tests/emul) against the same vectors, the Python surface, and
the error behaviour of the C ABI.

Tolerance: the reference computes in 80-bit long double; its own tests compare likelihoods with abs_tol = 1e-9
(whatshap/testhelpers.py:11-15).  Here: 1e-12 for the long double restatement, 1e-9 for the double-precision device
code (observed: ~4e-14)."""
import os

import numpy as np
import pytest

import emul_genotype
from oracle import checker
from whatshap_b200 import GenotypeDPTable, NumericSampleIds, Pedigree, PhredGenotypeLikelihoods, ReadSet, _lib, synth
from whatshap_b200._abi import FlatProblem
from whatshap_b200.testhelpers import canonic_index_to_biallelic_gt, string_to_readset

HERE = os.path.dirname(os.path.abspath(__file__))
TOL_DEVICE = 1e-9
TOL_PORT = 1e-12


def golden():
    z = np.load(os.path.join(HERE, "golden", "genotype.npz"))
    for i in range(int(z["n"])):
        g = lambda f: z[f"{i}.{f}"]
        prob = FlatProblem(positions=g("positions"), read_off=g("read_off"), ent_col=g("ent_col"), ent_allele=g("ent_allele"),
                           ent_phred=g("ent_phred"), read_ind=g("read_ind"), recombcost=g("recombcost"), n_ind=int(g("n_ind")),
                           trios=g("trios"), distrust=True, gl=g("gl"))
        stated = z[f"{i}.stated"] if f"{i}.stated" in z.files else None
        yield str(g("label")), prob, g("likelihoods"), stated


def close(a, b, tol):
    return a.shape == b.shape and bool(np.all((np.abs(a - b) <= tol) | (np.isnan(a) & np.isnan(b))))


def test_restatement_matches_the_reference_on_golden_vectors():
    n = 0
    for label, prob, want, stated in golden():
        got = checker.port().genotype(prob)
        assert close(got, want, TOL_PORT), (label, float(np.nanmax(np.abs(got - want))))
        if stated is not None:  # the numbers written in the reference's own test file
            assert close(got[0], stated, 1e-9), label
        n += 1
    assert n == 65


def test_restatement_matches_the_live_reference():
    ref = checker.reference()
    if ref is None:
        pytest.skip("compiled reference not available")
    rng = np.random.default_rng(7)
    for it in range(60):
        ped = ("single", "trio", "quartet", "two_unrelated")[it % 4]
        prob = synth.genotyping_problem(rng, int(rng.integers(2, 30)), int(rng.integers(2, 8 if ped == "single" else 5)), ped,
                                        prior=("uniform", "random", "sparse")[it % 3])
        assert close(checker.port().genotype(prob), ref.genotype(prob), TOL_PORT), (it, ped)


def test_kernel_code_matches_the_reference_on_golden_vectors():
    worst = 0.0
    for label, prob, want, _ in golden():
        got, _ = emul_genotype.genotype(prob)
        assert close(got, want, TOL_DEVICE), (label, float(np.nanmax(np.abs(got - want))))
        worst = max(worst, float(np.nanmax(np.abs(got - want))))
    assert worst < 1e-11  # doubles + max scaling vs long doubles + sum scaling: far inside the stated tolerance


def test_kernel_code_on_many_chains_and_wide_columns():
    """Single individual: every DP-independent chain is a table of its own (gl_pack.cpp); columns up to 14 reads."""
    rng = np.random.default_rng(11)
    for cov, n in ((3, 120), (9, 60), (14, 24)):
        prob = synth.genotyping_problem(rng, n, cov, "single", prior="random", burst=6, mean_len=6.0)
        got, _ = emul_genotype.genotype(prob)
        assert close(got, checker.port().genotype(prob), TOL_DEVICE), cov
        assert np.allclose(got.sum(axis=2), 1.0, atol=1e-12)


def test_groups_of_tables_give_the_same_likelihoods():
    """The launch schedule (one launch = one column of every table of a group, gl_schedule) under shrinking memory budgets
    (gl_groups): more and more groups, the same likelihoods; a budget below one table is refused."""
    from whatshap_b200._abi import Unsupported

    rng = np.random.default_rng(13)
    prob = synth.sliding_window(144, 5, block_len=12, seed=13, gap=0.1)  # 12 chains of 12 columns
    prob.gl = rng.random((1, prob.n_cols, 3)) + 0.05
    prob.recombcost = rng.integers(0, 30, prob.n_cols).astype(np.uint32)
    whole, info = emul_genotype.genotype(prob)
    assert info["groups"] == 1 and close(whole, checker.port().genotype(prob), TOL_DEVICE)
    seen, refused = set(), False
    for budget in (20000, 2000, 1000, 600, 400, 250, 8):  # doubles; the last one is below any table
        try:
            got, info = emul_genotype.genotype(prob, budget_doubles=budget)
        except Unsupported:
            refused = True
            continue
        assert not refused, "a smaller budget cannot be feasible after a larger one was refused"
        assert np.array_equal(got, whole, equal_nan=True), budget
        seen.add(info["groups"])
    assert refused and max(seen) > 3 and len(seen) >= 3, seen


@pytest.fixture
def emulated_backend(monkeypatch):
    monkeypatch.setattr(_lib, "genotype", emul_genotype.genotype)


def _single(reads, weights=None, priors=None):
    rs = string_to_readset(s=reads, w=weights, scale_quality=10)
    positions = rs.get_positions()
    ids = NumericSampleIds()
    ped = Pedigree(ids)
    gls = priors or [PhredGenotypeLikelihoods([1 / 3.0] * 3)] * len(positions)
    ped.add_individual("individual0", [canonic_index_to_biallelic_gt(1)] * len(positions), gls)
    return GenotypeDPTable(ids, rs, [1] * len(positions), ped), positions


def test_python_surface_known_answers(emulated_backend):
    """Values stated in the reference's tests/test_genotyping.py:113-190 (uniform priors and given priors)."""
    table, positions = _single("""
          11
           01
        """)
    want = [[0.06666666666666667, 0.3333333333333333, 0.6], [0.20930232558139536, 0.5813953488372093, 0.20930232558139536],
            [0.06666666666666667, 0.3333333333333333, 0.6]]
    for k in range(len(positions)):
        lk = table.get_genotype_likelihoods("individual0", k)
        assert isinstance(lk, PhredGenotypeLikelihoods)
        assert np.allclose(lk.as_vector(), want[k], rtol=0, atol=1e-9)
    table, positions = _single("""
          01
          11
        """, priors=[PhredGenotypeLikelihoods([0.1, 0.8, 0.1]), PhredGenotypeLikelihoods([0.1, 0.2, 0.7])])
    want = [[0.04257892641700095, 0.9148421471659981, 0.04257892641700095], [0.0016688611936185199, 0.05208684202468078, 0.9462442967817007]]
    for k in range(len(positions)):
        assert np.allclose(table.get_genotype_likelihoods("individual0", k).as_vector(), want[k], rtol=0, atol=1e-9)
    with pytest.raises(IndexError):
        table.get_genotype_likelihoods("individual0", len(positions))


def test_empty_readset_and_missing_priors(emulated_backend):
    ids = NumericSampleIds()
    ped = Pedigree(ids)
    ped.add_individual("individual0", [canonic_index_to_biallelic_gt(1)] * 2, [None, None])
    GenotypeDPTable(ids, ReadSet(), [1, 1], ped)  # no columns: nothing to do (reference: tests/test_genotyping.py:53-61)
    rs = string_to_readset("11\n01")
    with pytest.raises(RuntimeError, match="genotype likelihoods"):
        GenotypeDPTable(ids, rs, [1, 1], ped)


def test_c_abi_errors_need_no_gpu():
    """Host-detectable errors come back before CUDA is touched; with valid input and no device the call fails loudly."""
    prob = synth.genotyping_problem(np.random.default_rng(1), 10, 4)
    single = FlatProblem(positions=[10, 20, 30], read_off=[0, 2, 3], ent_col=[0, 1, 1], ent_allele=[0, 1, 1], ent_phred=[10, 10, 10],
                         read_ind=[0, 0], recombcost=[1, 1, 1], n_ind=1, distrust=True, gl=np.full((1, 3, 3), 1 / 3.0))
    with pytest.raises(RuntimeError, match="single variant"):
        _lib.genotype(single)
    no_priors = FlatProblem(positions=prob.positions, read_off=prob.read_off, ent_col=prob.ent_col, ent_allele=prob.ent_allele,
                            ent_phred=prob.ent_phred, read_ind=prob.read_ind, recombcost=prob.recombcost, n_ind=1)
    with pytest.raises(RuntimeError, match="priors"):
        _lib.genotype(no_priors)
    if _lib.device_count() == 0:
        with pytest.raises(RuntimeError, match="CUDA"):
            _lib.genotype(prob)


def test_prior_genotyper_is_bit_exact():
    """compute_genotypes (src/genotyper.cpp:12-54): host code on both sides, same operation order -> identical doubles."""
    z = np.load(os.path.join(HERE, "golden", "genotype.npz"))
    for i in range(int(z["n_prior"])):
        g = lambda f: z[f"prior.{i}.{f}"]
        prob = FlatProblem(positions=g("positions"), read_off=g("read_off"), ent_col=g("ent_col"), ent_allele=g("ent_allele"),
                           ent_phred=g("ent_phred"), read_ind=g("read_ind"), recombcost=g("recombcost"), n_ind=1)
        gl, gt = _lib.compute_genotypes(prob)
        assert np.array_equal(gl, g("gl")) and np.array_equal(gt, g("gt")), i


def test_prior_genotyper_python_surface():
    from whatshap_b200.core import Genotype, compute_genotypes

    rs = string_to_readset("\n".join(["11"] * 9 + ["01"]), "\n".join(["99"] * 9 + ["11"]))  # ten reads; error floor 0.05 per read
    genotypes, likelihoods = compute_genotypes(rs)
    assert len(genotypes) == len(likelihoods) == 2
    assert genotypes[1] == Genotype([1, 1]) and abs(sum(likelihoods[1]) - 1.0) < 1e-12
    assert all(isinstance(x, tuple) and len(x) == 3 for x in likelihoods)
    # a column with conflicting evidence stays uncalled (error probability >= 0.1 -> empty genotype)
    mixed, _ = compute_genotypes(string_to_readset("""
      11
      00
    """))
    assert all(g.is_none() for g in mixed)

"""Single-individual phasing through the drop-in `PedigreeDPTable`, compared with brute force.

Restates the reference's tests/test_phasing.py (read matrices :154-238, the four solver
set-ups :96-151, the comparison :39-76) against this package's classes."""
import pytest

from whatshap_b200 import NumericSampleIds, Pedigree, PedigreeDPTable, PhredGenotypeLikelihoods, ReadSet
from whatshap_b200.testhelpers import (
    brute_force_phase,
    canonic_index_list_to_biallelic_gt_list,
    canonic_index_to_biallelic_gt,
    string_to_readset,
)

pytestmark = pytest.mark.gpu

MATRICES = {
    "trivial": ("""
          11
           01
        """, None),
    "phase1": ("""
     10
     010
     010
    """, None),
    "phase2": ("""
      1  11010
      00 00101
      001 0101
    """, None),
    "phase3": ("""
      1  11010
      00 00101
      001 01010
    """, None),
    "phase4": ("""
      1  11010
      00 00101
      001 01110
       1    111
    """, None),
    "phase5": ("""
      0             0
      110111111111
      00100
           0001000000
           000
            10100
                  101
    """, None),
    "weighted1": ("""
      1  11010
      00 00101
      001 01110
       1    111
    """, """
      2  13112
      11 23359
      223 56789
       2    111
    """),
}


def test_phase_empty_readset(gpu):
    rs = ReadSet()
    pedigree = Pedigree(NumericSampleIds())
    pedigree.add_individual("individual0", canonic_index_list_to_biallelic_gt_list([1, 1]), [None, None])
    dp_table = PedigreeDPTable(rs, [1, 1], pedigree)
    superreads, transmission_vector = dp_table.get_super_reads()
    assert dp_table.get_optimal_cost() == 0
    assert transmission_vector == []
    assert len(superreads) == 1 and len(superreads[0]) == 2 and len(superreads[0][0]) == 0


def compare_with_brute_force(superreads, cost, partition, readset, all_heterozygous):
    assert len(superreads) == 2
    assert len(superreads[0]) == len(superreads[1])
    for v1, v2 in zip(*superreads):
        assert v1.position == v2.position
    haplotypes = tuple(sorted("".join(str(v.allele) for v in sr) for sr in superreads))
    exp_cost, exp_partition, solution_count, exp_h1, exp_h2 = brute_force_phase(readset, all_heterozygous)
    inverse = [1 - p for p in partition]
    assert partition == exp_partition or inverse == exp_partition
    assert solution_count == 1
    assert cost == exp_cost
    assert haplotypes in ((exp_h1, exp_h2), (exp_h2, exp_h1))


def build_pedigree_for(positions, all_heterozygous, trio):
    pedigree = Pedigree(NumericSampleIds())
    gls = [None if all_heterozygous else PhredGenotypeLikelihoods([0, 0, 0])] * len(positions)
    names = ["individual0", "individual1", "individual2"] if trio else ["individual0"]
    for name in names:
        pedigree.add_individual(name, [canonic_index_to_biallelic_gt(1) for _ in positions], gls)
    if trio:  # two relatives without reads must not change the answer
        pedigree.add_relationship("individual0", "individual1", "individual2")
    return pedigree


def solve(readset, positions, all_heterozygous, trio):
    pedigree = build_pedigree_for(positions, all_heterozygous, trio)
    return PedigreeDPTable(readset, [1] * len(positions), pedigree, distrust_genotypes=not all_heterozygous)


@pytest.mark.parametrize("name", sorted(MATRICES))
@pytest.mark.parametrize("trio", [False, True])
@pytest.mark.parametrize("all_heterozygous", [False, True])
def test_phasing_matches_brute_force(gpu, name, trio, all_heterozygous):
    reads, weights = MATRICES[name]
    readset = string_to_readset(reads, weights)
    positions = readset.get_positions()
    dp_table = solve(readset, positions, all_heterozygous, trio)
    superreads, transmission_vector = dp_table.get_super_reads()
    assert len(set(transmission_vector)) == 1
    compare_with_brute_force(superreads[0], dp_table.get_optimal_cost(), dp_table.get_optimal_partitioning(), readset, all_heterozygous)

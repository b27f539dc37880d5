"""Pedigree phasing (PedMEC) known-answer tests through the drop-in `PedigreeDPTable`.

Restates the cases of the reference's tests/test_pedigreephasing.py — expected costs
(:113,136,172,207,233,251,272,311,358,412,476), haplotypes, accepted transmission vectors
(:173-178,208) and the Mendelian allele-order check (:42-69) — as one table."""
from collections import defaultdict

import pytest

from whatshap_b200 import NumericSampleIds, Pedigree, PedigreeDPTable, PhredGenotypeLikelihoods, ReadSet
from whatshap_b200.testhelpers import canonic_index_list_to_biallelic_gt_list, string_to_readset_pedigree

pytestmark = pytest.mark.gpu

GL0 = [0, 0, 0]

CASES = {
    "trio1": dict(
        reads="""
          A 111
          A 010
          A 110
          B 001
          B 110
          B 101
          C 001
          C 010
          C 010
        """,
        genotypes=[[1, 2, 1], [1, 1, 1], [0, 1, 1]], trios=[(0, 1, 2)], recomb=[10, 10, 10],
        cost=2, constant_tv=True, haplotypes=[("111", "010"), ("001", "110"), ("010", "001")]),
    "trio2": dict(
        reads="""
          A 00
          A 00
          B 11
          B 11
          C 11
          C 00
        """,
        genotypes=[[2, 2], [0, 0], [1, 1]], trios=[(0, 1, 2)], recomb=[10, 10, 10],
        cost=8, constant_tv=True, haplotypes=[("11", "11"), ("00", "00"), ("00", "11")]),
    "trio3": dict(
        reads="""
          A 1111
          B 1010
          C 111000
          C 010101
          B 0101
          A  0000
          B  1010
          C  1010
          C  1100
          A   0000
          A   1111
          B   1010
          B    010
        """,
        genotypes=[[1] * 6, [1] * 6, [1, 2, 1, 1, 0, 1]], trios=[(0, 1, 2)], recomb=[3, 3, 3, 4, 3, 3],
        cost=4, tv_in=([0, 0, 0, 1, 1, 1], [1, 1, 1, 0, 0, 0], [2, 2, 2, 3, 3, 3], [3, 3, 3, 2, 2, 2]),
        haplotypes=[("111111", "000000"), ("010101", "101010"), ("111000", "010101")]),
    "trio4": dict(
        reads="""
          B 101
          B 101
          B 101
          A 111
          A 111
          A 111
          C 111
          C 111
          C 111
        """,
        genotypes=[[1, 1, 1]] * 3, trios=[(0, 1, 2)], recomb=[1, 1, 1],
        cost=2, tv_in=([0, 2, 0], [2, 0, 2], [1, 3, 1], [3, 1, 3]),
        haplotypes=[("111", "000"), ("101", "010"), ("111", "000")]),
    "trio5": dict(
        reads="""
          B 101
          B 101
          B 101
          A 111
          A 111
          A 111
          C 111
          C 111
          C 111
        """,
        genotypes=[[1, 1, 1]] * 3, trios=[(0, 1, 2)], recomb=[2, 2, 2],
        cost=3, constant_tv=True, haplotypes=[("111", "000"), ("111", "000"), ("111", "000")]),
    "trio_pure_genetic": dict(
        reads="", positions=[10, 20, 30, 40],
        genotypes=[[2, 1, 1, 0], [1, 2, 2, 1], [1, 1, 1, 0]], trios=[(0, 1, 2)], recomb=[2, 2, 2],
        cost=0, constant_tv=True, haplotypes=[("1110", "1000"), ("1111", "0110"), ("1000", "0110")]),
    "doubletrio_pure_genetic": dict(
        reads="", positions=[10, 20, 30, 40],
        genotypes=[[1, 2, 1, 0], [1, 0, 1, 1], [2, 1, 1, 0], [1, 2, 2, 1], [1, 1, 1, 0]],
        trios=[(0, 1, 2), (2, 3, 4)], recomb=[2, 2, 2],
        cost=0, constant_tv=True,
        haplotypes=[("0100", "1110"), ("0011", "1000"), ("1110", "1000"), ("1111", "0110"), ("1000", "0110")]),
    "quartet1": dict(
        reads="""
          A 111
          A 010
          A 110
          B 001
          B 110
          B 101
          C 001
          C 010
          C 010
          D 001
          D 010
          D 010
        """,
        genotypes=[[1, 2, 1], [1, 1, 1], [0, 1, 1], [0, 1, 1]], trios=[(0, 1, 2), (0, 1, 3)], recomb=[10, 10, 10],
        cost=2, constant_tv=True, haplotypes=[("111", "010"), ("001", "110"), ("001", "010"), ("001", "010")]),
    "quartet2": dict(
        reads="""
          A 111111
          A 000000
          B 010101
          B 101010
          C 000000
          C 010101
          D 000000
          D 010101
        """,
        genotypes=[[1] * 6, [1] * 6, [0, 1, 0, 1, 0, 1], [0, 1, 0, 1, 0, 1]], trios=[(0, 1, 2), (0, 1, 3)],
        recomb=[3] * 6, cost=0, constant_tv=True,
        haplotypes=[("111111", "000000"), ("010101", "101010"), ("000000", "010101"), ("000000", "010101")]),
    "quartet3": dict(
        reads="""
          A 1111
          A 0000
          B 1010
          C 111000
          C 010101
          D 000000
          D 010
          B 0101
          C  1100
          D  10010
          A   0000
          A   1111
          B   1010
          B   0101
        """,
        genotypes=[[1] * 6, [1] * 6, [1, 2, 1, 1, 0, 1], [0, 1, 0, 0, 1, 0]], trios=[(0, 1, 2), (0, 1, 3)],
        recomb=[3, 3, 3, 4, 3, 3], cost=8,
        haplotypes=[("111111", "000000"), ("010101", "101010"), ("111000", "010101"), ("000000", "010010")]),
    "trio_genotype_likelihoods": dict(
        reads="""
          A 111
          A 010
          A 110
          B 001
          B 110
          B 101
          C 001
          C 010
          C 010
        """,
        genotypes=[[0, 0, 0]] * 3, gls=[[[0, 0, 0], [0, 0, 1], [5, 0, 5]], [GL0] * 3, [GL0] * 3], distrust=True,
        trios=[(0, 1, 2)], recomb=[10, 10, 10],
        cost=3, constant_tv=True, haplotypes=[("111", "010"), ("001", "110"), ("001", "010")]),
}


def build_pedigree(case):
    pedigree = Pedigree(NumericSampleIds())
    for i, gts in enumerate(case["genotypes"]):
        gls = None
        if "gls" in case:
            gls = [PhredGenotypeLikelihoods(g) for g in case["gls"][i]]
        pedigree.add_individual("individual{}".format(i), canonic_index_list_to_biallelic_gt_list(gts), gls)
    for f, m, c in case["trios"]:
        pedigree.add_relationship("individual{}".format(f), "individual{}".format(m), "individual{}".format(c))
    return pedigree


def assert_trio_allele_order(father, mother, child, transmission_vector, n):
    for pos in range(n):
        tv = transmission_vector[pos]
        paternal = father[not (tv % 2)][pos].allele
        maternal = mother[not (tv // 2)][pos].allele
        assert paternal == child[0][pos].allele
        assert maternal == child[1][pos].allele


def split_transmission_vector(transmission_vector, n_trios):
    per_trio = defaultdict(list)
    for value in transmission_vector:
        for trio in range(n_trios):
            per_trio[trio].append(value % 4)
            value //= 4
    return per_trio


@pytest.mark.parametrize("name", sorted(CASES))
def test_pedigree_known_answers(gpu, name):
    case = CASES[name]
    pedigree = build_pedigree(case)
    rs = string_to_readset_pedigree(case["reads"]) if case["reads"].strip() else ReadSet()
    dp_table = PedigreeDPTable(rs, case["recomb"], pedigree, case.get("distrust", False), case.get("positions"))
    superreads_list, transmission_vector = dp_table.get_super_reads()
    n = len(case["genotypes"][0])
    assert dp_table.get_optimal_cost() == case["cost"]
    if case.get("constant_tv"):
        assert len(set(transmission_vector)) == 1
    if "tv_in" in case:
        assert transmission_vector in case["tv_in"]
    assert len(superreads_list) == len(case["genotypes"])
    for k, (superreads, expected) in enumerate(zip(superreads_list, case["haplotypes"])):
        assert len(superreads) == 2 and len(superreads[0]) == len(superreads[1]) == n
        assert superreads[0].name == "superread_0_{}".format(k) and superreads[1].name == "superread_1_{}".format(k)
        assert superreads[0].mapqs == (-1,) and superreads[0].source_id == -1 and superreads[0].sample_id == k
        got = tuple(sorted("".join(str(v.allele) for v in sr) for sr in superreads))
        assert got == tuple(sorted(expected))
    per_trio = split_transmission_vector(transmission_vector, len(case["trios"]))
    for t, (f, m, c) in enumerate(case["trios"]):
        assert_trio_allele_order(superreads_list[f], superreads_list[m], superreads_list[c], per_trio[t], n)


def test_phase_empty_trio(gpu):
    pedigree = Pedigree(NumericSampleIds())
    for i in range(3):
        pedigree.add_individual("individual{}".format(i), [])
    pedigree.add_relationship("individual0", "individual1", "individual2")
    dp_table = PedigreeDPTable(ReadSet(), [], pedigree)
    (father, mother, child), transmission_vector = dp_table.get_super_reads()
    assert transmission_vector == [] and dp_table.get_optimal_cost() == 0


def test_mendelian_conflict_raises(gpu):
    """Genotypes that no transmission can explain: RuntimeError('Error: Mendelian conflict')
    (src/pedigreedptable.cpp:301-303)."""
    pedigree = Pedigree(NumericSampleIds())
    pedigree.add_individual("individual0", canonic_index_list_to_biallelic_gt_list([0, 1]))
    pedigree.add_individual("individual1", canonic_index_list_to_biallelic_gt_list([0, 1]))
    pedigree.add_individual("individual2", canonic_index_list_to_biallelic_gt_list([2, 1]))
    pedigree.add_relationship("individual0", "individual1", "individual2")
    rs = string_to_readset_pedigree("""
      A 11
      B 01
      C 10
    """)
    with pytest.raises(RuntimeError, match="Mendelian conflict"):
        PedigreeDPTable(rs, [5, 5], pedigree)


def test_unsorted_readset_raises(gpu):
    from whatshap_b200 import Read

    rs = ReadSet()
    for name, start in (("late", 30), ("early", 10)):
        r = Read(name, 50, 0, 0)
        r.add_variant(start, 0, 1)
        r.add_variant(start + 10, 1, 1)
        rs.add(r)
    pedigree = Pedigree(NumericSampleIds())
    pedigree.add_individual("individual0", canonic_index_list_to_biallelic_gt_list([1, 1, 1, 1]))
    with pytest.raises(RuntimeError, match="not sorted"):
        PedigreeDPTable(rs, [1] * 4, pedigree)


def test_unknown_sample_raises(gpu):
    rs = string_to_readset_pedigree("""
      B 11
      B 01
    """)
    pedigree = Pedigree(NumericSampleIds())
    pedigree.add_individual("individual0", canonic_index_list_to_biallelic_gt_list([1, 1]))
    with pytest.raises(RuntimeError, match="not present in pedigree"):
        PedigreeDPTable(rs, [1, 1], pedigree)

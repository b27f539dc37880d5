"""BASELINE.json config 1: tests/test.matrix (and one string matrix) through `PedigreeDPTable`,
checked for self-consistency of cost / partitioning / super-reads.

Restates the reference's tests/test_verification.py:8-43 (helper whatshap/verification.py:4-50).
The 10-read matrix below is the content of the reference's tests/test.matrix."""
import pytest

from whatshap_b200 import NumericSampleIds, Pedigree, PedigreeDPTable, PhredGenotypeLikelihoods
from whatshap_b200.testhelpers import canonic_index_to_biallelic_gt, matrix_to_readset, string_to_readset
from whatshap_b200.verification import verify_mec_score_and_partitioning

pytestmark = pytest.mark.gpu

TEST_MATRIX = """\
1 2 1011
2 3 1001
3 3 011
4 3 011
5 4 0011
6 6 00
7 6 00
8 6 11
9 7 01
10 8 11110
""".splitlines()

STRING_MATRIX = """
      0             0
      110111111111
      00100
           0001000000
           000
            10100
                  101
    """


def verify(rs, all_heterozygous):
    positions = rs.get_positions()
    pedigree = Pedigree(NumericSampleIds())
    gls = [None if all_heterozygous else PhredGenotypeLikelihoods([0, 0, 0])] * len(positions)
    pedigree.add_individual("individual0", [canonic_index_to_biallelic_gt(1) for _ in positions], gls)
    dp_table = PedigreeDPTable(rs, [1] * len(positions), pedigree, distrust_genotypes=not all_heterozygous)
    verify_mec_score_and_partitioning(dp_table, rs)
    return dp_table


@pytest.mark.parametrize("all_heterozygous", [True, False])
def test_string(gpu, all_heterozygous):
    verify(string_to_readset(STRING_MATRIX), all_heterozygous)


@pytest.mark.parametrize("all_heterozygous", [True, False])
def test_matrix(gpu, all_heterozygous):
    verify(matrix_to_readset(TEST_MATRIX), all_heterozygous)

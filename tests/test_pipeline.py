"""The step before the DP feeding the DP, as `whatshap phase` chains them (whatshap/cli/phase.py:528-533,
604-612): `readselection` caps the coverage, `ReadSet.subset` keeps the chosen reads, `PedigreeDPTable`
phases them on the GPU; the result must equal the CPU checker's on the same selected reads."""
import numpy as np
import pytest

from whatshap_b200 import NumericSampleIds, Pedigree, PedigreeDPTable, Read, ReadSet
from whatshap_b200.core import _flatten
from whatshap_b200.readselect import readselection
from whatshap_b200.testhelpers import canonic_index_list_to_biallelic_gt_list


def deep_readset(seed, n_variants=120, n_reads=900):
    rng = np.random.default_rng(seed)
    truth = rng.integers(0, 2, n_variants)
    rs = ReadSet()
    for i in range(n_reads):
        start = int(rng.integers(0, n_variants - 2))
        hap = int(rng.integers(0, 2))
        read = Read("read%04d" % i, 60, 0, 0)
        for v in range(start, min(n_variants, start + 2 + int(rng.geometric(0.25)))):
            allele = int(truth[v] ^ hap ^ (rng.random() < 0.05))
            read.add_variant((v + 1) * 100, allele, int(rng.integers(1, 40)))
        rs.add(read)
    rs.sort()
    return rs


def selected_problem(seed, max_coverage):
    rs = deep_readset(seed)
    chosen = readselection(rs, max_coverage)
    kept = rs.subset(chosen)
    positions = kept.get_positions()
    pedigree = Pedigree(NumericSampleIds())
    pedigree.add_individual("sample", canonic_index_list_to_biallelic_gt_list([1] * len(positions)))
    return rs, kept, positions, pedigree


def test_selection_caps_the_active_reads():
    """CPU half: after the selection no column of the DP sees more than max_coverage active reads."""
    for max_coverage in (5, 15):
        rs, kept, positions, pedigree = selected_problem(11, max_coverage)
        assert 0 < len(kept) < len(rs)
        prob = _flatten(kept, [1] * len(positions), pedigree, False, None)
        active = np.zeros(prob.n_cols + 1, int)
        first = prob.ent_col[prob.read_off[:-1].astype(int)]
        last = prob.ent_col[prob.read_off[1:].astype(int) - 1]
        np.add.at(active, first, 1)
        np.add.at(active, last + 1, -1)
        assert np.cumsum(active).max() <= max_coverage


@pytest.mark.gpu
@pytest.mark.parametrize("max_coverage", [5, 15])
def test_select_then_phase(gpu, checker, max_coverage):
    rs, kept, positions, pedigree = selected_problem(12, max_coverage)
    table = PedigreeDPTable(kept, [1] * len(positions), pedigree)
    want = checker.solve(table._problem)
    assert table.get_optimal_cost() == int(want.cost)
    assert table.get_optimal_partitioning() == want.partition.tolist()
    superreads, transmission_vector = table.get_super_reads()
    assert [v.allele for v in superreads[0][0]] == want.sr_allele[0, 0].tolist()
    assert transmission_vector == [0] * len(positions)

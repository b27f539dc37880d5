"""The solver swap-in for the reference's OWN objects (whatshap_b200/adapters.py), checked in the
authoring container against the real `whatshap.core` extension built out-of-tree from /root/reference:

 * real ReadSet / Pedigree objects flatten to exactly the arrays this package's own containers give;
 * the adapter class (constructor + three methods of whatshap/types.py) returns, as real whatshap.core
   Read/ReadSet objects, exactly what the real PedigreeDPTable returns — with the CPU checker standing
   in for the per-call CUDA solve so that the test needs no GPU.

Skipped where /root/reference is absent (GPU box)."""
import sys

import numpy as np
import pytest

import test_pedigreephasing as tp
import test_phasing as ts
from whatshap_b200 import adapters
from whatshap_b200 import core as mine
from whatshap_b200.testhelpers import string_to_readset, string_to_readset_pedigree


@pytest.fixture(scope="module")
def ref_core():
    from oracle import build_pyref

    path = build_pyref.build()
    if not path:
        pytest.skip("reference tree not available: cannot build whatshap.core")
    sys.path.insert(0, path)
    import whatshap.core as core

    return core


def to_real(core, rs):
    """Copy one of this package's ReadSets into a real whatshap.core.ReadSet."""
    out = core.ReadSet()
    for r in rs:
        read = core.Read(r.name, r.mapqs[0], r.source_id, r.sample_id)
        for v in r:
            read.add_variant(v.position, v.allele, v.quality)
        out.add(read)
    return out


def real_pedigree(core, case, recording):
    cls = adapters.recording_pedigree(core.Pedigree) if recording else core.Pedigree
    ped = cls(core.NumericSampleIds())
    for i, gts in enumerate(case["genotypes"]):
        genotypes = [core.Genotype([0] * (2 - g) + [1] * g) if 0 <= g <= 2 else core.Genotype([]) for g in gts]
        gls = [core.PhredGenotypeLikelihoods(g) for g in case["gls"][i]] if "gls" in case else None
        ped.add_individual("individual{}".format(i), genotypes, gls)
    for f, m, c in case["trios"]:
        ped.add_relationship("individual{}".format(f), "individual{}".format(m), "individual{}".format(c))
    return ped


@pytest.mark.parametrize("recording", [True, False])
@pytest.mark.parametrize("name", sorted(tp.CASES))
def test_adapter_on_real_objects_equals_real_dp_table(ref_core, checker, name, recording):
    case = tp.CASES[name]
    my_rs = string_to_readset_pedigree(case["reads"]) if case["reads"].strip() else mine.ReadSet()
    real_rs = to_real(ref_core, my_rs)
    ped = real_pedigree(ref_core, case, recording)
    distrust, positions = case.get("distrust", False), case.get("positions")
    # 1. same flat arrays as this package's own containers
    flat, ids = adapters.flatten_objects(real_rs, case["recomb"], ped, distrust, positions)
    want = mine._flatten(my_rs, case["recomb"], tp.build_pedigree(case), distrust, positions)
    for field in ("positions", "read_off", "ent_col", "ent_allele", "ent_phred", "read_ind", "recombcost", "trios", "gt"):
        assert np.array_equal(getattr(flat, field), getattr(want, field)), field
    assert (flat.gl is None) == (want.gl is None) and (flat.gl is None or np.array_equal(flat.gl, want.gl))
    # 2. the swap-in class answers like the real one, in real objects
    Table = adapters.make_dp_table_class(ref_core, solver=checker.solve)
    a = Table(real_rs, case["recomb"], ped, distrust, positions)
    recomb = list(case["recomb"]) + [case["recomb"][-1]] * 4  # the real class reads recombcost[k] unchecked
    b = ref_core.PedigreeDPTable(real_rs, recomb, ped, distrust, positions)
    assert a.get_optimal_cost() == b.get_optimal_cost() == case["cost"]
    assert a.get_optimal_partitioning() == b.get_optimal_partitioning()
    (sa, ta), (sb, tb) = a.get_super_reads(), b.get_super_reads()
    assert ta == tb and len(sa) == len(sb)
    for ra, rb in zip(sa, sb):
        assert type(ra) is type(rb) and len(ra) == len(rb) == 2
        for x, y in zip(ra, rb):
            assert (x.name, x.mapqs, x.source_id, x.sample_id) == (y.name, y.mapqs, y.source_id, y.sample_id)
            assert [(v.position, v.allele, v.quality) for v in x] == [(v.position, v.allele, v.quality) for v in y]


def test_single_individual_matrices_through_real_objects(ref_core, checker):
    Table = adapters.make_dp_table_class(ref_core, solver=checker.solve)
    for name in sorted(ts.MATRICES):
        reads, weights = ts.MATRICES[name]
        real_rs = to_real(ref_core, string_to_readset(reads, weights))
        positions = real_rs.get_positions()
        for het in (True, False):
            ped = ref_core.Pedigree(ref_core.NumericSampleIds())
            gls = [None if het else ref_core.PhredGenotypeLikelihoods([0, 0, 0])] * len(positions)
            ped.add_individual("individual0", [ref_core.Genotype([0, 1])] * len(positions), gls)
            a = Table(real_rs, [1] * len(positions), ped, not het)
            b = ref_core.PedigreeDPTable(real_rs, [1] * len(positions), ped, not het)
            assert a.get_optimal_cost() == b.get_optimal_cost()
            assert a.get_optimal_partitioning() == b.get_optimal_partitioning()
            (sa, _), (sb, _) = a.get_super_reads(), b.get_super_reads()
            assert [[(v.position, v.allele, v.quality) for v in r] for r in sa[0]] == [[(v.position, v.allele, v.quality) for v in r] for r in sb[0]]

"""Phase-set (connected component) construction after the DP — the step that consumes the path's
result in `whatshap phase`.

Same contract as the reference's `ComponentFinder` (whatshap/graph.py:35-86: union-find whose
representative is the MINIMUM value of a set) and `find_components` (whatshap/cli/phase.py:71-113:
two variants share a component iff some read covers both; optional master block; optional
restriction to each sample's heterozygous positions)."""
from __future__ import annotations

from typing import Dict, Hashable, Iterable, Mapping, Optional, Sequence, Set


class ComponentFinder:
    """Disjoint sets over arbitrary comparable values; `find(x)` returns the smallest member of x's set."""

    def __init__(self, values: Iterable[Hashable]):
        self._index: Dict[Hashable, int] = {}
        self._values = []
        for v in values:
            if v not in self._index:
                self._index[v] = len(self._values)
                self._values.append(v)
        self._parent = list(range(len(self._values)))

    def _root(self, i: int) -> int:
        root = i
        while self._parent[root] != root:
            root = self._parent[root]
        while self._parent[i] != root:  # path compression
            self._parent[i], i = root, self._parent[i]
        return root

    def merge(self, x, y) -> None:
        assert x != y
        rx, ry = self._root(self._index[x]), self._root(self._index[y])
        if rx == ry:
            return
        # the root holding the smaller value stays root, so find() is always the set's minimum
        if self._values[rx] < self._values[ry]:
            self._parent[ry] = rx
        else:
            self._parent[rx] = ry

    def find(self, value):
        return self._values[self._root(self._index[value])]


def find_components(
    phased_positions: Sequence[int],
    reads,
    master_block: Optional[Sequence[int]] = None,
    heterozygous_positions: Optional[Mapping[int, Set[int]]] = None,
) -> Dict[int, int]:
    """Map every phased variant position to its component, named by the component's leftmost position."""
    assert list(phased_positions) == sorted(phased_positions)
    finder = ComponentFinder(phased_positions)
    phased = set(phased_positions)
    for read in reads:
        allowed = None if heterozygous_positions is None else heterozygous_positions[read.sample_id]
        covered = [v.position for v in read if v.position in phased and (allowed is None or v.position in allowed)]
        for position in covered[1:]:
            finder.merge(covered[0], position)
    if master_block is not None:
        for position in master_block[1:]:
            finder.merge(master_block[0], position)
    return {position: finder.find(position) for position in phased}


def compute_overall_components(
    accessible_positions: Sequence[int],
    all_reads,
    distrust_genotypes: bool,
    family: Sequence[str],
    genetic_haplotyping: bool,
    homozygous_positions: Sequence[int],
    numeric_sample_ids,
    superreads_list,
) -> Dict[int, int]:
    """The step `whatshap phase` runs on the DP's super-reads (whatshap/cli/phase.py:676-714): which positions are
    heterozygous is re-read from the super-reads when genotypes were distrusted (a site the DP called homozygous no
    longer connects anything), and with genetic haplotyping of a family the homozygous sites form one master block."""
    master_block = None
    het_by_sample: Optional[Dict[int, Set[int]]] = None
    accessible = set(accessible_positions)
    if distrust_genotypes:
        hom_anywhere: Set[int] = set()
        het_by_sample = {}
        for sample, (hap0, hap1) in zip(family, superreads_list):
            hets: Set[int] = set()
            for v0, v1 in zip(hap0, hap1):
                assert v0.position == v1.position
                if v0.position not in accessible:
                    continue
                call = (v0.allele, v1.allele)
                if call in ((0, 1), (1, 0)):
                    hets.add(v0.position)
                elif call in ((0, 0), (1, 1)):
                    hom_anywhere.add(v0.position)  # EQUAL_SCORES (3) calls are neither
            het_by_sample[numeric_sample_ids[sample]] = hets
        if len(family) > 1 and genetic_haplotyping:
            master_block = sorted(hom_anywhere)
    elif len(family) > 1 and genetic_haplotyping:
        master_block = sorted(set(homozygous_positions) & accessible)
    return find_components(accessible_positions, all_reads, master_block, het_by_sample)


def phase_calls(family: Sequence[str], superreads_list, overall_components: Mapping[int, int], numeric_sample_ids=None):
    """Hand-off to `PhasedVcfWriter.write(chromosome, sample_superreads, sample_components)`
    (whatshap/cli/phase.py:640-652, whatshap/vcf.py:1147-1188): returns the two dictionaries the writer takes —
    sample -> its two super-reads, sample -> the (shared) component map — and, per sample, the table the writer
    derives from them first: position -> (allele of haplotype 0, allele of haplotype 1) for every site whose two
    calls are 0 / 1 (sites called EQUAL_SCORES = 3 stay unphased), i.e. the GT it writes with PS = component + 1."""
    superreads, components, phases = {}, {}, {}
    for sample, pair in zip(family, superreads_list):
        assert len(pair) == 2
        if numeric_sample_ids is not None:
            assert pair[0].sample_id == pair[1].sample_id == numeric_sample_ids[sample]
        superreads[sample] = pair
        components[sample] = overall_components  # identical for all samples of a family
        table = {}
        for v0, v1 in zip(pair[0], pair[1]):
            if v0.allele in (0, 1) and v1.allele in (0, 1):
                table[v0.position] = (v0.allele, v1.allele)
        phases[sample] = table
    return superreads, components, phases

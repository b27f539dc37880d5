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

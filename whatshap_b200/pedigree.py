"""Recombination costs fed to the pedigree DP and the recombination events read off its result -- the
Python steps on either side of `PedigreeDPTable` in `whatshap phase` (whatshap/cli/phase.py:556-562,
650-674).  Same names and contracts as the reference's whatshap/pedigree.py:24-250 for:

  centimorgen_to_phred, recombination_cost_map, UniformRecombinationCostComputer,
  GeneticMapRecombinationCostComputer (incl. its map-file parser), mendelian_conflict,
  find_recombination, RecombinationEvent, RecombinationMapEntry, ParseError

`recombcost[k]` is the phred-scaled probability of a recombination between columns k-1 and k (what the DP adds per
changed transmission bit, src/pedigreedptable.cpp:283-287).  The floating-point steps keep the reference's
operation order (interpolation formula, math.log10 / math.exp, round-half-even) so that the integer costs agree.
PED-file parsing (PedReader) is I/O and out of scope.
"""
from __future__ import annotations

import logging
import math
from abc import ABC, abstractmethod
from bisect import bisect_left, bisect_right
from dataclasses import dataclass
from pathlib import Path
from typing import List, Mapping, Sequence, Union

import numpy as np

logger = logging.getLogger(__name__)

MINIMUM_GENETIC_DISTANCE: float = 1e-10  # cM


class ParseError(Exception):
    pass


@dataclass
class RecombinationMapEntry:
    position: int
    cum_distance: float


@dataclass(order=True)
class RecombinationEvent:
    position1: int
    position2: int
    transmitted_hap_father1: int
    transmitted_hap_father2: int
    transmitted_hap_mother1: int
    transmitted_hap_mother2: int
    recombination_cost: float


def centimorgen_to_phred(distance: float) -> float:
    """Phred-scaled recombination probability of a genetic distance in cM (Haldane's map function;
    below 1e-10 cM the linear limit, which avoids the cancellation in 1 - exp)."""
    assert distance >= 0
    if distance == 0:
        raise ValueError("Cannot convert genetic distance of zero to phred.")
    if distance < 1e-10:
        return -10.0 * (math.log10(distance) - 2.0)
    p = (1.0 - math.exp(-(2.0 * distance) / 100.0)) / 2.0
    return -10.0 * math.log10(p)


def _cumulative_distances(genetic_map: Sequence[RecombinationMapEntry], positions: Sequence[int]) -> np.ndarray:
    """Genetic distance from the chromosome start to every position: linear between map points, from (0, 0) up
    to the first one, at the map's average rate past the last one."""
    map_pos = np.array([e.position for e in genetic_map], np.int64)
    map_cum = np.array([e.cum_distance for e in genetic_map], np.float64)
    pos = np.asarray(positions, np.int64)
    left = np.searchsorted(map_pos, pos, side="right") - 1   # last map point at or before the position (-1: none)
    right = np.searchsorted(map_pos, pos, side="left")       # first map point at or after it (len: none)
    before, after = left < 0, right >= len(map_pos)
    lo = np.clip(left, 0, len(map_pos) - 1)
    hi = np.clip(right, 0, len(map_pos) - 1)
    start_pos = np.where(before, 0, map_pos[lo])
    start_val = np.where(before, 0.0, map_cum[lo])
    end_pos, end_val = map_pos[hi], map_cum[hi]
    span = end_pos - start_pos
    with np.errstate(divide="ignore", invalid="ignore"):
        inside = start_val + ((pos - start_pos) * (end_val - start_val) / span)
    inside = np.where(span == 0, start_val, inside)  # the position is a map point
    avg_rate = genetic_map[-1].cum_distance / genetic_map[-1].position
    beyond = map_cum[-1] + (pos - map_pos[-1]) * avg_rate
    return np.where(after & ~before, beyond, inside)


def _distances_to_costs(distances) -> List[int]:
    return [0] + [round(centimorgen_to_phred(max(float(d), MINIMUM_GENETIC_DISTANCE))) for d in distances]


def recombination_cost_map(genetic_map: Sequence[RecombinationMapEntry], positions: Sequence[int]) -> List[int]:
    """results[i] = phred-scaled recombination probability between positions[i-1] and positions[i] (results[0] = 0)."""
    assert len(genetic_map) > 0
    if len(positions) == 0:
        return [0]
    cum = _cumulative_distances(genetic_map, positions)
    return _distances_to_costs((cum[1:] - cum[:-1]).tolist())


class RecombinationCostComputer(ABC):
    @abstractmethod
    def compute(self, positions: Sequence[int]) -> Sequence[int]:
        pass


class GeneticMapRecombinationCostComputer(RecombinationCostComputer):
    def __init__(self, genetic_map_path):
        self._genetic_map = self.load_genetic_map(genetic_map_path)

    @staticmethod
    def load_genetic_map(filename: Union[str, Path]) -> List[RecombinationMapEntry]:
        """Three whitespace-separated columns after a header line: position, rate (unused), cumulative cM."""
        entries: List[RecombinationMapEntry] = []
        warned = False
        with open(filename) as f:
            for line_number, line in enumerate(f, 1):
                fields = line.strip().split()
                if line_number == 1 or not fields:
                    continue
                if len(fields) != 3:
                    raise ParseError(f"Error at line {line_number} of genetic map file '{filename}': Found {len(fields)} fields instead of 3")
                try:
                    entry = RecombinationMapEntry(position=int(fields[0]), cum_distance=float(fields[2]))
                except ValueError as e:
                    raise ParseError(f"Error at line {line_number} of genetic map file '{filename}': {e}")
                if entries and not warned and entries[-1].cum_distance == entry.cum_distance:
                    logger.warning("Zero genetic distances encountered in %s", filename)
                    warned = True
                entries.append(entry)
        return entries

    def compute(self, positions: Sequence[int]) -> List[int]:
        return recombination_cost_map(self._genetic_map, positions)


class UniformRecombinationCostComputer(RecombinationCostComputer):
    def __init__(self, recombination_rate: float):
        self._recombination_rate = recombination_rate

    @staticmethod
    def uniform_recombination_map(recombrate: float, positions) -> List[int]:
        """Constant rate in cM/Mb."""
        pos = [int(p) for p in positions]
        return [0] + [round(centimorgen_to_phred((b - a) * 1e-6 * recombrate)) for a, b in zip(pos, pos[1:])]

    def compute(self, positions: Sequence[int]) -> List[int]:
        return self.uniform_recombination_map(self._recombination_rate, positions)


def mendelian_conflict(genotypem, genotypef, genotypec) -> bool:
    """True if the child's two alleles cannot be drawn one from each parent."""
    m, f, c = genotypem.as_vector(), genotypef.as_vector(), genotypec.as_vector()
    return not ((c[0] in m and c[1] in f) or (c[1] in m and c[0] in f))


def find_recombination(transmission_vector: Sequence[int], components: Mapping[int, int], positions: Sequence[int],
                       recombcost: Sequence[int]) -> List[RecombinationEvent]:
    """Changes of the transmission value between consecutive phased variants of a block = recombination events
    (bit 0: haplotype passed on by the father, bit 1: by the mother).  As in the reference, a change between a block's
    first two variants is not reported."""
    assert len(transmission_vector) == len(positions) == len(recombcost)
    index_of = {p: i for i, p in enumerate(positions)}
    assert set(components.keys()).issubset(index_of.keys())
    blocks = {}
    for position, block_id in components.items():
        blocks.setdefault(block_id, []).append(index_of[position])
    events, total = [], 0
    for members in blocks.values():
        members.sort(key=lambda i: positions[i])
        for a, b in zip(members[1:], members[2:]):
            before, after = transmission_vector[a], transmission_vector[b]
            if before != after:
                events.append(RecombinationEvent(positions[a], positions[b], before % 2, after % 2, before // 2, after // 2, recombcost[b]))
                total += recombcost[b]
    logger.info("Cost accounted for by recombination events: %d", total)
    events.sort()
    return events

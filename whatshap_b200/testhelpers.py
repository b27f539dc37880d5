"""Fixture builders used by the parity tests; behaviour follows the reference's
`whatshap/testhelpers.py` (string_to_readset :17-43, string_to_readset_pedigree :46-59,
matrix_to_readset :62-82, brute_force_phase :130-181, canonic_index_to_biallelic_gt :184-205)
so that this repository's tests read like the reference's own."""
import textwrap
from collections import Counter
from itertools import product

from .core import Genotype, Read, ReadSet


def string_to_readset(s, w=None, sample_ids=None, source_id=0, scale_quality=None):
    """One read per line; column c of the text is position (c+1)*10; ' ' = not covered; digits in
    `w` (same shape) are the qualities (default 1)."""
    rows = textwrap.dedent(s).strip().split("\n")
    weights = textwrap.dedent(w).strip().split("\n") if w is not None else None
    rs = ReadSet()
    for index, line in enumerate(rows):
        if not line:
            continue
        sample = 0 if sample_ids is None else sample_ids[index]
        read = Read("Read {}".format(index + 1), 50, source_id, sample)
        for col, ch in enumerate(line):
            if ch == " ":
                continue
            q = int(weights[index][col]) if weights is not None else 1
            if scale_quality is not None:
                q *= scale_quality
            read.add_variant(position=(col + 1) * 10, allele=int(ch), quality=q)
        assert len(read) > 1, "Reads covering less than two variants are not allowed"
        rs.add(read)
    return rs


def string_to_readset_pedigree(s, w=None, scaling_quality=None):
    """Like string_to_readset, but every line starts with a letter naming the individual (A=0, B=1, ...)."""
    sources, body = [], []
    for line in textwrap.dedent(s).strip().split("\n"):
        if not line:
            continue
        individual = ord(line[0]) - ord("A")
        assert 0 <= individual < 26
        sources.append(individual)
        body.append(line[1:])
    return string_to_readset("\n".join(body), w=w, sample_ids=sources, scale_quality=scaling_quality)


def matrix_to_readset(lines):
    """'index offset bits [offset bits ...]' per line; position = (offset + i) * 10, quality 1."""
    rs = ReadSet()
    for expected, line in enumerate(lines, start=1):
        fields = line.split()
        assert len(fields) % 2 == 1, "Not in matrix format."
        assert int(fields[0]) == expected, "Not in matrix format."
        read = Read("Read {}".format(expected), 50)
        for i in range(len(fields) // 2):
            offset = int(fields[2 * i + 1])
            for pos, ch in enumerate(fields[2 * i + 2]):
                read.add_variant(position=(offset + pos) * 10, allele=int(ch), quality=1)
        rs.add(read)
    return rs


def _column_cost(variants, assignments):
    """Cheapest allele assignment of one position given the reads on each side; a haplotype whose
    allele differs among the cheapest assignments is reported as 3 (tie)."""
    costs = []
    for assignment in assignments:
        costs.append(sum(v.quality for side, allele in zip(variants, assignment) for v in side if v.allele != allele))
    best = min(costs)
    winners = [a for a, c in zip(assignments, costs) if c == best]
    order = sorted(range(len(assignments)), key=lambda i: (costs[i], i))
    call = list(assignments[order[0]])
    for h in range(2):
        if len({a[h] for a in winners}) > 1:
            call[h] = 3
    return best, call


def brute_force_phase(read_set, all_heterozygous):
    """Weighted MEC by enumerating all 2^reads bipartitions (< 10 reads).  Returns
    (cost, partition, number of optimal solutions up to inversion, haplotype1, haplotype2)."""
    assert len(read_set) < 10, "Too many reads for brute force"
    positions = read_set.get_positions()
    assignments = [(0, 1), (1, 0)] if all_heterozygous else [(0, 0), (0, 1), (1, 0), (1, 1)]
    per_position = {p: [(n, v) for n, read in enumerate(read_set) for v in read if v.position == p] for p in positions}
    best = None
    count = Counter()
    for partition in range(2 ** len(read_set)):
        cost, haps = 0, []
        for p in positions:
            sides = ([], [])
            for n, v in per_position[p]:
                sides[(partition >> n) & 1].append(v)
            c, call = _column_cost(sides, assignments)
            cost += c
            haps.append(call)
        count[cost] += 1
        if best is None or cost < best[0]:
            best = (cost, partition, haps)
    cost, partition, haps = best
    assert count[cost] % 2 == 0  # every bipartition has an equally good inverse
    return (
        cost,
        [(partition >> x) & 1 for x in range(len(read_set))],
        count[cost] // 2,
        "".join(str(a) for a, _ in haps),
        "".join(str(b) for _, b in haps),
    )


def canonic_index_to_biallelic_gt(num_alt, ploidy=2):
    """VCF index of a biallelic genotype (= number of ALT alleles) -> Genotype; out of range -> empty."""
    if 0 <= num_alt <= ploidy:
        return Genotype([0] * (ploidy - num_alt) + [1] * num_alt)
    return Genotype([])


def canonic_index_list_to_biallelic_gt_list(list_int, ploidy=2):
    return [canonic_index_to_biallelic_gt(i, ploidy) for i in list_int]

"""Post-DP phase-set construction (reference: whatshap/graph.py:35-86, whatshap/cli/phase.py:71-113,
tests/test_graph.py)."""
import random

from whatshap_b200.components import ComponentFinder, find_components
from whatshap_b200.testhelpers import string_to_readset


def test_component_finder_minimum_is_representative():
    cf = ComponentFinder([1, 2, 3, 4, 5, "a", "b"][:5])
    assert [cf.find(i) for i in range(1, 6)] == [1, 2, 3, 4, 5]
    cf.merge(5, 4)
    cf.merge(3, 4)
    assert cf.find(5) == 3 and cf.find(4) == 3 and cf.find(1) == 1
    cf.merge(2, 1)
    cf.merge(5, 1)
    assert {cf.find(i) for i in range(1, 6)} == {1}


def test_find_components_reads_master_block_and_het_filter():
    reads = string_to_readset("""
      11
        00
         11
            01
    """)
    positions = [10, 20, 30, 40, 50, 70, 80]
    assert find_components(positions, reads) == {10: 10, 20: 10, 30: 30, 40: 30, 50: 30, 70: 70, 80: 70}
    merged = find_components(positions, reads, master_block=[20, 70])
    assert merged[80] == 10 and merged[30] == 30
    # restricting to heterozygous positions removes the bridge at 40
    het = {0: {10, 20, 30, 50, 70, 80}}
    assert find_components([10, 20, 30, 50, 70, 80], reads, heterozygous_positions=het)[50] == 50


def test_find_components_equals_graph_search_on_random_reads():
    rng = random.Random(3)
    from whatshap_b200 import Read, ReadSet

    for _ in range(30):
        n = rng.randint(2, 25)
        positions = sorted(rng.sample(range(1, 200), n))
        rs = ReadSet()
        adj = {p: set() for p in positions}
        for r in range(rng.randint(0, 12)):
            cover = sorted(rng.sample(positions, rng.randint(2, min(4, n))))
            read = Read(f"r{r}", 10, 0, 0)
            for p in cover:
                read.add_variant(p, rng.randint(0, 1), 1)
            rs.add(read)
            for p in cover[1:]:
                adj[cover[0]].add(p)
                adj[p].add(cover[0])
        got = find_components(positions, rs)
        for p in positions:  # flood fill
            seen, stack = {p}, [p]
            while stack:
                q = stack.pop()
                for t in adj[q]:
                    if t not in seen:
                        seen.add(t)
                        stack.append(t)
            assert got[p] == min(seen)

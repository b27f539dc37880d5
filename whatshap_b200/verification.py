"""Self-consistency check of a solved DP table (reference: whatshap/verification.py:4-50):
the reported partitioning and super-reads must reproduce the reported MEC score."""


def verify_mec_score_and_partitioning(dp_table, reads):
    superreads_list, _ = dp_table.get_super_reads()
    assert len(superreads_list) == 1
    superreads = superreads_list[0]
    assert len(superreads) == 2 and len(superreads[0]) == len(superreads[1])
    # haplotype alleles with ties (3) resolved arbitrarily but differently on the two haplotypes
    haps = [{}, {}]
    for j in range(2):
        for v in superreads[j]:
            haps[j][v.position] = j if v.allele == 3 else v.allele
    partitioning = dp_table.get_optimal_partitioning()
    mec, swapped, decided = 0, False, 0
    for read_index, read in enumerate(reads):
        cost = [sum(v.quality for v in read if v.position in haps[j] and haps[j][v.position] != v.allele) for j in range(2)]
        mec += min(cost)
        if cost[0] == cost[1]:
            continue
        haplotype = 0 if (cost[0] < cost[1]) != swapped else 1
        if partitioning[read_index] != haplotype:
            assert decided == 0, "partitioning inconsistent with the super-reads"
            swapped = True
        decided += 1
    assert mec == dp_table.get_optimal_cost(), (mec, dp_table.get_optimal_cost())

"""Sharding of DP-independent blocks across the GPUs of one box.

For a single individual (T = 1) the DP decomposes exactly at columns that no read spans: cost adds,
partitioning / super-reads concatenate, tie-breaks are unaffected (a constant offset changes no `<`).
Those blocks are the unit of work: rank 0 broadcasts the flat problem, every rank (one process per
GPU, `torch.distributed` over NCCL) solves its share through the C ABI with no collective on the data
path, and the per-block results are gathered on rank 0.  Pedigrees (T > 1) couple the blocks through
the transmission vector and run on one GPU ("replicas only" this round).

The reference has no counterpart (single process, `whatshap/cli/phase.py:604-610` runs one
PedigreeDPTable per chromosome x family).
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

from ._abi import FlatProblem, FlatSolution


def independent_blocks(prob: FlatProblem) -> List[Tuple[int, int]]:
    """Maximal column ranges [lo, hi) that no read crosses (chains of the DP)."""
    n = prob.n_cols
    if n == 0:
        return []
    span = np.zeros(n + 2, np.int64)
    if prob.n_reads:
        starts = prob.read_off[:-1].astype(np.int64)
        ends = prob.read_off[1:].astype(np.int64) - 1
        first = prob.ent_col[starts].astype(np.int64)
        last = prob.ent_col[ends].astype(np.int64)
        np.add.at(span, first + 1, 1)
        np.add.at(span, last + 1, -1)
    crossing = np.cumsum(span)[1:n]  # crossing[k-1] > 0: some read is active in columns k-1 and k
    cuts = [0] + [int(k) for k in (np.nonzero(crossing == 0)[0] + 1)] + [n]
    return [(cuts[i], cuts[i + 1]) for i in range(len(cuts) - 1)]


def block_work(prob: FlatProblem, blocks: Sequence[Tuple[int, int]]) -> np.ndarray:
    """DP cells per block, sum_k 2^{a_k} (the quantity the sweep time is proportional to)."""
    n = prob.n_cols
    cov = np.zeros(n + 1, np.int64)
    if prob.n_reads:
        starts = prob.read_off[:-1].astype(np.int64)
        ends = prob.read_off[1:].astype(np.int64) - 1
        np.add.at(cov, prob.ent_col[starts].astype(np.int64), 1)
        np.add.at(cov, prob.ent_col[ends].astype(np.int64) + 1, -1)
    a = np.cumsum(cov)[:n]
    cells = np.exp2(np.minimum(a, 40).astype(np.float64))
    return np.array([cells[lo:hi].sum() for lo, hi in blocks])


def assign_blocks(work: np.ndarray, world: int) -> List[List[int]]:
    """Longest-processing-time-first assignment of blocks to ranks."""
    order = np.argsort(-work, kind="stable")
    load = np.zeros(world)
    shares: List[List[int]] = [[] for _ in range(world)]
    for b in order:
        r = int(np.argmin(load))
        shares[r].append(int(b))
        load[r] += work[b]
    for s in shares:
        s.sort()
    return shares


def merge_block_solutions(prob: FlatProblem, blocks, solutions) -> FlatSolution:
    """Concatenate per-block results in column order; costs add (T = 1)."""
    out = FlatSolution(prob.n_cols, prob.n_reads, prob.n_ind)
    first = prob.ent_col[prob.read_off[:-1].astype(np.int64)] if prob.n_reads else np.zeros(0, np.uint32)
    cost = 0
    for (lo, hi), sol in zip(blocks, solutions):
        cost += int(sol.cost)
        out.path_index[lo:hi] = sol.path_index
        out.path_tv[lo:hi] = sol.path_tv
        out.sr_allele[:, :, lo:hi] = sol.sr_allele
        out.sr_quality[:, lo:hi] = sol.sr_quality
        reads = np.nonzero((first >= lo) & (first < hi))[0]
        out.partition[reads] = sol.partition
    out.cost = cost & 0xFFFFFFFF
    return out


def solve_sharded(prob: Optional[FlatProblem], solver: Optional[Callable[[FlatProblem], FlatSolution]] = None,
                  group=None) -> Optional[FlatSolution]:
    """Solve `prob` (given on rank 0; other ranks pass None) on all ranks of `group`.

    Returns the merged solution on rank 0 and None elsewhere.  `solver` defaults to the CUDA path on
    this rank's current device; tests inject a CPU checker to exercise the sharding logic with gloo."""
    import torch.distributed as dist

    if solver is None:
        import torch

        from . import _lib

        device = torch.cuda.current_device()
        solver = lambda p: _lib.solve(p, device=device)[0]
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    box = [prob]
    dist.broadcast_object_list(box, src=0, group=group)  # "trivial broadcast of the block list"
    prob = box[0]
    if prob.n_trios > 0 or prob.n_cols == 0:  # transmission vectors couple the blocks: one GPU
        sol = solver(prob) if rank == 0 else None
        dist.barrier(group)
        return sol
    blocks = independent_blocks(prob)
    shares = assign_blocks(block_work(prob, blocks), world)
    mine = [(b, solver(prob.slice_columns(*blocks[b]))) for b in shares[rank]]
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(mine, gathered, dst=0, group=group)  # per-block super-reads back to rank 0
    if rank != 0:
        return None
    by_block = dict(pair for part in gathered for pair in part)
    return merge_block_solutions(prob, blocks, [by_block[b] for b in range(len(blocks))])

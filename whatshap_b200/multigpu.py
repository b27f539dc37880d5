"""Sharding of DP-independent blocks across the GPUs of one box.

For a single individual (T = 1) the DP decomposes exactly at columns that no read spans: cost adds,
partitioning / super-reads concatenate, tie-breaks are unaffected (a constant offset changes no `<`).
Those blocks are the unit of work: rank 0 scatters to every rank (one process per GPU,
`torch.distributed` over NCCL) the blocks of its share, each rank solves them through the C ABI with
no collective on the data path, and the per-block results are gathered on rank 0.  Pedigrees (T > 1)
couple the blocks through the transmission vector: every rank holds a SEGMENT of the table
(`whmec_segment_*`) and the ranks exchange T x T transfer matrices and exit tables (two all-gathers of a
few hundred bytes), bit-identical to the single-GPU solve.  `genotype_sharded` does the same by chains for the
forward-backward genotyping DP of a single individual.

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


def merge_block_solutions(prob: FlatProblem, blocks, solutions, cost: Optional[int] = None) -> FlatSolution:
    """Concatenate per-block results in column order; costs add (T = 1) unless `cost` is given."""
    out = FlatSolution(prob.n_cols, prob.n_reads, prob.n_ind)
    first = prob.ent_col[prob.read_off[:-1].astype(np.int64)] if prob.n_reads else np.zeros(0, np.uint32)
    total = 0
    for (lo, hi), sol in zip(blocks, solutions):
        total += int(sol.cost)
        out.path_index[lo:hi] = sol.path_index
        out.path_tv[lo:hi] = sol.path_tv
        out.sr_allele[:, :, lo:hi] = sol.sr_allele
        out.sr_quality[:, lo:hi] = sol.sr_quality
        reads = np.nonzero((first >= lo) & (first < hi))[0]
        out.partition[reads] = sol.partition
    out.cost = (total if cost is None else cost) & 0xFFFFFFFF
    return out


UMAX = 0xFFFFFFFF


def contiguous_shares(work: np.ndarray, world: int) -> List[Tuple[int, int]]:
    """Cut the block sequence into `world` contiguous runs [b0, b1) minimising the largest run's work
    (a run may be empty when there are fewer blocks than ranks)."""
    work = np.asarray(work, np.float64)
    n = len(work)

    def runs_for(limit):  # greedy: fewest runs with no run above `limit`
        cuts, acc = [0], 0.0
        for i, w in enumerate(work):
            if acc > 0 and acc + w > limit:
                cuts.append(i)
                acc = 0.0
            acc += w
        cuts.append(n)
        return cuts

    lo, hi = (float(work.max()) if n else 0.0), float(work.sum())
    for _ in range(60):  # bisection on the largest run (the classic linear-partition bound)
        mid = (lo + hi) / 2
        if len(runs_for(mid)) - 1 <= world:
            hi = mid
        else:
            lo = mid
    cuts = runs_for(hi * (1 + 1e-9))  # the bisection ends a rounding error below a feasible limit at worst
    if len(cuts) - 1 > world:
        cuts = cuts[:world] + [n]
    runs = [(cuts[i], cuts[i + 1]) for i in range(len(cuts) - 1) if cuts[i + 1] > cuts[i]]
    prefix = np.concatenate([[0.0], np.cumsum(work)])
    load = lambda run: prefix[run[1]] - prefix[run[0]]
    while len(runs) < world:  # idle ranks left: halve the heaviest run that still has two blocks
        candidates = [r for r in runs if r[1] - r[0] >= 2]
        if not candidates:
            break
        a, b = max(candidates, key=load)
        mid = min(range(a + 1, b), key=lambda m: max(prefix[m] - prefix[a], prefix[b] - prefix[m]))
        runs[runs.index((a, b))] = (a, mid)
        runs.append((mid, b))
        runs.sort()
    runs += [(n, n)] * (world - len(runs))  # more ranks than blocks: the rest hold nothing
    return runs


def minplus(vec: np.ndarray, matrix: np.ndarray) -> np.ndarray:
    """out[i] = min_u vec[u] + matrix[u, i] over u32 with 0xFFFFFFFF as +inf (the fold of
    `ped_prefix_kernel`, whatshap_b200/csrc/whmec.cu)."""
    v = vec.astype(np.uint64)[:, None]
    m = matrix.astype(np.uint64)
    s = v + m
    s[(v == UMAX) | (m == UMAX)] = UMAX
    return s.min(axis=0).astype(np.uint32)


def segment_inputs(matrices: Sequence[Optional[np.ndarray]]) -> List[Optional[np.ndarray]]:
    """True input vector of every segment from the segments' transfer matrices (None: rank without
    columns).  The first segment ignores its input: all rows of its matrix are its output."""
    inputs: List[Optional[np.ndarray]] = []
    vec = None
    for m in matrices:
        if m is None:
            inputs.append(None)
            continue
        inputs.append(vec)
        vec = m[0].copy() if vec is None else minplus(vec, m)
    return inputs


def segment_entries(exits: Sequence[Optional[np.ndarray]]) -> List[Optional[int]]:
    """Transmission value the backtrace enters every segment with, right to left; -1 for the segment
    that ends the table (it starts at the optimum and reports the same exit for every value)."""
    entries: List[Optional[int]] = [None] * len(exits)
    entry = -1
    for s in range(len(exits) - 1, -1, -1):
        if exits[s] is None:
            continue
        entries[s] = entry
        entry = int(exits[s][max(entry, 0)])
    return entries


def _default_segment_factory():
    import torch

    from . import _lib

    device = torch.cuda.current_device()
    return lambda p, continues: _lib.Segment(p, continues, device=device)


def segment_ranges(prob: FlatProblem, world: int) -> List[Optional[Tuple[int, int]]]:
    """Column range [lo, hi) of every rank's segment (None: no columns for that rank)."""
    blocks = independent_blocks(prob)
    shares = contiguous_shares(block_work(prob, blocks), world)
    return [(blocks[b0][0], blocks[b1 - 1][1]) if b1 > b0 else None for b0, b1 in shares]


def solve_pedigree_segments(prob: FlatProblem, n_segments: int, segment_factory=None) -> FlatSolution:
    """The multi-GPU pedigree scheme with all `n_segments` segments driven by ONE process on one
    device, in the order the ranks would run them -- the single-GPU check of `whmec_segment_*`
    (tests/test_gpu_sharded.py) and of this module's folding logic."""
    if segment_factory is None:
        segment_factory = _default_segment_factory()
    ranges = segment_ranges(prob, n_segments)
    segs = [None if r is None else segment_factory(prob.slice_columns(*r), r[0] > 0) for r in ranges]
    try:
        last_active = max(i for i, sg in enumerate(segs) if sg is not None)
        inputs = segment_inputs([sg.transfer() if sg else None for sg in segs])
        for sg, vec in zip(segs, inputs):
            if sg:
                sg.sweep(vec)
        entries = segment_entries([sg.exits(i == last_active) if sg else None for i, sg in enumerate(segs)])
        parts = [(ranges[i], sg.finish(entries[i])) for i, sg in enumerate(segs) if sg]
    finally:
        for sg in segs:
            if sg:
                sg.close()
    return merge_block_solutions(prob, [r for r, _ in parts], [sol for _, sol in parts], cost=int(parts[-1][1].cost))


def solve_pedigree_sharded(prob: Optional[FlatProblem], segment_factory=None, group=None, ranges=None,
                           my_slice: Optional[FlatProblem] = None) -> Tuple[bool, Optional[FlatSolution]]:
    """T > 1.  Every rank holds its segment of the table: either `prob` is known on every rank (each slices its own
    range), or `ranges` (all ranks' column ranges) and `my_slice` (this rank's columns, None for a rank without columns)
    are given and only rank 0 needs `prob` (for merging).  Returns (True, solution) on rank 0 and (True, None)
    elsewhere, or (False, None) on every rank if some segment is outside what the two-pass scheme handles (the
    caller then solves on one GPU).  Input errors (Mendelian conflict, unsorted reads) are raised on every rank."""
    import torch.distributed as dist

    from ._abi import Unsupported

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if segment_factory is None:
        segment_factory = _default_segment_factory()
    if ranges is None:
        ranges = segment_ranges(prob, world)
        my_slice = prob.slice_columns(*ranges[rank]) if ranges[rank] is not None else None
    last_active = max(r for r in range(world) if ranges[r] is not None)

    def everyone(value):
        box = [None] * world
        dist.all_gather_object(box, value, group=group)
        return box

    import os
    import time

    marks = [("start", time.perf_counter())]
    mark = lambda name: marks.append((name, time.perf_counter()))
    seg, status = None, ("ok", "")
    mine = ranges[rank]
    if mine is not None:
        try:
            seg = segment_factory(my_slice, mine[0] > 0)
        except Unsupported as e:
            status = ("unsupported", str(e))
        except RuntimeError as e:
            status = ("error", e)
    try:
        states = everyone(status)
        for kind, payload in states:
            if kind == "error":
                raise payload
        if any(kind == "unsupported" for kind, _ in states):
            return False, None
        mark("create")
        matrix = seg.transfer() if seg else None
        mark("transfer")
        matrices = everyone(matrix)
        mark("gather matrices")
        out_vec = seg.sweep(segment_inputs(matrices)[rank]) if seg else None
        mark("sweep")
        exits = everyone(seg.exits(rank == last_active) if seg else None)
        mark("exits + gather")
        entry = segment_entries(exits)[rank]
        part = (rank, seg.finish(entry), out_vec) if seg else None
        mark("finish")
    finally:
        if seg is not None:
            seg.close()
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(part, gathered, dst=0, group=group)
    mark("gather results")
    if rank == 0 and os.environ.get("WHMEC_TIMING"):
        print("[whmec] pedigree segments, rank 0: " + ", ".join(
            "%s %.1f ms" % (name, (t - t0) * 1e3) for (name, t), (_, t0) in zip(marks[1:], marks[:-1])), flush=True)
    if rank != 0:
        return True, None
    parts = sorted((p for p in gathered if p is not None), key=lambda p: p[0])
    return True, merge_block_solutions(prob, [ranges[r] for r, _, _ in parts], [sol for _, sol, _ in parts],
                                       cost=int(parts[-1][1].cost))


def solve_sharded(prob: Optional[FlatProblem], solver: Optional[Callable[[FlatProblem], FlatSolution]] = None,
                  group=None, segment_factory=None) -> Optional[FlatSolution]:
    """Solve `prob` (given on rank 0; other ranks pass None) on all ranks of `group`.

    Rank 0 cuts the problem and SCATTERS the pieces — a rank receives only the columns and reads it works on (the
    "trivial broadcast of the block list" of the north_star, without shipping every rank the whole ReadSet): the blocks of
    its share for a single individual, its segment of the table for a pedigree.  Returns the merged solution on rank 0
    and None elsewhere.  `solver` / `segment_factory` default to the CUDA path on this rank's current device; tests
    inject CPU stand-ins to exercise the sharding logic with gloo."""
    import torch.distributed as dist

    if solver is None:
        import torch

        from . import _lib

        device = torch.cuda.current_device()
        solver = lambda p: _lib.solve(p, device=device)[0]
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    payloads = None
    blocks = None
    if rank == 0:
        blocks = independent_blocks(prob)
        if prob.n_trios > 0 and world > 1 and len(blocks) > 1:  # transmission vectors couple the blocks: segments of the table
            ranges = segment_ranges(prob, world)
            payloads = [("segments", ranges, prob.slice_columns(*r) if r is not None else None) for r in ranges]
        elif prob.n_trios > 0 or prob.n_cols == 0:  # one chain, one rank, or no columns: one GPU
            payloads = [("single", None, None)] * world
        else:
            shares = assign_blocks(block_work(prob, blocks), world)
            payloads = [("blocks", None, [(b, prob.slice_columns(*blocks[b])) for b in share]) for share in shares]
    box = [None]
    dist.scatter_object_list(box, payloads, src=0, group=group)
    mode, ranges, piece = box[0]
    if mode == "segments":
        handled, sol = solve_pedigree_sharded(prob, segment_factory, group, ranges=ranges, my_slice=piece)
        if handled:
            return sol
        mode = "single"  # some segment is outside the two-pass scheme: the whole table on rank 0's GPU
    if mode == "single":
        sol = solver(prob) if rank == 0 else None
        dist.barrier(group)
        return sol
    mine = [(b, solver(sub)) for b, sub in piece]
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(mine, gathered, dst=0, group=group)  # per-block super-reads back to rank 0
    if rank != 0:
        return None
    by_block = dict(pair for part in gathered for pair in part)
    return merge_block_solutions(prob, blocks, [by_block[b] for b in range(len(blocks))])


def genotype_sharded(prob: Optional[FlatProblem], solver: Optional[Callable[[FlatProblem], np.ndarray]] = None,
                     group=None) -> Optional[np.ndarray]:
    """Genotype likelihoods (`whmec_genotype`, the reference's GenotypeDPTable) of `prob` (given on rank 0; other ranks
    pass None) on all ranks of `group`; returns [n_ind, n_cols, 3] on rank 0, None elsewhere.

    Without transmission values every DP-independent chain is a forward-backward table of its own (a chain boundary hands
    over one number, which cancels in each column's normalisation), so a single individual shards like the phasing DP:
    every rank takes a contiguous run of whole chains balanced by DP cells, no collective on the data path, likelihoods
    gathered on rank 0.  A pedigree is one table and runs on one GPU (replicas only)."""
    import torch.distributed as dist

    if solver is None:
        import torch

        from . import _lib

        device = torch.cuda.current_device()
        solver = lambda p: _lib.genotype(p, device=device)[0]
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    payloads = None
    if rank == 0:
        if prob.n_trios > 0 or prob.n_cols == 0 or world == 1:
            payloads = [("single", 0, 0, None)] * world
        else:  # every rank receives only its run of chains
            blocks = independent_blocks(prob)
            payloads = []
            for b0, b1 in contiguous_shares(block_work(prob, blocks), world):
                lo, hi = (blocks[b0][0], blocks[b1 - 1][1]) if b1 > b0 else (0, 0)
                payloads.append(("chains", lo, hi, prob.slice_columns(lo, hi) if hi > lo else None))
    box = [None]
    dist.scatter_object_list(box, payloads, src=0, group=group)
    mode, lo, hi, piece = box[0]
    if mode == "single":
        out = solver(prob) if rank == 0 else None
        dist.barrier(group)
        return out
    mine = (lo, hi, solver(piece)) if piece is not None else None
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(mine, gathered, dst=0, group=group)
    if rank != 0:
        return None
    out = np.zeros((prob.n_ind, prob.n_cols, 3), np.float64)
    for part in gathered:
        if part is not None:
            lo, hi, lk = part
            out[:, lo:hi, :] = lk
    return out

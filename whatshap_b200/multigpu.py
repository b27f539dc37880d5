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


def _read_spans(prob: FlatProblem) -> Tuple[np.ndarray, np.ndarray]:
    """(first column, last column) of every read, int64."""
    starts = prob.read_off[:-1].astype(np.int64)
    ends = prob.read_off[1:].astype(np.int64) - 1
    return prob.ent_col[starts].astype(np.int64), prob.ent_col[ends].astype(np.int64)


def independent_blocks(prob: FlatProblem, spans=None) -> List[Tuple[int, int]]:
    """Maximal column ranges [lo, hi) that no read crosses (chains of the DP)."""
    n = prob.n_cols
    if n == 0:
        return []
    span = np.zeros(n + 2, np.int64)
    if prob.n_reads:
        first, last = spans if spans is not None else _read_spans(prob)
        span += np.bincount(first + 1, minlength=n + 2) - np.bincount(last + 1, minlength=n + 2)
    crossing = np.cumsum(span)[1:n]  # crossing[k-1] > 0: some read is active in columns k-1 and k
    cuts = [0] + (np.nonzero(crossing == 0)[0] + 1).tolist() + [n]
    return list(zip(cuts[:-1], cuts[1:]))


def block_work(prob: FlatProblem, blocks: Sequence[Tuple[int, int]], spans=None) -> np.ndarray:
    """DP cells per block, sum_k 2^{a_k} (the quantity the sweep time is proportional to)."""
    n = prob.n_cols
    cov = np.zeros(n + 1, np.int64)
    if prob.n_reads:
        first, last = spans if spans is not None else _read_spans(prob)
        cov += np.bincount(first, minlength=n + 1) - np.bincount(last + 1, minlength=n + 1)
    a = np.cumsum(cov)[:n]
    cells = np.exp2(np.minimum(a, 40).astype(np.float64))
    prefix = np.concatenate([[0.0], np.cumsum(cells)])
    if not len(blocks):
        return np.zeros(0)
    bounds = np.asarray(blocks, np.int64)
    return prefix[bounds[:, 1]] - prefix[bounds[:, 0]]


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
    from bisect import bisect_right

    pre = [0.0] + np.cumsum(work).tolist()  # the loads are sums of powers of two: exact in float64 for any realistic table

    def runs_for(limit):  # greedy: fewest runs with no run above `limit` (a block heavier than the limit is a run of its own)
        cuts, s = [0], 0
        while s < n:
            e = max(bisect_right(pre, pre[s] + limit) - 1, s + 1)  # largest e with load(s, e) <= limit, at least one block
            s = min(e, n)
            cuts.append(s)
            if len(cuts) > world + 2:  # already more runs than ranks: the caller only compares with `world`
                cuts[-1] = n
                return cuts
        return cuts

    lo, hi = (float(work.max()) if n else 0.0), float(work.sum())
    for _ in range(48):  # bisection on the largest run (the classic linear-partition bound)
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
                           my_slice: Optional[FlatProblem] = None, comm=None, timings: Optional[dict] = None
                           ) -> Tuple[bool, Optional[FlatSolution]]:
    """T > 1.  Every rank holds its segment of the table: either `prob` is known on every rank (each slices its own
    range), or `ranges` (all ranks' column ranges) and `my_slice` (this rank's columns, None for a rank without columns)
    are given and only rank 0 needs `prob` (for merging).  Returns (True, solution) on rank 0 and (True, None)
    elsewhere, or (False, None) on every rank if some segment is outside what the two-pass scheme handles (the
    caller then solves on one GPU).  Every rank-local phase reports its status in the all-gather that follows it, so an
    error on one rank (Mendelian conflict, unsorted reads, a CUDA failure) is raised on EVERY rank."""
    import os
    import time

    from . import _wire

    comm = comm or _wire.Comm(group)
    rank, world = comm.rank, comm.world
    if segment_factory is None:
        segment_factory = _default_segment_factory()
    if ranges is None:
        ranges = segment_ranges(prob, world)
        my_slice = prob.slice_columns(*ranges[rank]) if ranges[rank] is not None else None
    last_active = max(r for r in range(world) if ranges[r] is not None)
    T = 1 << (2 * (my_slice.n_trios if my_slice is not None else 0))
    marks = [("start", time.perf_counter())]
    mark = lambda name: marks.append((name, time.perf_counter()))
    seg, mine = None, ranges[rank]

    def phase(fn):
        """Run a rank-local step; returns (status, result)."""
        try:
            return (0, ""), fn()
        except Exception as e:  # noqa: BLE001 - reported to every rank below
            return _wire.status_of(e), None

    try:
        status, seg = phase(lambda: segment_factory(my_slice, mine[0] > 0) if mine is not None else None)
        if _wire.raise_first_error(comm.all_status(*status), "create"):
            return False, None
        mark("create")
        status, matrix = phase(lambda: seg.transfer() if seg else None)
        mark("transfer")
        T_all = max(comm.lengths(T))  # ranks without columns do not know T
        width = 4 * T_all * T_all
        states = comm.all_status(*status, payload=None if matrix is None else np.ascontiguousarray(matrix, np.uint32), payload_width=width)
        _wire.raise_first_error(states, "transfer")
        matrices = [None if pl is None else pl.view(np.uint32).reshape(T_all, T_all).copy() for _, _, pl in states]
        mark("gather matrices")
        status, out_vec = phase(lambda: seg.sweep(segment_inputs(matrices)[rank]) if seg else None)
        _wire.raise_first_error(comm.all_status(*status), "sweep")
        mark("sweep")
        status, ex = phase(lambda: seg.exits(rank == last_active) if seg else None)
        states = comm.all_status(*status, payload=None if ex is None else np.ascontiguousarray(ex, np.uint32), payload_width=4 * T_all)
        _wire.raise_first_error(states, "exits")
        exits = [None if pl is None else pl.view(np.uint32).copy() for _, _, pl in states]
        mark("exits + gather")
        entry = segment_entries(exits)[rank]
        status, part = phase(lambda: seg.finish(entry) if seg else None)
        _wire.raise_first_error(comm.all_status(*status), "finish")
        mark("finish")
    finally:
        if seg is not None:
            seg.close()
    rows = comm.gather_rows(_wire.encode_solution(part, tag=rank) if part is not None else np.zeros(0, np.uint8))
    mark("gather results")
    if timings is not None:
        timings.update({name: (t - t0) * 1e3 for (name, t), (_, t0) in zip(marks[1:], marks[:-1])})
    if rank == 0 and os.environ.get("WHMEC_TIMING"):
        print("[whmec] pedigree segments, rank 0: " + ", ".join(
            "%s %.1f ms" % (name, (t - t0) * 1e3) for (name, t), (_, t0) in zip(marks[1:], marks[:-1])), flush=True)
    if rank != 0:
        return True, None
    parts = [(r, _wire.decode_solution(row)[0]) for r, row in enumerate(rows) if row.size]
    return True, merge_block_solutions(prob, [ranges[r] for r, _ in parts], [sol for _, sol in parts], cost=int(parts[-1][1].cost))


def solve_sharded(prob: Optional[FlatProblem], solver: Optional[Callable[[FlatProblem], FlatSolution]] = None,
                  group=None, segment_factory=None, comm=None, timings: Optional[dict] = None) -> Optional[FlatSolution]:
    """Solve `prob` (given on rank 0; other ranks pass None) on all ranks of `group`.

    Rank 0 cuts the problem and SCATTERS the pieces — a rank receives only the columns and reads it works on (the
    "trivial broadcast of the block list" of the north_star, without shipping every rank the whole ReadSet): the blocks of
    its share for a single individual, its segment of the table for a pedigree.  The pieces travel as flat byte rows through
    tensor collectives (`_wire`), the per-block results come back the same way.  Returns the merged solution on rank 0 and
    None elsewhere; an error on any rank is raised on every rank.  `solver` / `segment_factory` default to the CUDA path on
    this rank's current device; tests inject CPU stand-ins to exercise the sharding logic with gloo.  `timings` (optional
    dict) receives this rank's per-phase wall times in ms."""
    import time

    from . import _wire

    if solver is None:
        import torch

        from . import _lib

        device = torch.cuda.current_device()
        solver = lambda p: _lib.solve(p, device=device)[0]
    comm = comm or _wire.Comm(group)
    rank, world = comm.rank, comm.world
    t0 = time.perf_counter()
    rows = None
    blocks = None
    MODE = {"single": 0, "blocks": 1, "segments": 2}
    if rank == 0:
        spans = _read_spans(prob) if prob.n_reads else None
        blocks = independent_blocks(prob, spans)
        if prob.n_trios > 0 and world > 1 and len(blocks) > 1:  # transmission vectors couple the blocks: segments of the table
            ranges = segment_ranges(prob, world)
            flat = np.array([MODE["segments"]] + [x for r in ranges for x in (r if r is not None else (0, 0))], np.int64)
            rows = [[flat.view(np.uint8)] + ([enc] if enc is not None else []) for enc in _wire.encode_problem_slices(prob, ranges)]
        elif prob.n_trios > 0 or prob.n_cols == 0:  # one chain, one rank, or no columns: one GPU
            rows = [[np.array([MODE["single"]], np.int64).view(np.uint8)]] * world
        else:
            # every rank gets ONE sub-problem: a contiguous run of whole blocks with about 1 / world of the DP cells (one
            # whmec_solve per rank sweeps all its chains together; block-by-block calls would be launch-latency bound)
            runs = contiguous_shares(block_work(prob, blocks, spans), world)
            cuts = [(blocks[b0][0], blocks[b1 - 1][1]) if b1 > b0 else None for b0, b1 in runs]
            rows = [[np.array([MODE["blocks"]], np.int64).view(np.uint8)] + ([enc] if enc is not None else [])
                    for enc in _wire.encode_problem_slices(prob, cuts)]
    t1 = time.perf_counter()
    pieces = _wire.separate(comm.scatter_rows(rows))
    head = pieces[0].view(np.int64)
    mode = int(head[0])
    t2 = time.perf_counter()
    if timings is not None:
        timings.update({"cut + encode": (t1 - t0) * 1e3, "scatter": (t2 - t1) * 1e3})
    if mode == MODE["segments"]:
        ranges = [None if head[1 + 2 * r] == head[2 + 2 * r] else (int(head[1 + 2 * r]), int(head[2 + 2 * r])) for r in range(world)]
        piece = _wire.decode_problem(pieces[1])[0] if len(pieces) > 1 else None
        handled, sol = solve_pedigree_sharded(prob, segment_factory, group, ranges=ranges, my_slice=piece, comm=comm, timings=timings)
        if handled:
            return sol
        mode = MODE["single"]  # some segment is outside the two-pass scheme: the whole table on rank 0's GPU
    if mode == MODE["single"]:
        sol, err = None, None
        if rank == 0:
            try:
                sol = solver(prob)
            except Exception as e:  # noqa: BLE001
                err = e
        _wire.raise_first_error(comm.all_status(*_wire.status_of(err)), "solve")
        return sol
    mine, err = [], None
    try:
        for row in pieces[1:]:
            sub, r, lo = _wire.decode_problem(row)
            mine.append(_wire.encode_solution(solver(sub), tag=lo, extra=np.array([sub.n_cols], np.uint32)))
    except Exception as e:  # noqa: BLE001 - e.g. more active reads than supported, CUDA out of memory
        err = e
    t3 = time.perf_counter()
    _wire.raise_first_error(comm.all_status(*_wire.status_of(err)), "solve")
    gathered = comm.gather_rows(_wire.join(mine))  # per-rank super-reads back to rank 0
    t4 = time.perf_counter()
    if timings is not None:
        timings.update({"solve": (t3 - t2) * 1e3, "gather": (t4 - t3) * 1e3})
    if rank != 0:
        return None
    parts = []
    for row in gathered:
        for enc in _wire.separate(row):
            sol, lo, extra = _wire.decode_solution(enc)
            parts.append(((lo, lo + int(extra[0])), sol))
    parts.sort(key=lambda p: p[0][0])
    out = merge_block_solutions(prob, [sp for sp, _ in parts], [sol for _, sol in parts])
    if timings is not None:
        timings["merge"] = (time.perf_counter() - t4) * 1e3
    return out


def genotype_sharded(prob: Optional[FlatProblem], solver: Optional[Callable[[FlatProblem], np.ndarray]] = None,
                     group=None, comm=None) -> Optional[np.ndarray]:
    """Genotype likelihoods (`whmec_genotype`, the reference's GenotypeDPTable) of `prob` (given on rank 0; other ranks
    pass None) on all ranks of `group`; returns [n_ind, n_cols, 3] on rank 0, None elsewhere.

    Without transmission values every DP-independent chain is a forward-backward table of its own (a chain boundary hands
    over one number, which cancels in each column's normalisation), so a single individual shards like the phasing DP:
    every rank takes a contiguous run of whole chains balanced by DP cells, no collective on the data path, likelihoods
    gathered on rank 0.  A pedigree is one table and runs on one GPU (replicas only).  Errors are raised on every rank."""
    from . import _wire

    if solver is None:
        import torch

        from . import _lib

        device = torch.cuda.current_device()
        solver = lambda p: _lib.genotype(p, device=device)[0]
    comm = comm or _wire.Comm(group)
    rank, world = comm.rank, comm.world
    rows = None
    if rank == 0:
        if prob.n_trios > 0 or prob.n_cols == 0 or world == 1:
            rows = [_wire.join([np.array([0, 0, 0], np.int64).view(np.uint8)])] * world
        else:  # every rank receives only its run of chains
            blocks = independent_blocks(prob)
            rows = []
            for b0, b1 in contiguous_shares(block_work(prob, blocks), world):
                lo, hi = (blocks[b0][0], blocks[b1 - 1][1]) if b1 > b0 else (0, 0)
                rows.append(_wire.join([np.array([1, lo, hi], np.int64).view(np.uint8)] +
                                       ([_wire.encode_problem(prob.slice_columns(lo, hi), lo=lo)] if hi > lo else [])))
    pieces = _wire.separate(comm.scatter_rows(rows))
    mode, lo, hi = (int(x) for x in pieces[0].view(np.int64))
    out, err = None, None
    try:
        if mode == 0:
            out = solver(prob) if rank == 0 else None
        elif len(pieces) > 1:
            out = np.ascontiguousarray(solver(_wire.decode_problem(pieces[1])[0]), np.float64)
    except Exception as e:  # noqa: BLE001
        err = e
    _wire.raise_first_error(comm.all_status(*_wire.status_of(err)), "genotype")
    if mode == 0:
        return out
    head = np.array([lo, hi], np.int64).view(np.uint8)
    gathered = comm.gather_rows(_wire.join([head, out.reshape(-1).view(np.uint8)]) if out is not None else np.zeros(0, np.uint8))
    if rank != 0:
        return None
    full = np.zeros((prob.n_ind, prob.n_cols, 3), np.float64)
    for row in gathered:
        parts = _wire.separate(row)
        if parts:
            lo, hi = (int(x) for x in parts[0].view(np.int64))
            full[:, lo:hi, :] = parts[1].view(np.float64).reshape(prob.n_ind, hi - lo, 3)
    return full

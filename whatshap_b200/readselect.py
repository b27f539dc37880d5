"""Coverage-capping read selection, the step that runs right before the DP in `whatshap phase` and
sets the maximum number of active reads per column.

Same contract as the reference's `readselection(readset, max_cov, preferred_source_ids=None,
bridging=True) -> set of read indices` (whatshap/readselect.pyx:240-272; caller
whatshap/cli/phase.py:157-171): reads are taken greedily by the score
(variants not yet covered − gaps, variants − gaps, minimum base quality) until every variant is
covered or `max_cov` physical coverage is reached, slice after slice, with optional "bridging" reads
that connect the blocks of a slice.  It works on this package's `ReadSet` and on the reference's
(only the public container API is used).

The heuristic is sequential and full of ties, and the reference resolves them through three
incidental orders, all of which are reproduced so that the SAME reads are selected:
  * the sift rules of its positional binary heap      -> `whmec_selector_*` (csrc/readselect.cpp)
  * the iteration order of a std::unordered_set<int>   -> the same container, same place
  * the iteration order of CPython `set`s of read indices (which reads enter the heap in which
    order, readselect.pyx:95-105,200,219; in which order scores are lowered, :152-164)
                                                       -> real `set` objects put through the same
                                                          sequence of operations, below
Data is held as CSR numpy arrays; per-variant Python work is avoided except where a `set` must see it.
"""
from __future__ import annotations

import ctypes as C
from typing import Iterable, Optional, Set

import numpy as np

from . import _lib


def _reads_to_csr(readset):
    """(read_off u64, ent_pos i32, ent_quality i32, source_id i64[n]) from any ReadSet-like object."""
    off, pos, qual, src = [0], [], [], []
    for read in readset:
        p = getattr(read, "_pos", None)
        if p is not None:  # this package's Read: parallel lists
            pos.extend(p)
            qual.extend(read._quality)
        else:
            for v in read:
                pos.append(v.position)
                qual.append(v.quality)
        off.append(len(pos))
        src.append(read.source_id)
    return (np.array(off, np.uint64), np.array(pos, np.int32), np.array(qual, np.int32), np.array(src, np.int64))


class _State:
    """Flat data + the native selector (heap, fresh-position container, coverage, blocks)."""

    def __init__(self, read_off, ent_pos, ent_quality, max_cov, bridging):
        self.bridging = bridging
        self.n_reads = n_reads = len(read_off) - 1
        positions, rank = np.unique(ent_pos, return_inverse=True)  # sorted positions ("vcf indices"), rank of every entry
        self.positions = np.ascontiguousarray(positions, np.int32)
        self.ent_rank = np.ascontiguousarray(rank, np.uint32)
        n_variants = len(positions)
        off = read_off.astype(np.int64)
        first, last = rank[off[:-1]], rank[off[1:] - 1]
        # initial scores (readselect.pyx:55-92): good = variants of the read, bad = ranks it spans without covering
        good = np.diff(off)
        bad = (last - first + 1) - good
        minq = np.minimum.reduceat(ent_quality, off[:-1])
        self.scores = np.ascontiguousarray(np.stack([good - bad, good - bad, minq], axis=1), np.int32)
        # reads of every variant, ascending read index (readselect.pyx:33-35), as Python lists for set.update
        order = np.argsort(rank, kind="stable")
        var_reads = np.repeat(np.arange(n_reads), good)[order]
        var_off = np.concatenate([[0], np.cumsum(np.bincount(rank, minlength=n_variants))])
        self._var_reads, self._var_off = var_reads, var_off
        self._var_lists = {}
        self.read_off, self.ent_pos = read_off, ent_pos
        L = self._lib = _lib.lib()
        u32, i32 = C.POINTER(C.c_uint32), C.POINTER(C.c_int32)
        self._h = C.c_void_p(L.whmec_selector_create(n_reads, n_variants, read_off.ctypes.data_as(C.POINTER(C.c_uint64)),
                                                     ent_pos.ctypes.data_as(i32), self.ent_rank.ctypes.data_as(u32),
                                                     self.positions.ctypes.data_as(i32), int(max_cov)))
        self._scores_p = self.scores.ctypes.data_as(i32)
        self._fresh = np.zeros(max(1, int(good.max())), np.uint32)
        self._over = np.zeros(n_reads, np.uint32)
        self._removed = np.zeros(n_reads, np.uint32)
        self._taken = np.zeros(n_reads, np.uint8)
        self._fresh_p, self._over_p, self._removed_p = (a.ctypes.data_as(u32) for a in (self._fresh, self._over, self._removed))
        self._taken_p = self._taken.ctypes.data_as(C.POINTER(C.c_uint8))
        self._u32 = u32

    def reads_of_variant(self, v):
        lst = self._var_lists.get(v)
        if lst is None:
            lst = self._var_lists[v] = self._var_reads[self._var_off[v]:self._var_off[v + 1]].tolist()
        return lst

    def _order_of(self, a_set):
        arr = np.fromiter(a_set, np.uint32, len(a_set))  # the set's iteration order
        return arr, arr.ctypes.data_as(self._u32)

    def slice(self, undecided: Set[int]):
        """One pass over a fresh queue of the undecided reads: every variant gets covered once where reads and
        coverage allow (readselect.pyx:108-166).  Returns (reads taken, reads whose span is already full)."""
        L, h = self._lib, self._h
        arr, ptr = self._order_of(undecided)
        L.whmec_selector_begin_slice(h, ptr, self._scores_p, len(arr))
        in_slice, over_coverage = set(), set()
        read, n_fresh, n_over = C.c_uint32(0), C.c_uint32(0), C.c_uint32(0)
        while True:
            more = L.whmec_selector_next(h, C.byref(read), self._fresh_p, C.byref(n_fresh), self._over_p, C.byref(n_over))
            if n_over.value:
                over_coverage.update(self._over[:n_over.value].tolist())
            if not more:
                break
            in_slice.add(read.value)
            # every still-queued read sharing a freshly covered variant gets its score lowered -- in the
            # iteration order of a set filled variant by variant (readselect.pyx:150-164)
            to_update = set()
            for v in self._fresh[:n_fresh.value].tolist():
                to_update.update(self.reads_of_variant(v))
            lowered = to_update.difference(in_slice)
            arr, ptr = self._order_of(lowered)
            L.whmec_selector_rescore(h, ptr, len(arr))
        return in_slice, over_coverage

    def rounds(self, selected: Set[int], undecided: Set[int]):
        """Slices, each followed by reads that bridge the slice's blocks, until no read is undecided
        (readselect.pyx:183-237).  Mutates both sets like the reference does."""
        L, h = self._lib, self._h
        while len(undecided) > 0:
            in_slice, over_coverage = self.slice(undecided)
            selected.update(in_slice)
            undecided -= in_slice
            undecided -= over_coverage
            if not self.bridging:
                continue
            arr, ptr = self._order_of(undecided)
            taken_now, taken_ptr = self._order_of(in_slice)
            n = L.whmec_selector_bridge(h, ptr, self._scores_p, len(arr), taken_ptr, len(taken_now), self._removed_p, self._taken_p)
            for r, took in zip(self._removed[:n].tolist(), self._taken[:n].tolist()):
                undecided.remove(r)
                if took:
                    selected.add(r)
        return selected

    def close(self):
        if self._h:
            self._lib.whmec_selector_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()


def select_reads_csr(read_off, ent_pos, ent_quality, source_ids, max_cov, preferred_source_ids=None, bridging=True) -> Set[int]:
    """The selection on flat arrays: reads as CSR over (position, quality) entries in ReadSet order."""
    read_off = np.ascontiguousarray(read_off, np.uint64)
    ent_pos = np.ascontiguousarray(ent_pos, np.int32)
    ent_quality = np.ascontiguousarray(ent_quality, np.int32)
    n_reads = len(read_off) - 1
    if n_reads and int(np.diff(read_off.astype(np.int64)).min()) < 2:
        raise ValueError("readselection expects reads that cover at least two variants")
    preferred = set()
    if preferred_source_ids is not None:
        for index, source in enumerate(np.asarray(source_ids).tolist()):
            if source in preferred_source_ids:
                preferred.add(index)
    selected: Set[int] = set()
    undecided = set(range(n_reads))
    if n_reads == 0:
        return selected
    state = _State(read_off, ent_pos, ent_quality, max_cov, bridging)
    try:
        if len(preferred) > 0:
            chosen = state.rounds(selected, preferred)  # consumes `preferred`, like the reference
            selected.update(chosen)
            undecided -= preferred
        return state.rounds(selected, undecided)
    finally:
        state.close()


def readselection(readset, max_cov, preferred_source_ids=None, bridging=True) -> Set[int]:
    """Indices of the reads to keep so that no variant is covered by more than `max_cov` of them."""
    read_off, ent_pos, ent_quality, source_ids = _reads_to_csr(readset)
    return select_reads_csr(read_off, ent_pos, ent_quality, source_ids, max_cov, preferred_source_ids, bridging)

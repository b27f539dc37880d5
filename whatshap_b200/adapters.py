"""Running the CUDA solver on the reference's OWN objects.

Inside a real `whatshap phase` run the ReadSet / Pedigree are the Cython objects of `whatshap.core`
(`readselect.pyx:244` reaches into `ReadSet.thisptr`, so they cannot be replaced); only the solver is
swapped by patching the name imported at `whatshap/cli/phase.py:34-42`.  This module provides that
swap-in class.  It touches the foreign objects only through their public Python API
(`whatshap/core.pyx:62-361,419-466`), so it works for this package's containers as well.

    import whatshap.core, whatshap.cli.phase, whatshap_b200.adapters as a
    whatshap.cli.phase.Pedigree = a.recording_pedigree(whatshap.core.Pedigree)
    whatshap.cli.phase.PedigreeDPTable = a.make_dp_table_class(whatshap.core)

The real `Pedigree` has no getter for its individuals or trio relationships; `recording_pedigree`
returns a subclass that remembers what it was given (cdef classes are subclassable from Python).
Without it the adapter falls back to parsing `str(pedigree)` (src/pedigree.cpp:91-123), which prints
genotype likelihoods with 6 significant digits — exact for the integer phred GLs WhatsHap produces
(`whatshap/vcf.py:269-285`), lossy for arbitrary doubles.
"""
from __future__ import annotations

import re
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from ._abi import GT_OTHER, FlatProblem


def recording_pedigree(base):
    """Subclass of a `Pedigree` class that records individuals (name, numeric id, genotypes, GLs) and
    trio relationships as they are added."""

    class RecordingPedigree(base):
        def __init__(self, numeric_sample_ids):
            try:
                super().__init__(numeric_sample_ids)
            except TypeError:  # cdef class: __cinit__ already consumed the argument
                pass
            self._rec_ids = numeric_sample_ids
            self._rec_individuals = []  # (numeric id, genotypes, genotype likelihoods or None)
            self._rec_trios = []        # (father numeric id, mother numeric id, child numeric id)

        def add_individual(self, id, genotypes, genotype_likelihoods=None):
            genotypes = list(genotypes)
            gls = list(genotype_likelihoods) if genotype_likelihoods else None
            super().add_individual(id, genotypes, gls)
            self._rec_individuals.append((self._rec_ids[id], genotypes, gls))

        def add_relationship(self, father_id, mother_id, child_id):
            super().add_relationship(father_id, mother_id, child_id)
            self._rec_trios.append((self._rec_ids[father_id], self._rec_ids[mother_id], self._rec_ids[child_id]))

    RecordingPedigree.__name__ = "Recording" + base.__name__
    return RecordingPedigree


def _pedigree_from_recording(pedigree):
    ids = [numeric for numeric, _, _ in pedigree._rec_individuals]
    index = {numeric: i for i, numeric in enumerate(ids)}
    trios = [(index[f], index[m], index[c]) for f, m, c in pedigree._rec_trios]
    gts, gls = [], []
    for _, genotypes, likelihoods in pedigree._rec_individuals:
        gts.append([int(g.get_index()) if g.is_diploid_and_biallelic() else GT_OTHER for g in genotypes])
        if likelihoods is None:
            gls.append([None] * len(genotypes))
        else:
            gls.append([None if gl is None else [float(x) for x in list(gl)[:3]] for gl in likelihoods])
    return ids, trios, gts, gls


_GT_RE = re.compile(r"(\S+) \(GL:(None\)|PhredGenotypeLikelihoods\(([^\s)]*))")


def _pedigree_from_str(pedigree):
    text = str(pedigree)
    head = re.search(r"individuals \(index,id\):(.*)", text).group(1)
    ids = [int(tok.split(",")[1]) for tok in head.split()]
    trio_line = re.search(r"triples by index \(father,mother,child\):(.*)", text).group(1)
    trios = [tuple(int(x) for x in t.split(",")) for t in re.findall(r"\(([^)]*)\)", trio_line)]
    gts, gls = [], []
    blocks = re.split(r"individual index:\d+ / id:\d+:", text)[1:]
    for block in blocks:
        g_row, l_row = [], []
        for gt, none_or_gl, values in _GT_RE.findall(block):
            alleles = [] if gt == "." else [int(a) for a in gt.split("/")]
            ok = len(alleles) == 2 and all(a <= 1 for a in alleles)
            g_row.append(sum(alleles) if ok else GT_OTHER)
            l_row.append(None if none_or_gl.startswith("None") else [float(x) for x in values.split(",")][:3])
        gts.append(g_row)
        gls.append(l_row)
    return ids, trios, gts, gls


def flatten_objects(readset, recombcost, pedigree, distrust_genotypes: bool = False, positions=None) -> Tuple[FlatProblem, List[int]]:
    """ReadSet / Pedigree objects of `whatshap.core` (or of this package) -> (`whmec_problem` arrays,
    numeric sample id of every pedigree index)."""
    if hasattr(pedigree, "_rec_individuals"):
        ids, trios, gts, gls = _pedigree_from_recording(pedigree)
    else:
        ids, trios, gts, gls = _pedigree_from_str(pedigree)
    index_of = {numeric: i for i, numeric in enumerate(ids)}
    pos_list = list(readset.get_positions()) if positions is None else [int(p) for p in positions]
    col_of = {p: i for i, p in enumerate(pos_list)}
    n = len(pos_list)
    read_off, ent_col, ent_allele, ent_phred, read_ind = [0], [], [], [], []
    prev_first = None
    for read in readset:
        if read.sample_id not in index_of:
            raise RuntimeError("Individual with ID {} not present in pedigree.".format(read.sample_id))
        read_ind.append(index_of[read.sample_id])
        variants = list(read)
        if not variants:
            raise RuntimeError("No variants present")
        pos = [v.position for v in variants]
        if prev_first is not None and pos[0] < prev_first:
            raise RuntimeError("ColumnIterator: reads in ReadSet are not sorted.")
        if any(b <= a for a, b in zip(pos, pos[1:])):
            raise RuntimeError("ColumnIterator: encountered read with unsorted variants.")
        prev_first = pos[0]
        if pos[0] not in col_of or pos[-1] not in col_of:
            raise RuntimeError("read {!r}: first/last variant position is not among the given positions".format(read.name))
        for v in variants:
            c = col_of.get(v.position)
            if c is None:
                continue
            if v.allele not in (0, 1, 2):
                raise RuntimeError("read {!r}: allele {} is not 0, 1 or 2".format(read.name, v.allele))
            ent_col.append(c)
            ent_allele.append(v.allele)
            ent_phred.append(v.quality)
        read_off.append(len(ent_col))
    n_ind = len(ids)
    rc = [int(x) for x in recombcost]
    if len(rc) < n:
        rc = rc + [rc[-1] if rc else 0] * (n - len(rc))
    gt = np.full((n_ind, n), GT_OTHER, np.uint8)
    gl = np.zeros((n_ind, n, 3), np.float64) if distrust_genotypes else None
    for i in range(n_ind):
        if n and len(gts[i]) < n:
            raise RuntimeError("pedigree holds genotypes for {} variants but the DP has {} columns".format(len(gts[i]), n))
        gt[i, :] = gts[i][:n]
        if distrust_genotypes:
            for k in range(n):
                if gls[i][k] is None:
                    raise RuntimeError("distrust_genotypes requires genotype likelihoods for every variant")
                gl[i, k, :] = gls[i][k]
    return FlatProblem(
        positions=np.array(pos_list, np.uint32), read_off=np.array(read_off, np.uint64), ent_col=np.array(ent_col, np.uint32),
        ent_allele=np.array(ent_allele, np.uint8), ent_phred=np.array(ent_phred, np.uint32), read_ind=np.array(read_ind, np.uint32),
        recombcost=np.array(rc[:n], np.uint32), n_ind=n_ind, trios=np.array([x for t in trios for x in t], np.uint32),
        distrust=bool(distrust_genotypes), gt=gt, gl=gl,
    ), ids


def make_dp_table_class(core_module, solver=None):
    """A `PedigreeDPTable` look-alike (same constructor, same three methods, `whatshap/types.py:7-15`)
    that solves on the GPU and answers with `core_module`'s own Read / ReadSet objects.
    `solver(problem) -> FlatSolution` defaults to the CUDA path; tests may inject a CPU checker."""
    Read, ReadSet = core_module.Read, core_module.ReadSet

    class PedigreeDPTable:
        def __init__(self, readset, recombcost, pedigree, distrust_genotypes=False, positions=None):
            self.pedigree = pedigree
            self._problem, self._ids = flatten_objects(readset, recombcost, pedigree, distrust_genotypes, positions)
            solve = solver or (lambda p: _lib.solve(p)[0])
            self._solution = solve(self._problem)

        def get_super_reads(self):
            prob, sol = self._problem, self._solution
            positions = prob.positions.tolist()
            results = []
            for k, numeric in enumerate(self._ids):
                rs = ReadSet()
                for h in range(2):
                    read = Read("superread_{}_{}".format(h, k), -1, -1, numeric)
                    for p, a, q in zip(positions, sol.sr_allele[k, h].tolist(), sol.sr_quality[k].tolist()):
                        read.add_variant(p, a, q)
                    rs.add(read)
                results.append(rs)
            return results, sol.path_tv.tolist()

        def get_optimal_cost(self):
            c = int(self._solution.cost)
            return c - (1 << 32) if c >= (1 << 31) else c

        def get_optimal_partitioning(self):
            return self._solution.partition.tolist()

    return PedigreeDPTable

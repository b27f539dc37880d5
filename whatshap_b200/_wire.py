"""Flat byte encoding of problems / solutions and the tensor collectives that move them between the ranks of one box.

`torch.distributed`'s object collectives pickle every payload and, on NCCL, run a size exchange plus two host<->device
copies per object: for a 9 MB problem cut eight ways that was the whole cost of the sharded solve (round 1: 352 ms on
8 ranks against a 40 ms single-GPU solve).  Here every payload is ONE contiguous uint8 buffer (a small header followed by
the raw arrays of `whmec_problem` / `whmec_solution`, include/whmec.h) and travels through tensor collectives only:
`scatter` / `gather` of equal-size padded rows after one exchange of the row lengths, and `all_gather` for the few hundred
bytes the pedigree scheme shares (T x T transfer matrices, exit tables, per-rank status).  On NCCL the rows go pinned host
-> device -> NVLink -> device -> host; on gloo (CPU tests) they stay on the host.

Every rank-local phase reports (ok | error, text) through `all_status` before the next collective, so that a failure on one
rank is raised on every rank instead of leaving the others in a collective until the communicator times out."""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np

from ._abi import FlatProblem, FlatSolution

_MAGIC_PROBLEM = 0x50484D57   # "WMHP"
_MAGIC_SOLUTION = 0x53484D57  # "WMHS"
_ALIGN = 16


def _pad(n: int) -> int:
    return (n + _ALIGN - 1) // _ALIGN * _ALIGN


def _concat(header: Sequence[int], arrays: Sequence[np.ndarray]) -> np.ndarray:
    head = np.array(list(header) + [a.nbytes for a in arrays], np.uint64)
    total = _pad(head.nbytes) + sum(_pad(a.nbytes) for a in arrays)
    out = np.zeros(total, np.uint8)
    out[: head.nbytes] = head.view(np.uint8)
    off = _pad(head.nbytes)
    for a in arrays:
        out[off : off + a.nbytes] = np.ascontiguousarray(a).reshape(-1).view(np.uint8)
        off += _pad(a.nbytes)
    return out


def _split(buf: np.ndarray, n_header: int, dtypes: Sequence) -> Tuple[List[int], List[np.ndarray]]:
    head = buf[: 8 * (n_header + len(dtypes))].view(np.uint64)
    header = [int(x) for x in head[:n_header]]
    off = _pad(8 * (n_header + len(dtypes)))
    arrays = []
    for i, dt in enumerate(dtypes):
        nbytes = int(head[n_header + i])
        arrays.append(buf[off : off + nbytes].view(dt).copy())
        off += _pad(nbytes)
    return header, arrays


def encode_problem(p: FlatProblem, tag: int = 0, lo: int = 0) -> np.ndarray:
    """`tag` and `lo` are free header words (block id / first column of the slice)."""
    gl = p.gl if p.gl is not None else np.zeros(0, np.float64)
    return _concat(
        [_MAGIC_PROBLEM, p.n_cols, p.n_reads, p.n_ind, 1 if p.distrust else 0, 1 if p.gl is not None else 0, tag, lo],
        [p.positions, p.read_off, p.ent_col, p.ent_allele, p.ent_phred, p.read_ind, p.recombcost, p.trios, p.gt, gl],
    )


class Lazy:
    """A payload whose size is known before its bytes exist: `write(view)` fills a uint8 view of `nbytes` bytes.  `Comm.scatter_rows`
    lets such a payload write itself straight into the page-locked send buffer (one pass over the data on rank 0 instead of two)."""

    def __init__(self, nbytes: int, write):
        self.nbytes = int(nbytes)
        self.write = write

    def materialize(self) -> np.ndarray:
        out = np.empty(self.nbytes, np.uint8)
        self.write(out)
        return out


def encode_problem_slices(p: FlatProblem, spans: Sequence[Optional[Tuple[int, int]]]) -> List[Optional[Lazy]]:
    """`encode_problem(p.slice_columns(lo, hi), tag=i, lo=lo)` for every span (lo, hi) -- no read may cross a span's ends --
    without materialising the slices: read / entry ranges from one pass over the reads' first columns; every row is returned as a
    `Lazy` payload that writes each array once, straight into wherever the row is wanted."""
    m = p.n_reads
    first = p.ent_col[p.read_off[:-1].astype(np.int64)] if m else np.zeros(0, np.uint32)
    gl_all = p.gl
    rows: List[Optional[Lazy]] = []
    for i, sp in enumerate(spans):
        if sp is None:
            rows.append(None)
            continue
        lo, hi = sp
        r0, r1 = (int(x) for x in np.searchsorted(first, [lo, hi], side="left"))  # reads are sorted by first column
        e0, e1 = int(p.read_off[r0]), int(p.read_off[r1])
        n = hi - lo
        if e1 > e0 and int(p.ent_col[e0:e1].max()) >= hi:
            raise ValueError("a read crosses the requested cut")
        sizes = [4 * n, 8 * (r1 - r0 + 1), 4 * (e1 - e0), e1 - e0, 4 * (e1 - e0), 4 * (r1 - r0), 4 * n, p.trios.nbytes, p.n_ind * n,
                 (p.n_ind * n * 3 * 8) if gl_all is not None else 0]
        head = np.array([_MAGIC_PROBLEM, n, r1 - r0, p.n_ind, 1 if p.distrust else 0, 1 if gl_all is not None else 0, i, lo] + sizes, np.uint64)
        total = _pad(head.nbytes) + sum(_pad(x) for x in sizes)

        def write(out, lo=lo, hi=hi, r0=r0, r1=r1, e0=e0, e1=e1, head=head, sizes=sizes):
            out[: head.nbytes] = head.view(np.uint8)
            out[head.nbytes : _pad(head.nbytes)] = 0
            off = _pad(head.nbytes)

            def put(arr, dtype, nbytes):
                nonlocal off
                if nbytes:
                    out[off : off + nbytes].view(dtype)[:] = arr.reshape(-1)
                out[off + nbytes : off + _pad(nbytes)] = 0
                off += _pad(nbytes)

            put(p.positions[lo:hi], np.uint32, sizes[0])
            ro = out[off : off + sizes[1]].view(np.uint64)
            np.subtract(p.read_off[r0 : r1 + 1], p.read_off[r0], out=ro)
            out[off + sizes[1] : off + _pad(sizes[1])] = 0
            off += _pad(sizes[1])
            ec = out[off : off + sizes[2]].view(np.uint32)
            np.subtract(p.ent_col[e0:e1], np.uint32(lo), out=ec)
            out[off + sizes[2] : off + _pad(sizes[2])] = 0
            off += _pad(sizes[2])
            put(p.ent_allele[e0:e1], np.uint8, sizes[3])
            put(p.ent_phred[e0:e1], np.uint32, sizes[4])
            put(p.read_ind[r0:r1], np.uint32, sizes[5])
            put(p.recombcost[lo:hi], np.uint32, sizes[6])
            put(p.trios, np.uint32, sizes[7])
            put(np.ascontiguousarray(p.gt[:, lo:hi]), np.uint8, sizes[8])
            if gl_all is not None:
                put(np.ascontiguousarray(gl_all[:, lo:hi]), np.float64, sizes[9])
            else:
                put(np.zeros(0, np.float64), np.float64, 0)

        rows.append(Lazy(total, write))
    return rows


def decode_problem(buf: np.ndarray) -> Tuple[FlatProblem, int, int]:
    h, a = _split(buf, 8, [np.uint32, np.uint64, np.uint32, np.uint8, np.uint32, np.uint32, np.uint32, np.uint32, np.uint8, np.float64])
    assert h[0] == _MAGIC_PROBLEM, "not an encoded problem"
    n_cols, n_ind = h[1], h[3]
    prob = FlatProblem(positions=a[0], read_off=a[1], ent_col=a[2], ent_allele=a[3], ent_phred=a[4], read_ind=a[5], recombcost=a[6],
                       n_ind=n_ind, trios=a[7], distrust=bool(h[4]), gt=a[8].reshape(n_ind, n_cols), gl=a[9] if h[5] else None)
    return prob, h[6], h[7]


def encode_solution(s: FlatSolution, tag: int = 0, extra: Optional[np.ndarray] = None) -> np.ndarray:
    extra = np.zeros(0, np.uint32) if extra is None else np.asarray(extra, np.uint32)
    return _concat([_MAGIC_SOLUTION, s.n_cols, s.n_reads, s.n_ind, int(s.cost), tag],
                   [s.path_index, s.path_tv, s.partition, s.sr_allele, s.sr_quality, extra])


def decode_solution(buf: np.ndarray) -> Tuple[FlatSolution, int, np.ndarray]:
    h, a = _split(buf, 6, [np.uint32, np.uint32, np.uint8, np.uint8, np.uint32, np.uint32])
    assert h[0] == _MAGIC_SOLUTION, "not an encoded solution"
    s = FlatSolution(h[1], h[2], h[3])
    s.cost = h[4]
    s.path_index, s.path_tv, s.partition = a[0], a[1], a[2]
    s.sr_allele, s.sr_quality = a[3].reshape(h[3], 2, h[1]), a[4].reshape(h[3], h[1])
    return s, h[5], a[5]


def join(buffers: Sequence[np.ndarray]) -> np.ndarray:
    """Several encoded payloads in one row: [count, len_0, len_1, ...] then the payloads (each already 16-byte padded)."""
    head = np.array([len(buffers)] + [b.nbytes for b in buffers], np.uint64)
    out = np.zeros(_pad(head.nbytes) + sum(_pad(b.nbytes) for b in buffers), np.uint8)
    out[: head.nbytes] = head.view(np.uint8)
    off = _pad(head.nbytes)
    for b in buffers:
        out[off : off + b.nbytes] = b
        off += _pad(b.nbytes)
    return out


def separate(row: np.ndarray) -> List[np.ndarray]:
    if row.size == 0:
        return []
    count = int(row[:8].view(np.uint64)[0])
    lens = row[8 : 8 * (count + 1)].view(np.uint64)
    off = _pad(8 * (count + 1))
    out = []
    for n in lens:
        out.append(row[off : off + int(n)])
        off += _pad(int(n))
    return out


# ---- collectives ------------------------------------------------------------------------------------------------


class Comm:
    """The ranks of one process group with byte-row collectives.  Tensors live on this rank's CUDA device for NCCL and on
    the host for gloo."""

    def __init__(self, group=None):
        import torch
        import torch.distributed as dist

        self.torch, self.dist, self.group = torch, dist, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.cuda = dist.get_backend(group) == "nccl"
        self.device = torch.device("cuda", torch.cuda.current_device()) if self.cuda else torch.device("cpu")
        self.root = dist.get_global_rank(group, 0) if group is not None else 0

        self._pinned = {}  # page-locked staging buffers, kept for the life of the communicator (cudaHostAlloc costs milliseconds)

    def _staging(self, key: str, nbytes: int):
        """A page-locked uint8 tensor of at least `nbytes` (NCCL) or a plain one (gloo), reused between calls."""
        t = self._pinned.get(key)
        if t is None or t.numel() < nbytes:
            t = self.torch.empty(max(nbytes, 1 << 16), dtype=self.torch.uint8, pin_memory=self.cuda)
            self._pinned[key] = t
        return t[:nbytes]

    def _to_device(self, a: np.ndarray):
        t = self.torch.from_numpy(a)
        if self.cuda:
            t = t.pin_memory().to(self.device, non_blocking=True)
        return t

    def _to_host(self, t) -> np.ndarray:
        return t.cpu().numpy() if self.cuda else t.numpy()

    def lengths(self, mine: int) -> List[int]:
        t = self.torch.tensor([mine], dtype=self.torch.int64, device=self.device)
        box = [self.torch.zeros(1, dtype=self.torch.int64, device=self.device) for _ in range(self.world)]
        self.dist.all_gather(box, t, group=self.group)
        return [int(x.item()) for x in box]

    def scatter_rows(self, rows) -> np.ndarray:
        """Rank 0 passes one row per rank -- a uint8 array, or a list of uint8 arrays that `join` would concatenate (they are
        then written straight into the page-locked send buffer, one copy); every rank returns its row."""
        torch, dist = self.torch, self.dist
        sizes = torch.zeros(self.world, dtype=torch.int64, device=self.device)
        layouts = None
        if self.rank == 0:
            layouts = []
            for r in rows:
                if isinstance(r, np.ndarray):
                    layouts.append((None, r.nbytes))
                else:
                    head = np.array([len(r)] + [b.nbytes for b in r], np.uint64)
                    layouts.append((head, _pad(head.nbytes) + sum(_pad(b.nbytes) for b in r)))
            sizes = self._to_device(np.array([n for _, n in layouts], np.int64))
        dist.broadcast(sizes, src=self.root, group=self.group)
        sizes = [int(x) for x in self._to_host(sizes)]
        width = max(_pad(max(sizes)), _ALIGN)
        recv = torch.empty(width, dtype=torch.uint8, device=self.device)
        parts = None
        if self.rank == 0:
            stage = self._staging("scatter", self.world * width)
            big = stage.numpy().reshape(self.world, width)
            for i, r in enumerate(rows):
                head, n = layouts[i]
                if head is None:
                    big[i, :n] = r
                else:
                    big[i, : head.nbytes] = head.view(np.uint8)
                    off = _pad(head.nbytes)
                    for b in r:
                        if isinstance(b, Lazy):
                            b.write(big[i, off : off + b.nbytes])  # encoded in place: no intermediate copy of the payload
                        else:
                            big[i, off : off + b.nbytes] = b
                        off += _pad(b.nbytes)
                # (padding bytes are never read: every reader goes by the recorded lengths)
            dev = stage.to(self.device, non_blocking=True) if self.cuda else stage
            parts = list(dev.reshape(self.world, width).unbind(0))
        dist.scatter(recv, parts, src=self.root, group=self.group)
        if self.cuda:
            out = self._staging("scatter_recv", width)
            out.copy_(recv)
            torch.cuda.current_stream().synchronize()
            return out.numpy()[: sizes[self.rank]].copy()
        return recv.numpy()[: sizes[self.rank]].copy()

    def gather_rows(self, row: np.ndarray) -> Optional[List[np.ndarray]]:
        """Every rank passes one uint8 row; rank 0 returns all of them (others None)."""
        torch, dist = self.torch, self.dist
        sizes = self.lengths(row.nbytes)
        width = max(_pad(max(sizes)), _ALIGN)
        mine = np.zeros(width, np.uint8)
        mine[: row.nbytes] = row
        send = self._to_device(mine)
        parts = [torch.empty(width, dtype=torch.uint8, device=self.device) for _ in range(self.world)] if self.rank == 0 else None
        dist.gather(send, parts, dst=self.root, group=self.group)
        if self.rank != 0:
            return None
        return [self._to_host(parts[i])[: sizes[i]].copy() for i in range(self.world)]

    def all_rows(self, row: np.ndarray, width: int) -> List[np.ndarray]:
        """all_gather of one fixed-width uint8 row per rank (a few hundred bytes: matrices, exit tables, status)."""
        torch, dist = self.torch, self.dist
        mine = np.zeros(width, np.uint8)
        mine[: row.nbytes] = row
        send = self._to_device(mine)
        box = [torch.empty(width, dtype=torch.uint8, device=self.device) for _ in range(self.world)]
        dist.all_gather(box, send, group=self.group)
        return [self._to_host(b).copy() for b in box]

    STATUS_WIDTH = 512

    def all_status(self, kind: int, text: str = "", payload: Optional[np.ndarray] = None, payload_width: int = 0):
        """Every rank reports (kind, text[, payload]); returns the list of all ranks' reports.  kind: 0 ok, 1 unsupported,
        2 error (RuntimeError), 3 Mendelian conflict.  One all_gather."""
        width = self.STATUS_WIDTH + _pad(payload_width)
        row = np.zeros(width, np.uint8)
        msg = text.encode("utf-8", "replace")[: self.STATUS_WIDTH - 16]
        row[:8] = np.array([kind], np.uint32).view(np.uint8).tolist() + np.array([len(msg)], np.uint32).view(np.uint8).tolist()
        row[8:16] = np.array([0 if payload is None else payload.nbytes], np.uint64).view(np.uint8)
        row[16 : 16 + len(msg)] = np.frombuffer(msg, np.uint8)
        if payload is not None:
            row[self.STATUS_WIDTH : self.STATUS_WIDTH + payload.nbytes] = np.ascontiguousarray(payload).reshape(-1).view(np.uint8)
        out = []
        for r in self.all_rows(row, width):
            k, n = (int(x) for x in r[:8].view(np.uint32))
            nb = int(r[8:16].view(np.uint64)[0])
            out.append((k, bytes(r[16 : 16 + n]).decode("utf-8", "replace"), r[self.STATUS_WIDTH : self.STATUS_WIDTH + nb] if nb else None))
        return out

    def warm_up(self) -> None:
        """First use of a communicator builds its rings / trees (tens to hundreds of ms on NCCL): do it outside any timed region."""
        self.scatter_rows([np.zeros(16, np.uint8)] * self.world if self.rank == 0 else None)
        self.gather_rows(np.zeros(16, np.uint8))
        self.all_status(0)


def status_of(exc: Optional[BaseException]) -> Tuple[int, str]:
    from ._abi import MendelianConflict, Unsupported

    if exc is None:
        return 0, ""
    if isinstance(exc, Unsupported):
        return 1, str(exc)
    if isinstance(exc, MendelianConflict):
        return 3, str(exc)
    return 2, "%s: %s" % (type(exc).__name__, exc) if not isinstance(exc, RuntimeError) else str(exc)


def raise_first_error(states, where: str) -> bool:
    """Raises on EVERY rank the first error any rank reported (same exception type as the single-GPU call would raise);
    returns True if some rank reported 'unsupported' (the caller falls back to one GPU)."""
    from ._abi import MendelianConflict

    for rank, (kind, text, _) in enumerate(states):
        if kind == 3:
            raise MendelianConflict(text or "Error: Mendelian conflict")
        if kind == 2:
            raise RuntimeError(text if text else "rank %d failed during %s" % (rank, where))
    return any(kind == 1 for kind, _, _ in states)

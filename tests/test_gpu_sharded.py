"""Block sharding with the real CUDA solver: two ranks (gloo rendezvous, both on cuda:0 so that the
test runs on a single-GPU box) solve their shares through the C ABI, rank 0 merges; must equal the
unsharded CUDA solve bit for bit.  With more GPUs the same code runs one rank per GPU over NCCL
(`scripts/sharded_nccl_check.py`).  Pedigrees go through `whmec_segment_*`: segments of one table."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q, kind="single"):
    import sys

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from whatshap_b200 import _lib, multigpu, synth

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    prob = None
    if rank == 0:
        prob = synth.config("cfg2", 3000) if kind == "single" else synth.trio(1500, 4, block_len=100, seed=77)
    sol = multigpu.solve_sharded(prob, solver=lambda p: _lib.solve(p, device=0)[0],
                                 segment_factory=lambda p, continues: _lib.Segment(p, continues, device=0))
    if rank == 0:
        whole, _ = _lib.solve(prob, device=0)
        q.put((sol.same_as(whole), sol.diff(whole), int(sol.cost)))
    dist.destroy_process_group()


@pytest.mark.parametrize("kind", ["single", "trio"])
def test_two_ranks_share_one_problem(gpu, kind):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, kind)) for r in range(2)]
    for p in procs:
        p.start()
    ok, diff, cost = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert ok, diff
    assert cost > 0


def _segments_equal_whole(gpu, prob, n_segments, checker=None):
    from whatshap_b200 import multigpu

    whole, stats = gpu.solve(prob)
    got = multigpu.solve_pedigree_segments(prob, n_segments, lambda p, continues: gpu.Segment(p, continues, device=0))
    assert got.same_as(whole), (n_segments, got.diff(whole))
    if checker is not None:
        assert got.same_as(checker.solve(prob)), got.diff(checker.solve(prob))
    return stats


@pytest.mark.parametrize("n_segments", [2, 3, 8])
def test_pedigree_segments_equal_the_whole_table(gpu, checker, n_segments):
    """`whmec_segment_*` on one device, segments driven in rank order: transfer matrices (multi-RHS pass),
    folded inputs, pass 2 and the three-kernel backtrace must give the single sweep's answer."""
    from whatshap_b200 import synth

    stats = _segments_equal_whole(gpu, synth.trio(600, 3, block_len=40, seed=41), n_segments, checker)
    assert stats["path_kind"] == 3
    _segments_equal_whole(gpu, synth.config("cfg5", 4000), n_segments)


@pytest.mark.parametrize("pedigree", ["trio", "quartet", "three_generations", "trio_child_first"])
def test_pedigree_segments_random(gpu, checker, pedigree):
    """T = 4 (multi-RHS kernel) and T = 16 (one instance per unit vector), distrusted genotypes, conflicts."""
    from whatshap_b200 import multigpu, synth

    rng = np.random.default_rng(len(pedigree) * 7 + 1)
    done = 0
    for it in range(40):
        prob = synth.random_problem(rng, int(rng.integers(6, 40)), int(rng.integers(2, 6)), pedigree=pedigree,
                                    distrust=it % 3 == 0, conflict_free=it % 6 != 0, mean_len=float(rng.choice([1.5, 3.0])))
        if len(multigpu.independent_blocks(prob)) < 2:
            continue
        try:
            want, werr = checker.solve(prob), ""
        except RuntimeError as e:
            want, werr = None, str(e)
        for n_segments in (2, 5):
            try:
                got, gerr = multigpu.solve_pedigree_segments(prob, n_segments, lambda p, c: gpu.Segment(p, c, device=0)), ""
            except RuntimeError as e:
                got, gerr = None, str(e)
            assert gerr == werr, (it, n_segments, gerr, werr)
            if want is not None:
                assert got.same_as(want), (it, n_segments, got.diff(want))
        done += 1
    assert done >= 15


def test_segment_api_misuse(gpu):
    from whatshap_b200 import synth
    from whatshap_b200._abi import Unsupported

    with pytest.raises(Unsupported):  # single individual: chains are independent problems, no segments
        gpu.Segment(synth.sliding_window(40, 4, block_len=20, seed=1), False)
    seg = gpu.Segment(synth.trio(60, 2, block_len=20, seed=2), True)
    with pytest.raises(RuntimeError, match="before whmec_segment_transfer"):
        seg.sweep(np.zeros(4, np.uint32))
    seg.transfer()
    with pytest.raises(RuntimeError, match="needs an input vector"):
        seg.sweep(None)
    with pytest.raises(RuntimeError, match="before whmec_segment_sweep"):
        seg.exits(False)
    seg.close()

"""N > 1 host logic on CPU: two `gloo` ranks shard the DP-independent blocks of a problem, solve
them with the CPU checker standing in for the per-rank CUDA call, and rank 0 merges — the result
must equal the unsharded solve bit for bit."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from whatshap_b200 import multigpu, synth


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, kind, out_queue):
    import sys

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import checker

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ck = checker.best()
    prob = None
    if rank == 0:
        if kind in ("single", "rank1_fails"):
            prob = synth.sliding_window(120, 6, block_len=20, seed=5, gap=0.1)
        elif kind == "trio":
            prob = synth.trio(40, 2, block_len=10, seed=6)
        elif kind == "trio_two_blocks":  # fewer blocks than ranks: one rank holds no segment
            prob = synth.trio(24, 2, block_len=12, seed=8)
        else:  # "conflict": the error of one rank's segment is raised on every rank
            prob = synth.trio(40, 2, block_len=10, seed=6)
            prob.gt = prob.gt.copy()
            prob.gt[:, 33] = [0, 0, 2]
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from emul_segment import EmulSegment

    solver = ck.solve
    if kind == "rank1_fails" and rank == 1:  # a rank-local failure (e.g. CUDA out of memory) must not leave the others in a collective
        def solver(p):
            raise RuntimeError("CUDA failure: out of memory (injected)")
    # pedigrees: every rank holds a segment of the table (host emulation of the per-rank CUDA calls)
    try:
        sol = multigpu.solve_sharded(prob, solver=solver, segment_factory=EmulSegment)
    except RuntimeError as e:
        if kind == "rank1_fails":
            assert "out of memory (injected)" in str(e), str(e)
        else:
            assert kind == "conflict" and "Mendelian conflict" in str(e), (kind, str(e))
        if rank == 0:
            out_queue.put((True, "", 1))
        dist.destroy_process_group()
        return
    assert kind not in ("conflict", "rank1_fails")
    if rank == 0:
        want = ck.solve(prob)
        out_queue.put((sol.same_as(want), sol.diff(want), int(sol.cost)))
    else:
        assert sol is None
    dist.destroy_process_group()


@pytest.mark.parametrize("kind,world", [("single", 2), ("trio", 2), ("trio", 3), ("trio_two_blocks", 3), ("conflict", 2), ("rank1_fails", 3)])
def test_block_sharding_matches_unsharded(kind, world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, kind, q)) for r in range(world)]
    for p in procs:
        p.start()
    ok, diff, cost = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok, diff
    assert cost > 0


def test_block_partition_and_lpt():
    prob = synth.sliding_window(100, 5, block_len=25, seed=2)
    blocks = multigpu.independent_blocks(prob)
    assert blocks == [(0, 25), (25, 50), (50, 75), (75, 100)]
    work = multigpu.block_work(prob, blocks)
    shares = multigpu.assign_blocks(work, 3)
    assert sorted(b for s in shares for b in s) == [0, 1, 2, 3]
    assert max(len(s) for s in shares) == 2
    # a read that bridges two blocks fuses them
    prob2 = synth.sliding_window(60, 4, block_len=60, seed=3)
    assert multigpu.independent_blocks(prob2) == [(0, 60)]


def test_ragged_blocks_balance_over_ranks():
    """cfg3g (block lengths ~ Geometric(500)): LPT by DP cells keeps 8 ranks within 1 % of each other,
    and the best contiguous runs (pedigree segments) are as balanced as ~40 ragged blocks allow."""
    prob = synth.config("cfg3g", 20000)
    blocks = multigpu.independent_blocks(prob)
    lengths = np.array([hi - lo for lo, hi in blocks])
    assert lengths.sum() == 20000 and lengths.min() >= 2 and lengths.max() > 3 * lengths.mean()
    work = multigpu.block_work(prob, blocks)
    for world in (2, 4, 8):
        loads = np.array([work[s].sum() for s in multigpu.assign_blocks(work, world)])
        assert loads.max() / loads.mean() < 1.01, (world, loads)
        runs = np.array([work[a:b].sum() for a, b in multigpu.contiguous_shares(work, world)])
        assert runs.max() / runs.mean() < (1.05, 1.15, 1.3)[(2, 4, 8).index(world)], (world, runs)


def _genotype_worker(rank, world, port, out_queue):
    import sys

    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.dirname(here))
    sys.path.insert(0, here)
    import emul_genotype

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    prob = synth.genotyping_problem(np.random.default_rng(21), 160, 5, "single", prior="random", burst=3, mean_len=3.0) if rank == 0 else None
    got = multigpu.genotype_sharded(prob, solver=lambda p: emul_genotype.genotype(p)[0])
    if rank == 0:
        whole, _ = emul_genotype.genotype(prob)
        out_queue.put((bool(np.array_equal(got, whole, equal_nan=True)), len(multigpu.independent_blocks(prob))))
    else:
        assert got is None
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_genotyping_shards_by_chains(world):
    """Single individual: every rank runs the forward-backward DP on its run of chains (emulated kernels stand in for the
    per-rank CUDA call); the gathered likelihoods equal the unsharded ones exactly."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_genotype_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    same, n_blocks = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert same and n_blocks > world


def test_wire_encoding_round_trips():
    from whatshap_b200 import _wire

    for prob in (synth.trio(30, 2, block_len=10, seed=1), synth.sliding_window(0, 3), synth.random_problem(np.random.default_rng(3), 20, 4, "quartet", distrust=True)):
        back, tag, lo = _wire.decode_problem(_wire.encode_problem(prob, tag=7, lo=11))
        assert (tag, lo) == (7, 11) and back.n_ind == prob.n_ind and back.distrust == prob.distrust
        for f in ("positions", "read_off", "ent_col", "ent_allele", "ent_phred", "read_ind", "recombcost", "trios", "gt"):
            assert np.array_equal(getattr(back, f), getattr(prob, f)), f
        assert (back.gl is None) == (prob.gl is None) and (prob.gl is None or np.array_equal(back.gl, prob.gl))
    from oracle import checker

    prob = synth.trio(30, 2, block_len=10, seed=1)
    sol = checker.best().solve(prob)
    back, tag, extra = _wire.decode_solution(_wire.encode_solution(sol, tag=3, extra=np.arange(4)))
    assert back.same_as(sol) and tag == 3 and extra.tolist() == [0, 1, 2, 3]
    rows = _wire.separate(_wire.join([np.arange(5, dtype=np.uint8), np.zeros(0, np.uint8), np.arange(40, dtype=np.uint8)]))
    assert [r.tolist() for r in rows] == [list(range(5)), [], list(range(40))]


def test_contiguous_shares_minimise_the_largest_run():
    """The block runs handed to the ranks: contiguous, covering, at most `world` of them, and the heaviest run is as light as a
    brute-force search over all cut positions can make it (the bisection over prefix sums of multigpu.contiguous_shares)."""
    import itertools

    import numpy as np

    from whatshap_b200 import multigpu

    rng = np.random.default_rng(11)
    for _ in range(200):
        n, world = int(rng.integers(1, 9)), int(rng.integers(1, 6))
        work = np.exp2(rng.integers(0, 12, n)).astype(np.float64)
        runs = multigpu.contiguous_shares(work, world)
        assert len(runs) == world
        real = [r for r in runs if r[1] > r[0]]
        assert real[0][0] == 0 and real[-1][1] == n and all(a[1] == b[0] for a, b in zip(real, real[1:]))
        heaviest = max(work[a:b].sum() for a, b in real)
        best = min(max(work[a:b].sum() for a, b in zip((0,) + cuts, cuts + (n,)))
                   for k in range(min(world, n)) for cuts in itertools.combinations(range(1, n), k))
        assert heaviest == best, (work, world, runs)
        if n >= world:  # no rank idles while another holds two blocks
            assert len(real) == world

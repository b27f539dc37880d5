"""Full-size shapes on the GPU, checked through size-independent properties (the CPU reference cannot
finish these in test time): the reported optimum must be the cost of the reported path, and solving
DP-independent blocks one by one must reproduce the whole-problem result bit for bit."""
import numpy as np
import pytest

from whatshap_b200 import multigpu, synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,cols", [("cfg2", 10_000), ("cfg3", 3_000), ("cfg4", 160)])
def test_cost_is_cost_of_reported_path(gpu, name, cols):
    prob = synth.config(name, cols)
    sol, stats = gpu.solve(prob)
    assert stats["path_kind"] == 1  # tile kernel
    assert sol.cost == synth.het_path_cost(prob, sol.path_index)
    # partition and super-reads are consistent with the path: a read is on side 0 iff its bit is 0
    assert set(np.unique(sol.partition)) <= {0, 1}
    assert np.all((sol.sr_allele == 0) | (sol.sr_allele == 1) | (sol.sr_allele == 3))


def test_blocks_solved_separately_equal_the_whole(gpu):
    prob = synth.config("cfg3", 2_000)
    whole, _ = gpu.solve(prob)
    blocks = multigpu.independent_blocks(prob)
    assert len(blocks) == 4
    parts = [gpu.solve(prob.slice_columns(lo, hi))[0] for lo, hi in blocks]
    merged = multigpu.merge_block_solutions(prob, blocks, parts)
    assert merged.same_as(whole), merged.diff(whole)


def test_column_kernel_agrees_with_tile_kernel_at_coverage_20(gpu, monkeypatch):
    """Both CUDA paths on a 2^20-cell-per-column instance (the CPU checker would need minutes)."""
    prob = synth.config("cfg3", 700)
    tile, st1 = gpu.solve(prob)
    monkeypatch.setenv("WHMEC_FORCE_COLUMN_KERNEL", "1")
    col, st2 = gpu.solve(prob)
    assert (st1["path_kind"], st2["path_kind"]) == (1, 2)
    assert tile.same_as(col), tile.diff(col)


def test_trio_replay_is_deterministic(gpu):
    prob = synth.config("cfg5", 1_500)
    plan = gpu.Plan(prob)
    plan.sweep()
    a = plan.finish()
    plan.sweep()  # second sweep is replayed from the captured CUDA graph
    plan.sweep()
    b = plan.finish()
    plan.close()
    assert a.same_as(b)


def test_pedigree_batched_sweep_equals_sequential_sweep(gpu, monkeypatch):
    """Trio with 6 chains: the two-pass transfer-matrix sweep (all chains advance together) must write
    exactly the back-pointers of the column-by-column sweep."""
    prob = synth.config("cfg5", 3_000)
    batched, st1 = gpu.solve(prob)
    monkeypatch.setenv("WHMEC_PED_SEQUENTIAL", "1")
    sequential, st2 = gpu.solve(prob)
    assert (st1["path_kind"], st2["path_kind"]) == (3, 2)
    assert batched.same_as(sequential), batched.diff(sequential)
    assert st1["kernel_launches"] < st2["kernel_launches"] / 2


@pytest.mark.parametrize("name,cols,block", [("many chains", 3000, 500), ("one chain", 400, 400)])
def test_memory_bounded_sweep_equals_the_resident_one(gpu, monkeypatch, name, cols, block):
    """Back-pointers that do not fit the arena budget: the launch rounds are cut into segments, the forward sweep keeps a
    checkpoint of the projection state in front of every segment, the backtrace re-sweeps the earlier segments from their
    checkpoints (the reference's sqrt(n) checkpointing, pedigreedptable.cpp:103-134,146-173, with large segments).  Same
    result as with everything resident, for budgets that give 2 ... many segments; an impossible budget is refused."""
    from whatshap_b200._abi import Unsupported

    prob = synth.sliding_window(cols, 18, block_len=block, seed=cols)
    monkeypatch.delenv("WHMEC_TILE_ARENA_BUDGET", raising=False)
    whole, st = gpu.solve(prob)
    assert st["path_kind"] == 1
    resident = st["backptr_bytes"]
    seen = set()
    for divisor in (1.5, 3, 8, 16):
        monkeypatch.setenv("WHMEC_TILE_ARENA_BUDGET", str(int(resident / divisor)))  # bytes for the arena; checkpoints come on top
        got, st2 = gpu.solve(prob)
        assert got.same_as(whole), (name, divisor, got.diff(whole))
        assert st2["backptr_bytes"] < resident
        seen.add(st2["kernel_launches"])
    assert len(seen) >= 2  # different budgets re-sweep different numbers of rounds
    monkeypatch.setenv("WHMEC_TILE_ARENA_BUDGET", "4096")
    with pytest.raises((Unsupported, RuntimeError)):
        gpu.solve(prob)


@pytest.mark.parametrize("cov", [31, 32])
def test_up_to_thirty_two_active_reads(gpu, monkeypatch, cov):
    """The reference's own limit (graycodes.cpp:12: 32 reads, 2^32 cells per column).  The CPU reference needs ~3 minutes and
    48 GB per such column, so the check is internal: the reported optimum is the cost of the reported path, and (coverage 31)
    the general column kernel -- different device code, 64-bit entry counts -- finds the same solution."""
    n = cov + 6
    prob = synth.sliding_window(n, cov, block_len=n, seed=cov)
    monkeypatch.delenv("WHMEC_FORCE_COLUMN_KERNEL", raising=False)
    sol, stats = gpu.solve(prob)
    assert stats["max_active"] == cov and stats["path_kind"] == 1
    assert sol.cost == synth.het_path_cost(prob, sol.path_index)
    assert set(np.unique(sol.partition)) <= {0, 1}
    if cov == 31:
        monkeypatch.setenv("WHMEC_FORCE_COLUMN_KERNEL", "1")
        col, st2 = gpu.solve(prob)
        assert st2["path_kind"] == 2
        assert col.same_as(sol), col.diff(sol)

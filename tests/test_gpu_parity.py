"""GPU parity: the CUDA path, called through the C ABI, against the CPU checker (bit-exact)."""
import numpy as np
import pytest

from conftest import solve_or_error
from whatshap_b200 import synth

pytestmark = pytest.mark.gpu

PEDS = ["single", "single", "trio", "quartet", "two_unrelated", "trio_child_first", "three_generations"]


def assert_same(gpu, checker, prob, label="", want_path=None):
    want, werr = solve_or_error(checker.solve, prob)
    stats = {}

    def run(p):
        sol, st = gpu.solve(p)
        stats.update(st)
        return sol

    got, gerr = solve_or_error(run, prob)
    assert werr == gerr, (label, werr, gerr)
    if want is not None:
        assert got.same_as(want), (label, got.diff(want))
        if want_path is not None and prob.n_cols:
            assert stats["path_kind"] == want_path, (label, stats)


@pytest.fixture(params=["tile", "column"])
def kernel_path(request, monkeypatch):
    """Single-individual problems run on the tile kernel by default; the env hook forces the
    general column kernel so that both CUDA paths are held to the same oracle."""
    if request.param == "column":
        monkeypatch.setenv("WHMEC_FORCE_COLUMN_KERNEL", "1")
        return 2
    monkeypatch.delenv("WHMEC_FORCE_COLUMN_KERNEL", raising=False)
    return 1


@pytest.mark.parametrize("seed", range(8))
def test_fuzz_irregular_instances(gpu, checker, seed):
    """Random spans, gaps, blanks, ties, trusted/distrusted genotypes, Mendelian conflicts."""
    rng = np.random.default_rng(1000 + seed)
    for it in range(60):
        ped = PEDS[it % len(PEDS)]
        maxcov = 6 if ped in ("quartet", "three_generations") else 9
        prob = synth.random_problem(
            rng,
            int(rng.integers(1, 18)),
            int(rng.integers(1, maxcov)),
            ped,
            distrust=bool(rng.integers(0, 2)),
            conflict_free=bool(rng.integers(0, 4)),
            max_phred=int(rng.integers(1, 8)),
        )
        assert_same(gpu, checker, prob, f"seed={seed} it={it} ped={ped}")


@pytest.mark.parametrize(
    "n,c,stride,block,gap",
    [(120, 10, 1, 50, 0.0), (90, 8, 2, 45, 0.1), (64, 12, 1, 32, 0.0), (40, 14, 1, 20, 0.05), (30, 16, 1, 30, 0.0)],
)
def test_sliding_window_blocks(gpu, checker, kernel_path, n, c, stride, block, gap):
    """Coverage-capped single-individual ReadSets; chain ends drop up to c reads at once."""
    prob = synth.sliding_window(n, c, stride=stride, block_len=block, seed=n * 31 + c, gap=gap)
    assert_same(gpu, checker, prob, want_path=kernel_path)


@pytest.mark.parametrize("n,c,block", [(70, 17, 35), (40, 19, 40), (48, 20, 24)])
def test_sliding_window_multi_tile(gpu, checker, kernel_path, n, c, block):
    """Coverage above one tile (2^14 entries): the projection column is cut along global reads."""
    prob = synth.sliding_window(n, c, block_len=block, seed=n + c)
    assert_same(gpu, checker, prob, want_path=kernel_path)


@pytest.mark.parametrize("seed", range(4))
def test_fuzz_single_individual_both_paths(gpu, checker, kernel_path, seed):
    rng = np.random.default_rng(77 + seed)
    for it in range(60):
        prob = synth.random_problem(
            rng,
            int(rng.integers(1, 40)),
            int(rng.integers(1, 11)),
            "single",
            distrust=bool(rng.integers(0, 2)),
            conflict_free=bool(rng.integers(0, 4)),
            max_phred=int(rng.integers(1, 8)),
            mean_len=float(rng.choice([2, 4, 8])),
        )
        assert_same(gpu, checker, prob, f"seed={seed} it={it}", want_path=kernel_path)


def test_trio_blocks(gpu, checker):
    prob = synth.trio(90, 3, block_len=30, seed=5)
    assert_same(gpu, checker, prob)
    prob = synth.trio(40, 4, block_len=40, seed=6)
    assert_same(gpu, checker, prob)


def test_empty(gpu, checker):
    prob = synth.sliding_window(0, 5)
    sol, _ = gpu.solve(prob)
    assert sol.cost == 0


def _spans_problem(spans, n_cols, seed):
    """Single individual, all-het, reads given as (first, last) column spans without gaps."""
    from whatshap_b200._abi import FlatProblem

    rng = np.random.default_rng(seed)
    off, cols = [0], []
    for a, b in spans:
        cols.extend(range(a, b + 1))
        off.append(len(cols))
    nnz = len(cols)
    return FlatProblem(
        positions=(np.arange(n_cols) + 1) * 10, read_off=np.array(off, np.uint64), ent_col=np.array(cols, np.uint32),
        ent_allele=rng.integers(0, 2, nnz).astype(np.uint8), ent_phred=rng.integers(1, 30, nnz).astype(np.uint32),
        read_ind=np.zeros(len(spans), np.uint32), recombcost=np.zeros(n_cols, np.uint32), n_ind=1, gt=np.ones((1, n_cols), np.uint8),
    )


def test_tile_planner_declines_and_column_kernel_takes_over(gpu, checker):
    """16 reads end in the same column while 3 others continue: more simultaneous drops than a tile
    holds, so the whole problem must fall back to the general column kernel — same answer."""
    spans = [(0, 10)] * 16 + [(5, 20)] * 3
    prob = _spans_problem(spans, 21, seed=4)
    want = checker.solve(prob)
    got, stats = gpu.solve(prob)
    assert stats["path_kind"] == 2
    assert got.same_as(want), got.diff(want)


def test_many_reads_start_together_and_end_one_by_one(gpu, checker):
    """18 reads start in column 0 (some become tile-global bits at once) and end staggered."""
    spans = [(0, 3 + i) for i in range(18)]
    prob = _spans_problem(spans, 21, seed=5)
    want = checker.solve(prob)
    got, stats = gpu.solve(prob)
    assert stats["path_kind"] == 1
    assert got.same_as(want), got.diff(want)


@pytest.mark.parametrize("seed", range(3))
def test_fuzz_high_coverage_irregular_single_individual(gpu, checker, kernel_path, seed):
    """Irregular spans at coverage up to 13: columns with >= 2^10 outputs take the tile kernel's fast
    paths (one read ends, zero / one / several reads start) as well as its generic loops."""
    rng = np.random.default_rng(900 + seed)
    for it in range(12):
        prob = synth.random_problem(rng, int(rng.integers(20, 45)), 13, "single", distrust=bool(rng.integers(0, 2)),
                                    conflict_free=True, max_phred=int(rng.integers(1, 30)), mean_len=float(rng.choice([8, 14])),
                                    gap=0.05)
        assert_same(gpu, checker, prob, f"seed={seed} it={it}", want_path=kernel_path)


def test_pedigree_many_reads_end_mid_chain(gpu, checker):
    """Trio, two chains; inside the first chain 9 reads end in one column while others continue
    (d = 9: one thread block per projection entry, raw hand-over to the next column) — the batched
    two-pass sweep must still equal the reference."""
    from whatshap_b200._abi import FlatProblem

    rng = np.random.default_rng(12)
    spans = [(0, 6)] * 9 + [(3, 12)] * 3 + [(8, 12)] * 2 + [(13, 20)] * 4 + [(15, 20)] * 3
    n_cols = 21
    off, cols = [0], []
    for a, b in spans:
        cols.extend(range(a, b + 1))
        off.append(len(cols))
    nnz = len(cols)
    prob = FlatProblem(
        positions=(np.arange(n_cols) + 1) * 10, read_off=np.array(off, np.uint64), ent_col=np.array(cols, np.uint32),
        ent_allele=rng.integers(0, 2, nnz).astype(np.uint8), ent_phred=rng.integers(1, 9, nnz).astype(np.uint32),
        read_ind=(np.arange(len(spans)) % 3).astype(np.uint32), recombcost=np.full(n_cols, 3, np.uint32), n_ind=3,
        trios=np.array([0, 1, 2], np.uint32), gt=np.ones((3, n_cols), np.uint8),
    )
    want = checker.solve(prob)
    got, stats = gpu.solve(prob)
    assert stats["path_kind"] == 3
    assert got.same_as(want), got.diff(want)


def test_tiny_chain_before_a_multi_tile_chain(gpu, checker):
    """Regression: a one-read chain (2 state words) used to leave the next chain's state buffers off a
    16-byte boundary, which the vectorised tile-major hand-off needs (irregular data only)."""
    spans = [(0, 1)]
    for s in range(-15, 38):
        a, b = max(2, 2 + s), min(2 + s + 16, 41)
        if b - a >= 1:
            spans.append((a, b))
    spans.sort(key=lambda ab: ab[0])
    prob = _spans_problem(spans, 42, seed=8)
    want = checker.solve(prob)
    got, stats = gpu.solve(prob)
    assert stats["path_kind"] == 1 and stats["max_active"] == 17
    assert got.same_as(want), got.diff(want)

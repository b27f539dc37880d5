"""Host logic of the CUDA path without a GPU: the packer (pack.cpp), the tile planner
(tile_plan.cpp) and the __host__ __device__ per-thread functions the kernels execute
(dp_device.h, tile_device.h) are stepped serially by a TEST-ONLY harness (tests/emul) and held to
the reference's golden vectors.  The harness is not part of the product and is never a fallback."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import golden_io
from whatshap_b200 import synth
from whatshap_b200._abi import CProblem, CSolution, FlatSolution, raise_for

EMUL_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "emul")


@pytest.fixture(scope="module")
def emul():
    subprocess.run(["make", "-C", EMUL_DIR, "all"], check=True, capture_output=True)
    libs = {}
    for name in ("libwhemul.so", "libwhemul_small.so"):
        lib = C.CDLL(os.path.join(EMUL_DIR, name))
        lib.whemul_solve.argtypes = [C.POINTER(CProblem), C.POINTER(CSolution), C.c_uint32, C.c_char_p, C.c_size_t]
        lib.whemul_tile_solve.argtypes = [C.POINTER(CProblem), C.POINTER(CSolution), C.c_uint32, C.POINTER(C.c_uint32), C.c_char_p, C.c_size_t]
        lib.whemul_plan_info.argtypes = [C.POINTER(CProblem), C.POINTER(C.c_uint64)]
        libs[name] = lib
    return libs


def run_column(lib, prob, chunk):
    sol = FlatSolution(prob.n_cols, prob.n_reads, prob.n_ind)
    cp, cs, err = prob.as_c(), sol.as_c(), C.create_string_buffer(512)
    raise_for(lib.whemul_solve(C.byref(cp), C.byref(cs), chunk, err, len(err)), err.value.decode())
    sol.cost = int(cs.cost)
    return sol


def run_tile(lib, prob, chunk):
    sol = FlatSolution(prob.n_cols, prob.n_reads, prob.n_ind)
    cp, cs, err, npan = prob.as_c(), sol.as_c(), C.create_string_buffer(512), C.c_uint32(0)
    rc = lib.whemul_tile_solve(C.byref(cp), C.byref(cs), chunk, C.byref(npan), err, len(err))
    if rc == 100:
        return None
    raise_for(rc, err.value.decode())
    sol.cost = int(cs.cost)
    return sol


@pytest.mark.parametrize("group", golden_io.GROUPS)
@pytest.mark.parametrize("chunk", [0, 2])
def test_column_kernel_logic_on_golden_vectors(emul, group, chunk):
    golden_io.check(lambda p: run_column(emul["libwhemul.so"], p, chunk), group)


@pytest.mark.parametrize("libname", ["libwhemul.so", "libwhemul_small.so"])
def test_tile_kernel_logic_on_golden_vectors(emul, libname):
    """Single-individual cases through the tile planner + tile per-thread code; the `_small` build
    shrinks tiles to 2^4 entries so that every case is cut into many panels and tiles."""
    lib = emul[libname]
    n_tiled = 0
    for group in golden_io.GROUPS:
        for label, prob, want, error in golden_io.load(group):
            if prob.n_ind != 1:
                continue
            try:
                got, gerr = run_tile(lib, prob, 0), ""
            except RuntimeError as e:
                got, gerr = "err", str(e)
            if got is None:  # planner declined: the column kernel takes such problems
                continue
            assert gerr == error, (label, gerr, error)
            if want is not None:
                assert got.same_as(want), (label, got.diff(want))
                n_tiled += 1
    assert n_tiled > 40


def test_tile_logic_on_ragged_chains(emul, checker):
    """Chains of very different lengths (block lengths ~ Geometric, the cfg3g shape at small scale): some end
    after two columns, some span many panels; both tile sizes."""
    for libname, cov, mean in (("libwhemul.so", 6, 20.0), ("libwhemul_small.so", 7, 12.0), ("libwhemul_small.so", 9, 30.0)):
        for seed in (1, 2):
            prob = synth.sliding_window(300, cov, block_len=synth.geometric_blocks(300, mean, seed), seed=seed, gap=0.1 * (seed - 1))
            got = run_tile(emul[libname], prob, 0)
            assert got is not None and got.same_as(checker.solve(prob)), (libname, cov, seed)


def test_tile_planner_on_benchmark_shapes(emul):
    """Panels per chain and state traffic for the BASELINE.json shapes (planner only, no DP)."""
    lib = emul["libwhemul.so"]
    out = (C.c_uint64 * 8)()
    # tiles per panel: 2^g / 2 -- of every pair of mirror-image tiles only one is computed (Panel::half)
    for name, n, chains, tiles_per_panel in (("cfg2", 2000, 4, 1), ("cfg3", 1000, 2, 16), ("cfg4", 200, 1, 512)):
        prob = synth.config(name, n)
        cp = prob.as_c()
        assert lib.whemul_plan_info(C.byref(cp), out) == 0
        panels, rounds, tiles, max_tiles, state_w, bp_w, traffic, alg = list(out)
        assert max_tiles == chains * tiles_per_panel
        # a panel sweeps ~14 columns per pass over the state: >= 10x less state traffic than one
        # read + one write of the projection column per column (the "algorithmic" figure)
        if tiles_per_panel > 1:
            assert traffic * 10 < alg
        else:
            assert traffic == 0 and rounds == 1


# ---- segments of a pedigree table (multi-GPU scheme for T > 1) ----
def _segmented(prob, n_segments):
    from emul_segment import EmulSegment
    from whatshap_b200 import multigpu

    return multigpu.solve_pedigree_segments(prob, n_segments, EmulSegment)


def test_pedigree_segments_on_golden_vectors():
    """Every golden pedigree case that has at least two chains, cut into 2 and 3 segments: transfer
    matrices, folded input vectors, pass-2 back-pointers, exit tables and the stitched backtrace must
    reproduce the reference's whole-table answer (cost, path, transmission vector, super-reads)."""
    from whatshap_b200 import multigpu

    n = 0
    for group in golden_io.GROUPS:
        for label, prob, want, error in golden_io.load(group):
            if prob.n_trios == 0 or want is None or len(multigpu.independent_blocks(prob)) < 2:
                continue
            for n_segments in (2, 3):
                got = _segmented(prob, n_segments)
                assert got.same_as(want), (label, n_segments, got.diff(want))
            n += 1
    assert n >= 10


@pytest.mark.parametrize("pedigree", ["trio", "trio_child_first", "quartet", "three_generations"])
def test_pedigree_segments_random(pedigree):
    """Irregular pedigrees (T = 4 and 16, trusted and distrusted genotypes, Mendelian conflicts) against
    the checker; a conflict must surface as the same error."""
    import sys

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import checker
    from whatshap_b200 import multigpu

    ck = checker.best()
    rng = np.random.default_rng(abs(hash(pedigree)) % 1000 + 3)
    done = 0
    for it in range(60):
        prob = synth.random_problem(rng, int(rng.integers(4, 24)), int(rng.integers(2, 5)), pedigree=pedigree,
                                    distrust=it % 3 == 0, conflict_free=it % 5 != 0, mean_len=float(rng.choice([1.5, 3.0])))
        if len(multigpu.independent_blocks(prob)) < 2:
            continue
        try:
            want, werr = ck.solve(prob), ""
        except RuntimeError as e:
            want, werr = None, str(e)
        for n_segments in (2, 4):
            try:
                got, gerr = _segmented(prob, n_segments), ""
            except RuntimeError as e:
                got, gerr = None, str(e)
            assert gerr == werr, (it, n_segments, gerr, werr)
            if want is not None:
                assert got.same_as(want), (it, n_segments, got.diff(want))
        done += 1
    assert done >= 25


def test_segment_helpers():
    from whatshap_b200 import multigpu

    inf = multigpu.UMAX
    m = np.array([[1, inf], [5, 2]], np.uint32)
    assert multigpu.minplus(np.array([3, inf], np.uint32), m).tolist() == [4, inf]
    assert multigpu.minplus(np.array([3, 0], np.uint32), m).tolist() == [4, 2]
    first = np.array([[7, 9], [7, 9]], np.uint32)  # the first segment ignores its input: equal rows
    ins = multigpu.segment_inputs([None, first, None, m])
    assert ins[0] is None and ins[1] is None and ins[2] is None and ins[3].tolist() == [7, 9]
    # right to left: the last segment starts at the optimum (-1) and hands exits[0] on
    assert multigpu.segment_entries([np.array([1, 0]), None, np.array([0, 1]), np.array([1, 1])]) == [1, None, 1, -1]
    assert multigpu.contiguous_shares(np.array([1.0, 1, 1, 1]), 2) == [(0, 2), (2, 4)]
    assert sorted(multigpu.contiguous_shares(np.array([5.0]), 3)) == [(0, 1), (1, 1), (1, 1)]  # one block: one rank has it
    shares = multigpu.contiguous_shares(np.array([8.0, 1, 1, 1, 1, 4]), 3)
    assert shares[0][0] == 0 and shares[-1][1] == 6 and all(a[1] == b[0] for a, b in zip(shares, shares[1:]))


# ---- group-wise driver of whmec_solve (csrc/grouped.h): slicing and merging, emulated kernels as backend ----
def _grouped(lib, prob, groups, tiles):
    lib.whemul_grouped_solve.argtypes = [C.POINTER(CProblem), C.POINTER(CSolution), C.c_uint32, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_uint32)]
    sol = FlatSolution(prob.n_cols, prob.n_reads, prob.n_ind)
    cp, cs, handled, n_groups = prob.as_c(), sol.as_c(), C.c_int(0), C.c_uint32(0)
    assert lib.whemul_grouped_solve(C.byref(cp), C.byref(cs), groups, tiles, C.byref(handled), C.byref(n_groups)) == 0
    sol.cost = int(cs.cost)
    return (sol if handled.value else None), n_groups.value


def test_groups_of_chains_reproduce_the_whole_solve(emul, checker):
    """A single-individual problem cut into groups of whole chains, every group solved on its own slice of the
    input arrays, results written back at their offsets: identical to the reference on the whole problem.
    Problems the driver must decline (pedigrees, few chains, unsorted input, Mendelian conflicts) fall through."""
    lib = emul["libwhemul.so"]
    rng = np.random.default_rng(4)
    handled = 0
    for it in range(120):
        prob = synth.random_problem(rng, int(rng.integers(10, 80)), int(rng.integers(2, 7)), pedigree="single", distrust=it % 4 == 0,
                                    conflict_free=it % 7 != 0, mean_len=float(rng.choice([1.5, 3.0])))
        try:
            want = checker.solve(prob)
        except RuntimeError:
            want = None
        for groups in (2, 4):
            got, _ = _grouped(lib, prob, groups, it % 2)
            if got is None:
                continue
            assert want is not None and got.same_as(want), (it, groups, got.diff(want) if want is not None else "error expected")
            handled += 1
    assert handled > 100
    prob = synth.sliding_window(600, 12, block_len=50, seed=3)
    got, n_groups = _grouped(lib, prob, 4, 1)
    assert n_groups == 4 and got.same_as(checker.solve(prob))
    assert _grouped(lib, synth.trio(60, 2, block_len=10, seed=1), 2, 0)[0] is None      # transmission values couple the chains
    assert _grouped(lib, synth.sliding_window(60, 5, block_len=30, seed=1), 2, 1)[0] is None  # two chains: not worth cutting


def test_fused_pedigree_sweep(emul, checker):
    """Host mirror of the fused trio sweep (ped_fused_kernel, csrc/ped_fused.h): ONE unit sweep per chain -> the whole transfer
    matrix by symmetry (Mat[u][i] = row[i ^ u]) -> prefix -> true inputs; 16 register slots for the cost functions, byte tables,
    Gray steps, lanes-per-entry splitting with key merging on small columns, transition minima handed over inside a chain.  Must
    equal the reference on golden and random trios (trusted genotypes; distrusted ones have 16 assignments per transmission
    value and stay on the general sweep)."""
    lib = emul["libwhemul.so"]
    lib.whemul_ped_fused_solve.argtypes = [C.POINTER(CProblem), C.POINTER(CSolution), C.c_char_p, C.c_size_t]

    def run(prob):
        sol = FlatSolution(prob.n_cols, prob.n_reads, prob.n_ind)
        cp, cs, err = prob.as_c(), sol.as_c(), C.create_string_buffer(256)
        rc = lib.whemul_ped_fused_solve(C.byref(cp), C.byref(cs), err, len(err))
        if rc == 100:
            return None
        raise_for(rc, err.value.decode())
        sol.cost = int(cs.cost)
        return sol

    n = 0
    for group in golden_io.GROUPS:
        for label, prob, want, error in golden_io.load(group):
            if prob.n_trios != 1 or want is None:
                continue
            got = run(prob)
            if got is not None:
                assert got.same_as(want), (label, got.diff(want))
                n += 1
    assert n >= 15
    rng = np.random.default_rng(23)
    done = 0
    for it in range(90):
        prob = synth.random_problem(rng, int(rng.integers(4, 40)), int(rng.integers(2, 6)), pedigree=["trio", "trio_child_first"][it % 2],
                                    distrust=False, mean_len=float(rng.choice([1.5, 4.0, 8.0])), max_phred=int(rng.choice([1, 3, 40])),
                                    conflict_free=it % 5 != 0, gap=float(rng.choice([0.0, 0.2])))
        try:
            want = checker.solve(prob)
        except RuntimeError:
            continue
        got = run(prob)
        if got is not None:
            assert got.same_as(want), (it, got.diff(want))
            done += 1
    assert done >= 50
    for prob in (synth.trio(120, 5, block_len=60, seed=20250935), synth.trio(90, 4, block_len=30, seed=5, recomb_every=30, max_phred=3)):
        got = run(prob)
        assert got is not None and got.same_as(checker.solve(prob))


# ---- extremes of the value range and of the pedigree size (column path; SURVEY.md Appendix A: u32 arithmetic wraps) ----
def _extreme_problems():
    rng = np.random.default_rng(3)
    for max_phred in (10**6, 2**26, 2**30, 2**31 + 5):  # totals beyond 2^28 switch off every shortcut that assumes no wrap-around
        yield f"single max_phred={max_phred}", synth.random_problem(rng, 14, 5, pedigree="single", max_phred=max_phred)
        yield f"trio max_phred={max_phred}", synth.random_problem(rng, 10, 4, pedigree="trio", max_phred=max_phred)
    for rc in (2**31 - 1, 2**32 - 1):  # popcount * recombcost wraps
        prob = synth.random_problem(rng, 12, 4, pedigree="trio")
        prob.recombcost = np.full(prob.n_cols, rc, np.uint32)
        yield f"trio recombcost={rc}", prob
    prob = synth.random_problem(rng, 10, 4, pedigree="trio", distrust=True)
    prob.gl = prob.gl * 1e8  # `unsigned += double` on large likelihoods
    yield "trio gl*1e8", prob
    prob = synth.random_problem(rng, 10, 4, pedigree="single", distrust=True)
    prob.gl = prob.gl * 4e8
    yield "single gl*4e8", prob
    for pedigree in ("three_children", "four_children"):  # T = 64 and T = 256 (the supported maximum)
        for it in range(2):
            yield f"{pedigree} #{it}", synth.random_problem(rng, int(rng.integers(3, 6)), 3, pedigree=pedigree, distrust=it == 1, mean_len=2.0)


def test_extreme_values_and_pedigree_sizes(emul, checker):
    lib = emul["libwhemul.so"]
    n = 0
    for label, prob in _extreme_problems():
        want, got = checker.solve(prob), run_column(lib, prob, 0)
        assert got.same_as(want), (label, got.diff(want))
        tiled = run_tile(lib, prob, 0) if prob.n_ind == 1 else None
        assert tiled is None or tiled.same_as(want), (label, "tile", tiled.diff(want))  # the planner declines what it cannot do exactly
        n += 1
    assert n == 16


def test_more_than_thirty_two_active_reads_is_refused(emul):
    """32 active reads is the reference's own limit (graycodes.cpp:12): one more is refused by the packer."""
    from whatshap_b200._abi import Unsupported

    with pytest.raises(Unsupported, match="more than 32 reads are active"):
        run_column(emul["libwhemul.so"], synth.sliding_window(40, 33, block_len=40, seed=1), 0)


def test_host_worker_pool_serves_concurrent_callers(emul):
    """Packer / planner calls from several Python threads at once (ctypes releases the GIL): one caller owns the persistent
    pool, the others fall back to short-lived threads (hostpool.cpp); every call returns the same schedule digest."""
    import threading

    lib = emul["libwhemul.so"]
    lib.whemul_plan_digest.argtypes = [C.POINTER(CProblem), C.POINTER(C.c_uint64)]
    prob = synth.config("cfg3", 3000)
    cp = prob.as_c()
    want = C.c_uint64(0)
    assert lib.whemul_plan_digest(C.byref(cp), C.byref(want)) == 0
    results, errors = [], []

    def work():
        try:
            for _ in range(6):
                d = C.c_uint64(0)
                rc = lib.whemul_plan_digest(C.byref(cp), C.byref(d))
                results.append((rc, d.value))
        except Exception as e:  # pragma: no cover
            errors.append(e)

    threads = [threading.Thread(target=work) for _ in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=120)
    assert not errors and all(not t.is_alive() for t in threads)
    assert len(results) == 24 and all(r == (0, want.value) for r in results)


def test_steady_state_column_code_of_the_tile_kernel(emul, checker, monkeypatch):
    """`column_fast` (whatshap_b200/csrc/tile_fast.h) — the code that executes 60 % of the tile kernel's instructions — run
    thread by thread with its warp ballots assembled from the lanes' predicates: same cost, path and super-reads as the
    reference, and as the generic per-output loop (WHEMUL_TILE_FAST=0), tie-heavy weights and homozygous / distrusted
    columns (K0 active) included."""
    lib = emul["libwhemul.so"]
    lib.whemul_last_fast_columns.restype = C.c_uint64
    total_fast = 0
    for cov, n in ((11, 30), (12, 36), (15, 34), (16, 30), (17, 26)):
        for seed, max_phred in ((1, 40), (2, 2)):
            prob = synth.sliding_window(n, cov, block_len=n, seed=seed, gap=0.06 * seed, max_phred=max_phred)
            if seed == 2:  # homozygous sites make K0 finite (HASK0 variants)
                prob.gt = prob.gt.copy()
                prob.gt[0, ::3] = 0
            want = checker.solve(prob)
            monkeypatch.delenv("WHEMUL_TILE_FAST", raising=False)
            fast = run_tile(lib, prob, 0)
            n_fast = int(lib.whemul_last_fast_columns())
            monkeypatch.setenv("WHEMUL_TILE_FAST", "0")
            generic = run_tile(lib, prob, 0)
            assert int(lib.whemul_last_fast_columns()) == 0
            assert fast is not None and fast.same_as(want), (cov, seed, fast.diff(want))
            assert generic.same_as(want), (cov, seed)
            assert n_fast > 0 or cov < 12, (cov, n_fast)
            total_fast += n_fast
    assert total_fast > 300


@pytest.mark.parametrize("cov,n,seed", [(18, 44, 1), (20, 48, 3)])
def test_tile_kernel_code_at_benchmark_coverage(emul, checker, cov, n, seed):
    """Coverage of the BASELINE.json workloads (2^18 .. 2^20 cells per column): several tiles per panel, several panels with
    tile-major hand-offs, the steady-state column code on every tile — against the reference on a chain the CPU finishes in seconds."""
    lib = emul["libwhemul.so"]
    lib.whemul_last_fast_columns.restype = C.c_uint64
    prob = synth.sliding_window(n, cov, block_len=n, seed=seed, gap=0.05, max_phred=40 if seed % 4 == 1 else 2)
    got = run_tile(lib, prob, 0)
    assert got is not None and int(lib.whemul_last_fast_columns()) > 150
    assert got.same_as(checker.solve(prob)), got.diff(checker.solve(prob))


def test_thread_packed_back_pointers(emul, checker, monkeypatch):
    """Thread-packed back-pointer bits of the fast columns (the default; WHMEC_TILE_PACKED_BP=0 keeps warp ballots): each thread keeps the bits of its own
    outputs, formed from the sign of v1 - v0 - par by a funnel shift; tile_packed_bit_index decodes them in the backtrace):
    same results as the reference with both layouts, sliding windows (twin outputs) and irregular starts, K0 active or not."""
    lib = emul["libwhemul.so"]
    lib.whemul_last_fast_columns.restype = C.c_uint64
    lib.whemul_last_packed_columns.restype = C.c_uint64
    monkeypatch.setenv("WHMEC_TILE_PACKED_BP", "1")
    rng = np.random.default_rng(23)
    n_fast = n_packed = 0
    cases = [synth.sliding_window(n, cov, block_len=n, seed=seed, gap=0.05 * (seed % 3), max_phred=2 if seed % 2 == 0 else 40)
             for cov, n, seed in ((13, 30, 1), (14, 30, 2), (15, 34, 3), (16, 30, 4), (17, 28, 5), (19, 40, 6))]
    for prob in cases[1::2]:
        prob.gt = prob.gt.copy()
        prob.gt[0, ::4] = 2  # homozygous sites: K0 finite
    cases += [synth.random_problem(rng, 30, cov, "single", gap=0.1, mean_len=14.0, burst=5, max_phred=3) for cov in (13, 14, 15)]
    for prob in cases:
        got = run_tile(lib, prob, 0)
        assert got is not None and got.same_as(checker.solve(prob)), got.diff(checker.solve(prob))
        n_fast += int(lib.whemul_last_fast_columns())
        n_packed += int(lib.whemul_last_packed_columns())
    assert n_fast > 400 and n_packed > 300  # 8 or 16 outputs per thread: coverage >= 14


def test_host_worker_pool_rethrows_a_task_exception_on_the_caller(emul):
    """A throwing task (std::bad_alloc in a packer chunk ...) neither terminates a pool worker nor leaves workers with a dangling
    job: the first exception reaches the caller (and from there the guarded C-ABI wrapper), the pool keeps working."""
    lib = emul["libwhemul.so"]
    for n_threads in (1, 2, 8):
        assert lib.whemul_pool_throw_check(n_threads) == 0

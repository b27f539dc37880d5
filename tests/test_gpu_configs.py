"""GPU parity at the BASELINE.json shapes, against the compiled reference (oracle/_ref, `checker.best()`), through the C ABI.

The reference needs 40 ns per DP cell, so the full configurations cannot be checked in test time; these instances keep the
SHAPE of every configuration (coverage, tile / panel structure, number of global reads, pedigree batch shape) and cut the
column count to what the reference finishes in seconds:

* cfg4 (coverage 25, 2^25 cells per column, 11 global reads): a 10-column single block, and a coverage-23 block of 20 columns;
* cfg3 (coverage 20): one 80-column single block (>= 5 panels of the steady state, tile-major hand-offs) and one at coverage 18;
* cfg5 (trio, 5 reads per sample, a = 15, T = 4): 120 columns in 2 chains through the batched two-pass sweep and through
  the segment entry points (the multi-GPU scheme on one device).

Every kernel switch of DESIGN.md 7d is run over the same instances: a switch that cannot reproduce the reference is deleted."""
import numpy as np
import pytest

from conftest import solve_or_error
from whatshap_b200 import multigpu, synth

pytestmark = pytest.mark.gpu

SWITCHES = {
    "default": {},
    "ballot_bp": {"WHMEC_TILE_PACKED_BP": "0"},
    "mirror_off": {"WHMEC_TILE_MIRROR": "0"},
    "groups4": {"WHMEC_SOLVE_GROUPS": "4"},
    "pageable_upload": {"WHMEC_PINNED_UPLOAD": "0"},  # upload arrays on the heap instead of the page-locked pool
    "column": {"WHMEC_FORCE_COLUMN_KERNEL": "1"},
}
PED_SWITCHES = {
    "default": {},
    "ped_batched": {"WHMEC_PED_FUSED": "0"},
    "ped_one_cta": {"WHMEC_PED_CLUSTER": "1"},   # default: a cluster of up to 8 CTAs per chain when there are fewer chains than SMs
    "ped_cluster2": {"WHMEC_PED_CLUSTER": "2"},
    "ped_sequential": {"WHMEC_PED_SEQUENTIAL": "1"},
    "pageable_upload": {"WHMEC_PINNED_UPLOAD": "0"},
}
ALL_ENV = sorted({k for d in list(SWITCHES.values()) + list(PED_SWITCHES.values()) for k in d})

_want = {}


def reference_answer(checker, key, make):
    """The reference is slow at these sizes: one run per instance, shared by all switch parametrisations."""
    if key not in _want:
        prob = make()
        _want[key] = (prob, solve_or_error(checker.solve, prob))
    return _want[key]


def set_switch(monkeypatch, env):
    for k in ALL_ENV:
        monkeypatch.delenv(k, raising=False)
    for k, v in env.items():
        monkeypatch.setenv(k, v)


def check(gpu, checker, key, make):
    prob, (want, werr) = reference_answer(checker, key, make)
    assert werr is None, werr
    got, stats = gpu.solve(prob)
    assert got.same_as(want), (key, stats, got.diff(want))
    return stats


SINGLE = {
    # key: (generator, expected number of active reads)
    "cov25x10": (lambda: synth.sliding_window(10, 25, block_len=10, seed=20250925), 25),  # cfg4 shape: 2^25 cells per column, 11 global reads (reference: ~30 s)
    "cov23x20": (lambda: synth.sliding_window(20, 23, block_len=20, seed=23), 23),
    "cov20x80": (lambda: synth.sliding_window(80, 20, block_len=80, seed=20250920), 20),  # cfg3 shape, one chain, >= 5 panels
    "cov18x80": (lambda: synth.sliding_window(80, 18, block_len=80, seed=18), 18),
    "cov20x2blocks": (lambda: synth.sliding_window(96, 20, block_len=48, seed=7), 20),
    "cov16gaps": (lambda: synth.sliding_window(90, 16, block_len=45, seed=3, gap=0.1, max_phred=3), 16),  # tie-heavy
    "cov15x2000": (lambda: synth.sliding_window(2000, 15, block_len=500, seed=20250915), 15),  # cfg2 shape, whole chains
}


# the general column kernel is not the path of the coverage 25 / 23 shapes (minutes per instance): those two pairs are not generated
@pytest.mark.parametrize("key,switch", [(k, s) for k in SINGLE for s in SWITCHES if not (s == "column" and k in ("cov25x10", "cov23x20"))])
def test_single_individual_shapes(gpu, checker, monkeypatch, key, switch):
    make, active = SINGLE[key]
    set_switch(monkeypatch, SWITCHES[switch])
    stats = check(gpu, checker, key, make)
    assert stats["max_active"] == active, stats
    assert stats["path_kind"] == (2 if switch == "column" else 1), stats


TRIO = {
    "trio5x120": lambda: synth.trio(120, 5, block_len=60, seed=20250935),        # cfg5 shape: a = 15, T = 4, 2 chains
    "trio5x150r": lambda: synth.trio(150, 5, block_len=50, seed=11, recomb_every=40, max_phred=4),  # recombinations, ties
    "trio4x90": lambda: synth.trio(90, 4, block_len=30, seed=5),
}


@pytest.mark.parametrize("switch", list(PED_SWITCHES))
@pytest.mark.parametrize("key", list(TRIO))
def test_trio_shapes(gpu, checker, monkeypatch, key, switch):
    set_switch(monkeypatch, PED_SWITCHES[switch])
    stats = check(gpu, checker, key, TRIO[key])
    assert stats["path_kind"] == (2 if switch == "ped_sequential" else 3), stats
    if switch in ("default", "ped_one_cta", "ped_cluster2"):
        assert stats["kernel_launches"] == 3, stats  # the fused trio sweep: unit pass, prefix, true pass


@pytest.mark.parametrize("segments", [2, 3])
@pytest.mark.parametrize("key", list(TRIO))
def test_trio_shapes_through_segments(gpu, checker, monkeypatch, key, segments):
    """The multi-GPU scheme of SURVEY.md 8(e) for pedigrees (transfer matrices, true inputs, exits), all segments on one device."""
    set_switch(monkeypatch, {})
    prob, (want, werr) = reference_answer(checker, key, TRIO[key])
    assert werr is None
    got = multigpu.solve_pedigree_segments(prob, segments)
    assert got.same_as(want), (key, got.diff(want))


@pytest.mark.parametrize("switch", list(SWITCHES))
def test_golden_vectors_under_every_switch(gpu, monkeypatch, switch):
    import golden_io

    set_switch(monkeypatch, SWITCHES[switch])
    for group in golden_io.GROUPS:
        assert golden_io.check(lambda p: gpu.solve(p)[0], group) > 0


@pytest.mark.parametrize("switch", list(PED_SWITCHES))
def test_golden_vectors_under_every_pedigree_switch(gpu, monkeypatch, switch):
    import golden_io

    set_switch(monkeypatch, PED_SWITCHES[switch])
    for group in golden_io.GROUPS:
        assert golden_io.check(lambda p: gpu.solve(p)[0], group) > 0


@pytest.mark.parametrize("switch", ["default", "ballot_bp", "mirror_off", "groups4"])
def test_fuzz_high_coverage_under_switches(gpu, checker, monkeypatch, switch):
    """Sliding windows and irregular spans at coverage 14-19 (long runs of steady-state columns, homozygous sites, gaps,
    tie-heavy and wide weights): the thread-packed and the ballot column code against the reference."""
    set_switch(monkeypatch, SWITCHES[switch])
    rng = np.random.default_rng(4242)
    for it in range(24):
        cov = int(rng.integers(14, 20))
        if it % 2:
            prob = synth.random_problem(rng, int(rng.integers(12, 40)), cov, "single", distrust=bool(rng.integers(0, 3) == 0),
                                        conflict_free=True, max_phred=int(rng.integers(1, 40)), mean_len=float(rng.choice([10, 16, 24])),
                                        gap=float(rng.choice([0.0, 0.1])), burst=6)
        else:
            length = int(rng.integers(cov + 6, 56))
            prob = synth.sliding_window(length, cov, block_len=int(rng.integers(cov + 4, length + 1)), seed=int(rng.integers(1 << 30)),
                                        gap=float(rng.random() * 0.12), max_phred=int(rng.choice([1, 2, 40, 90])))
            if rng.random() < 0.3:
                prob.gt = prob.gt.copy()
                prob.gt[0, rng.random(prob.n_cols) < 0.1] = int(rng.integers(0, 3))
        key = ("fuzz", it)
        if key not in _want:
            _want[key] = (prob, solve_or_error(checker.solve, prob))
        prob, (want, werr) = _want[key]
        got, gerr = solve_or_error(lambda p: gpu.solve(p)[0], prob)
        assert gerr == werr, (it, gerr, werr)
        if want is not None:
            assert got.same_as(want), (switch, it, cov, got.diff(want))

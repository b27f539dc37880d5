"""CPU fuzz of the product's host code + the kernels' per-thread code (tests/emul: packer, planner, column kernel code, tile kernel code
incl. the steady-state column `column_fast`) against the compiled reference (oracle/_ref).  No GPU.  Authoring container only.
    python scripts/cpu_fuzz_emul.py <seed> <seconds>          # six pedigree shapes + single individuals, coverage <= 12
    python scripts/cpu_fuzz_emul.py <seed> <seconds> fast     # single individuals, coverage 11-15: every tile runs column_fast
    python scripts/cpu_fuzz_emul.py <seed> <seconds> mirror   # single individuals, coverage 15-19: mirrored multi-tile panels
    python scripts/cpu_fuzz_emul.py <seed> <seconds> fused    # trios through the fused pedigree sweep (csrc/ped_fused.h)
Set WHMEC_TILE_PACKED_BP=1 to fuzz the thread-packed back-pointer layout."""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import checker  # noqa: E402
from whatshap_b200 import synth  # noqa: E402
from whatshap_b200._abi import CProblem, CSolution, FlatSolution, raise_for  # noqa: E402

EMUL = os.path.join(ROOT, "tests", "emul")
libs = {}
for name in ("libwhemul.so", "libwhemul_small.so"):
    lib = C.CDLL(os.path.join(EMUL, name))
    lib.whemul_solve.argtypes = [C.POINTER(CProblem), C.POINTER(CSolution), C.c_uint32, C.c_char_p, C.c_size_t]
    lib.whemul_tile_solve.argtypes = [C.POINTER(CProblem), C.POINTER(CSolution), C.c_uint32, C.POINTER(C.c_uint32), C.c_char_p, C.c_size_t]
    lib.whemul_last_fast_columns.restype = C.c_uint64
    libs[name] = lib


def run(lib, prob, tile, chunk):
    sol = FlatSolution(prob.n_cols, prob.n_reads, prob.n_ind)
    cp, cs, err, npan = prob.as_c(), sol.as_c(), C.create_string_buffer(512), C.c_uint32(0)
    rc = (lib.whemul_tile_solve(C.byref(cp), C.byref(cs), chunk, C.byref(npan), err, 512) if tile
          else lib.whemul_solve(C.byref(cp), C.byref(cs), chunk, err, 512))
    if rc == 100:
        return None
    raise_for(rc, err.value.decode())
    sol.cost = int(cs.cost)
    return sol


def run_fused(prob):
    lib = libs["libwhemul.so"]
    lib.whemul_ped_fused_solve.argtypes = [C.POINTER(CProblem), C.POINTER(CSolution), C.c_char_p, C.c_size_t]
    sol = FlatSolution(prob.n_cols, prob.n_reads, prob.n_ind)
    cp, cs, err = prob.as_c(), sol.as_c(), C.create_string_buffer(512)
    rc = lib.whemul_ped_fused_solve(C.byref(cp), C.byref(cs), err, 512)
    if rc == 100:
        return None
    raise_for(rc, err.value.decode())
    sol.cost = int(cs.cost)
    return sol


seed, budget = int(sys.argv[1]), float(sys.argv[2])
if len(sys.argv) > 3 and sys.argv[3] == "fused":  # trios through the fused sweep's per-item code (csrc/ped_fused.h)
    rng = np.random.default_rng(seed)
    ref = checker.reference()
    t0 = time.time()
    n = done = 0
    while time.time() - t0 < budget:
        prob = synth.random_problem(rng, int(rng.integers(1, 60)), int(rng.integers(1, 7)), ["trio", "trio_child_first"][n % 2], distrust=False,
                                    gap=float(rng.random() * 0.3), mean_len=float(rng.uniform(1.5, 12)), burst=int(rng.integers(2, 6)),
                                    conflict_free=bool(rng.integers(0, 5)), max_phred=int(rng.choice([1, 5, 40])))
        if n % 7 == 0:
            prob.recombcost = rng.integers(0, 60, prob.n_cols).astype(np.uint32)
        n += 1
        try:
            want, werr = ref.solve(prob), ""
        except RuntimeError as e:
            want, werr = None, str(e)
        try:
            got, gerr = run_fused(prob), ""
        except RuntimeError as e:
            got, gerr = None, str(e)
        if got is None and not gerr:
            continue
        assert gerr == werr, (seed, n, gerr, werr)
        if want is not None:
            assert got.same_as(want), (seed, n, got.diff(want))
        done += 1
    print("seed", seed, "trios", n, "through the fused sweep", done, "OK")
    sys.exit(0)
fast_mode = len(sys.argv) > 3 and sys.argv[3] in ("fast", "mirror")
mirror_mode = len(sys.argv) > 3 and sys.argv[3] == "mirror"  # coverage 15-19: several tiles per panel, mirrored panels, both parities of km
rng = np.random.default_rng(seed)
ref = checker.reference()
assert ref is not None, "needs the compiled reference (oracle/_ref)"
peds = list(synth.PEDIGREES)
t0 = time.time()
n = tiles = fast_cols = 0
while time.time() - t0 < budget:
    if fast_mode:
        cov = int(rng.integers(15, 20)) if mirror_mode else int(rng.integers(11, 16))
        if n % 2:
            prob = synth.random_problem(rng, int(rng.integers(8, 40)), cov, "single", distrust=bool(rng.integers(0, 2)), gap=float(rng.random() * 0.2),
                                        mean_len=float(rng.uniform(8, 20)), burst=int(rng.integers(3, 7)), conflict_free=True, max_phred=int(rng.choice([1, 3, 40])))
        else:
            prob = synth.sliding_window(int(rng.integers(cov + 2, 40)), cov, block_len=int(rng.integers(cov + 2, 40)), seed=int(rng.integers(1 << 30)),
                                        gap=float(rng.random() * 0.15), max_phred=int(rng.choice([1, 3, 40])))
            if rng.random() < 0.5:
                prob.gt = prob.gt.copy()
                prob.gt[0, rng.random(prob.n_cols) < 0.3] = int(rng.integers(0, 3))
        combos = [("libwhemul.so", True)]
    else:
        ped = peds[n % len(peds)] if n % 3 == 0 else "single"
        prob = synth.random_problem(rng, int(rng.integers(1, 60)), int(rng.integers(1, 13 if ped == "single" else 5)), ped, distrust=bool(rng.integers(0, 2)),
                                    gap=float(rng.random() * 0.3), mean_len=float(rng.uniform(1.5, 10)), burst=int(rng.integers(2, 6)),
                                    conflict_free=bool(rng.integers(0, 5)), max_phred=int(rng.choice([1, 5, 40])))
        combos = [(name, tile) for name in libs for tile in ((False, True) if ped == "single" else (False,))]
    try:
        want, werr = ref.solve(prob), ""
    except RuntimeError as e:
        want, werr = None, str(e)
    for name, tile in combos:
        try:
            got, gerr = run(libs[name], prob, tile, 0 if fast_mode else int(rng.integers(0, 3))), ""
        except RuntimeError as e:
            got, gerr = None, str(e)
        if tile and got is None and not gerr:
            continue  # the planner declined: the column kernel handles it
        assert gerr == werr, (seed, n, name, tile, gerr, werr)
        if want is not None:
            assert got.same_as(want), (seed, n, name, tile, got.diff(want))
        tiles += tile
        if tile:
            fast_cols += int(libs[name].whemul_last_fast_columns())
    n += 1
print("seed", seed, "problems", n, "tile solves", tiles, "(tile, column) pairs through column_fast", fast_cols, "OK")

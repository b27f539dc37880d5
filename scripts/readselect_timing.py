"""Wall time of the coverage-capping read selection on a synthetic deep read set:
this package's `readselect.select_reads_csr` and -- where the reference can be built (authoring container,
oracle/build_pyref.py) -- the reference's `whatshap.readselect.readselection` on the same reads, with the two
selections compared.    python scripts/readselect_timing.py [n_variants] [n_reads] [max_cov]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import build_pyref  # noqa: E402
from whatshap_b200 import readselect  # noqa: E402

n_var = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000
n_reads = int(sys.argv[2]) if len(sys.argv) > 2 else 400_000
max_cov = int(sys.argv[3]) if len(sys.argv) > 3 else 15
rng = np.random.default_rng(5)
positions = np.sort(rng.choice(np.arange(1, 20 * n_var), n_var, replace=False)).astype(np.int32)
reads = []
for _ in range(n_reads):
    start = int(rng.integers(0, n_var - 1))
    idx = np.arange(start, min(n_var, start + 2 + int(rng.geometric(1 / 8.0))))
    if len(idx) > 2:
        keep = rng.random(len(idx)) >= 0.1
        keep[0] = keep[-1] = True
        idx = idx[keep]
    if len(idx) >= 2:
        reads.append(idx)
reads.sort(key=lambda r: int(r[0]))
off = np.cumsum([0] + [len(r) for r in reads]).astype(np.uint64)
ent_pos = positions[np.concatenate(reads)]
quality = rng.integers(1, 60, len(ent_pos)).astype(np.int32)
depth = len(ent_pos) / n_var
print(f"{len(reads)} reads, {len(ent_pos)} entries over {n_var} variants (mean depth {depth:.1f}), max_cov {max_cov}")
t = time.perf_counter()
mine = readselect.select_reads_csr(off, ent_pos, quality, np.zeros(len(reads), int), max_cov)
t_mine = time.perf_counter() - t
print(f"whatshap_b200.readselect: {t_mine:.2f} s, {len(mine)} reads selected ({len(reads) / t_mine / 1e3:.0f} k reads/s)")
path = build_pyref.build()
if path:
    sys.path.insert(0, path)
    import whatshap.core as wc
    from whatshap.readselect import readselection

    rs = wc.ReadSet()
    for i, idx in enumerate(reads):
        read = wc.Read("r%d" % i, 50, 0, 0)
        lo = int(off[i])
        for k, v in enumerate(idx.tolist()):
            read.add_variant(int(positions[v]), 0, int(quality[lo + k]))
        rs.add(read)
    t = time.perf_counter()
    ref = readselection(rs, max_cov, None, True)
    t_ref = time.perf_counter() - t
    print(f"reference whatshap.readselect: {t_ref:.2f} s, identical selection: {ref == mine}, speed-up {t_ref / t_mine:.1f} x")

"""First GPU run of whmec_genotype (written without GPU access): parity tests, then timings against the compiled
reference on a coverage-15 single-individual workload and a trio.  Run:  gpurun --timeout 600 -- 'python scripts/gpu_genotype_check.py'"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from oracle import checker  # noqa: E402
from whatshap_b200 import _lib, synth  # noqa: E402

env = dict(os.environ, WHMEC_GPU_GENOTYPE="1")
res = subprocess.run([sys.executable, "-m", "pytest", "-q", "-m", "gpu", "tests/test_zz_gpu_genotype.py"], cwd=ROOT, env=env,
                     capture_output=True, text=True)
print(res.stdout[-1500:], res.stderr[-500:])
rng = np.random.default_rng(2)
for label, prob in (("single cov15 n=2000", synth.genotyping_problem(rng, 2000, 15, "single", prior="random", burst=8, mean_len=10.0)),
                    ("trio cov5 n=2000", synth.genotyping_problem(rng, 2000, 5, "trio", prior="random"))):
    _lib.genotype(prob)  # warm-up
    t0 = time.perf_counter()
    got, stats = _lib.genotype(prob)
    wall = time.perf_counter() - t0
    ck = checker.best()
    sub = prob  # the CPU reference handles these sizes in seconds
    t0 = time.perf_counter()
    want = ck.genotype(sub)
    cpu = time.perf_counter() - t0
    print(f"{label}: GPU {wall * 1e3:.1f} ms (device {stats['sweep_ms']:.1f} ms, {stats['kernel_launches']} launches), "
          f"{ck.kind} CPU {cpu * 1e3:.1f} ms, max abs diff {np.nanmax(np.abs(got - want)):.2e}")

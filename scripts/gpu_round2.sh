#!/bin/bash
# First GPU call of the next round (single B200, ~15-25 min): everything that was written in round 1 after the
# GPU minutes ran out, plus the usual gates.  Run:  gpurun --timeout 2400 -- 'bash scripts/gpu_round2.sh'
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --timeout 400 -p no:cacheprovider 2>&1 | tail -60 | tee gpurun_out/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
# group-wise pipeline of whmec_solve (csrc/grouped.h): bit-equality with the ordinary solve + end-to-end times
timeout 300 python scripts/gpu_grouped_check.py 2>&1 | tail -12 | tee gpurun_out/grouped_check.log
# fused per-chain pedigree sweep (ped_chain_kernel): bit-equality + sweep time against the batched sweep
timeout 300 python scripts/gpu_ped_chain_check.py 2>&1 | tail -8 | tee gpurun_out/ped_chain_check.log
# extremes of the value range / pedigree size on the CUDA path
timeout 300 python scripts/gpu_extremes_check.py 2>&1 | tail -20 | tee gpurun_out/extremes_check.log
# forward-backward genotyping DP (whmec_genotype): first run on a device; parity first, then timings
timeout 500 python scripts/gpu_genotype_check.py 2>&1 | tail -12 | tee gpurun_out/genotype_check.log
# ncu: launch list of the genotyping DP + one full capture of its forward kernel (only if the parity run above was green)
if grep -q " passed" gpurun_out/genotype_check.log && ! grep -q "failed" gpurun_out/genotype_check.log; then
  cat > gpurun_out/_gl_run.py <<'PY'
import numpy as np
from whatshap_b200 import _lib, synth
prob = synth.genotyping_problem(np.random.default_rng(2), 1000, 15, "single", prior="random", burst=8, mean_len=10.0)
_lib.genotype(prob); out, st = _lib.genotype(prob); print(st)
PY
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_genotype.csv python gpurun_out/_gl_run.py > /dev/null 2>&1
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:gl_forward_kernel -s 40 -c 1 -o gpurun_out/r02_gl_forward python gpurun_out/_gl_run.py > /dev/null 2>&1
fi
# headline bench with and without the pipeline (look at "e2e")
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_cfg3.json 2> gpurun_out/bench_cfg3.err
WHMEC_SOLVE_GROUPS=4 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg3_groups4.json 2> gpurun_out/bench_cfg3_groups4.err
# thread-packed back-pointers: parity fuzz first, then the bench
WHMEC_TILE_PACKED_BP=1 timeout 200 python scripts/gpu_fuzz_highcov.py 90 14 18 2>&1 | tail -3 | tee gpurun_out/fuzz_packed_bp.log
WHMEC_TILE_PACKED_BP=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg3_packedbp.json 2> gpurun_out/bench_cfg3_packedbp.err
# packed 16-bit steady-state panels: parity fuzz at coverage 16-19 first, then the bench
WHMEC_TILE_U16=1 timeout 200 python scripts/gpu_fuzz_highcov.py 90 16 19 2>&1 | tail -3 | tee gpurun_out/fuzz_u16.log
WHMEC_TILE_U16=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg3_u16.json 2> gpurun_out/bench_cfg3_u16.err
WHMEC_PINNED_STAGING=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg3_pinned.json 2> gpurun_out/bench_cfg3_pinned.err
# ragged blocks (load balance over the persistent tile grid)
timeout 300 python bench.py --workload cfg3g --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg3g.json 2> gpurun_out/bench_cfg3g.err
for f in gpurun_out/bench_cfg3.json gpurun_out/bench_cfg3_groups4.json gpurun_out/bench_cfg3_packedbp.json gpurun_out/bench_cfg3_u16.json gpurun_out/bench_cfg3_pinned.json gpurun_out/bench_cfg3g.json; do
  python - "$f" <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], "value %.0f" % d["value"], "ms/step %.2f" % d["ms_per_step"], "e2e %.0f (%.1f ms)" % (d["e2e"]["value"], d["e2e"]["ms_per_step"]))
PY
done

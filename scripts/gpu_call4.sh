#!/bin/bash
TAG=${1:-r02d}
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_configs.py tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_large.py tests/test_gpu_sharded.py -q -m gpu --timeout 400 -p no:cacheprovider -x -k "trio or pedigree or golden or fuzz_irregular or two_ranks" 2>&1 | tail -15 | tee gpurun_out/${TAG}_pytest_ped.log
timeout 300 python bench.py --workload cfg5 --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_cfg5.json 2> gpurun_out/${TAG}_bench_cfg5.err
WHMEC_PED_FUSED=0 timeout 300 python bench.py --workload cfg5 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_cfg5_batched.json 2>> gpurun_out/${TAG}_bench_cfg5.err
WHMEC_TIMING=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_cfg3.json 2> gpurun_out/${TAG}_bench_cfg3_timing.err
WHMEC_TIMING=1 WHMEC_SOLVE_GROUPS=4 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_cfg3_groups4.json 2> gpurun_out/${TAG}_bench_cfg3_groups4_timing.err
WHMEC_TIMING=1 timeout 300 python bench.py --workload cfg5 --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2> gpurun_out/${TAG}_bench_cfg5_timing.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${TAG}_launches_bench_cfg5.csv python bench.py --workload cfg5 --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ped_fused_kernel -s 2 -c 1 -o gpurun_out/${TAG}_pedfused_cfg5 python bench.py --workload cfg5 --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
cat > gpurun_out/_san2.py <<'PY'
import sys
sys.path.insert(0, ".")
from oracle import checker
from whatshap_b200 import _lib, synth
ck = checker.best()
for prob in (synth.trio(60, 3, block_len=30, seed=3), synth.trio(40, 5, block_len=20, seed=4), synth.trio(24, 2, block_len=6, seed=5)):
    got, st = _lib.solve(prob)
    assert got.same_as(ck.solve(prob))
    print("ok", st["path_kind"], st["max_active"], st["kernel_launches"])
PY
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python gpurun_out/_san2.py 2>&1 | tail -6 | tee gpurun_out/${TAG}_memcheck_ped.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python gpurun_out/_san2.py 2>&1 | tail -6 | tee gpurun_out/${TAG}_racecheck_ped.log
for f in gpurun_out/${TAG}_bench_*.json; do
  python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value %.0f" % d["value"], "ms/step %.2f" % d["ms_per_step"], "e2e %.0f (%.1f ms)" % (d["e2e"]["value"], d["e2e"]["ms_per_step"]), "launches", d["roofline"]["launches_per_step"])
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
done
tail -30 gpurun_out/${TAG}_bench_cfg3_timing.err

#!/bin/bash
TAG=${1:-r02h}
set -x
mkdir -p gpurun_out
timeout -k 5 150 python scripts/_san2.py 2>&1 | tail -4 | tee gpurun_out/${TAG}_trio_smoke.log
timeout -k 5 500 python -m pytest tests/test_gpu_configs.py tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_sharded.py tests/test_gpu_large.py -q -m gpu --timeout 200 --timeout-method=thread -p no:cacheprovider -x -k "trio or pedigree or golden or fuzz_irregular or two_ranks or segments" 2>&1 | tail -15 | tee gpurun_out/${TAG}_pytest_ped.log
for cs in 1 2 4; do
  WHMEC_PED_CLUSTER=$cs timeout -k 5 200 python bench.py --workload cfg5 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_cfg5_cluster$cs.json 2> gpurun_out/${TAG}_bench_cfg5_cluster$cs.err
done
timeout -k 5 300 compute-sanitizer --tool racecheck --error-exitcode 9 python scripts/_san2.py 2>&1 | tail -6 | tee gpurun_out/${TAG}_racecheck_ped.log
timeout -k 5 300 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/_san2.py 2>&1 | tail -6 | tee gpurun_out/${TAG}_memcheck_ped.log
timeout -k 5 400 ncu --set full --clock-control none --import-source on -k regex:ped_fused -s 2 -c 2 -o gpurun_out/${TAG}_pedfused_cfg5 python bench.py --workload cfg5 --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
for f in gpurun_out/${TAG}_bench_*.json; do
  python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value %.0f" % d["value"], "ms/step %.2f" % d["ms_per_step"], "e2e %.0f (%.1f ms)" % (d["e2e"]["value"], d["e2e"]["ms_per_step"]), "sweep/bt", d["roofline"].get("sweep_ms"), d["roofline"].get("backtrace_ms"))
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
done

#!/bin/bash
# GPU call 12: page-locked pool with size classes, create without a synchronisation -> e2e breakdown + bench; the tests that drive
# plans / segments / switches
TAG=${1:-r02k}
set -x
mkdir -p gpurun_out
timeout -k 5 400 python -m pytest tests/test_gpu_configs.py tests/test_gpu_large.py tests/test_gpu_sharded.py tests/test_gpu_real_containers.py -q -m gpu --timeout 200 --timeout-method=thread -p no:cacheprovider -x 2>&1 | tail -6 | tee gpurun_out/${TAG}_pytest_gpu_subset.log
timeout -k 5 300 python scripts/e2e_breakdown.py cfg3 cfg2 cfg5 > gpurun_out/${TAG}_e2e_breakdown.log 2>&1
timeout -k 5 250 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_cfg3.json 2> gpurun_out/${TAG}_bench_cfg3.err
for w in cfg2 cfg5; do
  timeout -k 5 250 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_$w.json 2> gpurun_out/${TAG}_bench_$w.err
done
for f in gpurun_out/${TAG}_bench_*.json; do
  python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value %.0f" % d["value"], "ms/step %.2f" % d["ms_per_step"], "e2e %.0f (%.1f ms)" % (d["e2e"]["value"], d["e2e"].get("ms_per_step", 0)), "api", d.get("e2e_api") and d["e2e_api"]["ms_per_step"], "issue", d.get("roofline_issue") and d["roofline_issue"]["frac"], "sweep/bt", d.get("roofline", {}).get("sweep_ms"), d.get("roofline", {}).get("backtrace_ms"))
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
done
grep -E "==|RESULT|solve:|create:|plan:|pack:" gpurun_out/${TAG}_e2e_breakdown.log | tail -70

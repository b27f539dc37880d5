#!/bin/bash
# every step under a hard limit (a device-side hang cannot be interrupted by pytest's signal-based timeout)
TAG=${1:-r02f}
set -x
mkdir -p gpurun_out
timeout -k 5 150 python scripts/_san2.py 2>&1 | tail -4 | tee gpurun_out/${TAG}_trio_smoke.log
timeout -k 5 900 python -m pytest tests -q -m gpu --timeout 200 --timeout-method=thread -p no:cacheprovider -x 2>&1 | tail -15 | tee gpurun_out/${TAG}_pytest_new.log
WHMEC_TIMING=1 timeout -k 5 200 python bench.py --workload cfg5 --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_cfg5.json 2> gpurun_out/${TAG}_bench_cfg5_timing.err
timeout -k 5 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_cfg3.json 2> gpurun_out/${TAG}_bench_cfg3.err


timeout -k 5 400 ncu --set full --clock-control none --import-source on -k regex:ped_fused_kernel -s 2 -c 2 -o gpurun_out/${TAG}_pedfused_cfg5 python bench.py --workload cfg5 --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
for f in gpurun_out/${TAG}_bench_*.json; do
  python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value %.0f" % d["value"], "ms/step %.2f" % d["ms_per_step"], "e2e %.0f (%.1f ms)" % (d["e2e"]["value"], d["e2e"]["ms_per_step"]), "api", d.get("e2e_api") and d["e2e_api"]["ms_per_step"], "issue", d.get("roofline_issue") and d["roofline_issue"]["frac"], "sweep/bt", d["roofline"].get("sweep_ms"), d["roofline"].get("backtrace_ms"))
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
done
tail -4 gpurun_out/${TAG}_bench_cfg5_timing.err

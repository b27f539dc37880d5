#!/bin/bash
# GPU call 11: page-locked uploads / cached streams / parallel assemble; split column barrier + compact steady records; speculative
# tile backtrace; columnar ReadSet.  Tests first (tile + config + large + golden), then the e2e breakdown and the benches.
TAG=${1:-r02j}
set -x
mkdir -p gpurun_out
timeout -k 5 600 python -m pytest tests -q -m gpu --timeout 200 --timeout-method=thread -p no:cacheprovider -x 2>&1 | tail -8 | tee gpurun_out/${TAG}_pytest_gpu.log
timeout -k 5 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/${TAG}_smoke.log
timeout -k 5 300 python scripts/e2e_breakdown.py cfg3 cfg2 cfg5 > gpurun_out/${TAG}_e2e_breakdown.log 2>&1
timeout -k 5 250 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_cfg3.json 2> gpurun_out/${TAG}_bench_cfg3.err
for w in cfg2 cfg5; do
  timeout -k 5 250 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_$w.json 2> gpurun_out/${TAG}_bench_$w.err
done
timeout -k 5 300 python bench.py --workload cfg4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_cfg4.json 2> gpurun_out/${TAG}_bench_cfg4.err
timeout -k 5 400 ncu --set full --clock-control none --import-source on -k regex:tile_panel_kernel -s 20 -c 1 -o gpurun_out/${TAG}_tile_cfg3 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
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
grep -E "RESULT|solve:|create:|tiles.create|plan:|pack:|finish:" gpurun_out/${TAG}_e2e_breakdown.log | awk 'NR%1==0' | tail -60

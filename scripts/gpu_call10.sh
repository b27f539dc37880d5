#!/bin/bash
# GPU call 10 (round 2, session 3): HEAD on hardware after the backtrace / pdep changes, e2e phase breakdown, fresh ncu captures of HEAD
TAG=${1:-r02i}
set -x
mkdir -p gpurun_out
timeout -k 5 600 python -m pytest tests -q -m gpu --timeout 200 --timeout-method=thread -p no:cacheprovider -x 2>&1 | tail -8 | tee gpurun_out/${TAG}_pytest_gpu.log
timeout -k 5 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/${TAG}_smoke.log
timeout -k 5 400 python scripts/e2e_breakdown.py cfg3 cfg2 cfg5 > gpurun_out/${TAG}_e2e_breakdown.log 2>&1
timeout -k 5 250 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_cfg3.json 2> gpurun_out/${TAG}_bench_cfg3.err
for w in cfg2 cfg5; do
  timeout -k 5 250 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_$w.json 2> gpurun_out/${TAG}_bench_$w.err
done
timeout -k 5 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_bench_cfg3.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
timeout -k 5 400 ncu --set full --clock-control none --import-source on -k regex:tile_panel_kernel -s 20 -c 1 -o gpurun_out/${TAG}_tile_cfg3 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
timeout -k 5 400 ncu --set full --clock-control none --import-source on -k regex:ped_fused_cluster -s 2 -c 2 -o gpurun_out/${TAG}_pedfused_cfg5 python bench.py --workload cfg5 --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out | tail -20
for f in gpurun_out/${TAG}_bench_*.json; do
  python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value %.0f" % d["value"], "ms/step %.2f" % d["ms_per_step"], "e2e %.0f (%.1f ms)" % (d["e2e"]["value"], d["e2e"].get("ms_per_step", 0)), "api", d.get("e2e_api") and d["e2e_api"]["ms_per_step"], "issue", d.get("roofline_issue") and d["roofline_issue"]["frac"], "sweep/bt", d.get("roofline", {}).get("sweep_ms"), d.get("roofline", {}).get("backtrace_ms"), "clocks", d.get("clocks"))
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
done
cat gpurun_out/${TAG}_e2e_breakdown.log | tail -80

#!/bin/bash
# Final GPU call of round 2: every -m gpu test, smoke, ncu captures of HEAD (all four workloads) -> traffic.json, bench lines,
# reference arm, launch lists, compute-sanitizer, host -> device rate of the box.
TAG=${1:-r02m}
set -x
mkdir -p gpurun_out
timeout -k 5 600 python -m pytest tests -q -m gpu --timeout 200 --timeout-method=thread -p no:cacheprovider -x 2>&1 | tail -8 | tee gpurun_out/${TAG}_pytest_gpu.log
timeout -k 5 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/${TAG}_smoke.log
# ncu --set full of the dominant kernel of every workload (one launch from the steady state of a sweep)
timeout -k 5 300 ncu --set full --clock-control none --import-source on -k regex:tile_panel_kernel -s 20 -c 1 -o gpurun_out/${TAG}_tile_cfg3 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
timeout -k 5 300 ncu --set full --clock-control none --import-source on -k regex:ped_fused_cluster -s 2 -c 2 -o gpurun_out/${TAG}_pedfused_cfg5 python bench.py --workload cfg5 --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
timeout -k 5 200 ncu --set full --clock-control none --import-source on -k regex:tile_panel_kernel -s 3 -c 1 -o gpurun_out/${TAG}_tile_cfg2 python bench.py --workload cfg2 --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
timeout -k 5 300 ncu --set full --clock-control none --import-source on -k regex:tile_panel_kernel -s 2000 -c 1 -o gpurun_out/${TAG}_tile_cfg4 python bench.py --workload cfg4 --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
python profiles/summarize.py ${TAG} > gpurun_out/${TAG}_ncu_summaries.txt 2>&1
cp profiles/traffic.json gpurun_out/${TAG}_traffic.json
cp profiles/${TAG}_*.txt gpurun_out/ 2>/dev/null
# bench lines (the default run of the driver is cfg3)
timeout -k 5 250 python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_cfg3.json 2> gpurun_out/${TAG}_bench_cfg3.err
for w in cfg2 cfg5; do
  timeout -k 5 250 python bench.py --workload $w --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_$w.json 2> gpurun_out/${TAG}_bench_$w.err
done
timeout -k 5 300 python bench.py --workload cfg4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_cfg4.json 2> gpurun_out/${TAG}_bench_cfg4.err
timeout -k 5 250 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_cfg3_reference.json 2> gpurun_out/${TAG}_bench_cfg3_reference.err
timeout -k 5 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_bench_cfg3.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
timeout -k 5 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${TAG}_launches_bench_cfg5.csv python bench.py --workload cfg5 --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
timeout -k 5 60 python scripts/h2d_rate.py 2>&1 | tee gpurun_out/${TAG}_h2d_rate.log
timeout -k 5 120 python scripts/e2e_breakdown.py cfg3 2>&1 | grep -E "==|RESULT|solve:|create:" | tail -12 | tee gpurun_out/${TAG}_e2e_breakdown.log
timeout -k 5 200 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/_san.py 2>&1 | tail -8 | tee gpurun_out/${TAG}_memcheck.log
timeout -k 5 250 compute-sanitizer --tool racecheck --error-exitcode 9 python scripts/_san.py 2>&1 | tail -8 | tee gpurun_out/${TAG}_racecheck.log
for f in gpurun_out/${TAG}_bench_*.json; do
  python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value %.0f" % d["value"], "ms/step %.2f" % d["ms_per_step"], "e2e %.0f" % d["e2e"]["value"], d["e2e"].get("ms_per_step"), "api", d.get("e2e_api") and d["e2e_api"]["ms_per_step"], "issue", d.get("roofline_issue") and d["roofline_issue"]["frac"], "cpu", d.get("cpu_baseline", {}).get("value"))
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
done

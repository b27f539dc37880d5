#!/bin/bash
# Standard GPU call of round 2: parity gates, bench lines, ncu launch list + one full capture of the dominant kernel.
#   gpurun --timeout 1800 -- 'bash scripts/gpu_call.sh [tag]'
TAG=${1:-r02}
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --timeout 400 -p no:cacheprovider -x 2>&1 | tail -15 | tee gpurun_out/${TAG}_pytest_gpu.log
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_cfg3.json 2> gpurun_out/${TAG}_bench_cfg3.err
WHMEC_TILE_MIRROR=0 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_cfg3_nomirror.json 2>> gpurun_out/${TAG}_bench_cfg3.err
for w in cfg2 cfg4 cfg5; do
  timeout 600 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_$w.json 2> gpurun_out/${TAG}_bench_$w.err
done
# ncu: launch list of the bench command, then one full capture of a mid-sweep launch of the tile kernel
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_bench_cfg3.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tile_panel_kernel -s 20 -c 1 -o gpurun_out/${TAG}_tile_cfg3 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
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

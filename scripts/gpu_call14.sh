#!/bin/bash
# GPU call 14: racecheck on the split column barrier -- one arrival per warp (default build) vs one per thread (variant build) --
# with full logs; sweep time of both builds; host -> device rate of freshly written page-locked memory
TAG=${1:-r02n}
set -x
mkdir -p gpurun_out
timeout -k 5 250 compute-sanitizer --tool racecheck --racecheck-report analysis python scripts/_san.py > gpurun_out/${TAG}_racecheck_warp_arrive.log 2>&1
WHMEC_LIBRARY=$PWD/whatshap_b200/libwhmec_arriveall.so timeout -k 5 250 compute-sanitizer --tool racecheck --racecheck-report analysis python scripts/_san.py > gpurun_out/${TAG}_racecheck_thread_arrive.log 2>&1
tail -4 gpurun_out/${TAG}_racecheck_warp_arrive.log gpurun_out/${TAG}_racecheck_thread_arrive.log
timeout -k 5 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_cfg3_warp_arrive.json 2>/dev/null
WHMEC_LIBRARY=$PWD/whatshap_b200/libwhmec_arriveall.so timeout -k 5 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_cfg3_thread_arrive.json 2>/dev/null
WHMEC_LIBRARY=$PWD/whatshap_b200/libwhmec_arriveall.so timeout -k 5 200 python -m pytest tests/test_gpu_configs.py -q -m gpu -x -k "single_individual" --timeout 200 -p no:cacheprovider 2>&1 | tail -3 | tee gpurun_out/${TAG}_pytest_thread_arrive.log
timeout -k 5 60 python scripts/h2d_rate.py 2>&1 | tee gpurun_out/${TAG}_h2d_rate.log
for f in gpurun_out/${TAG}_bench_*.json; do
  python - "$f" <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], "value %.0f" % d["value"], "sweep", d["roofline"]["sweep_ms"], "bt", d["roofline"]["backtrace_ms"], "e2e", d["e2e"]["ms_per_step"])
PY
done

#!/bin/bash
# Last GPU call of round 2: the default build (one barrier arrival per thread) through every -m gpu test, racecheck + memcheck logs, one bench line
TAG=${1:-r02o}
set -x
mkdir -p gpurun_out
timeout -k 5 400 python -m pytest tests -q -m gpu --timeout 200 --timeout-method=thread -p no:cacheprovider -x 2>&1 | tail -5 | tee gpurun_out/${TAG}_pytest_gpu.log
timeout -k 5 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/${TAG}_smoke.log
timeout -k 5 100 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_cfg3.json 2> gpurun_out/${TAG}_bench_cfg3.err
timeout -k 5 150 compute-sanitizer --tool racecheck --racecheck-report analysis python scripts/_san.py > gpurun_out/${TAG}_racecheck.log 2>&1
timeout -k 5 100 compute-sanitizer --tool memcheck python scripts/_san.py > gpurun_out/${TAG}_memcheck.log 2>&1
grep -E "SUMMARY|^ok" gpurun_out/${TAG}_racecheck.log gpurun_out/${TAG}_memcheck.log
python - gpurun_out/${TAG}_bench_cfg3.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value %.0f" % d["value"], "sweep", d["roofline"]["sweep_ms"], "bt", d["roofline"]["backtrace_ms"], "e2e", d["e2e"]["ms_per_step"], "api", d["e2e_api"]["ms_per_step"])
PY

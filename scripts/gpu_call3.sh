#!/bin/bash
TAG=${1:-r02c}
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --timeout 400 -p no:cacheprovider -x 2>&1 | tail -15 | tee gpurun_out/${TAG}_pytest_gpu.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_cfg3.json 2> gpurun_out/${TAG}_bench_cfg3.err
for w in cfg2 cfg4 cfg3g; do
  timeout 600 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_$w.json 2> gpurun_out/${TAG}_bench_$w.err
done
cat > gpurun_out/_san.py <<'PY'
import sys
sys.path.insert(0, ".")
from oracle import checker
from whatshap_b200 import _lib, synth
ck = checker.best()
for prob in (synth.sliding_window(60, 17, block_len=60, seed=3), synth.sliding_window(48, 20, block_len=48, seed=5),
             synth.sliding_window(90, 16, block_len=45, seed=3, gap=0.1, max_phred=3), synth.sliding_window(300, 12, block_len=100, seed=1),
             synth.trio(60, 3, block_len=30, seed=3)):
    got, st = _lib.solve(prob)
    assert got.same_as(ck.solve(prob))
    print("ok", st["path_kind"], st["max_active"])
PY
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python gpurun_out/_san.py 2>&1 | tail -8 | tee gpurun_out/${TAG}_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python gpurun_out/_san.py 2>&1 | tail -8 | tee gpurun_out/${TAG}_racecheck.log
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

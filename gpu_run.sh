mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches_bench_cfg3.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
cat > prof_tmp.py <<'PY'
import sys
sys.path.insert(0,'.')
from whatshap_b200 import synth, _lib
name=sys.argv[1]
p=synth.config(name, int(sys.argv[2]))
plan=_lib.Plan(p)
for _ in range(2): plan.sweep()
plan.finish(); plan.close()
PY
ncu --set full --clock-control none --import-source on -k regex:tile_panel -s 50 -c 1 -o gpurun_out/r01_tile_cfg3 python prof_tmp.py cfg3 50000 > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tile_panel -s 1 -c 1 -o gpurun_out/r01_tile_cfg2 python prof_tmp.py cfg2 10000 > gpurun_out/ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tile_panel -s 30 -c 1 -o gpurun_out/r01_tile_cfg4 python prof_tmp.py cfg4 700 > gpurun_out/ncu3.log 2>&1
rm -f prof_tmp.py
ls -la gpurun_out | tail -8

mkdir -p gpurun_out
cat > /tmp/prof.py <<'PY'
import sys
sys.path.insert(0,'.')
from whatshap_b200 import synth, _lib
p=synth.config('cfg2',2000)
plan=_lib.Plan(p)
for _ in range(2): plan.sweep()
plan.finish(); plan.close()
PY
ncu --set full --clock-control none --import-source on -k regex:tile_panel -s 1 -c 1 -o gpurun_out/tile_r1c python /tmp/prof.py > gpurun_out/ncu_c.log 2>&1
tail -2 gpurun_out/ncu_c.log

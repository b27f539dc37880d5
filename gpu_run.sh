python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_large.py tests/test_pedigreephasing.py -x -q -m gpu 2>&1 | tail -6
python - <<'PY'
import time, sys, json
sys.path.insert(0,'.')
from whatshap_b200 import synth, _lib
for name, n in [('cfg5',20000)]:
    p=synth.config(name,n)
    plan=_lib.Plan(p)
    for i in range(3): plan.sweep()
    st=plan.stats(); sol=plan.finish(); plan.close()
    print(name,n,'sweep %.3f ms'%st['sweep_ms'], 'cols/s=%.0f'%(n/(st['sweep_ms']/1e3)), 'launches',st['kernel_launches'],'path',st['path_kind'],'cost',sol.cost, flush=True)
PY

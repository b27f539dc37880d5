python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py -x -q -m gpu 2>&1 | tail -4
python - <<'PY'
import time, sys, json
sys.path.insert(0,'.')
from whatshap_b200 import synth, _lib
from oracle import checker
# larger parity check of the multi-tile path against the compiled reference (coverage 18, 2 chains)
p=synth.sliding_window(120,18,block_len=60,seed=77)
a=checker.best().solve(p); b,_=_lib.solve(p); print('cov18 parity', a.same_as(b), a.cost)
for name, n in [('cfg2',10000),('cfg3',50000),('cfg4',1000)]:
    p=synth.config(name,n)
    plan=_lib.Plan(p)
    for _ in range(3): plan.sweep()
    st=plan.stats(); sol=plan.finish(); plan.close()
    print(name,n,'sweep %.3f ms'%st['sweep_ms'], 'cols/s=%.0f'%(n/(st['sweep_ms']/1e3)), 'launches',st['kernel_launches'],'cost',sol.cost, 'alg GB/s %.0f'%(st['algorithmic_bytes']/st['sweep_ms']/1e6), flush=True)
PY

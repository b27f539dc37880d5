#!/bin/bash
# gpurun --gpus N --timeout 900 -- 'bash scripts/gpu_multi.sh N'
N=${1:-2}
set -x
mkdir -p gpurun_out
nvidia-smi -L | head -8
for w in cfg3 cfg5; do
  timeout -k 5 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --workload $w --steps 3 --warmup 3 > gpurun_out/r02_scale_${w}_n$N.json 2> gpurun_out/r02_scale_${w}_n$N.err
  python - gpurun_out/r02_scale_${w}_n$N.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "n_gpus", d["n_gpus"], "value %.0f" % d["value"], "ms/step %.2f" % d["ms_per_step"], "e2e %.0f (%.1f ms)" % (d["e2e"]["value"], d["e2e"]["ms_per_step"]), "sharded", d.get("e2e_sharded"))
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
  tail -3 gpurun_out/r02_scale_${w}_n$N.err
done

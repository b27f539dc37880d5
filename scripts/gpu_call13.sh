#!/bin/bash
# GPU call 13 (2 GPUs): the driver's N > 1 launch of bench.py after the host-side changes (page-locked pool per rank, cached streams)
TAG=${1:-r02l}
set -x
mkdir -p gpurun_out
timeout -k 5 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/${TAG}_scale_cfg3_n2.json 2> gpurun_out/${TAG}_scale_cfg3_n2.err
timeout -k 5 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --workload cfg5 > gpurun_out/${TAG}_scale_cfg5_n2.json 2> gpurun_out/${TAG}_scale_cfg5_n2.err
tail -3 gpurun_out/${TAG}_scale_cfg3_n2.err gpurun_out/${TAG}_scale_cfg5_n2.err
for f in gpurun_out/${TAG}_scale_*.json; do
  python - "$f" <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith("{")][-1])
    print(sys.argv[1], "value %.0f" % d["value"], "ms/step %.2f" % d["ms_per_step"], "e2e %.0f (%.1f ms)" % (d["e2e"]["value"], d["e2e"].get("ms_per_step", 0)), "sharded", d.get("e2e_sharded") and (d["e2e_sharded"]["ms_per_step"], d["e2e_sharded"]["phases_ms_rank0"]), "clocks", d.get("clocks"))
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
done

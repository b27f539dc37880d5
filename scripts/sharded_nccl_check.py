"""Run under torchrun with one rank per GPU:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/sharded_nccl_check.py
Rank 0 builds a problem, `multigpu.solve_sharded` broadcasts it over NCCL, every rank solves its blocks
on its own GPU, rank 0 gathers and compares with the single-GPU solve."""
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from whatshap_b200 import _lib, multigpu, synth  # noqa: E402

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
prob = synth.config("cfg3", 8000) if rank == 0 else None
multigpu.solve_sharded(synth.config("cfg2", 1000) if rank == 0 else None)  # warm-up
torch.cuda.synchronize()
t0 = time.perf_counter()
sol = multigpu.solve_sharded(prob)
dt = time.perf_counter() - t0
if rank == 0:
    whole, _ = _lib.solve(prob, device=local)
    print("sharded == single-GPU:", sol.same_as(whole), "cost", sol.cost, "world", dist.get_world_size(), "%.1f ms" % (dt * 1e3))
    assert sol.same_as(whole)
# pedigree (T = 4): the ranks hold segments of ONE table and exchange transfer matrices / exit tables
ped = synth.config("cfg5") if rank == 0 else None
multigpu.solve_sharded(synth.config("cfg5", 2000) if rank == 0 else None)  # warm-up
torch.cuda.synchronize()
os.environ["WHMEC_TIMING"] = "1"
t0 = time.perf_counter()
sol = multigpu.solve_sharded(ped)
dt = time.perf_counter() - t0
os.environ.pop("WHMEC_TIMING")
if rank == 0:
    t1 = time.perf_counter()
    whole, _ = _lib.solve(ped, device=local)
    one = time.perf_counter() - t1
    print("pedigree segments == single-GPU:", sol.same_as(whole), "cost", sol.cost, "world", dist.get_world_size(),
          "%.1f ms (single GPU, host in/out: %.1f ms)" % (dt * 1e3, one * 1e3))
    assert sol.same_as(whole)
dist.destroy_process_group()

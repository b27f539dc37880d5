"""Block sharding with the real CUDA solver: two ranks (gloo rendezvous, both on cuda:0 so that the
test runs on a single-GPU box) solve their shares through the C ABI, rank 0 merges; must equal the
unsharded CUDA solve bit for bit.  With more GPUs the same code runs one rank per GPU over NCCL
(`scripts/sharded_nccl_check.py`)."""
import os
import socket

import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    import sys

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from whatshap_b200 import _lib, multigpu, synth

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    prob = synth.config("cfg2", 3000) if rank == 0 else None
    sol = multigpu.solve_sharded(prob, solver=lambda p: _lib.solve(p, device=0)[0])
    if rank == 0:
        whole, _ = _lib.solve(prob, device=0)
        q.put((sol.same_as(whole), sol.diff(whole), int(sol.cost)))
    dist.destroy_process_group()


def test_two_ranks_share_one_problem(gpu):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok, diff, cost = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert ok, diff
    assert cost > 0

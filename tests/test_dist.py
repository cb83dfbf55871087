"""CPU, world_size 2 over gloo: the N>1 plumbing (weight-arena broadcast, rank sharding, max-over-ranks)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from orv_b200 import dist as D
    r, lr, w = D.init_from_env("gloo")
    assert (r, w) == (rank, world)
    arena = torch.arange(1000, dtype=torch.float32).bfloat16() if rank == 0 else torch.zeros(1000, dtype=torch.bfloat16)
    D.broadcast_arena(arena, src=0)
    ok = torch.equal(arena, torch.arange(1000, dtype=torch.float32).bfloat16())
    mx = D.max_over_ranks(float(rank + 1), torch.device("cpu"))
    q.put((rank, ok, mx, D.shard_range(7, rank, world)))
    dist.barrier()
    dist.destroy_process_group()


def test_broadcast_and_sharding_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0] == (0, True, 2.0, (0, 3)) and res[1] == (1, True, 2.0, (3, 7))

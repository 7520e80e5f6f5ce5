"""world_size-2 gloo tests (CPU) of the N>1 host logic: replica sharding rule, the single gradient all-reduce,
max-over-ranks timing."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ultravox_b200 import dist_utils as du
    g = torch.full((1000,), float(rank + 1))
    du.allreduce_mean_(g)
    t = du.max_over_ranks(10.0 + rank)
    mine = du.shard_indices(11, du.rank(), du.world_size())
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    if rank == 0:
        out.put((float(g[0]), float(g.std()), t, gathered))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_allreduce_mean_and_sharding():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    mean, std, tmax, gathered = q.get(timeout=100)
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    assert mean == 1.5 and std == 0.0 and tmax == 11.0
    assert sorted(gathered[0] + gathered[1]) == list(range(11)) and gathered[0] == [0, 2, 4, 6, 8, 10]

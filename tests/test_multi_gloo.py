"""world_size-2 gloo test (CPU) of the multi-GPU host logic: every rank derives the SAME static LPT
partition from the job list, takes its own shard, and the only exchange is the final host gather —
no data-path collective."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _rank_main(rank, world, port, tmp):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from manifoldem_python_b200 import partition
    rng = np.random.default_rng(7)                      # same seed on every rank -> same job list
    nS = rng.integers(100, 2001, size=37)
    costs = [partition.pd_cost(int(n), 256) for n in nS]
    shards = partition.lpt_partition(costs, world)
    mine = shards[rank]
    for prD in mine:                                    # stand-in for the per-PD work: touch the marker
        open(os.path.join(tmp, str(prD)), 'a').close()
    pairs = float(sum(int(nS[i]) ** 2 for i in mine))
    t = torch.tensor([pairs, float(len(mine))], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)            # what bench.py does with its per-rank counters
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)              # the final host gather
    if rank == 0:
        flat = sorted(i for g in gathered for i in g)
        assert flat == list(range(37))
        assert sorted(int(f) for f in os.listdir(tmp)) == list(range(37))
        assert t[1].item() == 37 and t[0].item() == float(sum(int(n) ** 2 for n in nS))
        assert partition.imbalance(costs, shards) < 1.05
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_partition_and_gather(tmp_path):
    world = 2
    mp.spawn(_rank_main, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)

"""N > 1 host logic on CPU: sharding plan + the counter all-gather over gloo, world_size 2."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sonicsim_b200 import shard


def test_shard_plan_is_a_partition_and_balanced():
    # cfg3: 64 utterances x (2 moving + 2 static)
    costs = []
    for _ in range(64):
        costs += [shard.unit_cost(960000, 60, 6, 4096, True)] * 2 + [shard.unit_cost(960000, 1, 6, 4096, False)] * 2
    for world in (1, 2, 4, 8):
        parts = [shard.shard_units(costs, world, r) for r in range(world)]
        allu = sorted(i for p in parts for i in p)
        assert allu == list(range(len(costs)))
        tot = [sum(costs[i] for i in p) for p in parts]
        assert max(tot) / min(tot) < 1.02
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    costs = [3.0, 1.0, 3.0, 1.0, 2.0, 2.0]
    mine = shard.shard_units(costs, world, rank)
    c = shard.gather_counters(30.0 * len(mine), 1.0 + rank, 1e6 * len(mine))
    q.put((rank, mine, c.tolist()))
    dist.destroy_process_group()


def test_counter_all_gather_gloo_world2():
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
    (r0, m0, c0), (r1, m1, c1) = res
    assert sorted(m0 + m1) == list(range(6))
    assert c0 == c1                                  # every rank sees the same gathered table
    c = np.array(c0)
    thr, _ = shard.aggregate_throughput(c)
    assert abs(thr - 30.0 * 6 / 2.0) < 1e-9          # total audio / max elapsed over ranks

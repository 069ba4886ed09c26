"""N > 1 host logic on CPU: world_size-2 gloo (127.0.0.1 rendezvous)."""
import os
import socket

import pytest
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
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from mavlm_b200 import dist as D
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # video sharding: 5 videos over 2 ranks -> [0,1,2] and [3,4]; gathered in video order
        mine = D.shard_range(5, rank, world)
        local = torch.tensor([[float(v), 10.0 * v] for v in mine])
        counts = [len(D.shard_range(5, r, world)) for r in range(world)]
        full = D.all_gather_rows(local, counts)
        ok1 = torch.equal(full, torch.tensor([[float(v), 10.0 * v] for v in range(5)]))
        # equal shards take the single all-gather path
        local2 = torch.full((3, 4), float(rank))
        full2 = D.all_gather_rows(local2)
        ok2 = torch.equal(full2, torch.cat([torch.zeros(3, 4), torch.ones(3, 4)]))
        # timings are the max over ranks
        ok3 = D.max_over_ranks(1.0 + rank) == 2.0
        # empty shard (1 video on 2 ranks)
        mine1 = D.shard_range(1, rank, world)
        loc = torch.ones((len(mine1), 2))
        full3 = D.all_gather_rows(loc, [len(D.shard_range(1, r, world)) for r in range(world)])
        ok4 = full3.shape == (1, 2)
        # the dropout_frames coin (llava_arch.py:378-386): different local RNG states, one shared decision per draw
        torch.manual_seed(100 + 17 * rank)
        draws = [D.synced_dropout_decision(0.5) for _ in range(16)]
        gathered = [None, None]
        dist.all_gather_object(gathered, draws)
        ok5 = gathered[0] == gathered[1] and any(draws) and not all(draws)
        q.put((rank, ok1 and ok2 and ok3 and ok4 and ok5))
    finally:
        dist.destroy_process_group()


def test_shard_range_is_a_partition():
    from mavlm_b200 import dist as D
    for n in (0, 1, 5, 8, 64, 1024):
        for world in (1, 2, 4, 8):
            got = [i for r in range(world) for i in D.shard_range(n, r, world)]
            assert got == list(range(n))
            sizes = [len(D.shard_range(n, r, world)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.timeout(120)
def test_two_rank_gloo_gather_and_timing():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=100) for _ in range(2)]
    for p in procs:
        p.join(timeout=30)
    assert sorted(res) == [(0, True), (1, True)]

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
        # frame-sharded pre-pass of one long video: pieces of `world` chunks, one chunk per rank per piece, gathered in
        # place piece by piece; the buffer ends up in frame order on every rank (ragged tail: 70 frames, chunk 16)
        n_frames, chunk = 70, 16
        piece, sched = D.piece_schedule(n_frames, chunk, world)
        z = torch.full((len(sched) * piece, 3), -1.0)
        for j, row in enumerate(sched):
            s0, s1 = row[rank]
            slot = z[j * piece + rank * chunk: j * piece + (rank + 1) * chunk]
            slot[: s1 - s0] = torch.arange(s0, s1, dtype=torch.float32)[:, None]
            D.gather_piece(z[j * piece:(j + 1) * piece], slot)
        ok6 = torch.equal(z[:n_frames, 0], torch.arange(n_frames, dtype=torch.float32)) and piece == world * chunk
        q.put((rank, ok1 and ok2 and ok3 and ok4 and ok5 and ok6))
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


def test_piece_schedule_partitions_the_frames_chunk_by_chunk():
    from mavlm_b200 import dist as D
    for n in (1, 31, 32, 70, 256, 1024, 1000):
        for chunk in (8, 16, 32):
            for world in (1, 2, 4, 8):
                piece, sched = D.piece_schedule(n, chunk, world)
                assert piece == world * chunk and len(sched) == -(-n // piece)
                got = [i for row in sched for (a, b) in row for i in range(a, b)]
                assert got == list(range(n))                            # piece-major, rank-minor = time order
                for j, row in enumerate(sched):
                    for r, (a, b) in enumerate(row):
                        assert b - a <= chunk and (a == b or a == (j * world + r) * chunk)


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

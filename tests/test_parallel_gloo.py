"""N>1 path on CPU: world_size-2 gloo run of the batch sharding + score-ciphertext all-gather + max-over-ranks timing."""
import os
import socket

import numpy as np
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
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cryptonets_b200.parallel import batches_of_rank, gather_score_ciphertexts, max_over_ranks, row_slice
    mine = batches_of_rank(5, rank, world)
    # row-sharded dense layer on CPU: every rank sums its slice of "masked rows" modulo p, the partials are all-gathered and added locally
    # (modular addition is not a collective reduction) -- the exchange pattern of parallel.allreduce_ciphertext_sum
    p_mod, n_rows = 1000003, 5488
    first, count = row_slice(n_rows, rank, world)
    rows = (torch.arange(n_rows, dtype=torch.int64) * 7919 + 13) % p_mod
    partial = (rows[first:first + count].sum() % p_mod).reshape(1)
    parts = gather_score_ciphertexts(partial)
    total = int(sum(int(x) for x in parts) % p_mod)
    assert total == int(rows.sum() % p_mod)
    words = torch.arange(20, dtype=torch.int64) + 1000 * rank  # stands for 10 x P raw score ciphertext words
    got = gather_score_ciphertexts(words)
    t = max_over_ranks(1.0 + rank)
    q.put((rank, mine, [g.tolist() for g in got], t))
    dist.destroy_process_group()


def test_two_rank_gather_and_sharding():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(60)
    assert [r[1] for r in res] == [[0, 2, 4], [1, 3]]  # every batch owned exactly once
    from cryptonets_b200.parallel import row_slice
    assert [row_slice(5488, r, 4) for r in range(4)] == [(0, 1372), (1372, 1372), (2744, 1372), (4116, 1372)]
    assert [row_slice(10, r, 4) for r in range(4)] == [(0, 3), (3, 3), (6, 2), (8, 2)]
    for r in res:
        assert r[2][0] == list(range(20)) and r[2][1] == list(range(1000, 1020))
        assert r[3] == 2.0

"""CPU-only, world_size 2 over gloo: the multi-GPU host logic (contiguous ray shards + one all-gather of hit records)
reassembles exactly the single-process result."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, count, result):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from atlas_engine_b200 import sharding
    rng = np.random.default_rng(123)
    full = torch.from_numpy(rng.random((count, 12), dtype=np.float32))   # stands in for the traced PackedRay buffer
    b, e = sharding.shard_bounds(count, rank, world, 64)
    gathered = sharding.gather_hits_ragged(full[b:e].clone(), count, 64)
    ok = torch.equal(gathered, sharding.hit_records(full))
    if count % (64 * world) == 0:
        ok = ok and torch.equal(sharding.gather_hits(full[b:e].clone()), sharding.hit_records(full))
    result[rank] = bool(ok)
    dist.destroy_process_group()


def test_two_rank_gather_reassembles_global_order():
    for count in (64 * 2 * 50, 10_007):
        port = _free_port()
        mgr = mp.Manager()
        result = mgr.dict()
        mp.spawn(_worker, args=(2, port, count, result), nprocs=2, join=True)
        assert result[0] and result[1]


def test_blas_owner_round_robin():
    from atlas_engine_b200 import sharding
    owners = [sharding.blas_owner(i, 4) for i in range(10)]
    assert owners == [0, 1, 2, 3, 0, 1, 2, 3, 0, 1]


def _exchange_worker(rank, world, port, result):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from atlas_engine_b200 import sharding
    count = 7

    def tree(k):   # deterministic stand-in for a flattened BLAS of size depending on k (k == 3: a leaf-root BLAS, no nodes)
        g = torch.Generator().manual_seed(100 + k)
        n = 0 if k == 3 else 5 + 3 * k
        return (torch.randint(-2**31, 2**31 - 1, (n, 14), generator=g, dtype=torch.int32),
                torch.randint(0, 1000, (n + 1,), generator=g, dtype=torch.int32),
                torch.randint(0, 2, (n + 1,), generator=g, dtype=torch.uint8))
    local = {k: tree(k) for k in range(count) if sharding.blas_owner(k, world) == rank}
    got = sharding.exchange_flat_trees(local, count, torch.device("cpu"))
    ok = all(torch.equal(a, b) for k in range(count) for a, b in zip(got[k], tree(k)))
    result[rank] = bool(ok)
    dist.destroy_process_group()


def test_flat_trees_are_exchanged_between_ranks():
    port = _free_port()
    mgr = mp.Manager()
    result = mgr.dict()
    mp.spawn(_exchange_worker, args=(2, port, result), nprocs=2, join=True)
    assert result[0] and result[1]

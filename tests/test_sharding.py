"""N > 1 path on CPU: two gloo ranks shard a batch of passes round robin, "prove" them (here: the hashlib
expectations of the public signals, the host-side part of the job) and rank 0 reassembles the batch in order."""
import hashlib
import os
import socket
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_items, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    from nzcb_circom_b200 import nzcp_helpers as H
    from nzcb_circom_b200.sharding import gather_results, shard_indices

    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard_indices(n_items, rank, world)
    local = []
    for i in mine:
        p = H.synth_pass(1000 + i)
        local.append((i, hashlib.sha256(p["toBeSigned"]).hexdigest(), p["exp"]))
    dist.barrier()
    full = gather_results(local, n_items, rank, world, dist)
    if rank == 0:
        q.put(full)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_items", [7, 8])
def test_two_rank_sharding_gloo(n_items):
    import torch.multiprocessing as mp
    from nzcb_circom_b200 import nzcp_helpers as H

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_items, q)) for r in range(2)]
    for p in procs:
        p.start()
    full = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [f[0] for f in full] == list(range(n_items))
    for i, digest, exp in full:
        p = H.synth_pass(1000 + i)
        assert digest == hashlib.sha256(p["toBeSigned"]).hexdigest() and exp == p["exp"]


def test_shard_indices_cover_and_balance():
    from nzcb_circom_b200.sharding import shard_indices

    for n in (0, 1, 7, 8, 1024):
        for w in (1, 2, 4, 8):
            parts = [shard_indices(n, r, w) for r in range(w)]
            assert sorted(i for p in parts for i in p) == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    with pytest.raises(ValueError):
        shard_indices(4, 2, 2)

"""world_size-2 gloo test of the request sharding / ordered gather (the only multi-GPU logic: no in-loop collective)."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from instructany2pix_b200.parallel import gather_in_order, request_seed, run_sharded, shard_indices


def _work(indices):
    # deterministic per-request "trajectory": depends only on the GLOBAL request index
    out = []
    for i in indices:
        g = torch.Generator().manual_seed(request_seed(1000, i))
        out.append(torch.randn(4, 8, 8, generator=g))
    return torch.stack(out) if out else torch.empty(0, 4, 8, 8)


def _worker(rank, world, n_items, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    res = run_sharded(n_items, batch=2, work=_work)
    if rank == 0:
        q.put(res)
    dist.barrier()
    dist.destroy_process_group()


def test_shard_indices_cover_everything_once():
    for n in (0, 1, 7, 8, 9):
        for w in (1, 2, 4, 8):
            allidx = sorted(i for r in range(w) for i in shard_indices(n, r, w))
            assert allidx == list(range(n))
            sizes = [len(shard_indices(n, r, w)) for r in range(w)]
            assert max(sizes) - min(sizes) <= 1
            for block in (2, 4):
                per_rank = [shard_indices(n, r, w, block) for r in range(w)]
                assert sorted(i for idx in per_rank for i in idx) == list(range(n))
                # whole batches: every chunk of `block` local requests is one aligned run of consecutive global indices
                for idx in per_rank:
                    for c in range(0, len(idx), block):
                        chunk = idx[c:c + block]
                        assert chunk[0] % block == 0 and chunk == list(range(chunk[0], chunk[0] + len(chunk)))


def test_batches_are_the_same_for_every_gpu_count():
    """bit-identical output for 1..8 GPUs needs every request to run in the same batch, in the same slot, for every world size"""
    n, block = 37, 4
    ref = {tuple(range(b, min(b + block, n))) for b in range(0, n, block)}
    for w in (1, 2, 4, 8):
        got = set()
        for r in range(w):
            idx = shard_indices(n, r, w, block)
            got |= {tuple(idx[c:c + block]) for c in range(0, len(idx), block)}
        assert got == ref


def test_two_rank_result_equals_single_process():
    n_items = 7
    ref = gather_in_order(_work(list(range(n_items))), list(range(n_items)), n_items)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, n_items, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert torch.equal(res, ref)

"""CPU, world_size 2, gloo: the clip-sharding / gather logic of the multi-GPU path."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n_clips, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pyaudioanalysis_b200.dist import feature_extraction_sharded, shard_bounds
    full = torch.arange(n_clips * 6, dtype=torch.float32).reshape(n_clips, 6)
    fake = lambda x: torch.stack([x * 2.0, x + 1.0], dim=1)       # [n, 2, 6] stand-in for [n, F, T]
    got = feature_extraction_sharded(lambda lo, hi: full[lo:hi], n_clips, 16000, 800, 400, compute=fake)
    lo, hi = shard_bounds(n_clips, rank, world)
    local = feature_extraction_sharded(lambda lo, hi: full[lo:hi], n_clips, 16000, 800, 400, compute=fake, gather_to=None)
    ok = local.shape[0] == hi - lo
    if rank == 0:
        ok = ok and torch.equal(got, fake(full))
    else:
        ok = ok and got is None
    q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_clips", [7, 8, 1])
def test_shard_and_gather_gloo(n_clips):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + n_clips
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_clips, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def test_shard_bounds_cover():
    from pyaudioanalysis_b200.dist import shard_bounds
    for n in (0, 1, 7, 8, 1000, 100000):
        for w in (1, 2, 4, 8):
            spans = [shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1

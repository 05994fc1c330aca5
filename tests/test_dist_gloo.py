"""world_size-2 gloo test of the frequency sharding / gather plumbing on CPU."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, nfreq, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from casingsimulations_b200.sweep import gather_solutions, shard_frequencies, shard_indices
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    freqs = np.logspace(-1, 3, nfreq)
    mine = shard_frequencies(freqs, rank, world)
    # stand-in "solution" of frequency f: a vector that encodes f (the solve itself needs a GPU)
    local = torch.tensor(np.stack([f * (np.arange(5) + 1j) for f in mine]))
    full = gather_solutions(local, nfreq, rank, world)
    ok = np.allclose(full.numpy(), np.stack([f * (np.arange(5) + 1j) for f in freqs]))
    t = torch.tensor([float(len(mine))], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)          # the bench's max-over-ranks reduction
    q.put((rank, ok, len(shard_indices(nfreq, rank, world)), float(t)))
    dist.destroy_process_group()


@pytest.mark.parametrize("nfreq", [32, 7])
def test_frequency_sharding_gloo(nfreq):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + nfreq
    procs = [ctx.Process(target=_worker, args=(r, 2, port, nfreq, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] for r in res)
    assert sum(r[2] for r in res) == nfreq
    assert all(r[3] == (nfreq + 1) // 2 for r in res)


def test_shard_is_a_partition():
    from casingsimulations_b200.sweep import shard_indices
    for n in (1, 5, 32):
        for w in (1, 2, 4, 8):
            parts = np.concatenate([shard_indices(n, r, w) for r in range(w)])
            assert sorted(parts.tolist()) == list(range(n))

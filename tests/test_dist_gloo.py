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


def _slab_worker(rank, world, port, q):
    import torch.distributed as dist
    import casingsimulations_b200 as cs
    from casingsimulations_b200 import slab
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        for nth in (1, 4):
            hy = np.r_[2 * np.pi] if nth == 1 else np.ones(nth) * 2 * np.pi / nth
            mesh = cs.CylMesh([np.r_[1.0, 2.0, 3.0], hy, np.arange(1.0, 8.0)], x0=[0, 0, -3.0])
            nz = int(mesh.vnC[2])
            zsplit = slab.split_levels(nz, world)
            assert zsplit[0] == 0 and zsplit[-1] == nz and all(b - a >= 2 for a, b in zip(zsplit[:-1], zsplit[1:]))
            k0, k1 = slab.local_range(zsplit, rank)
            lm = slab.local_mesh(mesh, k0, k1)
            assert np.array_equal(lm.vectorNz, mesh.vectorNz[k0:k1 + 1])          # node coordinates of the global mesh, bit for bit
            for fam, n in ((0, mesh.nC), (1, mesh.nE), (2, mesh.nF)):
                g = np.random.default_rng(fam).standard_normal(n)                   # same global vector on every rank
                loc = slab.to_local(g, mesh, fam, k0, k1)
                assert loc.shape[0] == (lm.nC, lm.nE, lm.nF)[fam]
                own = slab.owned_part(loc, mesh, fam, zsplit, rank)
                parts = [None] * world
                dist.all_gather_object(parts, own)
                assert np.array_equal(slab.assemble(parts, mesh, fam), g)             # every entity has exactly one owner
            # grids of the local mesh are the matching part of the global grids (sources / painters see the same points)
            for col in (0, 2):
                gz = np.r_[mesh.gridFx[:, col], mesh.gridFy[:, col], mesh.gridFz[:, col]]
                lz = np.r_[lm.gridFx[:, col], lm.gridFy[:, col], lm.gridFz[:, col]]
                assert np.array_equal(slab.to_local(gz, mesh, 2, k0, k1), lz)
        owner, pos = slab.mode_owner(32, world)
        assert owner[1] == owner[31] and owner[0] == 0 and owner[16] == 16 % world   # modes m and nth - m share a factor
        for r in range(world):
            assert sorted(p for o, p in zip(owner, pos) if o == r) == list(range(owner.count(r)))
        q.put((rank, "ok"))
    except Exception as exc:                                                          # report instead of hanging the peer
        q.put((rank, repr(exc)))
    finally:
        dist.destroy_process_group()


def test_z_slab_partition_two_ranks():
    """the host-side bookkeeping of the z-slab solve (split, global <-> local layout, ownership, mode owners) with two
    gloo ranks: the owned parts of all ranks reassemble every global cell / edge / face vector exactly"""
    import multiprocessing as mpc
    ctx = mpc.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 400) + 17
    procs = [ctx.Process(target=_slab_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res

"""ONE 3-D FDEM-h system over N GPUs: z-slab decomposition + NCCL halo exchange + mode-sharded factors.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/dist_fdem.py NTH CSZ NPADZ [compare]

Every rank owns a slab of z levels.  With `compare` rank 0 afterwards solves the same system on its own GPU alone
(when it fits) and the relative difference of the gathered solution is printed (the parity of the partitioned solve)."""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import casingsimulations_b200 as cs
from casingsimulations_b200 import _lib, slab



def solve_one_system(nth, csz, npadz, compare=False, reps=2, freq=1.0):
    """all ranks call this; returns the result dictionary on rank 0 (None elsewhere)"""
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); lr = int(os.environ.get("LOCAL_RANK", "0"))

    mp = cs.model.CasingInHalfspace(src_a=np.r_[0.0, 0.0, 0.0], src_b=np.r_[1000.0, 0.0, 0.0], mur_casing=100.0, freqs=np.r_[freq])
    mg = cs.CasingMeshGenerator(modelParameters=mp, npadx=10, npadz=npadz, csz=csz, hy=np.ones(nth) * 2 * np.pi / nth)
    m = mg.mesh
    th = m.vectorCCy[nth // 2]
    mp.src_a, mp.src_b = np.r_[0.0, th, 0.0], np.r_[1000.0, th, 0.0]
    nz = int(m.vnC[2])
    # the source is a handful of wire faces: rank 0 paints it on the global grids and ships (index, value) pairs
    se_sparse = [None]
    if rank == 0:
        src = cs.sources.TopCasingSrc(modelParameters=mp, meshGenerator=mg, physics="fdem")
        se = np.real(src.s_e)
        nzi = np.flatnonzero(se)
        se_sparse = [(nzi, se[nzi])]
        if not compare:
            del se, src
            m.__dict__.pop("_cache", None)                   # drop the global grids (GBs at config-5 size)
    if world > 1:
        dist.broadcast_object_list(se_sparse, src=0)
    zsplit = slab.split_levels(nz, world)

    ctx = _lib.Context(lr)
    uid = None
    if world > 1:
        t = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            t = torch.frombuffer(bytearray(_lib.Context.comm_unique_id()), dtype=torch.uint8).cuda()
        dist.broadcast(t, 0)
        uid = bytes(t.cpu().numpy().tobytes())
    ctx.comm_init(uid, rank, world, zsplit[rank], zsplit[rank + 1])
    ctx.set_mesh(m.hx, m.hy, m.hz, m.x0[2])
    k0, k1 = ctx.local_range()
    assert (k0, k1) == slab.local_range(zsplit, rank)
    params = mp.paint_params()
    se_l = ctx.to_device(slab.to_local_sparse(se_sparse[0][0], se_sparse[0][1], m, 2, k0, k1))
    times = []
    for rep in range(reps):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ctx.paint(params)
        u_l = ctx.solve_fdem("h", [freq], se_l, rtol=1e-10)
        e1.record()
        torch.cuda.synchronize()
        tt = torch.tensor([e0.elapsed_time(e1) * 1e-3], dtype=torch.float64, device="cuda")     # device time, max over ranks
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        times.append(float(tt[0]))
    info = {k: float(v) for k, v in ctx.info.items()}
    free, total = torch.cuda.mem_get_info()
    u_own = slab.owned_part(u_l[0].cpu().numpy(), m, 1, zsplit, rank)
    if world > 1:
        gathered = [None] * world if rank == 0 else None
        dist.gather_object(u_own, gathered, dst=0)
    else:
        gathered = [u_own]
    ctx.close()
    if rank == 0:
        u = slab.assemble(gathered, m, 1)
        out = {"mesh": [int(v) for v in m.vnC], "cells": int(m.nC), "edges": int(m.nE), "n_gpus": world, "zsplit": zsplit,
               "s_per_frequency": min(times), "first_s": times[0], "info": {k: round(v, 6) for k, v in info.items()},
               "gpu_mem_used_GB_rank0": round((total - free) / 1e9, 1)}
        if compare:
            c1 = _lib.Context(lr)
            c1.set_mesh(m.hx, m.hy, m.hz, m.x0[2])
            c1.paint(params)
            u1 = c1.solve_fdem("h", [freq], c1.to_device(se), rtol=1e-10)[0].cpu().numpy()
            out["rel_diff_vs_single_gpu_h"] = float(np.linalg.norm(u - u1) / np.linalg.norm(u1))
            sed = c1.to_device(se)
            j1 = c1.fields_fdem("h", "j", freq, c1.to_device(u1), sed).cpu().numpy()
            j = c1.fields_fdem("h", "j", freq, c1.to_device(u), sed).cpu().numpy()
            out["rel_diff_vs_single_gpu_j"] = float(np.linalg.norm(j - j1) / np.linalg.norm(j1))
            out["single_gpu_info"] = {k: round(float(v), 6) for k, v in c1.info.items()}
            c1.close()
        return out
    return None


if __name__ == "__main__":
    nth = int(sys.argv[1]); csz = float(sys.argv[2]); npadz = int(sys.argv[3])
    compare = len(sys.argv) > 4 and sys.argv[4] == "compare"
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); lr = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lr)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    res = solve_one_system(nth, csz, npadz, compare, int(os.environ.get("REPS", "2")), float(os.environ.get("FREQ", "1.0")))
    if res is not None:
        print(json.dumps(res), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()

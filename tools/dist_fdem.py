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

nth = int(sys.argv[1]); csz = float(sys.argv[2]); npadz = int(sys.argv[3])
compare = len(sys.argv) > 4 and sys.argv[4] == "compare"
freq = float(os.environ.get("FREQ", "1.0"))
rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))

mp = cs.model.CasingInHalfspace(src_a=np.r_[0.0, 0.0, 0.0], src_b=np.r_[1000.0, 0.0, 0.0], mur_casing=100.0, freqs=np.r_[freq])
mg = cs.CasingMeshGenerator(modelParameters=mp, npadx=10, npadz=npadz, csz=csz, hy=np.ones(nth) * 2 * np.pi / nth)
m = mg.mesh
th = m.vectorCCy[nth // 2]
mp.src_a, mp.src_b = np.r_[0.0, th, 0.0], np.r_[1000.0, th, 0.0]
src = cs.sources.TopCasingSrc(modelParameters=mp, meshGenerator=mg, physics="fdem")
se = np.real(src.s_e)
nz = int(m.vnC[2])
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
se_l = ctx.to_device(slab.to_local(se, m, 2, k0, k1))
times = []
for rep in range(int(os.environ.get("REPS", "2"))):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    ctx.paint(params)
    u_l = ctx.solve_fdem("h", [freq], se_l, rtol=1e-10)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    times.append(time.perf_counter() - t0)
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
    print(json.dumps(out), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()

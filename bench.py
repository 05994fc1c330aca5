#!/usr/bin/env python
"""Benchmark of the hot path: FDEM casing solves on the cylindrical mesh.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1], SURVEY.md section 8d "C2"): FDEM h-formulation
on the 109 x 1 x 717 casing mesh (78 153 cells), CasingInHalfspace with
sigma_casing = 5.5e6 S/m, TopCasingSrc, 32 frequencies 0.1 Hz - 1 kHz.  One "step"
is one full sweep: paint -> assemble -> factor -> solve for 32 frequencies per GPU
(weak scaling: every rank gets its own 32 frequencies, no data-path collective).

metric  : seconds per frequency (lower is better), whole job = max-over-ranks time
          divided by the frequencies all ranks solved.
value   : inputs resident in HBM.      e2e: host buffers in, host solutions out.
roofline: the stencil apply the metric names (K4, complex128 curl-curl + i w mass on
          the 3-D 109 x 32 x 1394 mesh, 144 B/cell), timed live with CUDA events.
cpu_baseline / --impl reference: the oracle (scipy sparse assembly + SuperLU, the
          reference's documented fallback solver, run.py:19-24) on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "fdem_casing_solve_seconds_per_frequency"
UNIT = "s/frequency"
NFREQ_PER_GPU = 32
NCU_TRAFFIC_K4 = 705.5e6      # bytes per launch of k_edge_op<cplx> on the 109x32x1394 mesh (497.7 MB read + 207.8 MB written)
WORKLOAD = ("C2: FDEM-h, CasingInHalfspace (sigma_casing 5.5e6 S/m), CasingMeshGenerator 109x1x717 = 78153 cells, "
            "TopCasingSrc, 32 freqs 0.1 Hz-1 kHz per GPU")


def build_inputs(nfreq_total):
    import casingsimulations_b200 as cs
    mp = cs.model.CasingInHalfspace(src_a=np.r_[0.0, 0.0, 0.0], src_b=np.r_[1000.0, 0.0, 0.0],
                                    freqs=np.logspace(-1, 3, nfreq_total))
    mg = cs.CasingMeshGenerator(modelParameters=mp, npadx=10, npadz=20, csz=1.5)
    src = cs.sources.TopCasingSrc(modelParameters=mp, meshGenerator=mg, physics="fdem")
    return mp, mg, src


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for n, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def blas_threads():
    """threads the CPU arm can actually use: SuperLU itself is sequential, its supernodal BLAS
    calls go through scipy's OpenBLAS thread pool"""
    try:
        from threadpoolctl import threadpool_info
        n = [i.get("num_threads", 1) for i in threadpool_info()]
        return max(n) if n else 1
    except Exception:
        return os.cpu_count() or 1


def oracle_sample(mg, mp, src_se, freqs):
    """scipy assembly + SuperLU factor/solve, one frequency at a time (the reference's loop)."""
    from oracle.cyl_oracle import OracleCylMesh, OracleFDEM, paint_casing_in_halfspace
    m = mg.mesh
    t0 = time.perf_counter()
    om = OracleCylMesh([m.hx, m.hy, m.hz], m.x0)
    sig, mu = paint_casing_in_halfspace(om, sigma_back=mp.sigma_back, sigma_air=mp.sigma_air, sigma_casing=mp.sigma_casing,
                                        sigma_inside=mp.sigma_inside, mur_casing=mp.mur_casing, casing_a=mp.casing_a,
                                        casing_b=mp.casing_b, casing_z=tuple(mp.casing_z), surface_z=mp.surface_z)
    P = OracleFDEM(om, sig, mu, "h")
    t_setup = time.perf_counter() - t0
    sols = []
    t1 = time.perf_counter()
    for f in freqs:
        sols.append(P.solve(f, src_se))
    t_solve = time.perf_counter() - t1
    return t_setup, t_solve, sols


_REF_STATE = {}


def _ref_worker(f):
    """one frequency in a forked worker: SuperLU factor + solve on the matrices assembled by the parent"""
    try:
        from threadpoolctl import threadpool_limits
        with threadpool_limits(1):
            return _REF_STATE["P"].solve(f, _REF_STATE["se"]).shape[0]
    except ImportError:
        return _REF_STATE["P"].solve(f, _REF_STATE["se"]).shape[0]


def run_reference(args, rank):
    """--impl reference: the reference's CPU path (oracle port; SimPEG/Pardiso are not installable offline) with all
    the host cores it can use: the frequencies of a sweep are independent systems (SimPEG loops over them,
    sources.py:119-123), so a step solves one frequency per core in forked workers that share the assembled
    operators; SuperLU itself is sequential."""
    if rank != 0:
        return
    import multiprocessing as mpc
    from oracle.cyl_oracle import OracleCylMesh, OracleFDEM, paint_casing_in_halfspace
    mp, mg, src = build_inputs(NFREQ_PER_GPU)
    se = np.real(src.s_e)
    ncores = max(1, min(os.cpu_count() or 1, NFREQ_PER_GPU))
    allf = np.logspace(-1, 3, NFREQ_PER_GPU)
    sample = [float(f) for f in allf[np.linspace(0, NFREQ_PER_GPU - 1, ncores).round().astype(int)]]
    m = mg.mesh
    t0 = time.perf_counter()
    om = OracleCylMesh([m.hx, m.hy, m.hz], m.x0)
    sig, mu = paint_casing_in_halfspace(om, sigma_back=mp.sigma_back, sigma_air=mp.sigma_air, sigma_casing=mp.sigma_casing,
                                        sigma_inside=mp.sigma_inside, mur_casing=mp.mur_casing, casing_a=mp.casing_a,
                                        casing_b=mp.casing_b, casing_z=tuple(mp.casing_z), surface_z=mp.surface_z)
    P = OracleFDEM(om, sig, mu, "h")
    P.solve(sample[0], se)                                  # builds the cached operators before the fork
    t_setup = time.perf_counter() - t0
    _REF_STATE.update(P=P, se=se)
    ts, t_seq = [], None
    with mpc.get_context("fork").Pool(ncores) as pool:
        for _ in range(max(args.warmup, 0) and 1):
            pool.map(_ref_worker, sample, chunksize=1)
        for _ in range(args.steps):
            t1 = time.perf_counter()
            pool.map(_ref_worker, sample, chunksize=1)
            ts.append((time.perf_counter() - t1) / len(sample) + t_setup / NFREQ_PER_GPU)
    t1 = time.perf_counter()
    P.solve(sample[len(sample) // 2], se)
    t_seq = time.perf_counter() - t1
    v = float(np.mean(ts))
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": v * NFREQ_PER_GPU * 1e3, "higher_is_better": False, "scaling": "weak",
            "vs_baseline": None, "dtype": "c128", "data": "synthetic",
            "config": {"workload": WORKLOAD},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": ncores, "kind": "port",
                             "sample": "{} of the 32 frequencies per step, one per core in forked workers; scipy assembly "
                                       "({:.2f} s) amortised over the sweep + SuperLU factor+solve per frequency (the "
                                       "reference's fallback SolverLU, run.py:19-24; one frequency alone on one core: "
                                       "{:.3f} s); MKL PARDISO not installable offline".format(len(sample), t_setup, t_seq),
                             "single_core_s_per_frequency": t_seq},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def stencil_roofline(torch, dev, reps=20):
    """Live roofline of the stencil applies on the 109x32x1394 mesh (C3 size, inputs >> L2):
    K4 edge form in complex128 (144 B/cell) and K3 DC in fp64 (40 B/cell), CUDA events on the
    library's stream around `reps` back-to-back launches."""
    import casingsimulations_b200 as cs
    from casingsimulations_b200 import _lib
    mp = cs.model.CasingInHalfspace(src_a=np.r_[0.0, 0.0, 0.0], src_b=np.r_[1000.0, 0.0, 0.0], mur_casing=100.0)
    mg = cs.CasingMeshGenerator(modelParameters=mp, npadx=10, npadz=25, csz=0.75, hy=np.ones(32) * 2 * np.pi / 32)
    m = mg.mesh
    ctx = _lib.Context(dev)
    ctx.set_mesh(m.hx, m.hy, m.hz, m.x0[2])
    ctx.paint(mp.paint_params())
    peak, src = measured_peak()
    op = ctx.op(_lib.KIND_EDGE_H, plan=False)
    x = torch.randn(ctx.nE, dtype=torch.complex128, device=ctx.device)
    torch.cuda.synchronize()
    t = op.apply_timed(x, 2j * np.pi, reps) * 1e-3
    alg = 144.0 * ctx.nC
    out = {"bound": "hbm", "kernel": "k_edge_op<cplx> (K4: C^T MfRho C + i w MeMu, 3-D 109x32x1394, 4.86 M cells)",
           "achieved": alg / t / 1e9, "peak": peak, "peak_source": src, "unit": "GB/s", "frac": alg / t / 1e9 / peak,
           "traffic": NCU_TRAFFIC_K4, "traffic_source": "profiles/r01_ncu_k_edge_op.md (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum, one launch)",
           "bytes_per_launch": alg, "us_per_launch": t * 1e6, "cells": ctx.nC}
    opd = ctx.op(_lib.KIND_DC, plan=False)
    xd = torch.randn(ctx.nC, dtype=torch.float64, device=ctx.device)
    torch.cuda.synchronize()
    td = opd.apply_timed(xd, 0.0, reps) * 1e-3
    out["dc_apply"] = {"kernel": "k_dc_apply<double> (K3, 40 B/cell)", "achieved": 40.0 * ctx.nC / td / 1e9,
                       "frac": 40.0 * ctx.nC / td / 1e9 / peak, "us_per_launch": td * 1e6}
    del x, xd
    ctx.close()
    return out


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: casingsimulations_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from casingsimulations_b200 import _lib
    stream = torch.cuda.Stream(device=local_rank)
    torch.cuda.set_stream(stream)               # torch copies, events and our kernels share one stream

    mp, mg, src = build_inputs(NFREQ_PER_GPU * world)
    m = mg.mesh
    freqs = np.ascontiguousarray(mp.freqs[rank::world])          # this rank's 32 frequencies
    se_host = torch.from_numpy(np.ascontiguousarray(src.s_e)).pin_memory()
    params = mp.paint_params()
    ctx = _lib.Context(local_rank, stream=stream.cuda_stream)
    ctx.set_mesh(m.hx, m.hy, m.hz, m.x0[2])
    se_dev = se_host.to(ctx.device)
    out_host = torch.empty((len(freqs), ctx.nE), dtype=torch.complex128).pin_memory()
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=ctx.device)   # > 126 MB L2

    def step_device():
        ctx.paint(params)                      # K1 (drops cached operators -> full re-assembly, as the reference does)
        return ctx.solve_fdem("h", freqs, se_dev, rtol=1e-10)

    def step_e2e():
        sd = se_host.to(ctx.device, non_blocking=True)          # H2D from pinned memory
        ctx.paint(params)
        u = ctx.solve_fdem("h", freqs, sd, rtol=1e-10)
        out_host.copy_(u, non_blocking=True)                    # D2H of the solutions
        torch.cuda.synchronize()
        return u

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        tot = 0.0
        for _ in range(steps):
            flush.zero_()                                       # L2 flush between timed iterations
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            fn()
            e1.record(stream)
            torch.cuda.synchronize()
            dt = e0.elapsed_time(e1) * 1e-3       # CUDA events on the stream every kernel of the step runs on
            if os.environ.get("BENCH_DEBUG"):
                print("rank", rank, fn.__name__, "step", round(dt * 1e3, 2), "ms", {k: round(v, 1) for k, v in ctx.info.items() if k.startswith("ms_")}, file=sys.stderr, flush=True)
            tot += dt
        return tot

    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = ctx.launch_count()
    t_dev = timed(step_device, args.steps)
    launches = ctx.launch_count() - l0
    info = dict(ctx.info)
    barrier()
    t_e2e = timed(step_e2e, args.steps)
    barrier()
    clocks = sampler.stop()

    tt = torch.tensor([t_dev, t_e2e], dtype=torch.float64, device=ctx.device)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_dev, t_e2e = float(tt[0]), float(tt[1])
    nfreq_total = NFREQ_PER_GPU * world

    if rank == 0:
        roof = stencil_roofline(torch, local_rank)
        t_setup, t_solve, sols = oracle_sample(mg, mp, src.s_e, [freqs[0], freqs[-1]])
        cpu_v = (t_setup / NFREQ_PER_GPU * 2 + t_solve) / 2
        cpu_base = {"value": cpu_v, "unit": UNIT, "cores": 1, "kind": "port",
                    "sample": "2 of the 32 frequencies ({:.3g} Hz, {:.3g} Hz), one after the other: scipy assembly (amortised "
                              "over the sweep) + SuperLU factor + solve (sequential); MKL PARDISO not installable "
                              "offline".format(freqs[0], freqs[-1])}
        # the all-cores figure comes from the reference arm itself, in a clean process (it forks workers)
        try:
            outp = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "2", "--warmup", "1"],
                                  capture_output=True, text=True, timeout=600, env=dict(os.environ, RANK="0", WORLD_SIZE="1", LOCAL_RANK="0"))
            ref = json.loads([l for l in outp.stdout.splitlines() if l.startswith("{")][-1])
            cpu_base = dict(ref["cpu_baseline"], one_after_the_other_s_per_frequency=cpu_v)
        except Exception as exc:                  # keep the sequential figure
            cpu_base["note"] = "all-cores run failed: {}".format(exc)
        # parity of the timed path against the oracle on the sampled frequencies
        u = step_device().cpu().numpy()
        par = [float(np.linalg.norm(u[i] - s) / np.linalg.norm(s)) for i, s in zip((0, len(freqs) - 1), sols)]
        value = t_dev / (args.steps * nfreq_total)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": t_dev / args.steps * 1e3, "higher_is_better": False,
                "scaling": "weak", "vs_baseline": None, "dtype": "c128", "data": "synthetic",
                "config": {"workload": WORKLOAD, "frequencies_per_gpu": NFREQ_PER_GPU, "unknowns_per_frequency": ctx.nE,
                           "l2": "256 MB buffer written between timed iterations", "rtol": 1e-10},
                "frequencies_per_s": 1.0 / value,
                "e2e": {"value": t_e2e / (args.steps * nfreq_total), "unit": UNIT,
                        "h2d_bytes_per_step": int(se_host.numel() * 8), "d2h_bytes_per_step": int(out_host.numel() * 16)},
                "gpu_launches": int(launches),
                "solve_info": {"krylov_iters": info.get("iters"), "relres_scaled": info.get("relres"),
                               "ms_assemble": info.get("ms_assemble"), "ms_factor": info.get("ms_factor"),
                               "ms_solve": info.get("ms_solve"), "rel_l2_vs_oracle_h": par},
                "roofline": roof,
                "cpu_baseline": cpu_base,
                "clocks": clocks}
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Benchmark of the hot path: 3-D FDEM casing solves on the cylindrical mesh.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json metric "3D cyl FDEM solve s/frequency"; largest 3-D FDEM configuration that fits one
B200): FDEM h-formulation on the 109 x 32 x 1394 casing mesh (4 862 272 cells, 14 595 186 complex edge unknowns;
the BASELINE configs[2] mesh), CasingInHalfspace with sigma_casing = 5.5e6 S/m, mu_r = 100, TopCasingSrc.  One "step"
is one full pass of the path for NFREQ_PER_GPU frequencies: paint -> assemble (inner products + probing of the
17 theta-mode band matrices) -> factor -> Krylov solve, per GPU (weak scaling over independent frequencies: every
rank gets its own frequencies, no data-path collective).

metric  : seconds per frequency (lower is better), whole job = max-over-ranks time / frequencies all ranks solved.
value   : inputs resident in HBM.     e2e: SimulationFDEM(...).run(save=False) -- numpy in, numpy Fields out.
roofline: the dominant GPU work of the timed step, the band LU of the theta-mode systems (FP64 FMA pipe), timed live
          by CUDA events on the library's stream, against the live-measured DFMA peak; the stencil applies the metric
          also names (K4 / K4' / K3, HBM-bound) are reported beside it against the measured HBM copy peak.
cpu_baseline / --impl reference: the oracle (scipy sparse assembly + SuperLU, the reference's documented fallback
          solver, run.py:19-24), one frequency after the other like the reference's loop, on the largest 3-D mesh of the
          same family the host factors in seconds (109 x 4 x 40); the 4.86 M-cell system is infeasible for a host sparse
          direct solver.  `same_mesh_as_reference_arm` in our line is the apples-to-apples figure: our path
          on that very mesh.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "fdem3d_casing_solve_seconds_per_frequency"
UNIT = "s/frequency"
NFREQ_PER_GPU = 2
NFREQ_STRONG = 16                                      # --mode strong: frequencies of the whole sweep, sharded over the GPUs
MESH_OURS = dict(nth=32, csz=0.75, npadz=25)          # 109 x 32 x 1394 (C3 mesh)
MESH_REF = dict(nth=4, csz=50.0, npadz=5)             # 109 x 4 x 40: ~8 s SuperLU per frequency and core
WORKLOAD = ("3-D FDEM-h, CasingInHalfspace (sigma_casing 5.5e6 S/m, mu_r 100, 1 km casing), CasingMeshGenerator "
            "109x32x1394 = 4862272 cells (14595186 complex edge unknowns), TopCasingSrc, {} frequencies 1-10 Hz per GPU"
            .format(NFREQ_PER_GPU))


def mesh_name(mg):
    m = mg.mesh
    return "{}x{}x{} = {} cells".format(int(m.vnC[0]), int(m.vnC[1]), int(m.vnC[2]), int(m.nC))


def build_inputs(mesh_kw, freqs):
    import casingsimulations_b200 as cs
    nth = mesh_kw["nth"]
    mp = cs.model.CasingInHalfspace(src_a=np.r_[0.0, 0.0, 0.0], src_b=np.r_[1000.0, 0.0, 0.0], mur_casing=100.0,
                                    freqs=np.asarray(freqs, dtype=float))
    mg = cs.CasingMeshGenerator(modelParameters=mp, npadx=10, npadz=mesh_kw["npadz"], csz=mesh_kw["csz"],
                                hy=np.ones(nth) * 2 * np.pi / nth)
    th = mg.mesh.vectorCCy[nth // 2]                   # a cell-centre azimuth (sources.py:779-781)
    mp.src_a, mp.src_b = np.r_[0.0, th, 0.0], np.r_[1000.0, th, 0.0]
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


# ---------------------------------------------------------------- reference arm ------------------------------------
_REF_STATE = {}


def _oracle_problem(mp, mg):
    from oracle.cyl_oracle import OracleCylMesh, OracleFDEM, paint_casing_in_halfspace
    m = mg.mesh
    om = OracleCylMesh([m.hx, m.hy, m.hz], m.x0)
    sig, mu = paint_casing_in_halfspace(om, sigma_back=mp.sigma_back, sigma_air=mp.sigma_air, sigma_casing=mp.sigma_casing,
                                        sigma_inside=mp.sigma_inside, mur_casing=mp.mur_casing, casing_a=mp.casing_a,
                                        casing_b=mp.casing_b, casing_z=tuple(mp.casing_z), surface_z=mp.surface_z)
    return OracleFDEM(om, sig, mu, "h")


def _ref_worker(f):
    """one frequency in a forked worker: scipy assembly of A(f) + SuperLU factor + solve (sequential code)"""
    try:
        from threadpoolctl import threadpool_limits
        with threadpool_limits(1):
            return _REF_STATE["P"].solve(f, _REF_STATE["se"]).shape[0]
    except ImportError:
        return _REF_STATE["P"].solve(f, _REF_STATE["se"]).shape[0]


def run_reference(args, rank):
    """--impl reference: the reference's CPU path (oracle port: SimPEG / Pardiso are not installable offline) for the same
    metric on the largest mesh of the same family the host factors in seconds.  The reference loops over the frequencies
    of a survey one after the other in one process (run.py:205 -> SimPEG `for freq in survey.freqs`) and its fallback
    solver SuperLU is sequential, so a step is ONE frequency on one core: scipy assembly of A(f) + factor + solve, every
    step measured.  What a hand-parallelised CPU sweep could do (one frequency per core in forked workers, which the
    reference does not do) is measured as well and reported in cpu_baseline.all_cores_concurrent."""
    if rank != 0:
        return
    import multiprocessing as mpc
    ncores = max(1, os.cpu_count() or 1)
    freqs = [float(f) for f in np.logspace(0, 1, max(ncores, 2))]
    mp, mg, src = build_inputs(MESH_REF, freqs)
    se = np.real(src.s_e)
    t0 = time.perf_counter()
    P = _oracle_problem(mp, mg)
    P.getA(freqs[0])                                        # builds the cached operators
    t_setup = time.perf_counter() - t0
    _REF_STATE.update(P=P, se=se)
    warm = 1 if args.warmup > 0 else 0                      # a direct solver has nothing to warm beyond first touch
    for _ in range(warm):
        _ref_worker(freqs[0])
    ts = []
    for k in range(args.steps):
        t1 = time.perf_counter()
        _ref_worker(freqs[k % len(freqs)])
        ts.append(time.perf_counter() - t1)
    step = float(np.mean(ts))
    # hand-parallelised variant (not the reference's code path): one frequency per core, two rounds
    conc = None
    try:
        with mpc.get_context("fork").Pool(ncores) as pool:
            tc = []
            for _ in range(2):
                t1 = time.perf_counter()
                pool.map(_ref_worker, freqs[:ncores], chunksize=1)
                tc.append(time.perf_counter() - t1)
        conc = {"cores": ncores, "s_per_frequency": float(min(tc)) / ncores, "s_per_round_of_one_frequency_per_core": float(min(tc))}
    except Exception as exc:
        conc = {"error": str(exc)}
    sample = ("mesh {} ({} complex edge unknowns; the 4862272-cell system of the main arm is infeasible for a host sparse "
              "direct solver), one frequency per step, sequential like the reference's loop (run.py:205): scipy assembly of A(f) + "
              "SuperLU factor + solve (the reference's fallback SolverLU, run.py:19-24, single-threaded; MKL PARDISO not "
              "installable offline); shared operator assembly {:.2f} s not counted".format(mesh_name(mg), mg.mesh.nE, t_setup))
    line = {"impl": "reference", "metric": METRIC, "value": step, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": warm, "ms_per_step": step * 1e3, "higher_is_better": False, "scaling": "weak",
            "vs_baseline": None, "dtype": "c128", "data": "synthetic",
            "config": {"workload": WORKLOAD, "reference_sample_mesh": mesh_name(mg), "frequencies_per_step": 1},
            "cpu_baseline": {"value": step, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample,
                             "mesh": mesh_name(mg), "all_cores_concurrent": conc},
            "e2e": {"value": step, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------- our arm ------------------------------------------
def stencil_roofline(torch, dev, reps=20):
    """Live HBM roofline of the stencil applies on the bench mesh (109x32x1394, inputs >> L2): K4 edge form in
    complex128 (144 B/cell), K4' face form in fp64 (96 B/cell), K3 DC in fp64 (40 B/cell); CUDA events on the
    library's stream around `reps` back-to-back launches."""
    import casingsimulations_b200 as cs
    from casingsimulations_b200 import _lib
    mp, mg, _ = build_inputs(MESH_OURS, [1.0])
    m = mg.mesh
    ctx = _lib.Context(dev)
    ctx.set_mesh(m.hx, m.hy, m.hz, m.x0[2])
    ctx.paint(mp.paint_params())
    peak, src = measured_peak()
    out = {"bound": "hbm", "peak": peak, "peak_source": src, "unit": "GB/s", "cells": ctx.nC,
           "ncu": "profiles/r02_ncu_stencils.md"}

    def one(kind, n, cplx, bpc, name):
        op = ctx.op(kind, plan=False)
        x = torch.randn(n, dtype=torch.complex128 if cplx else torch.float64, device=ctx.device)
        torch.cuda.synchronize()
        t = op.apply_timed(x, 2j * np.pi if cplx else 1e3, reps) * 1e-3
        del x
        return {"kernel": name, "bytes_per_cell": bpc, "us_per_launch": t * 1e6, "achieved": bpc * ctx.nC / t / 1e9,
                "frac": bpc * ctx.nC / t / 1e9 / peak}
    out["k4_edge_c128"] = one(_lib.KIND_EDGE_H, ctx.nE, True, 144.0, "k_edge_op<cplx> (K4: C^T MfRho C + i w MeMu)")
    out["k4_face_f64"] = one(_lib.KIND_FACE_J, ctx.nF, False, 96.0, "k_face_axis + k_face_op_ring<double> (K4': MfRho (C MeMuI C^T MfRho + 1/dt))")
    out["k3_dc_f64"] = one(_lib.KIND_DC, ctx.nC, False, 40.0, "k_dc_apply<double> (K3: D~ MfRhoI D~^T)")
    ctx.close()
    return out


def small_mesh_check(torch, dev):
    """Our path on the reference arm's mesh: time per frequency (apples to apples with cpu_baseline) and parity of the
    solution against the oracle's extended-precision refined SuperLU solve."""
    from casingsimulations_b200 import _lib
    mp, mg, src = build_inputs(MESH_REF, [1.0, 10.0])
    m = mg.mesh
    ctx = _lib.Context(dev)
    ctx.set_mesh(m.hx, m.hy, m.hz, m.x0[2])
    se = ctx.to_device(np.real(src.s_e))
    params = mp.paint_params()
    for _ in range(2):
        ctx.paint(params)
        ctx.solve_fdem("h", mp.freqs, se)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 5
    for _ in range(reps):
        ctx.paint(params)
        u = ctx.solve_fdem("h", mp.freqs, se)
    torch.cuda.synchronize()
    t = (time.perf_counter() - t0) / (reps * len(mp.freqs))
    out = {"mesh": mesh_name(mg), "s_per_frequency": t, "relres_scaled": ctx.info.get("relres")}
    try:
        # parity against the EXACTLY assembled discrete system (oracle.exact_mesh, long-double assembly + refinement) on the
        # quantities the fp64 data determine; the plain L2 norm of j beside what ONE fp64 rounding of the
        # right-hand side does to it
        from oracle.cyl_oracle import (OracleCylMesh, OracleFDEM, paint_casing_in_halfspace, exact_mesh, energy_rel,
                                       casing_currents, rhs_rounding_sensitivity)
        m = mg.mesh
        om = OracleCylMesh([m.hx, m.hy, m.hz], m.x0)
        sig, mu = paint_casing_in_halfspace(om, sigma_back=mp.sigma_back, sigma_air=mp.sigma_air, sigma_casing=mp.sigma_casing,
                                            sigma_inside=mp.sigma_inside, mur_casing=mp.mur_casing, casing_a=mp.casing_a,
                                            casing_b=mp.casing_b, casing_z=tuple(mp.casing_z), surface_z=mp.surface_z)
        PX = OracleFDEM(exact_mesh(om), sig, mu, "h")
        f = float(mp.freqs[1])
        se_np = np.real(src.s_e)
        h_ref = PX.solve(f, se_np, refine=True)
        j_ref = np.asarray(PX.j_from(f, h_ref, se_np), dtype=complex)
        j = ctx.fields_fdem("h", "j", f, u[1], se).cpu().numpy()
        rho = np.asarray(om.getFaceInnerProduct(sig, invProp=True).diagonal())
        cas = (mp.casing_a, mp.casing_b, tuple(mp.casing_z))
        iz, iz_ref = casing_currents(j, om, *cas)["z"][1], casing_currents(j_ref, om, *cas)["z"][1]
        out["parity_10Hz_vs_exactly_assembled_system"] = {
            "j_energy_norm": energy_rel(j, j_ref, rho), "casing_current_Iz": float(np.linalg.norm(iz - iz_ref) / np.linalg.norm(iz_ref)),
            "j_plain_l2": float(np.linalg.norm(j - j_ref) / np.linalg.norm(j_ref)),
            "j_plain_l2_moved_by_one_fp64_rounding_of_the_rhs": rhs_rounding_sensitivity(PX.getA(f), PX.getRHS(f, se_np), lambda x: PX.j_from(f, x, se_np))}
    except Exception as exc:                                  # parity is a report here; the tests assert it
        out["parity_error"] = str(exc)
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
    import casingsimulations_b200 as cs
    from casingsimulations_b200 import _lib
    stream = torch.cuda.Stream(device=local_rank)
    torch.cuda.set_stream(stream)               # torch copies, events and our kernels share one stream

    strong = getattr(args, "mode", "weak") == "strong"
    nfreq_total = NFREQ_STRONG if strong else NFREQ_PER_GPU * world
    all_freqs = np.logspace(0, 1, nfreq_total)
    freqs = np.ascontiguousarray(all_freqs[rank::world])          # this rank's frequencies (round robin)
    mp, mg, src = build_inputs(MESH_OURS, freqs)
    m = mg.mesh
    params = mp.paint_params()
    se_np = np.ascontiguousarray(np.real(src.s_e))
    ctx = _lib.Context(local_rank, stream=stream.cuda_stream)
    ctx.set_mesh(m.hx, m.hy, m.hz, m.x0[2])
    se_dev = torch.from_numpy(se_np).to(ctx.device)
    nE, nF = ctx.nE, ctx.nF
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=ctx.device)   # > 126 MB L2

    def step_device():
        ctx.paint(params)                      # K1 (drops cached operators -> full re-assembly, as the reference does)
        return ctx.solve_fdem("h", freqs, se_dev, rtol=1e-10)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    infos = []

    def timed(fn, steps, info_of=None):
        tot = 0.0
        for _ in range(steps):
            flush.zero_()                                       # L2 flush between timed iterations
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            fn()
            e1.record(stream)
            torch.cuda.synchronize()
            dt = e0.elapsed_time(e1) * 1e-3       # CUDA events on the stream the step's kernels are ordered on
            if info_of is not None:
                infos.append(dict(info_of()))
            if os.environ.get("BENCH_DEBUG"):
                print("rank", rank, fn.__name__, "step", round(dt * 1e3, 2), "ms", file=sys.stderr, flush=True)
            tot += dt
        return tot

    warm = max(args.warmup, 3)
    for _ in range(warm):
        step_device()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = ctx.launch_count()
    t_dev = timed(step_device, args.steps, lambda: ctx.info)
    launches = ctx.launch_count() - l0
    barrier()
    fp64_peak = ctx.fp64_peak_tflops()
    ctx.close()                                                  # the e2e arm owns its own context (and 130 GB of band storage)
    del se_dev

    # ---- end to end through the reference-facing API: SimulationFDEM(...).run(save=False) (run.py:166-216) ----
    tmpdir = tempfile.mkdtemp(prefix="bench_sim_")
    sim = cs.run.SimulationFDEM(modelParameters=mp, meshGenerator=mg, src=src, directory=tmpdir, device=local_rank)
    sink = open(os.devnull, "w")

    def step_e2e():
        so = sys.stdout
        sys.stdout = sink                                        # run() prints progress like the reference
        try:
            f = sim.run(save=False)                              # numpy s_e in (H2D), numpy Fields out (D2H)
        finally:
            sys.stdout = so
        return f
    for _ in range(2):
        step_e2e()
    barrier()
    t_e2e = timed(step_e2e, args.steps)
    barrier()
    clocks = sampler.stop()
    e2e_info = dict(sim.prob.info)
    sim.prob.clean()

    tt = torch.tensor([t_dev, t_e2e], dtype=torch.float64, device=torch.device("cuda", local_rank))
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_dev, t_e2e = float(tt[0]), float(tt[1])

    if rank == 0:
        def avg(key):
            return float(np.mean([i.get(key, 0.0) for i in infos])) if infos else None
        ms_factor, ms_solve, ms_asm = avg("ms_factor"), avg("ms_solve"), avg("ms_assemble")
        # band LU of the theta-mode systems: nth/2+1 complex band matrices of order n2, half bandwidth p, per frequency
        nx, nth, nz = (int(v) for v in m.vnC)
        n2, p, nsys = (3 * nx + 1) * (nz + 1), 3 * nx + 3, nth // 2 + 1
        # EXECUTED flops of the symmetric step LU (8 per complex FMA): per 16-column block step the row solves of the panel
        # (p rows x 16^2/2) and the 32x32 tiles on and below the diagonal of the trailing window (T (T+1)/2 tiles x 32x32x16,
        # T = ceil(p / 32)); a general LU of the same band would execute 8 n2 p^2 (reported beside it)
        tl = (p + 31) // 32
        flops = 8.0 * (n2 / 16.0) * (p * 128.0 + tl * (tl + 1) / 2 * 32 * 32 * 16) * nsys * len(freqs)
        flops_lu = 8.0 * n2 * float(p) ** 2 * nsys * len(freqs)
        ach = flops / (ms_factor * 1e-3) / 1e12
        # two-front elimination: the fronts meet at the middle z level, both advance in the same launches
        lrow, levels = 3 * nx + 1, nz + 1
        nt_ = lrow * (levels // 2)
        steps2 = ((max(nt_, n2 - nt_ - lrow) + 15) // 16 * 16 + lrow + 15) // 16
        roof = {"bound": "fp64", "kernel": "symmetric two-front band LU of the {} theta-mode systems per frequency (k_lu_panel4 + k_lu_update on the lower "
                                            "triangle, {} dependent 16-column block steps of both fronts replayed as one CUDA graph): {:.0f} % of the step's device time"
                                            .format(nsys, steps2, 100.0 * ms_factor / (t_dev / args.steps * 1e3)),
                "achieved": ach, "peak": fp64_peak, "peak_source": "live DFMA microbenchmark (cs_fp64_peak), 148 SMs",
                "unit": "TFLOP/s", "frac": ach / fp64_peak, "flops_per_step": flops,
                "general_lu_equivalent": {"flops_per_step": flops_lu, "tflops": flops_lu / (ms_factor * 1e-3) / 1e12},
                "ms_factor_per_step": ms_factor, "traffic": None,
                "traffic_note": "per-launch DRAM bytes of k_lu_panel4 / k_lu_update (0.7 / 9.6 MB: the windows are L2 resident) in profiles/r02_ncu_lu_final.md",
                "launch_shares": "profiles/r02_launch_list_3d.md"}
        try:
            roof["stencil_apply"] = stencil_roofline(torch, local_rank)
        except Exception as exc:
            roof["stencil_apply"] = {"error": str(exc)}
        cpu_base, same_mesh = None, None
        if world == 1:
            # the all-cores CPU figure comes from the reference arm itself, in a clean process (it forks workers)
            try:
                outp = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "2", "--warmup", "0"],
                                      capture_output=True, text=True, timeout=900, env=dict(os.environ, RANK="0", WORLD_SIZE="1", LOCAL_RANK="0"))
                ref = json.loads([l for l in outp.stdout.splitlines() if l.startswith("{")][-1])
                cpu_base = ref["cpu_baseline"]
            except Exception as exc:
                cpu_base = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": "reference arm failed: {}".format(exc)}
            try:
                same_mesh = small_mesh_check(torch, local_rank)
            except Exception as exc:
                same_mesh = {"error": str(exc)}
        value = t_dev / (args.steps * nfreq_total)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": warm, "ms_per_step": t_dev / args.steps * 1e3, "higher_is_better": False,
                "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "c128", "data": "synthetic",
                "config": {"workload": WORKLOAD if not strong else WORKLOAD.replace("{} frequencies 1-10 Hz per GPU".format(NFREQ_PER_GPU),
                                                                                     "{} frequencies 1-10 Hz in total".format(NFREQ_STRONG)),
                           "mesh": mesh_name(mg), "frequencies_per_gpu": len(freqs),
                           "unknowns_per_frequency": nE, "l2": "256 MB buffer written between timed iterations", "rtol": 1e-10,
                           "reference_arm_mesh": "109x4x40 = 17440 cells (largest the host solves in seconds; see cpu_baseline.sample)"},
                "frequencies_per_s": 1.0 / value,
                "e2e": {"value": t_e2e / (args.steps * nfreq_total), "unit": UNIT,
                        "api": "SimulationFDEM(modelParameters, meshGenerator, src).run(save=False)",
                        "h2d_bytes_per_step": int(nF * 8), "d2h_bytes_per_step": int(len(freqs) * nE * 16),
                        "relres_scaled": e2e_info.get("relres")},
                "gpu_launches": int(launches),
                "solve_info": {"krylov_iters": avg("iters"), "relres_scaled": avg("relres"), "ms_assemble": ms_asm,
                               "ms_factor": ms_factor, "ms_solve": ms_solve},
                "roofline": roof,
                "cpu_baseline": cpu_base,
                "same_mesh_as_reference_arm": same_mesh,
                "clocks": clocks}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# one system too big for one GPU, per GPU count: (nth, csz, npadz) -> mesh; the last is BASELINE configs[4]
ONE_SYSTEM = {1: (32, 0.75, 25), 2: (64, 0.75, 25), 4: (128, 0.75, 25), 8: (128, 0.3, 25)}


def run_one_system(args, rank, world, local_rank):
    """--mode one-system: ONE 3-D FDEM-h system over all N GPUs (z-slab decomposition, NCCL halo exchange, mode-sharded
    factors; tools/dist_fdem.py).  N = 2: 109x64x1394 (9.7 M cells), 4: 109x128x1394 (19.4 M), 8: 109x128x3394 = BASELINE
    configs[4] (47.4 M cells, 142 M complex unknowns) -- each needs more band storage than one B200 has."""
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from dist_fdem import solve_one_system
    nth, csz, npadz = ONE_SYSTEM.get(world, ONE_SYSTEM[8])
    res = solve_one_system(nth, csz, npadz, compare=False, reps=max(args.steps, 1) + 1)
    if rank == 0:
        line = {"metric": METRIC, "mode": "one-system", "value": res["s_per_frequency"], "unit": UNIT, "n_gpus": world,
                "steps": max(args.steps, 1), "warmup": 1, "ms_per_step": res["s_per_frequency"] * 1e3, "higher_is_better": False,
                "scaling": "one system partitioned over all GPUs (size grows with N)", "vs_baseline": None, "dtype": "c128", "data": "synthetic",
                "config": {"workload": "ONE 3-D FDEM-h casing system at 1 Hz, mesh {}x{}x{} = {} cells, {} complex edge unknowns, z slabs {}"
                           .format(res["mesh"][0], res["mesh"][1], res["mesh"][2], res["cells"], res["edges"], res["zsplit"])},
                "solve_info": res["info"], "gpu_mem_used_GB_rank0": res["gpu_mem_used_GB_rank0"]}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="weak", choices=["weak", "strong", "one-system"],
                    help="weak (default, the driver's contract): NFREQ_PER_GPU frequencies per GPU; strong: NFREQ_STRONG frequencies "
                         "in total, sharded over the GPUs; one-system: ONE system too big for one GPU, z slabs over all GPUs")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if args.mode == "one-system":
        run_one_system(args, rank, world, local_rank)
        return
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py -- RK4 cell-steps/s of the fused tendency+RK4 path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload NAME] [--dtype f64|f32]

A "step" is one RungeKutta4 step (four fused stage kernels) over the whole mesh.  Workloads
(BASELINE.json configs): igw4096 = configs[3], the 16.8 M-cell mesh the metric's 1/2/4/8-GPU scaling is quoted
on and the default at every N (so the per-N values are one strong-scaling series; it fits one GPU: 18 GB of
180 GB); igw2048 = configs[2] (single-B200 roofline run), igw512 = configs[1], igw64 = configs[0],
kelvin1024 = configs[4].
Inputs (10 GB of mesh + state per stage pass at 4096x4096) are far larger than the 126 MB L2, so no flush is needed.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "mpas-ocean.jl_b200"))

WORKLOADS = {"igw64": 64, "igw512": 512, "igw1024": 1024, "igw2048": 2048, "igw4096": 4096, "kelvin1024": 1024,
             # unstructured: periodic Voronoi mesh of a jittered lattice (~0.5 % pentagons, ~0.5 % heptagons, all metrics different)
             "voronoi64": 64, "voronoi1024": 1024, "voronoi2048": 2048,
             # spherical: quasi-uniform Voronoi mesh of the sphere (nx*nx cells), fEdge = 2 Omega sin(lat), geostrophic zonal flow + noise
             "sphere64": 64, "sphere1024": 1024, "sphere2048": 2048}
# algorithmic bytes per cell per RK4 step on a planar hex mesh (SURVEY.md 8d / BASELINE.md section 3): every distinct
# array element moved once per stage, edgesOnEdge included
ALGO_BYTES_PER_CELL_STEP = {"f64": 2400.0, "f32": 1536.0}
# the same tally for the kernel that rebuilds edgesOnEdge from edgesOnCell: per stage and cell
# 3 edges x (40 B of indices -> 1 position byte) less
ALGO_BYTES_PER_CELL_STEP_DERIVED = {"f64": 2400.0 - 4 * 3 * 39, "f32": 1536.0 - 4 * 3 * 39}


def algo_bytes_per_cell_step(dtype: str, derived_fraction: float) -> float:
    return derived_fraction * ALGO_BYTES_PER_CELL_STEP_DERIVED[dtype] + (1.0 - derived_fraction) * ALGO_BYTES_PER_CELL_STEP[dtype]


def algo_bytes_per_cell_step_general(m: dict, dtype: str, derived_fraction: float) -> float:
    """The same tally for an arbitrary mesh, from its live row lengths: per RK stage every edge moves
    cellsOnEdge 8 + g/dc R + dvEdge R + nEdgesOnEdge x (weight R + index 4, or 1 position byte where the row is rebuilt), every
    cell nEdgesOnCell x 4 + 1/area R + restingThickness R, every degree of freedom 4R on average over the four stages."""
    R = 8.0 if dtype == "f64" else 4.0
    nee, nec = float(np.sum(m["nEdgesOnEdge"])), float(np.sum(m["nEdgesOnCell"]))
    nE, nC = float(m["nEdges"]), float(m["nCells"])
    idx = derived_fraction * nE * 1.0 + (1.0 - derived_fraction) * 4.0 * nee
    per_stage = nE * (8.0 + 2.0 * R) + R * nee + idx + 4.0 * nec + 2.0 * R * nC + 4.0 * R * (nE + nC)
    return 4.0 * per_stage / nC


def workload_label(name: str, nx: int) -> str:
    """config.workload: the SAME string in both arms (--impl b200 / reference), one per BASELINE.json config."""
    if name.startswith("kelvin"):
        return f"coastal Kelvin wave, {nx}x{nx} channel hex mesh with boundary-edge masks, RK4"
    if name.startswith("sphere"):
        return f"geostrophic zonal flow + noise, quasi-uniform spherical Voronoi mesh of {nx * nx} cells, fEdge = 2 Omega sin(lat), RK4"
    if name.startswith("voronoi"):
        return f"inertial gravity wave, {nx}x{nx} periodic planar Voronoi mesh of a jittered lattice, RK4"
    return f"inertial gravity wave, {nx}x{nx} periodic planar hex mesh, RK4"


def parity_against_oracle(m, state, dt, nsteps, got_ssh, got_u, dtype, workload):
    """The bench line's `parity` object: the result of `nsteps` RK4 steps from the initial state, as this run's GPU path
    computed it, against the C oracle (oracle/moka_oracle.c, the CPU restatement of the reference path -- here only as the
    checker) on the same undecomposed mesh.  Tolerance: BASELINE.json's relative L2 1e-12 (Float64) / 1e-5 (Float32)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import moka_oracle_c as OC
    mm = OC.apply_boundary_mask(m) if workload.startswith("kelvin") else m
    if "edgeSignOnCell" not in mm:
        OC.sign_index_fields(mm)
    ssh, u, h = state
    om = OC.OracleModel(mm, ssh, u, h)
    om.run_loop(dt, nsteps, "RungeKutta4")
    rel = lambda a, b: float(np.linalg.norm(np.asarray(a, np.float64) - b) / np.linalg.norm(b))
    tol = 1e-12 if dtype == "f64" else 1e-5
    e_ssh, e_u = rel(got_ssh, om.ssh[1]), rel(got_u, om.normalVelocity[1])
    return {"vs": "C oracle (oracle/moka_oracle.c: CPU restatement of the reference path) on the undecomposed mesh",
            "steps": nsteps, "rel_l2_ssh": e_ssh, "rel_l2_normalVelocity": e_u, "tolerance": tol,
            "bit_identical": bool(dtype == "f64" and np.array_equal(got_ssh, om.ssh[1]) and np.array_equal(got_u, om.normalVelocity[1])),
            "ok": bool(e_ssh <= tol and e_u <= tol),
            "oracle_pinning": "operators pinned by the reference's six golden vectors; tendency / RK4 field values unpinned by the reference (DESIGN.md section 8)"}


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device, self.rows, self.proc = device, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.t.join(timeout=2)
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": reasons}


def build_case(nx: int, dtype: str):
    import moka_b200 as mb
    t0 = time.time()
    if dtype == "kelvin":
        m = mb.channel_hex(nx, nx, 1.0e7 / nx)
        ssh, u, h = mb.kelvinWave(m).initial_state()
    elif dtype == "sphere":
        m = mb.spherical_voronoi(nx * nx, with_dual=False)
        ssh, u, h = mb.geostrophic_zonal_flow(m)
        u = u + 0.1 * np.random.default_rng(0).standard_normal(m["nEdges"])
        return m, (ssh, u, h), 0.25 * float(m["dcEdge"].min()) / float(np.sqrt(9.80616 * 1000.0)), time.time() - t0
    elif dtype == "voronoi":
        m = mb.periodic_voronoi(nx, nx, 1.0e7 / nx, jitter=0.25, seed=2, allow_obtuse=True, with_dual=False)
        ssh, u, h = mb.inertialGravityWave(m).initial_state()
        return m, (ssh, u, h), 0.25 * mb.cfl_dt(m["dc"]), time.time() - t0      # the shortest dcEdge is about half the mean
    else:
        m = mb.periodic_hex(nx, nx, 1.0e7 / nx, with_dual=False)
        ssh, u, h = mb.inertialGravityWave(m).initial_state()
    return m, (ssh, u, h), mb.cfl_dt(m["dc"]), time.time() - t0


def run_b200(args):
    import moka_b200 as mb
    from moka_b200 import _lib as L
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        from moka_b200 import multi_gpu
        return multi_gpu.bench_main(args, rank, world, local)

    nx = WORKLOADS[args.workload]
    npdt = np.float64 if args.dtype == "f64" else np.float32
    kelvin, sphere = args.workload.startswith("kelvin"), args.workload.startswith("sphere")
    voronoi = args.workload.startswith("voronoi") or sphere           # unstructured: byte accounting from the actual rows
    m, (ssh, u, h), dt, t_gen = build_case(nx, "kelvin" if kelvin else "sphere" if sphere else "voronoi" if voronoi else args.dtype)
    nC, nE = m["nCells"], m["nEdges"]
    backend = mb.B200(local)
    t0 = time.time()
    mesh = mb.Mesh(m, backend, explicit_eoe=args.explicit_eoe)
    t_mesh = time.time() - t0
    nblk, nder = mesh.derived_blocks()
    prog = mb.PrognosticVars(ssh.astype(npdt), u.astype(npdt), h.astype(npdt), 2, mesh)
    diag, tend = mb.DiagnosticVars(prog), mb.TendencyVars(prog)
    K, W = args.steps, max(args.warmup, 3)

    # ---- device-resident throughput (inputs already in HBM) -----------------------------------------
    mb.ocn_run_loop(dt, prog, diag, tend, None, mb.RungeKutta4, W)
    backend.synchronize()
    # the timed region is K steps repeated until it lasts >= 0.5 s (at 6 ms/step 20 steps are 0.12 s: one hiccup would be 5 %)
    backend.timer_start()
    mb.ocn_timestep(dt, prog, diag, tend, None, mb.RungeKutta4, nsteps=K)
    reps = 1 if args.quick else int(max(1, np.ceil(500.0 / max(backend.timer_stop(), 1e-3))))
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    l0 = backend.launch_count()
    backend.timer_start()
    for _ in range(reps):
        mb.ocn_timestep(dt, prog, diag, tend, None, mb.RungeKutta4, nsteps=K)
    ms = backend.timer_stop()
    launches_total = backend.launch_count() - l0
    steps_timed = K * reps
    stage_launches = 4 * steps_timed                        # + 2 ssh refresh kernels at the end of every call
    # keep the GPU under the same load while nvidia-smi gets its samples (not part of the number)
    t_end = time.time() + (0.0 if args.quick else 1.0)
    while time.time() < t_end:
        mb.ocn_timestep(dt, prog, diag, tend, None, mb.RungeKutta4, nsteps=10)
        backend.synchronize()
    clocks = sampler.stop()
    value = nC * steps_timed / (ms * 1e-3)

    # ---- end to end through the public API with HOST buffers ----------------------------------------------
    # every step: H2D of that step's (normalVelocity, layerThickness) from pinned memory, one RK4 step, D2H of ssh.
    # The copies go through the pipelined upload/download of the API, so the PCIe transfer of step n+1 and
    # n-1 overlap the kernels of step n; two alternating sets of host buffers stand for distinct inputs.
    # The step's result that comes back is (ssh, normalVelocity) of the new state; layerThickness = ssh + restingThicknessSum is
    # redundant with ssh (restingThicknessSum is static and already on the host) and stays on the device.  Float32 states
    # upload ssh instead of layerThickness (the perturbation is their prognostic variable, include/moka_b200.h).
    f32 = npdt == np.float32
    hin = [(backend.pinned(nE, npdt), backend.pinned(nC, npdt)) for _ in range(2)]
    hout = [(backend.pinned(nC, npdt), backend.pinned(nE, npdt)) for _ in range(2)]
    for hu, hh in hin:
        hu[:], hh[:] = u.astype(npdt), (ssh if f32 else h).astype(npdt)
    def e2e_steps(n):
        for i in range(n):
            if f32:
                prog.upload_async(normalVelocity=hin[i & 1][0], ssh=hin[i & 1][1])
            else:
                prog.upload_async(normalVelocity=hin[i & 1][0], layerThickness=hin[i & 1][1])
            mb.ocn_timestep(dt, prog, diag, tend, None, mb.RungeKutta4, nsteps=1)
            prog.download_async(ssh=hout[i & 1][0], normalVelocity=hout[i & 1][1])
        prog.synchronize()
    e2e_steps(1 if args.quick else 3)
    Ke = 2 if args.quick else max(3, min(K, 50))
    t0 = time.perf_counter()
    e2e_steps(Ke)
    e2e_s = (time.perf_counter() - t0) / Ke
    # what the same transfers cost with no step between them (explains the leg: PCIe-bound when the two agree)
    def copy_only(n):
        for i in range(n):
            if f32:
                prog.upload_async(normalVelocity=hin[i & 1][0], ssh=hin[i & 1][1])
            else:
                prog.upload_async(normalVelocity=hin[i & 1][0], layerThickness=hin[i & 1][1])
            prog.download_async(ssh=hout[~i & 1][0], normalVelocity=hout[~i & 1][1])
        prog.synchronize()
    copy_only(1)
    t0 = time.perf_counter()
    copy_only(Ke)
    copy_s = (time.perf_counter() - t0) / Ke
    e2e_steps(2)                                            # both result buffers hold a step's output again for the check below
    item = np.dtype(npdt).itemsize
    # the result of the last e2e step is one RK4 step from the uploaded state: check it against a device-resident step
    prog_chk = mb.PrognosticVars(ssh.astype(npdt), u.astype(npdt), h.astype(npdt), 2, mesh)
    mb.ocn_timestep(dt, prog_chk, diag, tend, None, mb.RungeKutta4, nsteps=1)
    e2e_ok = bool(np.array_equal(prog_chk.ssh, hout[(Ke - 1) & 1][0]) and np.array_equal(prog_chk.normalVelocity, hout[(Ke - 1) & 1][1]))
    # ---- parity: two RK4 steps from the initial state against the C oracle (checker only) --------------------------------
    parity = None
    if not args.no_parity:
        mb.ocn_timestep(dt, prog_chk, diag, tend, None, mb.RungeKutta4, nsteps=1)
        parity = parity_against_oracle(m, (ssh, u, h), dt, 2, prog_chk.ssh, prog_chk.normalVelocity, args.dtype, args.workload)
    del prog_chk

    # ---- roofline of the dominant kernel (k_rk_stage) -----------------------------------------------------------
    peak, peak_src = measured_peak_gbs()
    per_cell_step = (algo_bytes_per_cell_step_general(m, args.dtype, nder / nblk) if voronoi
                     else algo_bytes_per_cell_step(args.dtype, nder / nblk))
    algo_bytes_per_launch = per_cell_step / 4.0 * nC
    avg_launch_s = (ms * 1e-3) / stage_launches
    achieved = algo_bytes_per_launch / avg_launch_s / 1e9
    survey_rate = ALGO_BYTES_PER_CELL_STEP[args.dtype] / 4.0 * nC / avg_launch_s / 1e9
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            key = f"{args.workload}_{args.dtype}" + ("_explicit" if args.explicit_eoe else "")
            traffic = (json.load(f).get(key) or {}).get("bytes_per_launch")
    except Exception:
        pass

    out = {
        "metric": "RK4 cell-steps/sec", "value": value, "unit": "cell-steps/s", "n_gpus": 1, "steps": K, "warmup": W,
        "ms_per_step": ms / steps_timed, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": args.dtype, "data": "synthetic",
        "config": {"workload": workload_label(args.workload, nx),
                   "detail": f"{nC} cells, {nE} edges, {'Float64' if args.dtype == 'f64' else 'Float32'}, dt={dt:.4g}s"
                             + (f", polygons by side count from 5: {np.bincount(m['nEdgesOnCell'])[5:].tolist()}; device rows "
                                f"{mesh.maxEdges2} / {mesh.maxEdges}" if voronoi else ""),
                   "timed_steps": steps_timed, "repeats_of_steps": reps,
                   "name": args.workload, "l2": "inputs larger than L2 (no flush)" if nx >= 1024 else "fits in L2",
                   "mesh_gen_s": round(t_gen, 2), "mesh_upload_s": round(t_mesh, 2),
                   "blocks_rebuilding_edgesOnEdge": [int(nder), int(nblk)],
                   # which build / stage-kernel variant ran (tools/gpu_sweep_variants.sh; the defaults when unset)
                   "variant": {"lib": os.environ.get("MOKAB_LIB", "libmoka_b200.so"),
                               "stage_tma": L.get_option("stage_tma"), "stage_prefetch": L.get_option("stage_prefetch"),
                               "stage_prefetch_distance": L.get_option("stage_prefetch_distance"),
                               "stage_auto": L.get_option("stage_auto"), "stage_pdl": L.get_option("stage_pdl")}},
        "clocks": clocks,
        "e2e": {"value": nC / e2e_s, "unit": "cell-steps/s", "h2d_bytes_per_step": int((nE + nC) * item),
                "d2h_bytes_per_step": int((nC + nE) * item), "ms_per_step": e2e_s * 1e3, "steps": Ke,
                "pipelined": True, "matches_device_resident_step": e2e_ok, "copies_only_ms_per_step": copy_s * 1e3,
                "returns": "ssh + normalVelocity of the new state (layerThickness = ssh + restingThicknessSum stays on the device)"},
        "gpu_launches": int(launches_total),
        "parity": parity,
        "roofline": {"bound": "hbm", "kernel": "k_rk_stage", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                     "traffic_source": "profiles/traffic.json (dram__bytes_read.sum + dram__bytes_write.sum of one committed `ncu --set full` capture, per launch; not measured in this run)" if traffic else None,
                     "algorithmic_bytes_per_launch": algo_bytes_per_launch, "avg_launch_ms": avg_launch_s * 1e3,
                     "bytes_per_cell_step": algo_bytes_per_launch * 4.0 / nC,
                     "rate_if_edgesOnEdge_were_read": survey_rate,
                     "note": "achieved = bytes this kernel must move (edgesOnEdge rebuilt from edgesOnCell where the mesh allows, "
                             "DESIGN.md section 5) / launch time; rate_if_edgesOnEdge_were_read applies SURVEY.md 8d's 2400 B (F64) "
                             "/ 1536 B (F32) per cell-step to the same time"},
    }
    if not args.no_cpu:
        out["cpu_baseline"] = cpu_baseline(m, (ssh, u, h), dt, budget_s=args.cpu_budget, kelvin=kelvin)
    print(json.dumps(out))


def cpu_baseline(m, state, dt, budget_s=15.0, kelvin=False):
    """The C/OpenMP restatement of the reference path (oracle/), timed on this box's host cores on a
    bounded sample: RK4 steps on the same mesh until ~budget_s of CPU work."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import moka_oracle_c as OC
    ssh, u, h = state
    if kelvin:
        m = OC.apply_boundary_mask(m)
    OC.set_num_threads(os.cpu_count() or 1)
    om = OC.OracleModel(m, ssh, u, h)
    om.run_loop(dt, 1, "RungeKutta4")                       # warm-up (thread pool, page faults)
    t0 = time.perf_counter()
    om.run_loop(dt, 1, "RungeKutta4")
    t1 = time.perf_counter() - t0
    n = int(max(1, min(50, budget_s / max(t1, 1e-6))))
    chunk, best, tt, done = max(1, n // 3), 0.0, 0.0, 0
    while done < n:                                          # best chunk: the host is shared with the GPU process's threads
        k = min(chunk, n - done)
        t0 = time.perf_counter()
        om.run_loop(dt, k, "RungeKutta4")
        t = time.perf_counter() - t0
        best, tt, done = max(best, m["nCells"] * k / t), tt + t, done + k
    return {"value": best, "unit": "cell-steps/s", "cores": om.num_threads(), "kind": "port",
            "sample": f"{n} RK4 steps of the same {m['nCells']}-cell mesh in chunks of {chunk} (best chunk reported), C/OpenMP "
                      f"restatement of the reference's unfused per-stage kernel sequence (no Julia in this image), {tt:.1f} s"}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path.  The reference is Julia and
    Julia is not installed, so this is the oracle port (kind "port") with all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    nx = WORKLOADS[args.workload]
    kelvin, sphere, voronoi = (args.workload.startswith(k) for k in ("kelvin", "sphere", "voronoi"))
    m, (ssh, u, h), dt, _ = build_case(nx, "kelvin" if kelvin else "sphere" if sphere else "voronoi" if voronoi else "f64")
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import moka_oracle_c as OC
    if kelvin:
        m = OC.apply_boundary_mask(m)
    if sphere or voronoi:
        OC.sign_index_fields(m)
    # each "step" is a bounded sample: one RK4 step on the workload mesh (capped so the run ends in minutes)
    # all the host cores, explicitly: under torchrun the environment carries OMP_NUM_THREADS=1
    OC.set_num_threads(os.cpu_count() or 1)
    om = OC.OracleModel(m, ssh, u, h)
    K, W = args.steps, args.warmup
    om.run_loop(dt, 1, "RungeKutta4")
    t0 = time.perf_counter()
    om.run_loop(dt, 1, "RungeKutta4")
    t1 = time.perf_counter() - t0
    K = int(max(1, min(K, 120.0 / max(t1, 1e-6))))
    W = int(max(0, min(W, 20.0 / max(t1, 1e-6))))
    if W:
        om.run_loop(dt, W, "RungeKutta4")
    t0 = time.perf_counter()
    om.run_loop(dt, K, "RungeKutta4")
    tt = time.perf_counter() - t0
    v = m["nCells"] * K / tt
    print(json.dumps({
        "impl": "reference", "metric": "RK4 cell-steps/sec", "value": v, "unit": "cell-steps/s", "n_gpus": world,
        "steps": K, "warmup": W, "ms_per_step": tt / K * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_label(args.workload, nx), "detail": f"{m['nCells']} cells, Float64",
                   "name": args.workload},
        "cpu_baseline": {"value": v, "unit": "cell-steps/s", "cores": om.num_threads(), "kind": "port",
                         "sample": f"{K} RK4 steps on the full {m['nCells']}-cell mesh, C/OpenMP restatement of the "
                                   "reference's per-stage kernel sequence (Julia absent)"},
        "e2e": {"value": v, "unit": "cell-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS))
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-parity", dest="no_parity", action="store_true", help="skip the parity check against the C oracle")
    ap.add_argument("--cpu-budget", type=float, default=15.0)
    ap.add_argument("--no-overlap", dest="no_overlap", action="store_true", help="multi-GPU: exchange after each full stage")
    ap.add_argument("--no-graph", dest="no_graph", action="store_true", help="multi-GPU: do not capture steps into a CUDA graph")
    ap.add_argument("--halo", default="p2p", choices=["nccl", "p2p", "p2p_fused", "p2p_ll"],
                    help="multi-GPU halo exchange: direct stores into the peers' memory with push / wait kernels (default: measured "
                         "at N = 2 and N = 8, profiles/README.md r02c / r02e), packed NCCL send/recv, direct stores from inside the "
                         "boundary launch, or flag-in-data packets (p2p_ll: no fence on the sending side, +22 %% on 131 k-cell parts "
                         "at N = 2, equal on large ones, r02p; not yet run at N = 8)")
    ap.add_argument("--explicit-eoe", dest="explicit_eoe", action="store_true",
                    help="ablation: read edgesOnEdge from memory instead of rebuilding it from edgesOnCell")
    ap.add_argument("--quick", action="store_true", help="profiling runs: no clock-sampling load loop, one e2e step")
    args = ap.parse_args()
    if args.workload is None:
        args.workload = "igw4096"
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()

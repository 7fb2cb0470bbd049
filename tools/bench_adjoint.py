#!/usr/bin/env python
"""Throughput of the reverse mode (tape + adjoint sweep) of the fused RK4 path on one B200.

  python tools/bench_adjoint.py [--workload igw2048] [--steps 10] [--dtype f64] [--stepper rk4|fe]

Prints one JSON line: forward-with-tape and reverse cell-steps/s and the algorithmic HBM rate of the reverse
sweep (bytes per reversed cell-step: DESIGN.md section 5)."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mpas-ocean.jl_b200"))
sys.path.insert(0, ROOT)

# per reversed cell-step on a hex mesh: 3 recomputed forward stages + 4 adjoint stages
FE_ADJ_BYTES = 3 * (8 + 40 + 80 + 8 + 16 + 8 + 16) + (24 + 8 + 16 + 24)      # ForwardEuler (Float64 only): 600 B per reversed cell-step
ADJ_BYTES = {"f64": 3 * 472 + (96 + 160 + 160) + 4 * 464 + (4 + 6 + 6 + 3) * 32, "f32": 3 * 320 + (48 + 80 + 80) + 4 * 312 + (4 + 6 + 6 + 3) * 16}


def main():
    import moka_b200 as mb
    from bench import WORKLOADS, measured_peak_gbs
    from moka_b200 import _lib as L
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="igw2048", choices=sorted(WORKLOADS))
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"])
    ap.add_argument("--stepper", default="rk4", choices=["rk4", "fe"])
    args = ap.parse_args()
    fe = args.stepper == "fe"
    if fe and args.dtype != "f64":
        ap.error("ForwardEuler is Float64 only")
    nx = WORKLOADS[args.workload]
    npdt = np.float64 if args.dtype == "f64" else np.float32
    m = mb.periodic_hex(nx, nx, 1.0e7 / nx, with_dual=False)
    ssh, u, h = mb.inertialGravityWave(m).initial_state()
    dt = mb.cfl_dt(m["dc"])
    backend = mb.B200(0)
    mesh = mb.Mesh(m, backend)
    prog = mb.PrognosticVars(ssh.astype(npdt), u.astype(npdt), h.astype(npdt), 2, mesh)
    d_prog = mb.ocn_init_shadows(prog)
    K, lib, hd = args.steps, L.lib(), prog.dev.handle
    mb.autodiff_reverse_run_loop(dt, prog, d_prog, None, None, None, mb.ForwardEuler if fe else mb.RungeKutta4, 2)     # warm-up: builds the transposed stencil
    backend.synchronize()
    best = {"fwd": 1e30, "rev": 1e30}
    for _ in range(3):
        L.check(lib.mokab_tape_begin(hd, K))
        backend.timer_start()
        L.check(lib.mokab_timestep_forward_euler(hd, dt, K) if fe else lib.mokab_timestep_rk4(hd, dt, K, L.RK4_FUSED))
        best["fwd"] = min(best["fwd"], backend.timer_stop())
        L.check(lib.mokab_adjoint_seed(hd, L.SUM_SSH2))
        l0 = backend.launch_count()
        backend.timer_start()
        L.check(lib.mokab_adjoint_forward_euler(hd) if fe else lib.mokab_adjoint_rk4(hd))
        best["rev"] = min(best["rev"], backend.timer_stop())
        launches = backend.launch_count() - l0
    nC = m["nCells"]
    peak, src = measured_peak_gbs()
    rev = nC * K / (best["rev"] * 1e-3)
    g = d_prog.layerThickness
    nbytes = FE_ADJ_BYTES if fe else ADJ_BYTES[args.dtype]
    print(json.dumps({
        "metric": f"reverse-mode {'ForwardEuler' if fe else 'RK4'} cell-steps/sec", "value": rev, "unit": "cell-steps/s", "n_gpus": 1, "steps": K,
        "dtype": args.dtype, "workload": args.workload, "ms_per_reversed_step": best["rev"] / K,
        "forward_with_tape_cell_steps_per_s": nC * K / (best["fwd"] * 1e-3), "ms_per_taped_forward_step": best["fwd"] / K,
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernels": "1 x k_fe_step_adj per reversed step" if fe else
                     "3 x k_rk_stage (recompute) + 4 x k_rk_stage_adj per reversed step",
                     "algorithmic_bytes_per_cell_step": nbytes, "achieved": nbytes * rev / 1e9,
                     "peak": peak, "unit": "GB/s", "frac": nbytes * rev / 1e9 / peak, "peak_source": src},
        "gradient_finite": bool(np.all(np.isfinite(g))), "gradient_l2": float(np.linalg.norm(g)),
    }))


if __name__ == "__main__":
    main()

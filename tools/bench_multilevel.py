#!/usr/bin/env python
"""Throughput of the fused RK4 path on multi-level states (nVertLevels = K) on one B200.

  python tools/bench_multilevel.py [--workload igw1024] [--levels 1,4,10,20] [--steps 20]

One JSON line per K: level-cell-steps/s (= nCells * K * steps / time), the bytes a stage must move per cell (static data read
once per column: 355 + 128 K + 16 for the ssh round trip) and the fraction of the measured HBM peak."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mpas-ocean.jl_b200"))
sys.path.insert(0, ROOT)


def main():
    import moka_b200 as mb
    from bench import WORKLOADS, measured_peak_gbs
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="igw1024", choices=sorted(w for w in WORKLOADS if w.startswith("igw")))
    ap.add_argument("--levels", default="1,4,10,20")
    ap.add_argument("--steps", type=int, default=20)
    args = ap.parse_args()
    nx = WORKLOADS[args.workload]
    m = mb.periodic_hex(nx, nx, 1.0e7 / nx, with_dual=False)
    ssh, u, h = mb.inertialGravityWave(m).initial_state()
    dt = mb.cfl_dt(m["dc"])
    backend = mb.B200(0)
    peak, src = measured_peak_gbs()
    nC = m["nCells"]
    for K in [int(k) for k in args.levels.split(",")]:
        mk = dict(m)
        mk["restingThickness"], mk["nVertLevels"] = np.full((nC, K), 1000.0 / K), K
        mesh = mb.Mesh(mk, backend)
        uk = np.outer(u, np.ones(K)) if K > 1 else u
        hk = np.outer(h, np.full(K, 1.0 / K)) if K > 1 else h
        prog = mb.PrognosticVars(ssh, uk, hk, 2, mesh)
        mb.ocn_timestep(dt, prog, None, None, None, mb.RungeKutta4, nsteps=4)
        backend.synchronize()
        best = 1e30
        for _ in range(3):
            backend.timer_start()
            mb.ocn_timestep(dt, prog, None, None, None, mb.RungeKutta4, nsteps=args.steps)
            best = min(best, backend.timer_stop())
        # explicit edgesOnEdge (the multi-level kernel reads it): static 472 B per cell and stage, dynamic 128 B per level, ssh 16 B
        per_stage = (483.0 if K == 1 else 472.0 + 16.0 + 128.0 * K)
        rate = nC * args.steps / (best * 1e-3)
        print(json.dumps({"workload": args.workload, "levels": K, "ms_per_step": best / args.steps, "cell_steps_per_s": rate,
                          "level_cell_steps_per_s": rate * K, "bytes_per_cell_stage": per_stage,
                          "achieved_gbs": 4 * per_stage * rate / 1e9, "frac_of_peak": 4 * per_stage * rate / 1e9 / peak, "peak_source": src,
                          "finite": bool(np.all(np.isfinite(prog.ssh)))}), flush=True)
        del prog, mesh


if __name__ == "__main__":
    main()

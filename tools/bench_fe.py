#!/usr/bin/env python
"""Throughput of the reference's live stepper (ForwardEuler, reference operation order incl. its quirks) on one B200.

  python tools/bench_fe.py [--workload igw2048] [--steps 50] [--no-dual]

One JSON line: ForwardEuler cell-steps/s (fused kernel and the kernel-per-reference-kernel sequence) next to the fused
RK4 rate on the same mesh.  Algorithmic bytes of a fused FE step on a hex mesh: 576 B per cell."""
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
    from bench import WORKLOADS
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="igw2048", choices=sorted(WORKLOADS))
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--no-dual", action="store_true", help="mesh without vertex arrays (no relativeVorticity diagnostic)")
    args = ap.parse_args()
    nx = WORKLOADS[args.workload]
    m = mb.periodic_hex(nx, nx, 1.0e7 / nx, with_dual=not args.no_dual)
    ssh, u, h = mb.inertialGravityWave(m).initial_state()
    dt = 0.2 * mb.cfl_dt(m["dc"])
    backend = mb.B200(0)
    mesh = mb.Mesh(m, backend)
    out = {}
    for name, stepper, fused in (("ForwardEuler", mb.ForwardEuler, True), ("ForwardEuler_unfused", mb.ForwardEuler, False),
                                 ("RungeKutta4", mb.RungeKutta4, True)):
        prog = mb.PrognosticVars(ssh, u, h, 2, mesh)
        mb.ocn_timestep(dt, prog, None, None, None, stepper, nsteps=5, fused=fused)
        backend.synchronize()
        l0 = backend.launch_count()
        backend.timer_start()
        mb.ocn_timestep(dt, prog, None, None, None, stepper, nsteps=args.steps, fused=fused)
        ms = backend.timer_stop()
        out[name] = {"cell_steps_per_s": m["nCells"] * args.steps / (ms * 1e-3), "ms_per_step": ms / args.steps,
                     "launches_per_step": (backend.launch_count() - l0) / args.steps}
    print(json.dumps({"workload": args.workload, "dual": not args.no_dual, "steps": args.steps, **out}))


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""A/B of the run-time variants of the fused RK4 stage kernel in ONE process: the mesh is generated and uploaded once, every
variant (mokab_set_option: stage_prefetch, stage_prefetch_distance, stage_tma) then runs W warm-up + K timed steps on the
same state, timed with CUDA events on the launching stream (mokab_timer_*), and its result is compared bit for bit with the
default variant's (all variants keep the operation order).  Prints one JSON line per variant and, last, the best one per
precision ({"best": ...}) -- tools/gpu_round2.sh feeds that to bench.py through the environment.

  python tools/stage_sweep.py [--workload igw2048] [--steps 40] [--warmup 5] [--dtypes f64,f32] [--explicit-eoe]
                              [--variants prefetch:distance:tma[:0[:pdl[:auto]]],...]
Build variants (block size, resident blocks) are separate libraries: select with MOKAB_LIB and run this once per library.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mpas-ocean.jl_b200"))
sys.path.insert(0, ROOT)


def main():
    import bench
    import moka_b200 as mb
    from moka_b200 import _lib as L
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="igw2048", choices=sorted(bench.WORKLOADS))
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--dtypes", default="f64,f32")
    ap.add_argument("--explicit-eoe", dest="explicit_eoe", action="store_true")
    ap.add_argument("--variants", default="", help="comma list of prefetch:distance:tma triples; default = the built-in sweep")
    ap.add_argument("--edge-orders", default="slot_major", help="comma list of slot_major (default numbering) / by_cell (MOKAB_MESH_EDGES_BY_CELL)")
    args = ap.parse_args()
    nx = bench.WORKLOADS[args.workload]
    kelvin = args.workload.startswith("kelvin")
    m, (ssh, u, h), dt, _ = bench.build_case(nx, "kelvin" if kelvin else "f64")
    nC = m["nCells"]
    backend = mb.B200(0)
    peak, _ = bench.measured_peak_gbs()
    if args.variants:
        variants = [tuple(int(x) for x in v.split(":")) for v in args.variants.split(",")]     # prefetch:distance:tma[:0[:pdl[:auto]]]
    else:
        variants = [(0, 0, 0), (1, 0, 0), (2, 0, 0), (3, 0, 0), (2, 888, 0), (3, 888, 0), (2, 1184, 0), (3, 1184, 0),
                    (0, 0, 1), (1, 0, 1), (0, 0, 2), (1, 0, 2), (3, 0, 2)]
    best = {}
    for order in args.edge_orders.split(","):
        mesh = mb.Mesh(m, backend, explicit_eoe=args.explicit_eoe, edges_by_cell=(order == "by_cell"))
        nblk, nder = mesh.derived_blocks()
        for dtype in args.dtypes.split(","):
            npdt = np.float64 if dtype == "f64" else np.float32
            per_cell_step = bench.algo_bytes_per_cell_step(dtype, nder / nblk)
            ref = None
            for var in variants:
                pf, dist, tma = var[:3]
                pdl = var[4] if len(var) > 4 else 0              # (field 3 was the shared-memory flux experiment of r02h: removed, -9 %)
                auto = var[5] if len(var) > 5 else 0            # (explicit variants are forced; ...:1 = let the library choose per launch)
                L.set_option("stage_auto", auto)
                L.set_option("stage_prefetch", pf)
                L.set_option("stage_prefetch_distance", dist)
                L.set_option("stage_tma", tma)
                L.set_option("stage_pdl", pdl)
                prog = mb.PrognosticVars(ssh.astype(npdt), u.astype(npdt), h.astype(npdt), 2, mesh)
                try:
                    mb.ocn_timestep(dt, prog, None, None, None, mb.RungeKutta4, nsteps=args.warmup)
                    backend.synchronize()
                    times = []
                    for _ in range(3):                                   # best of three timed batches
                        backend.timer_start()
                        mb.ocn_timestep(dt, prog, None, None, None, mb.RungeKutta4, nsteps=args.steps)
                        times.append(backend.timer_stop())
                    ms = min(times)
                    out = (prog.ssh, prog.normalVelocity)
                except mb.MokaError as ex:
                    print(json.dumps({"dtype": dtype, "prefetch": pf, "distance": dist, "tma": tma, "pdl": pdl, "auto": auto, "error": str(ex)}), flush=True)
                    continue
                if ref is None:
                    ref = out
                same = bool(np.array_equal(out[0], ref[0]) and np.array_equal(out[1], ref[1]))
                value = nC * args.steps / (ms * 1e-3)
                frac = per_cell_step * value / 1e9 / peak
                rec = {"dtype": dtype, "prefetch": pf, "distance": dist, "tma": tma, "pdl": pdl, "auto": auto, "ms_per_step": ms / args.steps,
                    "cell_steps_per_s": value, "roofline_frac": frac, "bit_identical_to_default": same,
                    "workload": args.workload, "lib": os.environ.get("MOKAB_LIB", "libmoka_b200.so"),
                    "explicit_eoe": args.explicit_eoe, "edge_order": order}
                print(json.dumps(rec), flush=True)
                if same and (dtype not in best or value > best[dtype]["cell_steps_per_s"]):
                    best[dtype] = rec
                del prog
    print(json.dumps({"best": best}), flush=True)


if __name__ == "__main__":
    t0 = time.time()
    main()
    print(f"# stage_sweep: {time.time() - t0:.1f} s", file=sys.stderr)

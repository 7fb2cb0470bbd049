#!/usr/bin/env python
"""Timeline of the stage and halo kernels of a decomposed run, from the records the TRACE build of the library writes
(csrc/common.cuh: every block appends %globaltimer at entry / after its wait for the peers / at exit).  nsys is not in the image and
ncu serialises kernels; this answers "where do the ~28 us per stage of a 131 k-cell part go?".

  MOKAB_LIB=libmoka_b200_trace.so python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \\
      tools/trace_stages.py --workload igw512 [--halo p2p|p2p_fused|nccl] [--steps 20] [--no-graph] [--no-overlap]
Rank 0 prints, for the launches of the last traced steps, start / end relative to the first launch (us), the duration, the time the
launch spent waiting for its peers and the gap to the launch before it on the same list; then one JSON line with the per-kind medians."""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mpas-ocean.jl_b200"))
sys.path.insert(0, ROOT)

REC = np.dtype([("kind", "<u4"), ("block", "<u4"), ("grid", "<u4"), ("pad", "<u4"), ("t0", "<u8"), ("t1", "<u8"), ("t2", "<u8")])
PARTS = {0: "all", 1: "interior", 2: "boundary", 3: "boundary+push", 4: "all+push"}


def label(kind: int) -> str:
    if kind == 100:
        return "halo push"
    if kind == 101:
        return "halo wait"
    if kind == 102:
        return "push (flag-in-data)"
    if kind == 103:
        return "wait + unpack (f-i-d)"
    if kind == 110:
        return "pack"
    if kind == 111:
        return "unpack"
    return f"stage{kind & 15} {PARTS.get(kind >> 4, kind >> 4)}"


def launches(rec: np.ndarray):
    """Group the block records into launches: per kind, sorted by entry time, consecutive runs of `grid` records."""
    out = []
    for kind in np.unique(rec["kind"]):
        r = rec[rec["kind"] == kind]
        r = r[np.argsort(r["t0"], kind="stable")]
        i = 0
        while i < r.size:
            g = int(r["grid"][i])
            chunk = r[i:i + g]
            gate = chunk["t2"][chunk["t2"] > 0]
            out.append({"kind": int(kind), "grid": g, "t0": int(chunk["t0"].min()), "t1": int(chunk["t1"].max()),
                        "gate": int(gate.max()) if gate.size else 0, "complete": chunk.size == g})
            i += g
    out.sort(key=lambda x: x["t0"])
    return out


def main():
    import torch
    import torch.distributed as dist

    import bench
    import moka_b200 as mb
    from moka_b200 import _lib as L
    from moka_b200 import multi_gpu
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="igw512", choices=sorted(bench.WORKLOADS))
    ap.add_argument("--halo", default="p2p", choices=["nccl", "p2p", "p2p_fused", "p2p_ll"])
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--dtype", default="f64")
    ap.add_argument("--no-graph", dest="no_graph", action="store_true")
    ap.add_argument("--no-overlap", dest="no_overlap", action="store_true")
    ap.add_argument("--show", type=int, default=2, help="how many steps of launches to print")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")
    nx = bench.WORKLOADS[args.workload]
    npdt = np.float64 if args.dtype == "f64" else np.float32
    loc, state, _ = multi_gpu._share_locals(args, rank, world, nx, npdt)
    dt = mb.cfl_dt(1.0e7 / nx)
    backend = mb.B200(local)
    rt = multi_gpu.TorchRuntime(local, device="cpu")
    comm = multi_gpu.Communicator(backend, rt)
    model = multi_gpu.DecomposedModel(loc, state, backend, local, dtype=npdt, overlap=not args.no_overlap, graph=not args.no_graph,
                                      runtime=rt, halo=args.halo, comm=comm)
    model.step(dt, 10)
    model.finish()
    comm.barrier()
    cap = 1 << 20
    L.check(L.lib().mokab_trace_begin(backend.handle, cap))
    backend.timer_start()
    model.step(dt, args.steps)
    ms = backend.timer_stop()
    model.finish()
    buf = np.zeros(cap, REC)
    n = C.c_int64()
    L.check(L.lib().mokab_trace_read(backend.handle, buf.ctypes.data_as(C.c_void_p), cap, C.byref(n)))
    L.check(L.lib().mokab_trace_begin(backend.handle, 0))
    rec = buf[:min(n.value, cap)]
    if rank == 0:
        la = [x for x in launches(rec) if x["complete"]]
        per_step = max(1, len(la) // args.steps)
        t_ref = la[-per_step * args.show]["t0"] if len(la) >= per_step * args.show else la[0]["t0"]
        print(f"# {args.workload} over {world} GPUs, halo={args.halo}, {'graphs' if not args.no_graph else 'host-launched'}, blocks interior/boundary "
              f"{model.mesh.block_counts()}: {ms / args.steps * 1e3:.1f} us per step, {ms / args.steps / 4 * 1e3:.1f} us per stage; {n.value} records, "
              f"{len(la)} launches, {per_step} per step")
        print(f"# {'launch':24s} {'blocks':>6s} {'start':>8s} {'end':>8s} {'dur':>7s} {'waited':>7s}   (us; start relative to the first launch shown)")
        prev_end = {}
        for x in la[-per_step * args.show:]:
            waited = (x["gate"] - x["t0"]) / 1e3 if x["gate"] else 0.0
            print(f"  {label(x['kind']):24s} {x['grid']:6d} {(x['t0'] - t_ref) / 1e3:8.1f} {(x['t1'] - t_ref) / 1e3:8.1f} {(x['t1'] - x['t0']) / 1e3:7.1f} {waited:7.1f}")
        med = {}
        for x in la:
            med.setdefault(label(x["kind"]), []).append(((x["t1"] - x["t0"]) / 1e3, (x["gate"] - x["t0"]) / 1e3 if x["gate"] else 0.0))
        summary = {k: {"launches": len(v), "median_duration_us": float(np.median([a for a, _ in v])), "median_wait_us": float(np.median([b for _, b in v]))}
                   for k, v in med.items()}
        print(json.dumps({"workload": args.workload, "n_gpus": world, "halo": args.halo, "graph": not args.no_graph, "us_per_stage": ms / args.steps / 4 * 1e3,
                          "rank0_blocks_interior_boundary": list(model.mesh.block_counts()), "kinds": summary}))
    model.close()
    comm.destroy()
    dist.barrier()
    torch.cuda.synchronize()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""One-GPU companion of diag_graph.py: P ranks emulated in one process (messages moved with device copies instead of
NCCL), the staged schedule (boundary -> pack -> interior -> copy -> unpack) launched on one stream, eagerly and as a
captured 2-step CUDA graph.  If the graph replay differs from the eager run here, the fault is in this library's staged
entry points under capture; if it is identical, the N=8 discrepancy involves NCCL inside the graph.

  python tools/diag_graph_emulated.py [--nx 96] [--parts 8] [--steps 20]"""
import argparse
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "mpas-ocean.jl_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]


def main():
    import torch

    import moka_b200 as mb
    from moka_b200 import _lib as L
    from moka_b200 import partition
    from test_gpu_decomposed import _Rank
    ap = argparse.ArgumentParser()
    ap.add_argument("--nx", type=int, default=96)
    ap.add_argument("--parts", type=int, default=8)
    ap.add_argument("--steps", type=int, default=20)
    args = ap.parse_args()
    backend = mb.B200(0)
    lib = L.lib()
    m = mb.periodic_hex(args.nx, args.nx, 1.0e7 / args.nx, with_dual=False)
    state = mb.inertialGravityWave(m).initial_state()
    dt = mb.cfl_dt(m["dc"])
    locs = partition.decompose(m, args.parts)
    stream = torch.cuda.Stream()
    sp = C.c_void_p(stream.cuda_stream)

    def make():
        ranks = [_Rank(backend, loc, state, args.parts) for loc in locs]
        backend.synchronize()       # `stream` below is not ordered with the context's stream the uploads ran on
        torch.cuda.synchronize()
        return ranks

    def enqueue(ranks, nsteps):
        with torch.cuda.stream(stream):
            for _ in range(nsteps):
                for s in (1, 2, 3, 4):
                    for r in ranks:
                        L.check(lib.mokab_rk4_stage(r.h, dt, s, L.PART_BOUNDARY, sp))
                        L.check(lib.mokab_halo_pack(r.h, s, C.c_void_p(r.send.data_ptr()), sp))
                        L.check(lib.mokab_rk4_stage(r.h, dt, s, L.PART_INTERIOR, sp))
                    for r in ranks:
                        ro = 0
                        for q, cr in enumerate(r.rcnt):
                            if cr:
                                so = sum(ranks[q].scnt[:r.loc["rank"]])
                                r.recv.t[ro:ro + cr].copy_(ranks[q].send.t[so:so + cr])
                            ro += cr
                    for r in ranks:
                        L.check(lib.mokab_halo_unpack(r.h, s, C.c_void_p(r.recv.data_ptr()), sp))
                for r in ranks:
                    L.check(lib.mokab_rk4_finish_step(r.h))

    def result(ranks):
        stream.synchronize()
        backend.synchronize()
        return [(r.prog.normalVelocity.copy(), r.prog.layerThickness.copy()) for r in ranks]

    eager = make()
    enqueue(eager, args.steps)
    want = result(eager)
    graphed = make()
    enqueue(graphed, 2)
    stream.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=stream):
        enqueue(graphed, 2)
    with torch.cuda.stream(stream):
        for _ in range((args.steps - 2) // 2):
            g.replay()
    got = result(graphed)
    same = all(np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) for a, b in zip(want, got))
    err = max(float(np.linalg.norm(a[0] - b[0]) / np.linalg.norm(a[0])) for a, b in zip(want, got))
    blocks = [r.mesh.block_counts() for r in eager]
    print(f"emulated nx={args.nx} parts={args.parts} steps={args.steps} blocks(interior,boundary) per rank={blocks}: "
          f"graph replay {'identical to' if same else 'DIFFERENT from'} the eager schedule (max rel-L2 u {err:.2e})", flush=True)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Diagnostic for the open graph-schedule issue; run under torchrun on N GPUs:

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 \
         tools/diag_graph.py --nx 96 192 384 --steps 20

For every mesh size it steps the same initial state with (a) the stream-launched overlapped schedule, (b) the
stream-launched serial schedule, (c) the captured 2-step graph of the overlapped schedule, (d) the captured graph of the
serial schedule, and prints on rank 0 whether (b)-(d) reproduce (a) bit for bit on every rank, plus the block counts
(interior / boundary) per rank.  Nothing here is used by the product path; `DecomposedModel.validate_graph` is what keeps
wrong graphs out of results."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "mpas-ocean.jl_b200")]


def main():
    import torch
    import torch.distributed as dist

    import moka_b200 as mb
    from moka_b200 import multi_gpu, partition
    ap = argparse.ArgumentParser()
    ap.add_argument("--nx", type=int, nargs="+", default=[96, 192, 384])
    ap.add_argument("--steps", type=int, default=20)
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    backend = mb.B200(local)
    for nx in args.nx:
        m = mb.periodic_hex(nx, nx, 1.0e7 / nx, with_dual=False)
        state = mb.inertialGravityWave(m).initial_state()
        dt = mb.cfl_dt(m["dc"])
        loc = partition.decompose(m, world)[rank]
        results, blocks = {}, None
        for name, overlap, graph in (("stream/overlap", True, False), ("stream/serial", False, False),
                                     ("graph/overlap", True, True), ("graph/serial", False, True)):
            model = multi_gpu.DecomposedModel(loc, multi_gpu.local_state(loc, *state), backend, local, overlap=overlap, graph=False)
            blocks = model.mesh.block_counts()
            if graph:                                   # bypass validate_graph: this tool wants to SEE the raw graph result
                model._build_graph(dt)                  # two warm-up steps + capture
                with torch.cuda.stream(model.compute):
                    for _ in range((args.steps - 2) // 2):
                        model._graph.replay()
            else:
                model._enqueue_steps(dt, args.steps)
            model.finish()
            results[name] = (model.prog.normalVelocity.copy(), model.prog.layerThickness.copy())
            model.close()
            del model
        ref = results["stream/overlap"]
        line = []
        for name, (u, h) in results.items():
            same = bool(np.array_equal(u, ref[0]) and np.array_equal(h, ref[1]))
            err = float(np.linalg.norm(u - ref[0]) / np.linalg.norm(ref[0]))
            t = torch.tensor([1.0 if same else 0.0, err], dtype=torch.float64, device=f"cuda:{local}")
            mn, mx = t.clone(), t.clone()
            dist.all_reduce(mn, op=dist.ReduceOp.MIN)
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            line.append(f"{name}: {'identical' if mn[0].item() == 1.0 else 'DIFFERENT'} (max rel-L2 u {mx[1].item():.2e})")
        cnt = torch.tensor([float(blocks[0]), float(blocks[1])], dtype=torch.float64, device=f"cuda:{local}")
        lo = cnt.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        if rank == 0:
            print(f"nx={nx} world={world} steps={args.steps} min interior/boundary blocks per rank = {int(lo[0].item())}/{int(lo[1].item())}: "
                  + "; ".join(line), flush=True)
    dist.barrier()
    torch.cuda.synchronize()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

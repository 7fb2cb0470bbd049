#!/usr/bin/env python
"""torchrun entry: throughput of ForwardEuler (the reference's live stepper) on a decomposed mesh, one process per GPU.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29620 \\
         tools/bench_fe_decomposed.py [--workload igw4096] [--steps 100] [--warmup 5] [--no-overlap]

One JSON line from rank 0: cell-steps/s over all ranks (device time, max over ranks) and the per-rank halo bytes per step.
Each step is one fused ForwardEuler launch per block part + two packed exchanges ((h, u) and (ssh, layerThicknessEdge))."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mpas-ocean.jl_b200"))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist

    import moka_b200 as mb
    from bench import WORKLOADS
    from moka_b200 import multi_gpu
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="igw4096", choices=sorted(w for w in WORKLOADS if w.startswith(("igw", "kelvin"))))
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--no-overlap", dest="no_overlap", action="store_true")
    ap.add_argument("--no-graph", dest="no_graph", action="store_true")
    args = ap.parse_args()
    args.dtype = "f64"
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")                              # control plane only; the exchange is NCCL inside the library
    nx = WORKLOADS[args.workload]
    loc, state, t_setup = multi_gpu._share_locals(args, rank, world, nx, np.float64)
    dt = 0.2 * mb.cfl_dt(1.0e7 / nx)
    backend = mb.B200(local)
    model = multi_gpu.DecomposedModel(loc, state, backend, local, overlap=not args.no_overlap, graph=not args.no_graph,
                                      runtime=multi_gpu.TorchRuntime(local, device="cpu"))
    model.step(dt, max(args.warmup, 3), stepper=mb.ForwardEuler)
    model.finish()
    model.comm.barrier()
    torch.cuda.synchronize()
    backend.timer_start()
    model.step(dt, args.steps, stepper=mb.ForwardEuler)
    ms = float(model.comm.allreduce(backend.timer_stop(), "max")[0])
    model.finish()
    mass = model.reduce("mass")
    nsend = sum(len(v) for v in loc["halo"]["send"].values())
    if rank == 0:
        print(json.dumps({"metric": "ForwardEuler cell-steps/sec", "value": nx * nx * args.steps / (ms * 1e-3), "unit": "cell-steps/s",
                          "n_gpus": world, "steps": args.steps, "ms_per_step": ms / args.steps, "workload": args.workload,
                          "overlap": model.overlap, "graph": model.use_graph, "rank0_halo_bytes_per_step": 2 * 8 * nsend, "setup_s": round(t_setup, 1),
                          "mass": mass}))
    model.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

#!/bin/bash
# 8-GPU box pass: the N=8 bench line of the default workload (igw4096, strong scaling) and the torchrun parity check.
set -u
tag=${1:-r01o}
out=gpurun_out
mkdir -p $out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512"
timeout 400 $TR bench.py --gpus 8 --steps 60 --warmup 4 > $out/bench_n8_$tag.json 2> $out/bench_n8_$tag.err; echo "bench n8 rc=$?"; tail -n 1 $out/bench_n8_$tag.json; tail -n 3 $out/bench_n8_$tag.err
timeout 200 $TR tests/multi_gpu_check.py > $out/mgcheck_n8_$tag.log 2>&1; echo "mgcheck rc=$?"; tail -n 2 $out/mgcheck_n8_$tag.log

#!/bin/bash
set -u
tag=${1:-r01e}
out=gpurun_out
mkdir -p $out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR tests/multi_gpu_check.py > $out/mgcheck_$tag.log 2>&1; echo "mgcheck rc=$?"; tail -n 3 $out/mgcheck_$tag.log
timeout 400 $TR bench.py --gpus 2 --steps 40 --warmup 4 > $out/bench_n2_$tag.json 2> $out/bench_n2_$tag.err; echo "bench n2 rc=$?"; tail -n 2 $out/bench_n2_$tag.json; tail -n 3 $out/bench_n2_$tag.err
timeout 300 python -m pytest tests/test_gpu_decomposed.py -x -q > $out/pytest_n2_$tag.log 2>&1; echo "pytest rc=$?"; tail -n 3 $out/pytest_n2_$tag.log

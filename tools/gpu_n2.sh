#!/bin/bash
# 2-GPU box pass: full GPU test suite (incl. the real-NCCL 2-rank test), the torchrun parity check, the N=2 bench line.
set -u
tag=${1:-r01k}
out=gpurun_out
mkdir -p $out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 python -m pytest tests -m gpu -x -q > $out/pytest_n2_$tag.log 2>&1; echo "pytest rc=$?"; tail -n 12 $out/pytest_n2_$tag.log
python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke_$tag.log 2>&1; echo "smoke rc=$?"; tail -n 2 $out/smoke_$tag.log
timeout 300 $TR tests/multi_gpu_check.py > $out/mgcheck_$tag.log 2>&1; echo "mgcheck rc=$?"; tail -n 3 $out/mgcheck_$tag.log
timeout 400 $TR bench.py --gpus 2 --steps 40 --warmup 4 > $out/bench_n2_$tag.json 2> $out/bench_n2_$tag.err; echo "bench n2 rc=$?"; tail -n 1 $out/bench_n2_$tag.json; tail -n 3 $out/bench_n2_$tag.err
timeout 300 $TR bench.py --gpus 2 --steps 100 --warmup 4 --workload kelvin1024 > $out/bench_n2_kelvin_$tag.json 2>> $out/bench_n2_$tag.err; echo "bench n2 kelvin rc=$?"; tail -n 1 $out/bench_n2_kelvin_$tag.json

#!/bin/bash
# GPU-box pass for the reverse mode: parity tests, throughput, ncu launch list + one full capture of the adjoint stage.
set -u
tag=${1:-r01c}
out=gpurun_out
mkdir -p $out
python -m pytest tests/test_gpu_adjoint.py -m gpu -x -q > $out/pytest_adj_$tag.log 2>&1; echo "pytest adjoint rc=$?"; tail -n 25 $out/pytest_adj_$tag.log
python tools/bench_adjoint.py --workload igw2048 --steps 10 > $out/adjoint_${tag}_igw2048_f64.json 2> $out/adjoint_$tag.err; echo "rc=$?"; cat $out/adjoint_${tag}_igw2048_f64.json; tail -n 5 $out/adjoint_$tag.err
python tools/bench_adjoint.py --workload igw2048 --steps 10 --dtype f32 > $out/adjoint_${tag}_igw2048_f32.json 2>> $out/adjoint_$tag.err; cat $out/adjoint_${tag}_igw2048_f32.json
ncu --set full --clock-control none --import-source on -k regex:k_rk_stage_adj -s 8 -c 4 -f -o $out/prof_adj_$tag \
    python tools/bench_adjoint.py --workload igw2048 --steps 3 > $out/ncu_adj_$tag.log 2>&1; echo "ncu full rc=$?"

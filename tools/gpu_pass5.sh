#!/bin/bash
set -u
tag=${1:-r01g}
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -x -q > $out/pytest_$tag.log 2>&1; echo "pytest rc=$?"; tail -n 15 $out/pytest_$tag.log
python bench.py --workload igw2048 --no-cpu > $out/bench_${tag}_igw2048.json 2> $out/bench_$tag.err; cat $out/bench_${tag}_igw2048.json
python bench.py --workload igw2048 --no-cpu --derive-eoe > $out/bench_${tag}_igw2048_derive.json 2>> $out/bench_$tag.err; cat $out/bench_${tag}_igw2048_derive.json
python bench.py --workload igw2048 --no-cpu --dtype f32 > $out/bench_${tag}_igw2048_f32.json 2>> $out/bench_$tag.err; cat $out/bench_${tag}_igw2048_f32.json
python bench.py --workload igw2048 --no-cpu --dtype f32 --derive-eoe > $out/bench_${tag}_igw2048_f32_derive.json 2>> $out/bench_$tag.err; cat $out/bench_${tag}_igw2048_f32_derive.json
tail -n 5 $out/bench_$tag.err

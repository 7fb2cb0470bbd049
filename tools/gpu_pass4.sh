#!/bin/bash
set -u
tag=${1:-r01f}
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -x -q > $out/pytest_$tag.log 2>&1; echo "pytest rc=$?"; tail -n 25 $out/pytest_$tag.log
for wl in igw2048; do
  python bench.py --workload $wl --no-cpu > $out/bench_${tag}_${wl}_derived.json 2> $out/bench_$tag.err; cat $out/bench_${tag}_${wl}_derived.json
  python bench.py --workload $wl --no-cpu --explicit-eoe > $out/bench_${tag}_${wl}_explicit.json 2>> $out/bench_$tag.err; cat $out/bench_${tag}_${wl}_explicit.json
  python bench.py --workload $wl --no-cpu --dtype f32 > $out/bench_${tag}_${wl}_f32_derived.json 2>> $out/bench_$tag.err; cat $out/bench_${tag}_${wl}_f32_derived.json
done
python bench.py --workload kelvin1024 --no-cpu > $out/bench_${tag}_kelvin.json 2>> $out/bench_$tag.err; cat $out/bench_${tag}_kelvin.json
tail -n 5 $out/bench_$tag.err
ncu --set full --clock-control none --import-source on -k regex:k_rk_stage -s 12 -c 4 -f -o $out/prof_$tag \
    python bench.py --workload igw2048 --steps 3 --warmup 3 --no-cpu --quick > $out/ncu_full_$tag.log 2>&1; echo "ncu full rc=$?"

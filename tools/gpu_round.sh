#!/bin/bash
# One GPU-box pass: parity tests, bench lines, ncu launch list and full capture of the stage kernel.
# usage (through gpurun): bash tools/gpu_round.sh <tag>
set -u
tag=${1:-r01}
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -x -q > $out/pytest_$tag.log 2>&1; echo "pytest rc=$?"; tail -n 3 $out/pytest_$tag.log
python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke_$tag.log 2>&1; echo "smoke rc=$?"; tail -n 2 $out/smoke_$tag.log
python bench.py > $out/bench_$tag.json 2> $out/bench_$tag.err; echo "bench rc=$?"; cat $out/bench_$tag.json
python bench.py --workload igw2048 --no-cpu > $out/bench_${tag}_igw2048.json 2>> $out/bench_$tag.err; cat $out/bench_${tag}_igw2048.json
python bench.py --workload igw2048 --dtype f32 --no-cpu > $out/bench_${tag}_igw2048_f32.json 2>> $out/bench_$tag.err; cat $out/bench_${tag}_igw2048_f32.json
python bench.py --workload kelvin1024 --no-cpu > $out/bench_${tag}_kelvin1024.json 2>> $out/bench_$tag.err; cat $out/bench_${tag}_kelvin1024.json
python bench.py --explicit-eoe --no-cpu > $out/bench_${tag}_explicit_eoe.json 2>> $out/bench_$tag.err; cat $out/bench_${tag}_explicit_eoe.json
python bench.py --impl reference --steps 5 --warmup 1 > $out/bench_${tag}_reference.json 2>> $out/bench_$tag.err; cat $out/bench_${tag}_reference.json
# launch list of the default bench command (short), then one full capture of the four stage launches of a step
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $out/launches_$tag.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu --quick > $out/ncu_launches_$tag.log 2>&1; echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_rk_stage -s 12 -c 4 -f -o $out/prof_$tag \
    python bench.py --steps 3 --warmup 3 --no-cpu --quick > $out/ncu_full_$tag.log 2>&1; echo "ncu full rc=$?"
ls -la $out | tail -n 12

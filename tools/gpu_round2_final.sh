#!/bin/bash
# Round 2, final GPU pass (one GPU): the GPU suite, smoke(), the default bench line, the ncu launch list of the same command and
# ncu --set full summaries of the stage kernel as the defaults now run it (cp.async kernel at 4096x4096 / 2048x2048, the plain kernel
# chosen by "stage_auto" at 512x512), condensed on the box with tools/ncu_summary.py.
set -u
tag=${1:-r02o}
out=gpurun_out
mkdir -p $out
timeout 600 python -m pytest tests -m gpu -x -q > $out/pytest_$tag.log 2>&1; echo "pytest rc=$?"; tail -n 3 $out/pytest_$tag.log
python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke_$tag.log 2>&1; echo "smoke rc=$?"; tail -n 1 $out/smoke_$tag.log
python bench.py > $out/bench_igw4096_f64_$tag.json 2> $out/bench_$tag.err; echo "bench rc=$?"; cut -c1-260 $out/bench_igw4096_f64_$tag.json
cap() {   # cap <name> <kernel regex> <skip> <count> <command...>: one ncu --set full capture -> summary csv, report deleted
    name=$1; rx=$2; skip=$3; cnt=$4; shift 4
    ncu --set full --clock-control none --import-source on -k regex:"$rx" -s $skip -c $cnt -o /tmp/ncu_$name -f "$@" > $out/ncu_${name}_$tag.log 2>&1
    python tools/ncu_summary.py /tmp/ncu_$name.ncu-rep > $out/ncu_${name}_summary_$tag.csv 2>> $out/bench_$tag.err; echo "ncu $name rc=$? rows=$(wc -l < $out/ncu_${name}_summary_$tag.csv)"
    rm -f /tmp/ncu_$name.ncu-rep; tail -n 2 $out/ncu_${name}_$tag.log > $out/ncu_${name}_$tag.tail; rm -f $out/ncu_${name}_$tag.log
}
cap stage_f64_igw2048 k_rk_stage 12 4 python bench.py --workload igw2048 --steps 3 --warmup 3 --no-cpu --no-parity --quick
cap stage_f64_igw512_plain k_rk_stage 12 4 python bench.py --workload igw512 --steps 3 --warmup 3 --no-cpu --no-parity --quick
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $out/launches_igw4096_f64_$tag.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu --no-parity --quick > /dev/null 2>&1; echo "ncu launch list rc=$?"
python bench.py --workload igw2048 --dtype f32 --no-cpu > $out/bench_igw2048_f32_$tag.json 2>> $out/bench_$tag.err; cut -c1-200 $out/bench_igw2048_f32_$tag.json
tail -n 5 $out/bench_$tag.err
ls -la $out | tail -n 14

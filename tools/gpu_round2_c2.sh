#!/bin/bash
# Round 2, GPU pass C2 (one GPU): bench lines of the final default kernels + ncu captures of every hot kernel, condensed ON THE
# BOX with tools/ncu_summary.py (the .ncu-rep files are ~20 MB each and gpurun_out/ only carries 64 MiB back).
set -u
tag=${1:-r02f}
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q > $out/pytest_$tag.log 2>&1; echo "pytest rc=$?"; tail -n 3 $out/pytest_$tag.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke_$tag.log 2>&1; echo "smoke rc=$?"; tail -n 1 $out/smoke_$tag.log
python bench.py > $out/bench_$tag.json 2> $out/bench_$tag.err; echo "bench rc=$?"; cut -c1-260 $out/bench_$tag.json
python bench.py --workload igw2048 --no-cpu > $out/bench_igw2048_f64_$tag.json 2>> $out/bench_$tag.err; cut -c1-260 $out/bench_igw2048_f64_$tag.json
python bench.py --workload igw2048 --dtype f32 --no-cpu > $out/bench_igw2048_f32_$tag.json 2>> $out/bench_$tag.err; cut -c1-260 $out/bench_igw2048_f32_$tag.json
python bench.py --workload igw512 --no-cpu > $out/bench_igw512_f64_$tag.json 2>> $out/bench_$tag.err; cut -c1-260 $out/bench_igw512_f64_$tag.json
python bench.py --workload kelvin1024 --no-cpu > $out/bench_kelvin1024_f64_$tag.json 2>> $out/bench_$tag.err; cut -c1-260 $out/bench_kelvin1024_f64_$tag.json
python bench.py --workload voronoi1024 --no-cpu > $out/bench_voronoi1024_$tag.json 2>> $out/bench_$tag.err; cut -c1-260 $out/bench_voronoi1024_$tag.json
python bench.py --workload sphere1024 --no-cpu > $out/bench_sphere1024_$tag.json 2>> $out/bench_$tag.err; cut -c1-260 $out/bench_sphere1024_$tag.json
python bench.py --impl reference --steps 5 --warmup 1 > $out/bench_reference_$tag.json 2>> $out/bench_$tag.err; cut -c1-260 $out/bench_reference_$tag.json
python tools/bench_fe.py --workload igw4096 --no-dual > $out/fe_igw4096_$tag.json 2>> $out/bench_$tag.err
python tools/bench_adjoint.py > $out/adjoint_rk4_f64_$tag.json 2>> $out/bench_$tag.err
python tools/bench_adjoint.py --dtype f32 > $out/adjoint_rk4_f32_$tag.json 2>> $out/bench_$tag.err
python tools/bench_adjoint.py --stepper fe > $out/adjoint_fe_$tag.json 2>> $out/bench_$tag.err
python tools/bench_multilevel.py > $out/multilevel_$tag.jsonl 2>> $out/bench_$tag.err; cut -c1-200 $out/multilevel_$tag.jsonl
timeout 300 python tools/stage_sweep.py --workload igw2048 --variants 3:0:0,0:0:0,0:0:3,1:0:3 > $out/sweep_$tag.jsonl 2>> $out/bench_$tag.err
cap() {   # cap <name> <kernel regex> <skip> <count> <command...>: one ncu --set full capture -> summary csv, report deleted
    name=$1; rx=$2; skip=$3; cnt=$4; shift 4
    ncu --set full --clock-control none --import-source on -k regex:"$rx" -s $skip -c $cnt -o /tmp/ncu_$name -f "$@" > $out/ncu_${name}_$tag.log 2>&1
    python tools/ncu_summary.py /tmp/ncu_$name.ncu-rep > $out/ncu_${name}_summary_$tag.csv 2>> $out/bench_$tag.err; echo "ncu $name rc=$? rows=$(wc -l < $out/ncu_${name}_summary_$tag.csv)"
    rm -f /tmp/ncu_$name.ncu-rep; tail -n 2 $out/ncu_${name}_$tag.log > $out/ncu_${name}_$tag.tail; rm -f $out/ncu_${name}_$tag.log
}
cap stage_f64_igw4096 k_rk_stage 12 4 python bench.py --steps 3 --warmup 3 --no-cpu --no-parity --quick
cap stage_f64_igw2048 k_rk_stage 12 4 python bench.py --workload igw2048 --steps 3 --warmup 3 --no-cpu --no-parity --quick
cap stage_f32_igw2048 k_rk_stage 12 4 python bench.py --workload igw2048 --dtype f32 --steps 3 --warmup 3 --no-cpu --no-parity --quick
cap stage_voronoi1024 k_rk_stage 12 4 python bench.py --workload voronoi1024 --steps 3 --warmup 3 --no-cpu --no-parity --quick
cap fe_step k_fe_step 6 2 python tools/bench_fe.py --workload igw2048 --no-dual --steps 5
cap adj_rk4 "k_rk_stage_adj" 8 4 python tools/bench_adjoint.py --steps 3
cap adj_fe k_fe_step_adj 2 2 python tools/bench_adjoint.py --stepper fe --steps 3
cap stage_ml k_rk_stage_ml 8 4 python tools/bench_multilevel.py --levels 10 --steps 3
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $out/launches_igw4096_f64_$tag.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu --no-parity --quick > /dev/null 2>&1; echo "ncu launch list rc=$?"
tail -n 5 $out/bench_$tag.err
du -sh $out; ls $out | wc -l

#!/bin/bash
# Round 2, GPU pass C1 (one GPU): occupancy-side variants of the stage kernel on the slot-major numbering (weights through
# per-thread cp.async, 9 blocks of 128 threads, block-major weights), the multi-level kernel, and the evidence round 1 owed:
# bench lines + ncu captures of the ForwardEuler step, the adjoints and the (12, 7) kernels (Voronoi / sphere workloads).
set -u
tag=${1:-r02d}
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q > $out/pytest_$tag.log 2>&1; echo "pytest rc=$?"; tail -n 3 $out/pytest_$tag.log
for lib in libmoka_b200.so libmoka_b200_cpa67.so libmoka_b200_bc128.so libmoka_b200_bc128_o9.so; do
    MOKAB_LIB=$lib timeout 300 python tools/stage_sweep.py --workload igw2048 --variants 0:0:0,0:0:3,1:0:3 > $out/sweep_${lib%.so}_$tag.jsonl 2>> $out/sweep_$tag.err
done
MOKAB_LIB=libmoka_b200_wfb.so MOKAB_STAGE_WF_BLOCK_MAJOR=1 timeout 300 python tools/stage_sweep.py --workload igw2048 --variants 0:0:0 > $out/sweep_wfb_$tag.jsonl 2>> $out/sweep_$tag.err
timeout 300 python tools/stage_sweep.py --workload igw4096 --variants 0:0:0,0:0:3 --dtypes f64 --steps 20 > $out/sweep_igw4096_$tag.jsonl 2>> $out/sweep_$tag.err
python - $out/sweep_*_$tag.jsonl <<'PY'
import json, sys
for f in sys.argv[1:]:
    for line in open(f):
        d = json.loads(line)
        if "best" in d:
            continue
        if "error" in d:
            print(f, d); continue
        print(f"{f.split('/')[-1]:44s} {d['workload']:8s} {d['dtype']} pf={d['prefetch']} tma={d['tma']} "
              f"{d['cell_steps_per_s'] / 1e9:7.3f} G  frac {d['roofline_frac']:.3f}  same={d['bit_identical_to_default']}")
PY
timeout 600 python tools/bench_multilevel.py > $out/multilevel_$tag.jsonl 2> $out/multilevel_$tag.err; echo "multilevel rc=$?"; cut -c1-330 $out/multilevel_$tag.jsonl
# ---- evidence: the other kernels -----------------------------------------------------------------------------------------
python tools/bench_fe.py --workload igw4096 --no-dual > $out/fe_igw4096_$tag.json 2> $out/evid_$tag.err; cut -c1-400 $out/fe_igw4096_$tag.json
python tools/bench_adjoint.py > $out/adjoint_rk4_f64_$tag.json 2>> $out/evid_$tag.err; cut -c1-300 $out/adjoint_rk4_f64_$tag.json
python tools/bench_adjoint.py --dtype f32 > $out/adjoint_rk4_f32_$tag.json 2>> $out/evid_$tag.err; cut -c1-300 $out/adjoint_rk4_f32_$tag.json
python tools/bench_adjoint.py --stepper fe > $out/adjoint_fe_$tag.json 2>> $out/evid_$tag.err; cut -c1-300 $out/adjoint_fe_$tag.json
python bench.py --workload voronoi1024 --no-cpu > $out/bench_voronoi1024_$tag.json 2>> $out/evid_$tag.err; cut -c1-300 $out/bench_voronoi1024_$tag.json
python bench.py --workload sphere1024 --no-cpu > $out/bench_sphere1024_$tag.json 2>> $out/evid_$tag.err; cut -c1-300 $out/bench_sphere1024_$tag.json
ncu --set full --clock-control none --import-source on -k regex:k_fe_step -s 6 -c 2 -o $out/ncu_fe_step_$tag -f \
    python tools/bench_fe.py --workload igw2048 --no-dual --steps 5 > $out/ncu_fe_$tag.log 2>&1; echo "ncu fe rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"k_rk_stage_adj|k_fe_step_adj" -s 8 -c 4 -o $out/ncu_adj_rk4_$tag -f \
    python tools/bench_adjoint.py --steps 3 > $out/ncu_adj_$tag.log 2>&1; echo "ncu adj rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_fe_step_adj -s 2 -c 2 -o $out/ncu_adj_fe_$tag -f \
    python tools/bench_adjoint.py --stepper fe --steps 3 > $out/ncu_adjfe_$tag.log 2>&1; echo "ncu adj fe rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_rk_stage -s 12 -c 4 -o $out/ncu_stage_voronoi_$tag -f \
    python bench.py --workload voronoi1024 --steps 3 --warmup 3 --no-cpu --no-parity --quick > $out/ncu_vor_$tag.log 2>&1; echo "ncu voronoi rc=$?"
MOKAB_STAGE_TMA=3 ncu --set full --clock-control none --import-source on -k regex:k_rk_stage -s 12 -c 4 -o $out/ncu_stage_cpa_f64_$tag -f \
    python bench.py --workload igw2048 --steps 3 --warmup 3 --no-cpu --no-parity --quick > $out/ncu_cpa_$tag.log 2>&1; echo "ncu cpa f64 rc=$?"
MOKAB_STAGE_TMA=3 ncu --set full --clock-control none --import-source on -k regex:k_rk_stage -s 12 -c 4 -o $out/ncu_stage_cpa_f32_$tag -f \
    python bench.py --workload igw2048 --dtype f32 --steps 3 --warmup 3 --no-cpu --no-parity --quick > $out/ncu_cpa32_$tag.log 2>&1; echo "ncu cpa f32 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_rk_stage_ml -s 8 -c 4 -o $out/ncu_stage_ml_$tag -f \
    python tools/bench_multilevel.py --levels 10 --steps 3 > $out/ncu_ml_$tag.log 2>&1; echo "ncu ml rc=$?"
tail -n 5 $out/evid_$tag.err
ls -la $out | tail -n 24

#!/bin/bash
# Round 2, GPU pass B1 (one GPU): A/B of the slot-major edge numbering (mesh.cuh) against the cell-by-cell one, both block
# sizes, both precisions; bench + ncu of the new default.
set -u
tag=${1:-r02b}
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q > $out/pytest_$tag.log 2>&1; echo "pytest rc=$?"; tail -n 3 $out/pytest_$tag.log
for lib in libmoka_b200.so libmoka_b200_bc128.so; do
    MOKAB_LIB=$lib timeout 600 python tools/stage_sweep.py --workload igw2048 --variants 0:0:0,1:0:0 --edge-orders slot_major,by_cell > $out/sweep_${lib%.so}_$tag.jsonl 2>> $out/sweep_$tag.err
done
MOKAB_LIB=libmoka_b200.so timeout 600 python tools/stage_sweep.py --workload igw2048 --variants 0:0:0 --explicit-eoe --edge-orders slot_major,by_cell > $out/sweep_explicit_$tag.jsonl 2>> $out/sweep_$tag.err
timeout 600 python tools/stage_sweep.py --workload kelvin1024 --variants 0:0:0 --edge-orders slot_major,by_cell --dtypes f64 > $out/sweep_kelvin_$tag.jsonl 2>> $out/sweep_$tag.err
python - $out/sweep_*_$tag.jsonl <<'PY'
import json, sys
for f in sys.argv[1:]:
    for line in open(f):
        d = json.loads(line)
        if "best" in d or "error" in d:
            continue
        print(f"{d['lib']:24s} {d['workload']:10s} {d['dtype']} {d['edge_order']:10s} pf={d['prefetch']} expl={int(d['explicit_eoe'])} "
              f"{d['cell_steps_per_s'] / 1e9:7.3f} G  frac {d['roofline_frac']:.3f}  same={d['bit_identical_to_default']}")
PY
python bench.py > $out/bench_$tag.json 2> $out/bench_$tag.err; echo "bench rc=$?"; cut -c1-300 $out/bench_$tag.json
python bench.py --workload igw2048 --dtype f32 --no-cpu > $out/bench_igw2048_f32_$tag.json 2>> $out/bench_$tag.err; cut -c1-300 $out/bench_igw2048_f32_$tag.json
ncu --set full --clock-control none --import-source on -k regex:k_rk_stage -s 12 -c 4 -o $out/ncu_stage_f64_$tag -f \
    python bench.py --workload igw2048 --steps 3 --warmup 3 --no-cpu --no-parity --quick > $out/ncu_full_f64_$tag.log 2>&1; echo "ncu full f64 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_rk_stage -s 12 -c 4 -o $out/ncu_stage_f32_$tag -f \
    python bench.py --workload igw2048 --dtype f32 --steps 3 --warmup 3 --no-cpu --no-parity --quick > $out/ncu_full_f32_$tag.log 2>&1; echo "ncu full f32 rc=$?"
ls -la $out | tail -n 12

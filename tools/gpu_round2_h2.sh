#!/bin/bash
# (As run for r02i at commit 342f12a: coherent state loads have been the default build since; libmoka_b200_ldg.so is the read-only variant now.)
# Round 2, GPU pass H2 (one GPU): (1) the per-launch choice between the two tuned stage kernels by the wave-quantisation model
# ("stage_auto") against either kernel forced, on every BASELINE mesh size; (2) the build that reads the state through plain
# coherent loads (libmoka_b200_coh.so) against the default, and programmatic dependent launch on it ("stage_pdl"; r02h showed the
# read-only path is not safe under it); (3) the GPU suite on both builds; (4) bench lines with the new defaults.
set -u
tag=${1:-r02i}
out=gpurun_out
mkdir -p $out
V=1:0:3:0:0:0,0:0:0:0:0:0,1:0:3:0:0:1
for w in igw512 kelvin1024 igw2048; do
    timeout 400 python tools/stage_sweep.py --workload $w --variants $V --steps 400 > $out/sweep_auto_${w}_$tag.jsonl 2>> $out/sweep_$tag.err
done
timeout 400 python tools/stage_sweep.py --workload igw4096 --variants $V --dtypes f64 --steps 100 > $out/sweep_auto_igw4096_$tag.jsonl 2>> $out/sweep_$tag.err
VC=1:0:3:0:0:0,1:0:3:0:1:0,0:0:0:0:0:0,0:0:0:0:1:0
for w in igw512 kelvin1024 igw2048; do
    MOKAB_LIB=libmoka_b200_coh.so timeout 400 python tools/stage_sweep.py --workload $w --variants $VC --steps 400 > $out/sweep_coh_${w}_$tag.jsonl 2>> $out/sweep_$tag.err
done
python - $out/sweep_*_$tag.jsonl <<'PY'
import json, sys
for f in sys.argv[1:]:
    for line in open(f):
        d = json.loads(line)
        if "best" in d:
            continue
        if "error" in d:
            print(f, d); continue
        print(f"{f.split('/')[-1]:40s} {d['dtype']} pf={d['prefetch']} tma={d['tma']} pdl={d['pdl']} auto={d['auto']} "
              f"{d['cell_steps_per_s'] / 1e9:7.3f} G  frac {d['roofline_frac']:.3f}  same={d['bit_identical_to_default']}")
PY
MOKAB_LIB=libmoka_b200_coh.so MOKAB_STAGE_PDL=1 timeout 900 python -m pytest tests -m gpu -x -q > $out/pytest_coh_pdl_$tag.log 2>&1; echo "pytest(coh+pdl) rc=$?"; tail -n 3 $out/pytest_coh_pdl_$tag.log
timeout 900 python -m pytest tests -m gpu -x -q > $out/pytest_$tag.log 2>&1; echo "pytest rc=$?"; tail -n 3 $out/pytest_$tag.log
for w in igw512 kelvin1024 igw2048; do
    python bench.py --workload $w --no-cpu > $out/bench_${w}_f64_$tag.json 2>> $out/bench_$tag.err; cut -c1-200 $out/bench_${w}_f64_$tag.json
done
python bench.py --workload igw2048 --dtype f32 --no-cpu > $out/bench_igw2048_f32_$tag.json 2>> $out/bench_$tag.err; cut -c1-200 $out/bench_igw2048_f32_$tag.json
python bench.py > $out/bench_igw4096_f64_$tag.json 2>> $out/bench_$tag.err; cut -c1-200 $out/bench_igw4096_f64_$tag.json
tail -n 5 $out/sweep_$tag.err $out/bench_$tag.err

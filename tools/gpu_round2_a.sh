#!/bin/bash
# Round 2, GPU pass A (one GPU): the whole gpu suite (incl. the direct-store halo tests and the 1024x1024 Kelvin tests, first
# hardware run), smoke, the in-process A/B of the stage-kernel variants (L2 prefetch, bulk-copy variants, 128-cell blocks),
# bench lines with the default and with the best variant, ncu launch list + full capture of the winner.
# usage (through gpurun): bash tools/gpu_round2_a.sh <tag>
set -u
tag=${1:-r02a}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $out/gpu_$tag.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > $out/pytest_$tag.log 2>&1; echo "pytest rc=$?"; tail -n 4 $out/pytest_$tag.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke_$tag.log 2>&1; echo "smoke rc=$?"; tail -n 2 $out/smoke_$tag.log
# ---- A/B of the stage kernel variants, one process per library -------------------------------------------------------------
timeout 600 python tools/stage_sweep.py --workload igw2048 --variants 0:0:0,1:0:0,2:0:0,3:0:0,2:888:0,3:888:0,2:1184:0,3:1184:0,2:2368:0,3:2368:0 > $out/sweep_$tag.jsonl 2> $out/sweep_$tag.err; echo "sweep rc=$?"
# the bulk-copy (TMA) variants have never run on hardware: their own process, under a short timeout
timeout 180 python tools/stage_sweep.py --workload igw2048 --variants 0:0:0,0:0:1,1:0:1,0:0:2,1:0:2,3:0:2 > $out/sweep_tma_$tag.jsonl 2>> $out/sweep_$tag.err; echo "sweep tma rc=$?"
timeout 600 python tools/stage_sweep.py --workload igw2048 --explicit-eoe --dtypes f64 --variants 0:0:0,1:0:0,2:0:0,3:0:0 > $out/sweep_explicit_$tag.jsonl 2>> $out/sweep_$tag.err
if [ -f mpas-ocean.jl_b200/libmoka_b200_bc128.so ]; then
    MOKAB_LIB=libmoka_b200_bc128.so timeout 600 python tools/stage_sweep.py --workload igw2048 --variants 0:0:0,1:0:0,2:0:0,3:0:0,3:2368:0 > $out/sweep_bc128_$tag.jsonl 2>> $out/sweep_$tag.err
fi
python - $out/sweep_$tag.jsonl $out/sweep_tma_$tag.jsonl $out/sweep_explicit_$tag.jsonl $out/sweep_bc128_$tag.jsonl <<'PY'
import json, sys
for f in sys.argv[1:]:
    try:
        for line in open(f):
            d = json.loads(line)
            if "best" in d:
                print("BEST", {k: (v["prefetch"], v["distance"], v["tma"], round(v["cell_steps_per_s"] / 1e9, 3)) for k, v in d["best"].items()})
            elif "error" in d:
                print("ERR ", d)
            else:
                print(f"{d['lib']:24s} {d['dtype']} pf={d['prefetch']} dist={d['distance']:5d} tma={d['tma']} expl={int(d['explicit_eoe'])} "
                      f"{d['cell_steps_per_s'] / 1e9:7.3f} G  frac {d['roofline_frac']:.3f}  same={d['bit_identical_to_default']}")
    except Exception as e:
        print(f, "unreadable:", e)
PY
# the best run-time variant per precision -> environment of the bench runs below
eval $(python - $out/sweep_$tag.jsonl $out/sweep_tma_$tag.jsonl <<'PY'
import json, sys
best = {}
for f in sys.argv[1:]:
    try:
        for line in open(f):
            d = json.loads(line)
            for k, v in (d.get("best") or {}).items():
                if k not in best or v["cell_steps_per_s"] > best[k]["cell_steps_per_s"]:
                    best[k] = v
    except Exception:
        pass
for k in ("f64", "f32"):
    b = best.get(k, {"prefetch": 0, "distance": 0, "tma": 0})
    print(f"export BEST_{k.upper()}='MOKAB_STAGE_PREFETCH={b['prefetch']} MOKAB_STAGE_PREFETCH_DISTANCE={b['distance']} MOKAB_STAGE_TMA={b['tma']}'")
PY
)
echo "best f64: $BEST_F64 ; best f32: $BEST_F32"
# ---- bench lines ---------------------------------------------------------------------------------------------------------
python bench.py > $out/bench_default_$tag.json 2> $out/bench_$tag.err; echo "bench (default variant) rc=$?"; cut -c1-400 $out/bench_default_$tag.json
env $BEST_F64 python bench.py --no-cpu > $out/bench_best_$tag.json 2>> $out/bench_$tag.err; echo "bench (best variant) rc=$?"; cut -c1-400 $out/bench_best_$tag.json
env $BEST_F32 python bench.py --workload igw2048 --dtype f32 --no-cpu > $out/bench_igw2048_f32_best_$tag.json 2>> $out/bench_$tag.err; cut -c1-400 $out/bench_igw2048_f32_best_$tag.json
env $BEST_F64 python bench.py --workload kelvin1024 --no-cpu > $out/bench_kelvin1024_best_$tag.json 2>> $out/bench_$tag.err; cut -c1-300 $out/bench_kelvin1024_best_$tag.json
python bench.py --impl reference --steps 5 --warmup 1 > $out/bench_reference_$tag.json 2>> $out/bench_$tag.err; cut -c1-300 $out/bench_reference_$tag.json
# ---- ncu: launch list of the bench command, full capture of the stage kernel (f64 best, f64 default, f32 best) ----------------
env $BEST_F64 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $out/launches_best_$tag.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu --no-parity --quick > $out/ncu_launches_$tag.log 2>&1; echo "ncu launch list rc=$?"
env $BEST_F64 ncu --set full --clock-control none --import-source on -k regex:k_rk_stage -s 12 -c 4 -o $out/ncu_stage_best_$tag -f \
    python bench.py --workload igw2048 --steps 3 --warmup 3 --no-cpu --no-parity --quick > $out/ncu_full_best_$tag.log 2>&1; echo "ncu full (best f64) rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_rk_stage -s 12 -c 4 -o $out/ncu_stage_default_$tag -f \
    python bench.py --workload igw2048 --steps 3 --warmup 3 --no-cpu --no-parity --quick > $out/ncu_full_default_$tag.log 2>&1; echo "ncu full (default f64) rc=$?"
env $BEST_F32 ncu --set full --clock-control none --import-source on -k regex:k_rk_stage -s 12 -c 4 -o $out/ncu_stage_f32_best_$tag -f \
    python bench.py --workload igw2048 --dtype f32 --steps 3 --warmup 3 --no-cpu --no-parity --quick > $out/ncu_full_f32_$tag.log 2>&1; echo "ncu full (best f32) rc=$?"
ls -la $out | tail -n 25

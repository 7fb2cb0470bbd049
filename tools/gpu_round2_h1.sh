#!/bin/bash
# (As run for r02h at commit ff81503: the shared-memory flux variant it measures was removed afterwards -- 9 % slower.)
# Round 2, GPU pass H1 (one GPU): A/B of the two latency-side changes written after r02g -- the thickness flux of a block's own
# edges through shared memory ("stage_flux_smem") and programmatic dependent launch of the stage kernels ("stage_pdl") -- as
# 40-step bursts and as >= 0.5 s sustained batches (the live bench is power-capped), on 2048x2048 in both precisions and on the
# small meshes where the launch-to-launch gap is a visible share of a stage; then the whole GPU suite with both switches on.
set -u
tag=${1:-r02h}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $out/gpu_$tag.txt 2>&1
V=1:0:3,1:0:3:1,1:0:3:0:1,1:0:3:1:1,0:0:0,0:0:0:0:1
timeout 400 python tools/stage_sweep.py --workload igw2048 --variants $V > $out/sweep_burst_igw2048_$tag.jsonl 2>> $out/sweep_$tag.err
timeout 400 python tools/stage_sweep.py --workload igw2048 --variants $V --steps 400 > $out/sweep_sustained_igw2048_$tag.jsonl 2>> $out/sweep_$tag.err
timeout 200 python tools/stage_sweep.py --workload igw512 --variants $V --dtypes f64 --steps 400 > $out/sweep_igw512_$tag.jsonl 2>> $out/sweep_$tag.err
timeout 200 python tools/stage_sweep.py --workload kelvin1024 --variants $V --dtypes f64 --steps 400 > $out/sweep_kelvin1024_$tag.jsonl 2>> $out/sweep_$tag.err
timeout 400 python tools/stage_sweep.py --workload igw4096 --variants 1:0:3,1:0:3:1,1:0:3:1:1 --dtypes f64 --steps 100 > $out/sweep_sustained_igw4096_$tag.jsonl 2>> $out/sweep_$tag.err
python - $out/sweep_*_$tag.jsonl <<'PY'
import json, sys
for f in sys.argv[1:]:
    for line in open(f):
        d = json.loads(line)
        if "best" in d:
            continue
        if "error" in d:
            print(f, d); continue
        print(f"{f.split('/')[-1]:44s} {d['dtype']} pf={d['prefetch']} tma={d['tma']} fx={d['flux_smem']} pdl={d['pdl']} "
              f"{d['cell_steps_per_s'] / 1e9:7.3f} G  frac {d['roofline_frac']:.3f}  same={d['bit_identical_to_default']}")
PY
MOKAB_STAGE_FLUX_SMEM=1 MOKAB_STAGE_PDL=1 timeout 900 python -m pytest tests -m gpu -x -q > $out/pytest_fx_pdl_$tag.log 2>&1; echo "pytest(fx+pdl) rc=$?"; tail -n 3 $out/pytest_fx_pdl_$tag.log
timeout 900 python -m pytest tests -m gpu -x -q > $out/pytest_$tag.log 2>&1; echo "pytest rc=$?"; tail -n 3 $out/pytest_$tag.log
tail -n 5 $out/sweep_$tag.err

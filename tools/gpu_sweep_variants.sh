#!/bin/bash
# A/B of the kernel build variants on one GPU (build them HERE first: `make -C mpas-ocean.jl_b200 variants` -- the .so files
# travel with the snapshot): the default bench workload and the 2048^2 roofline configurations for each library.
# usage (through gpurun): bash tools/gpu_sweep_variants.sh <tag>
set -u
tag=${1:-sweep}
out=gpurun_out
mkdir -p $out
# the TMA variant of the stage kernel (weight rows by bulk asynchronous copies) is a run-time switch of the default library:
# check it first (bit-identical results expected), under a timeout -- it has never run on hardware
MOKAB_STAGE_TMA=1 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "config1_f64 or derived_edges or fused_f32 or variable_coriolis or full_size" > $out/pytest_tma_$tag.log 2>&1
echo "pytest with MOKAB_STAGE_TMA=1 rc=$?"; tail -n 2 $out/pytest_tma_$tag.log
for args in "--workload igw2048" "--workload igw2048 --dtype f32" "--workload igw2048 --explicit-eoe" ""; do
    name=$(echo "$args" | tr -d ' -' ); name=${name:-igw4096}
    MOKAB_STAGE_TMA=1 timeout 600 python bench.py $args --no-cpu > $out/sweep_${tag}_tma_${name}.json 2>> $out/sweep_$tag.err
    python - "$out/sweep_${tag}_tma_${name}.json" "MOKAB_STAGE_TMA=1 $args" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(f"{sys.argv[2]:60s} {d['value'] / 1e9:7.3f} G cell-steps/s   {d['ms_per_step']:7.3f} ms/step   roofline {d['roofline']['frac']:.3f}")
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
done
for lib in libmoka_b200.so libmoka_b200_bc128.so libmoka_b200_mb8.so; do
    [ -f mpas-ocean.jl_b200/$lib ] || { echo "$lib not built"; continue; }
    v=${lib%.so}; v=${v#libmoka_b200}; v=${v:-_default}
    for args in "--workload igw2048" "--workload igw2048 --dtype f32" "--workload igw2048 --explicit-eoe" "--workload kelvin1024" ""; do
        name=$(echo "$args" | tr -d ' -' ); name=${name:-igw4096}
        MOKAB_LIB=$lib python bench.py $args --no-cpu > $out/sweep_${tag}${v}_${name}.json 2>> $out/sweep_$tag.err
        python - "$out/sweep_${tag}${v}_${name}.json" "$lib $args" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(f"{sys.argv[2]:60s} {d['value'] / 1e9:7.3f} G cell-steps/s   {d['ms_per_step']:7.3f} ms/step   roofline {d['roofline']['frac']:.3f}")
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
    done
done

#!/bin/bash
# A/B of the kernel variants on one GPU: the build variants (make them HERE first: `make -C mpas-ocean.jl_b200 variants` -- the .so
# files travel with the snapshot; selected with MOKAB_LIB) crossed with the TMA variants of the stage kernel (a run-time switch:
# MOKAB_STAGE_TMA=1 stages the slot-major weight rows, ten bulk copies per block; =2 a block-major copy of the weights, one
# bulk copy per block), over the roofline configurations.  The TMA variants have never run on hardware: each is checked for
# bit-identical results first, under a timeout.
# usage (through gpurun): bash tools/gpu_sweep_variants.sh <tag>
set -u
tag=${1:-sweep}
out=gpurun_out
mkdir -p $out
declare -A tma_rc; tma_rc[0]=0
for tma in 1 2; do
    MOKAB_STAGE_TMA=$tma timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "config1_f64 or derived_edges or fused_f32 or variable_coriolis or full_size" > $out/pytest_tma${tma}_$tag.log 2>&1
    tma_rc[$tma]=$?; echo "pytest with MOKAB_STAGE_TMA=$tma rc=${tma_rc[$tma]}"; tail -n 2 $out/pytest_tma${tma}_$tag.log
done
for lib in libmoka_b200.so libmoka_b200_bc128.so libmoka_b200_mb8.so; do
    [ -f mpas-ocean.jl_b200/$lib ] || { echo "$lib not built"; continue; }
    v=${lib%.so}; v=${v#libmoka_b200}; v=${v:-_default}
    for tma in 0 1 2; do
        [ "${tma_rc[$tma]}" != 0 ] && continue                        # this TMA variant failed its check: do not time it
        [ "$tma" != 0 ] && [ "$lib" = libmoka_b200_mb8.so ] && continue
        for args in "--workload igw2048" "--workload igw2048 --dtype f32" "--workload igw2048 --explicit-eoe" "--workload kelvin1024" ""; do
            name=$(echo "$args" | tr -d ' -' ); name=${name:-igw4096}
            f=$out/sweep_${tag}${v}_tma${tma}_${name}.json
            MOKAB_LIB=$lib MOKAB_STAGE_TMA=$tma timeout 900 python bench.py $args --no-cpu > $f 2>> $out/sweep_$tag.err
            python - "$f" "$lib tma=$tma $args" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(f"{sys.argv[2]:64s} {d['value'] / 1e9:7.3f} G cell-steps/s   {d['ms_per_step']:7.3f} ms/step   roofline {d['roofline']['frac']:.3f}")
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
        done
    done
done

#!/bin/bash
set -u
tag=${1:-r01h}
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -x -q > $out/pytest_$tag.log 2>&1; echo "pytest rc=$?"; tail -n 15 $out/pytest_$tag.log
run() { # lib, extra flags, label
  MOKAB_LIB=$1 python bench.py --workload igw2048 --no-cpu --steps 60 $2 > $out/bench_${tag}_$3.json 2>> $out/bench_$tag.err
  python - <<PY
import json
d=json.loads(open("$out/bench_${tag}_$3.json").read().strip().splitlines()[-1])
print("$3", "value %.4g"%d["value"], "ms/step %.4f"%d["ms_per_step"], "e2e %.4g"%d["e2e"]["value"], "clk", d["clocks"]["sm_mhz"], d["config"]["blocks_rebuilding_edgesOnEdge"])
PY
}
run libmoka_b200.so "" v1_f64
run libmoka_b200.so "--derive-eoe" der5_f64
run libmoka_b200_der4.so "--derive-eoe" der4_f64
run libmoka_b200_der6.so "--derive-eoe" der6_f64
run libmoka_b200.so "--dtype f32" v1_f32
run libmoka_b200.so "--dtype f32 --derive-eoe" der5_f32
run libmoka_b200_der4.so "--dtype f32 --derive-eoe" der4_f32
run libmoka_b200_der6.so "--dtype f32 --derive-eoe" der6_f32
tail -n 5 $out/bench_$tag.err

#!/bin/bash
set -u
tag=${1:-r01j}
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
run libmoka_b200.so "" base_f64
run libmoka_b200_pf.so "" pf_f64
run libmoka_b200.so "--dtype f32" base_f32
run libmoka_b200_pf.so "--dtype f32" pf_f32
run libmoka_b200.so "--explicit-eoe" base_explicit_f64
run libmoka_b200_pf.so "--explicit-eoe" pf_explicit_f64
run libmoka_b200.so "" base_f64_again
run libmoka_b200_pf.so "" pf_f64_again
tail -n 5 $out/bench_$tag.err

#!/bin/bash
set -u
tag=${1:-r01d}
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -x -q > $out/pytest_$tag.log 2>&1; echo "pytest rc=$?"; tail -n 25 $out/pytest_$tag.log
for lib in libmoka_b200.so libmoka_b200_adj5.so; do
  echo "== $lib"
  MOKAB_LIB=$lib python tools/bench_adjoint.py --workload igw2048 --steps 10 > $out/adjoint_${tag}_${lib%.so}_f64.json 2>> $out/adjoint_$tag.err; cat $out/adjoint_${tag}_${lib%.so}_f64.json
  MOKAB_LIB=$lib python tools/bench_adjoint.py --workload igw2048 --steps 10 --dtype f32 > $out/adjoint_${tag}_${lib%.so}_f32.json 2>> $out/adjoint_$tag.err; cat $out/adjoint_${tag}_${lib%.so}_f32.json
done
tail -n 5 $out/adjoint_$tag.err

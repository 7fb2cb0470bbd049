#!/bin/bash
# Round 2, GPU pass H4 (gpurun --gpus 2): tests/multi_gpu_check.py with the reverse mode on the decomposed mesh (both steppers) on
# real NCCL / CUDA IPC, and the two tuned stage kernels forced at the per-GPU grid sizes of the multi-GPU runs (where the block-count
# rule of "stage_auto" was extrapolated from single-GPU data).
set -u
tag=${1:-r02k}
out=gpurun_out
mkdir -p $out
n=$(nvidia-smi -L | wc -l)
run="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1"
timeout 400 $run --master-port 29601 tests/multi_gpu_check.py > $out/mgcheck_$tag.log 2>&1; echo "mgcheck rc=$?"; tail -n 1 $out/mgcheck_$tag.log
show() { python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith("{")][-1])
    p = d.get("parity") or {}
    print(f"{sys.argv[2]:44s} {d['value'] / 1e9:7.3f} G  {d['ms_per_step']:.4f} ms/step  blocks={d['config'].get('rank0_blocks_interior_boundary')} parity={p.get('ok')} sm={d['clocks']['sm_mhz']}")
except Exception as ex:
    print(sys.argv[2], "FAILED", ex)
PY
}
i=0
one() {   # one <label> <env...> -- <bench args...>
    label=$1; shift; envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
    i=$((i+1)); f=$out/h4_${i}_$tag.json
    env "${envs[@]}" timeout 600 $run --master-port $((29620+i)) bench.py --gpus $n "$@" > $f 2>> $out/bench_$tag.err; show $f "$label"
}
one "igw2048 cp.async forced"   MOKAB_STAGE_AUTO=0 -- --workload igw2048 --steps 50 --warmup 5 --no-parity
one "igw2048 plain forced"      MOKAB_STAGE_TMA=0 MOKAB_STAGE_PREFETCH=0 -- --workload igw2048 --steps 50 --warmup 5 --no-parity
one "igw2048 defaults (rule)"   X=1 -- --workload igw2048 --steps 50 --warmup 5
one "kelvin1024 cp.async forced" MOKAB_STAGE_AUTO=0 -- --workload kelvin1024 --steps 100 --warmup 5 --no-parity
one "kelvin1024 plain forced"   MOKAB_STAGE_TMA=0 MOKAB_STAGE_PREFETCH=0 -- --workload kelvin1024 --steps 100 --warmup 5 --no-parity
one "kelvin1024 defaults (rule)" X=1 -- --workload kelvin1024 --steps 100 --warmup 5
one "igw512 defaults (rule)"    X=1 -- --workload igw512 --steps 200 --warmup 5
one "igw1024 cp.async forced"   MOKAB_STAGE_AUTO=0 -- --workload igw1024 --steps 100 --warmup 5 --no-parity
one "igw1024 plain forced"      MOKAB_STAGE_TMA=0 MOKAB_STAGE_PREFETCH=0 -- --workload igw1024 --steps 100 --warmup 5 --no-parity
grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" $out/bench_$tag.err | tail -n 10

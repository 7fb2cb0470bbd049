#!/bin/bash
# Round 2, GPU pass H3 (gpurun --gpus 2): (1) tests/multi_gpu_check.py on real NCCL / CUDA IPC with the reverse mode on the
# decomposed mesh added, and again with programmatic dependent launch of the stage kernels; (2) what the per-launch kernel choice
# ("stage_auto") and "stage_pdl" give at the per-GPU sizes of the 8-GPU runs: 2048x2048 over 2 = 2.1 M cells per GPU (igw4096 over 8),
# 512x512 over 2 = 131 k cells per GPU (the 1024x1024 channel over 8), for the push / wait kernels and for the exchange folded into
# the boundary launch; (3) one default igw4096 line for the halves of the end-to-end leg.
set -u
tag=${1:-r02j}
out=gpurun_out
mkdir -p $out
n=$(nvidia-smi -L | wc -l)
run="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1"
timeout 400 $run --master-port 29601 tests/multi_gpu_check.py > $out/mgcheck_$tag.log 2>&1; echo "mgcheck rc=$?"; tail -n 1 $out/mgcheck_$tag.log
MOKAB_STAGE_PDL=1 timeout 400 $run --master-port 29602 tests/multi_gpu_check.py > $out/mgcheck_pdl_$tag.log 2>&1; echo "mgcheck(pdl) rc=$?"; tail -n 1 $out/mgcheck_pdl_$tag.log
show() { python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith("{")][-1])
    p = d.get("parity") or {}
    e = d["e2e"]
    print(f"{sys.argv[2]:44s} {d['value'] / 1e9:7.3f} G  {d['ms_per_step']:.4f} ms/step  e2e {e['value'] / 1e9:.3f} G ({e['ms_per_step']:.3f} ms; copies only "
          f"{e.get('copies_only_ms_per_step', 0):.3f}, steps only {e.get('one_step_calls_only_ms_per_step', 0):.3f})  parity={p.get('ok')}")
except Exception as ex:
    print(sys.argv[2], "FAILED", ex)
PY
}
i=0
one() {   # one <label> <env...> -- <bench args...>
    label=$1; shift; envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
    i=$((i+1)); f=$out/h3_${i}_$tag.json
    env "${envs[@]}" timeout 600 $run --master-port $((29620+i)) bench.py --gpus $n "$@" > $f 2>> $out/bench_$tag.err; show $f "$label"
}
one "igw2048 p2p (defaults: auto)"          X=1 -- --workload igw2048 --steps 50 --warmup 5
one "igw2048 p2p auto off"                  MOKAB_STAGE_AUTO=0 -- --workload igw2048 --steps 50 --warmup 5 --no-parity
one "igw2048 p2p pdl"                       MOKAB_STAGE_PDL=1 -- --workload igw2048 --steps 50 --warmup 5
one "igw512 p2p (defaults)"                 X=1 -- --workload igw512 --steps 200 --warmup 5
one "igw512 p2p auto off"                   MOKAB_STAGE_AUTO=0 -- --workload igw512 --steps 200 --warmup 5 --no-parity
one "igw512 p2p pdl"                        MOKAB_STAGE_PDL=1 -- --workload igw512 --steps 200 --warmup 5 --no-parity
one "igw512 p2p_fused"                      X=1 -- --workload igw512 --steps 200 --warmup 5 --no-parity --halo p2p_fused
one "igw512 p2p_fused pdl"                  MOKAB_STAGE_PDL=1 -- --workload igw512 --steps 200 --warmup 5 --halo p2p_fused
one "igw512 nccl"                           X=1 -- --workload igw512 --steps 200 --warmup 5 --no-parity --halo nccl
one "kelvin1024 p2p (defaults)"             X=1 -- --workload kelvin1024 --steps 100 --warmup 5
one "kelvin1024 p2p pdl"                    MOKAB_STAGE_PDL=1 -- --workload kelvin1024 --steps 100 --warmup 5 --no-parity
one "igw4096 p2p (defaults)"                X=1 -- --steps 20 --warmup 5
grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" $out/bench_$tag.err | tail -n 10

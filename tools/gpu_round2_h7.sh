#!/bin/bash
# Round 2, GPU pass H7 (gpurun --gpus 2): the decomposed graphs instantiated with cudaGraphInstantiateFlagUseNodePriority (r02m: the
# priority attribute of the halo stream's launches alone changed nothing) -- timelines and bench lines.
set -u
tag=${1:-r02n}
out=gpurun_out
mkdir -p $out
n=$(nvidia-smi -L | wc -l)
run="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1"
i=0
one() { i=$((i+1)); MOKAB_LIB=libmoka_b200_trace.so timeout 300 $run --master-port $((29640+i)) tools/trace_stages.py "$@" > $out/trace_${i}_$tag.txt 2>> $out/trace_$tag.err; echo "== $*"; grep -v "^{" $out/trace_${i}_$tag.txt | head -n 24; }
one --workload igw2048 --halo p2p --show 1
one --workload kelvin1024 --halo p2p --show 1
show() { python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith("{")][-1])
    p = d.get("parity") or {}
    print(f"{sys.argv[2]:44s} {d['value'] / 1e9:7.3f} G  {d['ms_per_step']:.4f} ms/step  blocks={d['config'].get('rank0_blocks_interior_boundary')} parity={p.get('ok')} graphs={d['config']['detail'][-60:]}")
except Exception as ex:
    print(sys.argv[2], "FAILED", ex)
PY
}
j=0
bench() {   # bench <label> <env...> -- <bench args...>
    label=$1; shift; envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
    j=$((j+1)); f=$out/h7_${j}_$tag.json
    env "${envs[@]}" timeout 600 $run --master-port $((29660+j)) bench.py --gpus $n "$@" > $f 2>> $out/bench_$tag.err; show $f "$label"
}
bench "igw2048 p2p"                    X=1 -- --workload igw2048 --steps 50 --warmup 5
bench "igw2048 p2p no launch priority" MOKAB_LAUNCH_PRIORITY=0 -- --workload igw2048 --steps 50 --warmup 5 --no-parity
bench "kelvin1024 p2p"                 X=1 -- --workload kelvin1024 --steps 100 --warmup 5
bench "igw512 p2p"                     X=1 -- --workload igw512 --steps 200 --warmup 5
grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" $out/bench_$tag.err $out/trace_$tag.err | tail -n 10

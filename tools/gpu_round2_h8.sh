#!/bin/bash
# Round 2, GPU pass H8 (gpurun --gpus 2, the last GPU minutes of the round): the flag-in-data halo exchange (MOKAB_HALO_P2P_LL) on real
# NVLink -- the torchrun check restricted to it, then its stage time on the 131 k-cell parts next to the push / wait kernels.
set -u
tag=${1:-r02p}
out=gpurun_out
mkdir -p $out
n=$(nvidia-smi -L | wc -l)
run="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1"
MOKAB_P2P_TIMEOUT_S=5 MOKAB_CHECK_HALO=p2p_ll timeout 120 $run --master-port 29601 tests/multi_gpu_check.py > $out/mgcheck_ll_$tag.log 2>&1; echo "mgcheck(p2p_ll) rc=$?"; tail -n 2 $out/mgcheck_ll_$tag.log | cut -c1-600
show() { python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith("{")][-1])
    p = d.get("parity") or {}
    print(f"{sys.argv[2]:28s} {d['value'] / 1e9:7.3f} G  {d['ms_per_step']:.4f} ms/step  blocks={d['config'].get('rank0_blocks_interior_boundary')} parity={p.get('ok')} bit_identical={p.get('bit_identical')}")
except Exception as ex:
    print(sys.argv[2], "FAILED", ex)
PY
}
j=0
bench() { label=$1; shift; j=$((j+1)); f=$out/h8_${j}_$tag.json; MOKAB_P2P_TIMEOUT_S=5 timeout 120 $run --master-port $((29660+j)) bench.py --gpus $n "$@" > $f 2>> $out/bench_$tag.err; show $f "$label"; }
bench "igw512 p2p_ll"     --workload igw512 --steps 200 --warmup 5 --halo p2p_ll
bench "igw512 p2p"        --workload igw512 --steps 200 --warmup 5 --no-parity
MOKAB_LIB=libmoka_b200_trace.so MOKAB_P2P_TIMEOUT_S=5 timeout 100 $run --master-port 29650 tools/trace_stages.py --workload igw512 --halo p2p_ll --show 1 > $out/trace_ll_$tag.txt 2>> $out/bench_$tag.err; grep -v "^{" $out/trace_ll_$tag.txt | head -n 22
bench "kelvin1024 p2p_ll" --workload kelvin1024 --steps 100 --warmup 5 --halo p2p_ll
bench "igw2048 p2p_ll"    --workload igw2048 --steps 50 --warmup 5 --halo p2p_ll --no-parity
grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" $out/bench_$tag.err | tail -n 8

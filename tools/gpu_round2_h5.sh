#!/bin/bash
# Round 2, GPU pass H5 (gpurun --gpus 2): stage timelines of the decomposed schedule from the TRACE build (tools/trace_stages.py):
# where the ~28 us per stage of a 131 k-cell part go, for every halo path; the same for the 2.1 M-cell part.
set -u
tag=${1:-r02l}
out=gpurun_out
mkdir -p $out
n=$(nvidia-smi -L | wc -l)
export MOKAB_LIB=libmoka_b200_trace.so
run="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1"
i=0
one() { i=$((i+1)); timeout 300 $run --master-port $((29640+i)) tools/trace_stages.py "$@" > $out/trace_${i}_$tag.txt 2>> $out/trace_$tag.err; echo "== $*"; grep -v "^{" $out/trace_${i}_$tag.txt | head -n 40; }
one --workload igw512 --halo p2p
one --workload igw512 --halo p2p_fused
one --workload igw512 --halo nccl
one --workload igw512 --halo p2p --no-graph
one --workload igw512 --halo p2p --no-overlap
one --workload kelvin1024 --halo p2p --show 1
one --workload igw2048 --halo p2p --show 1
grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" $out/trace_$tag.err | tail -n 10

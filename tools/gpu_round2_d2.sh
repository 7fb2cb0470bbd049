#!/bin/bash
# Round 2, GPU pass D2 (gpurun --gpus 2): where the ~25 us per stage of the decomposed schedule go (2.1 M cells per GPU, the
# per-GPU size of igw4096 over 8), and whether NUMA binding lifts the host-bound end-to-end leg.
set -u
tag=${1:-r02g}
out=gpurun_out
mkdir -p $out
n=$(nvidia-smi -L | wc -l)
run="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1"
nvidia-smi topo -m > $out/topo_$tag.txt 2>&1; numactl -H >> $out/topo_$tag.txt 2>&1; lscpu | head -n 25 >> $out/topo_$tag.txt 2>&1
show() { python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith("{")][-1])
    p = d.get("parity") or {}
    print(f"{sys.argv[2]:58s} {d['value'] / 1e9:7.3f} G  {d['ms_per_step']:.4f} ms/step  e2e {d['e2e']['value'] / 1e9:.3f} G ({d['e2e']['ms_per_step']:.3f} ms)  parity={p.get('ok')}  numa={d['config'].get('rank0_numa_node')}")
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
}
i=0
one() {   # one <label> <env...> -- <bench args...>
    label=$1; shift; envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
    i=$((i+1)); f=$out/d2_${i}_$tag.json
    env "${envs[@]}" timeout 600 $run --master-port $((29620+i)) bench.py --gpus $n "$@" > $f 2>> $out/bench_$tag.err; show $f "$label"
}
one "igw2048 p2p default"                       X=1 -- --workload igw2048 --steps 50 --warmup 5
one "igw2048 p2p plain kernel (tma=0 pf=0)"     MOKAB_STAGE_TMA=0 MOKAB_STAGE_PREFETCH=0 -- --workload igw2048 --steps 50 --warmup 5 --no-parity
one "igw2048 p2p no overlap"                    X=1 -- --workload igw2048 --steps 50 --warmup 5 --no-parity --no-overlap
one "igw2048 p2p_fused"                         X=1 -- --workload igw2048 --steps 50 --warmup 5 --no-parity --halo p2p_fused
one "kelvin1024 p2p"                            X=1 -- --workload kelvin1024 --steps 100 --warmup 5 --no-parity
one "kelvin1024 p2p_fused one launch per stage" MOKAB_DECOMP_SERIAL_BLOCKS=1000000 -- --workload kelvin1024 --steps 100 --warmup 5 --halo p2p_fused
one "igw512 p2p"                                X=1 -- --workload igw512 --steps 200 --warmup 5 --no-parity
one "igw512 p2p_fused one launch per stage"     X=1 -- --workload igw512 --steps 200 --warmup 5 --halo p2p_fused
one "igw4096 p2p default (numa bind)"           X=1 -- --steps 20 --warmup 5
one "igw4096 p2p no numa bind"                  MOKAB_NO_NUMA_BIND=1 -- --steps 20 --warmup 5 --no-parity
# which of the two default switches costs the extra DRAM traffic (r02f: +7 % over the must-move bytes, r02b: +3.8 %): one capture without the prefetch
MOKAB_STAGE_PREFETCH=0 ncu --set full --clock-control none -k regex:k_rk_stage -s 12 -c 4 -o /tmp/ncu_pf0 -f python bench.py --workload igw2048 --steps 3 --warmup 3 --no-cpu --no-parity --quick > /dev/null 2>&1
python tools/ncu_summary.py /tmp/ncu_pf0.ncu-rep > $out/ncu_stage_f64_igw2048_tma3_pf0_summary_$tag.csv; rm -f /tmp/ncu_pf0.ncu-rep
grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" $out/bench_$tag.err | tail -n 10
head -n 20 $out/topo_$tag.txt

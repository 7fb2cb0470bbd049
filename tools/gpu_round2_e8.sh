#!/bin/bash
# Round 2, 8-GPU pass (gpurun --gpus 8; charged 8x: keep it short): every halo path against the oracle at N = 8 on the 96x96
# mesh (every block a boundary block), then bench lines of the three halo paths on the scaling workload and on Kelvin 1024x1024.
set -u
tag=${1:-r02e}
out=gpurun_out
mkdir -p $out
n=$(nvidia-smi -L | wc -l)
run="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1"
timeout 300 $run --master-port 29611 tests/multi_gpu_check.py > $out/mgcheck_n${n}_$tag.log 2>&1; echo "mgcheck rc=$?"; grep -E "MULTI_GPU_CHECK|^\[\(\(" $out/mgcheck_n${n}_$tag.log | cut -c1-2500
show() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    p = d.get("parity") or {}
    print(f"{sys.argv[1]}: {d['value'] / 1e9:.3f} G cell-steps/s, {d['ms_per_step']:.4f} ms/step, e2e {d['e2e']['value'] / 1e9:.3f} G ({d['e2e']['ms_per_step']:.3f} ms), "
          f"roofline {d['roofline']['frac']:.3f}, parity ok={p.get('ok')} bit={p.get('bit_identical')}, setup {d['config']['setup_s']} s")
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
for halo in p2p nccl; do
    f=$out/bench_n${n}_kelvin1024_${halo}_$tag.json
    timeout 300 $run --master-port 29613 bench.py --gpus $n --workload kelvin1024 --steps 200 --warmup 5 --halo $halo > $f 2>> $out/bench_$tag.err; show $f
done
for halo in p2p nccl; do
    f=$out/bench_n${n}_igw4096_${halo}_$tag.json
    timeout 400 $run --master-port 29612 bench.py --gpus $n --steps 50 --warmup 5 --halo $halo > $f 2>> $out/bench_$tag.err; show $f
done
grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" $out/bench_$tag.err | tail -n 15
ls -la $out | tail -n 10

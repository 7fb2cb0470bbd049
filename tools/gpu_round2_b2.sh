#!/bin/bash
# Round 2, GPU pass B2 (gpurun --gpus 2 or more): the in-library decomposed path on real NCCL / real CUDA IPC -- every halo path
# and schedule against the single-domain oracle, the YAML -> NetCDF driver, then bench lines per halo path.
set -u
tag=${1:-r02c}
out=gpurun_out
mkdir -p $out
n=$(nvidia-smi -L | wc -l)
run="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1"
NCCL_DEBUG=WARN timeout 600 $run --master-port 29611 tests/multi_gpu_check.py > $out/mgcheck_n${n}_$tag.log 2>&1; echo "mgcheck rc=$?"; tail -n 4 $out/mgcheck_n${n}_$tag.log | cut -c1-1500
timeout 300 $run --master-port 29614 tests/multi_gpu_driver_check.py nccl > $out/mgdriver_n${n}_$tag.log 2>&1; echo "driver rc=$?"; tail -n 2 $out/mgdriver_n${n}_$tag.log
timeout 300 $run --master-port 29615 tests/multi_gpu_driver_check.py nccl ForwardEuler > $out/mgdriver_fe_n${n}_$tag.log 2>&1; echo "driver (ForwardEuler) rc=$?"; tail -n 2 $out/mgdriver_fe_n${n}_$tag.log
timeout 300 $run --master-port 29616 tests/multi_gpu_driver_check.py p2p_fused > $out/mgdriver_p2pf_n${n}_$tag.log 2>&1; echo "driver (p2p_fused) rc=$?"; tail -n 2 $out/mgdriver_p2pf_n${n}_$tag.log
show() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    p = d.get("parity") or {}
    print(f"{sys.argv[1]}: {d['value'] / 1e9:.3f} G cell-steps/s, {d['ms_per_step']:.4f} ms/step, e2e {d['e2e']['value'] / 1e9:.3f} G ({d['e2e']['ms_per_step']:.3f} ms), "
          f"roofline {d['roofline']['frac']:.3f}, parity ok={p.get('ok')} bit={p.get('bit_identical')}, launches {d['gpu_launches']}")
    print("   ", d["config"]["detail"])
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
for halo in nccl p2p p2p_fused; do
    f=$out/bench_n${n}_kelvin1024_${halo}_$tag.json
    timeout 600 $run --master-port 29613 bench.py --gpus $n --workload kelvin1024 --steps 100 --warmup 5 --halo $halo > $f 2>> $out/bench_$tag.err; show $f
    f=$out/bench_n${n}_igw2048_${halo}_$tag.json
    timeout 600 $run --master-port 29612 bench.py --gpus $n --workload igw2048 --steps 50 --warmup 5 --halo $halo > $f 2>> $out/bench_$tag.err; show $f
done
f=$out/bench_n${n}_igw4096_nccl_$tag.json
timeout 900 $run --master-port 29617 bench.py --gpus $n --steps 20 --warmup 5 > $f 2>> $out/bench_$tag.err; show $f
timeout 900 $run --master-port 29618 bench.py --gpus $n --impl reference --steps 3 --warmup 1 > $out/bench_n${n}_reference_$tag.json 2>> $out/bench_$tag.err; cut -c1-400 $out/bench_n${n}_reference_$tag.json
tail -n 20 $out/bench_$tag.err
ls -la $out | tail -n 14

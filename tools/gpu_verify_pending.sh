#!/bin/bash
# First GPU pass of a round that follows CPU-only work: everything that was written or fixed on the simulated runtime
# (tests/sim) and has not seen hardware yet.  One GPU unless noted.
# usage (through gpurun): bash tools/gpu_verify_pending.sh <tag>          (gpurun --gpus 2 or 8: adds the multi-rank checks)
set -u
tag=${1:-r02a}
out=gpurun_out
mkdir -p $out
ngpu=$(nvidia-smi -L | wc -l)
# 1. the whole suite (incl. the ForwardEuler adjoint tests), then the direct-store halo exchange tests (opt-in until verified)
timeout 1500 python -m pytest tests -m gpu -x -q > $out/pytest_$tag.log 2>&1; echo "pytest rc=$?"; tail -n 3 $out/pytest_$tag.log
MOKAB_TEST_P2P=1 timeout 600 python -m pytest tests -m gpu -q -k direct_store > $out/pytest_p2p_$tag.log 2>&1; echo "pytest p2p rc=$?"; tail -n 3 $out/pytest_p2p_$tag.log
# the TMA variant of the stage kernel (a run-time switch; bit-identical results expected); under its own timeout: first hardware run
for tma in 1 2; do MOKAB_STAGE_TMA=$tma timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "config1_f64 or derived_edges or fused_f32 or variable_coriolis" > $out/pytest_tma${tma}_$tag.log 2>&1; echo "pytest tma=$tma rc=$?"; tail -n 2 $out/pytest_tma${tma}_$tag.log; done
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke_$tag.log 2>&1; echo "smoke rc=$?"; tail -n 2 $out/smoke_$tag.log
# 2. the one-GPU reproduction of last round's "graph mismatch" (expected now: identical)
timeout 300 python tools/diag_graph_emulated.py > $out/diag_emulated_$tag.log 2>&1; echo "diag rc=$?"; cat $out/diag_emulated_$tag.log
# 3. bench lines (default workload; the reverse modes incl. the new ForwardEuler adjoint)
python bench.py > $out/bench_$tag.json 2> $out/bench_$tag.err; echo "bench rc=$?"; cat $out/bench_$tag.json
# unstructured mesh (pentagons / hexagons / heptagons): the compile-time (12, 7) kernels with and without the edgesOnEdge rebuild
python bench.py --workload voronoi1024 --no-cpu > $out/bench_voronoi1024_$tag.json 2>> $out/bench_$tag.err; cat $out/bench_voronoi1024_$tag.json
python bench.py --workload voronoi1024 --no-cpu --explicit-eoe > $out/bench_voronoi1024_explicit_$tag.json 2>> $out/bench_$tag.err; cat $out/bench_voronoi1024_explicit_$tag.json
# the sphere: quasi-uniform spherical Voronoi mesh, variable Coriolis parameter (folded (12, 7) kernels)
python bench.py --workload sphere1024 --no-cpu > $out/bench_sphere1024_$tag.json 2>> $out/bench_$tag.err; cat $out/bench_sphere1024_$tag.json
python tools/bench_adjoint.py --stepper fe > $out/adjoint_fe_$tag.json 2>> $out/bench_$tag.err; cat $out/adjoint_fe_$tag.json
python tools/bench_adjoint.py > $out/adjoint_rk4_$tag.json 2>> $out/bench_$tag.err; cat $out/adjoint_rk4_$tag.json
if [ "$ngpu" -ge 2 ]; then
    n=$ngpu
    run="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1"
    # 4. real NCCL + real IPC: all schedules of both halo paths against the single-domain oracle (96x96: at 8 ranks every block
    #    is a boundary block -- the decomposition that exposed the unordered initialisation)
    MOKAB_CHECK_P2P=1 MOKAB_CHECK_FE=1 timeout 600 $run --master-port 29611 tests/multi_gpu_check.py > $out/mgcheck_$tag.log 2>&1; echo "mgcheck rc=$?"; tail -n 3 $out/mgcheck_$tag.log
    timeout 300 $run --master-port 29614 tests/multi_gpu_driver_check.py nccl > $out/mgdriver_$tag.log 2>&1; echo "driver rc=$?"; tail -n 2 $out/mgdriver_$tag.log
    timeout 300 $run --master-port 29615 tests/multi_gpu_driver_check.py nccl ForwardEuler > $out/mgdriver_fe_$tag.log 2>&1; echo "driver (ForwardEuler) rc=$?"; tail -n 2 $out/mgdriver_fe_$tag.log
    timeout 600 $run --master-port 29616 tools/bench_fe_decomposed.py --steps 100 > $out/bench_fe_n${n}_$tag.json 2>> $out/bench_$tag.err; cat $out/bench_fe_n${n}_$tag.json
    for halo in nccl p2p p2p_fused; do
        timeout 900 $run --master-port 29612 bench.py --gpus $n --steps 50 --warmup 5 --halo $halo > $out/bench_n${n}_${halo}_$tag.json 2>> $out/bench_$tag.err
        cat $out/bench_n${n}_${halo}_$tag.json
        timeout 600 $run --master-port 29613 bench.py --gpus $n --workload kelvin1024 --steps 100 --warmup 5 --halo $halo > $out/bench_n${n}_kelvin_${halo}_$tag.json 2>> $out/bench_$tag.err
        cat $out/bench_n${n}_kelvin_${halo}_$tag.json
    done
fi
ls -la $out | tail -n 14

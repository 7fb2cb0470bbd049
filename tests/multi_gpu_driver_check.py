"""torchrun entry: the decomposed DRIVER path (driver.ocn_run_decomposed: YAML -> NetCDF mesh / initial state -> clock and
alarms -> RK4 over the ranks -> NetCDF output on rank 0) against the single-domain CPU oracle.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29614 tests/multi_gpu_driver_check.py [halo] [RK4|ForwardEuler]"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path[:0] = [os.path.join(ROOT, "mpas-ocean.jl_b200"), os.path.join(ROOT, "oracle")]
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import moka_b200 as mb  # noqa: E402
import moka_oracle_c as OC  # noqa: E402

YAML = """
omega:
  time_management:
    config_start_time: 0001-01-01_00:00:00
    config_stop_time: none
    config_run_duration: 0000-00-00_03:00:00
    config_restart_timestamp_name: Restart_timestamp
    config_do_restart: false
  time_integration:
    config_dt: 0000-00-00_00:15:00
    config_number_of_time_levels: 2
    config_time_integrator: {stepper}
  streams:
    mesh:
      filename_template: {mesh}
    input:
      filename_template: {mesh}
    output:
      filename_template: {out}
      reference_time: 0001-01-01_00:00:00
      output_interval: 0000-00-00_01:00:00
"""


def main():
    from scipy.io import netcdf_file
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    halo = sys.argv[1] if len(sys.argv) > 1 else "nccl"
    stepper = sys.argv[2] if len(sys.argv) > 2 else "RK4"
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")                      # control plane only; the exchange is NCCL inside the library
    tag = f"/dev/shm/mokab_driver_{os.environ.get('MASTER_PORT', '0')}"
    mesh_fp, out_fp, cfg = tag + "_mesh.nc", tag + "_out.nc", tag + "_cfg.yml"
    m = mb.periodic_hex(48, 48, 300.0e3)                 # dc = 300 km: the reference's dt rule gives 900 s
    state = mb.inertialGravityWave(m).initial_state()
    if rank == 0:
        mb.write_mesh_netcdf(mesh_fp, m, state)
        with open(cfg, "w") as f:
            f.write(YAML.format(mesh=mesh_fp, out=out_fp, stepper=stepper))
    dist.barrier()
    series = []
    from moka_b200 import multi_gpu
    Setup, model, nsteps = mb.driver.ocn_run_decomposed(cfg, mb.B200(local), local, halo=halo, series=series,
                                                        runtime=multi_gpu.TorchRuntime(local, device="cpu"))
    status = model.graph_status
    model.close()
    ok = True
    if rank == 0:
        OC.sign_index_fields(m)
        om = OC.OracleModel(m, *state)
        om.run_loop(900.0, nsteps, "RungeKutta4" if stepper == "RK4" else "ForwardEuler")
        with netcdf_file(out_fp, "r", mmap=False) as ds:
            f_u, f_h = np.array(ds.variables["normalVelocity"][:]).reshape(-1), np.array(ds.variables["layerThickness"][:]).reshape(-1)
        rel = lambda a, b: float(np.linalg.norm(a - b) / np.linalg.norm(b))          # noqa: E731
        e = (rel(f_u, om.normalVelocity[0]), rel(f_h, om.layerThickness[0]))         # the file holds the state one step before the last
        drift = abs(series[-1]["mass"] - series[0]["mass"]) / series[0]["mass"]
        ok = nsteps == 12 and max(e) <= 1e-12 and len(series) == 3 and drift <= 1e-13
        print(f"halo={halo} stepper={stepper} steps={nsteps} rel-L2 of the output file vs the oracle {e} mass drift {drift:.1e} graph: {status}")
        print("MULTI_GPU_DRIVER_OK" if ok else "MULTI_GPU_DRIVER_FAILED")
        for p in (mesh_fp, out_fp, cfg):
            try:
                os.remove(p)
            except OSError:
                pass
    sys.stdout.flush()
    dist.barrier()
    torch.cuda.synchronize()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()

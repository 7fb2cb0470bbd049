"""Model set-up, clocked run loop and driver, mirror of src/forward/init.jl, src/forward/run_loop.jl and
src/driver/mpas_ocean.jl over the B200 backend.

  ModelSetup                          src/infra/ModelSetup.jl:4-9
  ocn_init / ocn_setup_mesh / ocn_setup_clock / ocn_init_alarms     src/forward/init.jl:3-30,43-127
  ocn_run_loop(timestep, Prog, Diag, Tend, Setup, Stepper, clock, simulationAlarm, outputAlarm)   run_loop.jl:8-45
  ocn_run(config_fp)                  src/driver/mpas_ocean.jl:20-52

Two things the reference hard-codes are configuration here: the time stepper
(`omega.time_integration.config_time_integrator`: "ForwardEuler" | "RK4"; the reference driver passes ForwardEuler,
mpas_ocean.jl:39) and the backend (`B200(device)`, the reference edits mpas_ocean.jl:28).  The host loop does not
visit the device once per step: the clock is advanced on the host to find how many steps separate now from the next
ringing alarm, and that many steps run as one device-resident call (CUDA graphs inside libmoka_b200).
"""
from __future__ import annotations

import numpy as np

from . import api, io_netcdf
from ._lib import MokaError
from .config import ConfigGet, ConfigRead, GlobalConfig
from .time_manager import (Clock, OneTimeAlarm, PeriodicAlarm, Second, advance, attachAlarm, changeTimeStep, isRinging,
                           mpas_create_clock, reset)


class ModelSetup:
    """ModelSetup(config, mesh, timeManager) (ModelSetup.jl:4-9); `mesh_fields` keeps the host arrays write_netcdf needs."""

    def __init__(self, config: GlobalConfig, mesh: api.Mesh, timeManager: Clock, mesh_fields: dict | None = None):
        self.config, self.mesh, self.timeManager, self.mesh_fields = config, mesh, timeManager, mesh_fields


def ocn_setup_mesh(Config: GlobalConfig, backend: api.B200):
    """init.jl:43-56: streams.mesh.filename_template -> ReadHorzMesh + VerticalMesh."""
    mesh_fp = ConfigGet(ConfigGet(Config.streams, "mesh"), "filename_template")
    fields = io_netcdf.ReadHorzMesh(mesh_fp, backend=backend)
    return io_netcdf.VerticalMesh(mesh_fp, fields, backend=backend), fields


def ocn_setup_clock(Config: GlobalConfig) -> Clock:
    """init.jl:58-109."""
    outputConfig = ConfigGet(Config.streams, "output")
    tm = ConfigGet(Config.namelist, "time_management")
    ti = ConfigGet(Config.namelist, "time_integration")
    dt = ConfigGet(ti, "config_dt")
    stop_time, start_time = ConfigGet(tm, "config_stop_time"), ConfigGet(tm, "config_start_time")
    run_duration = ConfigGet(tm, "config_run_duration")
    ConfigGet(tm, "config_restart_timestamp_name")                      # read (and required) by the reference, unused
    output_reference_time = ConfigGet(outputConfig, "reference_time")
    output_interval = ConfigGet(outputConfig, "output_interval")
    if run_duration != "none":
        clock = mpas_create_clock(dt, start_time, runDuration=run_duration)
        if stop_time != "none":
            if start_time + run_duration != stop_time:                  # as in the reference the configured stop_time stays
                print("Warning: config_run_duration and config_stop_time are inconsitent: using config_run_duration.")
        else:
            stop_time = start_time + run_duration
    elif stop_time != "none":
        clock = mpas_create_clock(dt, start_time, stopTime=stop_time)
    else:
        raise MokaError("Error: Neither config_run_duration nor config_stop_time were specified.")
    attachAlarm(clock, OneTimeAlarm("simulation_end", stop_time))
    attachAlarm(clock, PeriodicAlarm("outputAlarm", output_interval, output_reference_time))
    return clock


def _stepper_from_config(Config: GlobalConfig):
    try:
        name = ConfigGet(ConfigGet(Config.namelist, "time_integration"), "config_time_integrator")
    except KeyError:
        return api.ForwardEuler                                         # what the reference driver runs (mpas_ocean.jl:39)
    table = {"forwardeuler": api.ForwardEuler, "forward_euler": api.ForwardEuler, "rk4": api.RungeKutta4, "rungekutta4": api.RungeKutta4}
    if str(name).lower() not in table:
        raise MokaError(f"unknown config_time_integrator {name}")
    return table[str(name).lower()]


def ocn_init(Config_filepath: str, backend: api.B200 | None = None):
    """ocn_init(Config_filepath; backend) (init.jl:3-30): returns (Setup, Diag, Tend, Prog)."""
    if backend is None:
        backend = api.B200(0)
    Config = ConfigRead(Config_filepath)
    mesh, fields = ocn_setup_mesh(Config, backend)
    clock = ocn_setup_clock(Config)
    Setup = ModelSetup(Config, mesh, clock, fields)
    # PrognosticVars(config, mesh; backend) (PrognosticVars.jl:59-106)
    tm = ConfigGet(Config.namelist, "time_management")
    if ConfigGet(tm, "config_do_restart"):
        raise MokaError("restart not yet supported")
    input_filename = ConfigGet(ConfigGet(Config.streams, "input"), "filename_template")
    nTimeLevels = ConfigGet(ConfigGet(Config.namelist, "time_integration"), "config_number_of_time_levels")
    ssh, u, h = io_netcdf.read_initial_state(input_filename, mesh.nCells, mesh.nEdges, mesh.nVertLevels)
    Prog = api.PrognosticVars(ssh, u, h, int(nTimeLevels), mesh)
    return Setup, api.DiagnosticVars(Prog), api.TendencyVars(Prog), Prog


def ocn_init_alarms(Setup: ModelSetup, dt_seconds: float | None = None):
    """ocn_init_alarms(Setup) (init.jl:111-127): overrides the configured time step with the hard-coded rule
    floor(2 * (mean(dcEdge)/1e3) * mean(dcEdge)/200e3) seconds -- which is 0 s below dc ~ 10 km, so a caller may pass
    `dt_seconds` (whole seconds, the clock's resolution) instead."""
    dt = api.reference_dt(Setup.mesh) if dt_seconds is None else float(dt_seconds)
    if dt <= 0 or dt != np.floor(dt):
        raise MokaError(f"ocn_init_alarms: time step {dt} s is not a positive whole number of seconds "
                        "(the reference rule init.jl:118 floors to 0 on fine meshes; pass dt_seconds)")
    changeTimeStep(Setup.timeManager, Second(int(dt)))
    clock = Setup.timeManager
    return clock, clock.alarms["simulation_end"], clock.alarms["outputAlarm"]


def ocn_run_loop(timestep, Prog, Diag, Tend, Setup, stepper, clock: Clock, simulationAlarm, outputAlarm, sum_ssh2: bool = False,
                 on_output=None, series: list | None = None):
    """run_loop.jl:8-45: `while !isRinging(simulationAlarm): advance!(clock); ocn_timestep(...); outputAlarm handling`.
    Steps between two alarm events run as one device-resident call.  `on_output(clock)` is called where the reference
    has its "should be doing i/o in here" placeholder (run_loop.jl:16-19).  With `sum_ssh2` the second method's
    squared-SSH sum is returned (run_loop.jl:24-44).  When `series` is a list, one record {time, steps, mass, energy,
    ssh2} (device reductions over the current state) is appended at every output alarm.  Returns the number of steps
    taken (the reference's global `i`)."""
    dt = float(np.asarray(timestep).reshape(-1)[0]) if not isinstance(timestep, (int, float)) else float(timestep)
    i = 0
    while not isRinging(simulationAlarm):
        n = 0
        while True:                                                     # batch the steps up to the next alarm event
            advance(clock)
            n += 1
            if isRinging(simulationAlarm) or isRinging(outputAlarm):
                break
            if clock.currTime > simulationAlarm.ringTime:
                raise MokaError("ocn_run_loop: the clock stepped over the simulation_end alarm without hitting it "
                                "(the reference loop would not terminate: updateStatus! tests equality, TimeManager.jl:130-132)")
        api.ocn_timestep(dt, Prog, Diag, Tend, Setup, stepper, nsteps=n)
        i += n
        if isRinging(outputAlarm):
            if series is not None:
                series.append({"time": clock.currTime, "steps": i, "mass": api.reduce_sum(Prog, "mass"),
                               "energy": api.reduce_sum(Prog, "energy"), "ssh2": api.reduce_sum(Prog, "ssh2")})
            if on_output is not None:
                on_output(clock)
            reset(outputAlarm)
    return (i, api.reduce_sum(Prog, "ssh2")) if sum_ssh2 else i


def ocn_run(config_fp: str, backend: api.B200 | None = None, stepper=None, dt_seconds: float | None = None):
    """ocn_run(config_fp) (mpas_ocean.jl:20-52): init, alarms, run loop, write_netcdf at the end."""
    Setup, Diag, Tend, Prog = ocn_init(config_fp, backend=backend)
    clock, simulationAlarm, outputAlarm = ocn_init_alarms(Setup, dt_seconds)
    timestep = Setup.timeManager.timeStep.seconds()
    stepper = stepper or _stepper_from_config(Setup.config)
    nsteps = ocn_run_loop(timestep, Prog, Diag, Tend, Setup, stepper, clock, simulationAlarm, outputAlarm)
    io_netcdf.write_netcdf(Setup, Diag, Prog)
    print("Moka.jl ran on GPU")                                         # mpas_ocean.jl:48-51
    print(clock.currTime)
    return Setup, Diag, Tend, Prog, nsteps


class _GatheredProg:
    """What write_netcdf reads of a PrognosticVars, filled from the ranks' owned parts."""

    def __init__(self, end, prev):
        (self.ssh, self.normalVelocity, self.layerThickness), (self.ssh_prev, self.normalVelocity_prev, self.layerThickness_prev) = end, prev


def ocn_run_decomposed(config_fp: str, backend: api.B200, device_index: int = 0, runtime=None, dt_seconds: float | None = None,
                       halo: str = "nccl", graph: bool = True, series: list | None = None):
    """`ocn_run` with the mesh decomposed over the ranks of a process group (one process per GPU; under torchrun every rank
    calls this with its own `B200(LOCAL_RANK)`): the same YAML, the same NetCDF mesh / initial state, the same clock and
    alarms, the configured stepper (ForwardEuler, the reference driver's, or RungeKutta4) through multi_gpu.DecomposedModel
    (halo exchange per stage / per step overlapped with the interior blocks), and on rank 0 the same output file as the single-device run writes -- bit for bit.  The reference has no multi-device
    driver.  Every rank reads the mesh and derives the same partition from it (deterministic), then keeps
    only its own part.  Returns (Setup, model, nsteps)."""
    from . import multi_gpu, partition
    Config = ConfigRead(config_fp)
    stepper = _stepper_from_config(Config)
    if stepper is api.ForwardEuler and halo != "nccl":
        raise MokaError("ocn_run_decomposed: ForwardEuler steps use the packed exchange (halo='nccl')")
    rt = runtime if runtime is not None else multi_gpu.TorchRuntime(device_index)
    rank, world = rt.rank_and_size()
    mesh_fp = ConfigGet(ConfigGet(Config.streams, "mesh"), "filename_template")
    fields = io_netcdf.read_mesh_fields(mesh_fp)
    clock = ocn_setup_clock(Config)
    input_filename = ConfigGet(ConfigGet(Config.streams, "input"), "filename_template")
    ssh, u, h = io_netcdf.read_initial_state(input_filename, fields["nCells"], fields["nEdges"], int(fields.get("nVertLevels", 1)))
    loc = partition.decompose(fields, world)[rank]
    model = multi_gpu.DecomposedModel(loc, multi_gpu.local_state(loc, ssh, u, h), backend, device_index, graph=graph, runtime=rt, halo=halo)
    Setup = ModelSetup(Config, model.mesh, clock, fields)
    dt = float(np.floor(2 * (np.mean(fields["dcEdge"]) / 1e3) * np.mean(fields["dcEdge"]) / 200e3)) if dt_seconds is None else float(dt_seconds)
    if dt <= 0 or dt != np.floor(dt):
        raise MokaError(f"ocn_run_decomposed: time step {dt} s is not a positive whole number of seconds (pass dt_seconds)")
    changeTimeStep(clock, Second(int(dt)))                              # ocn_init_alarms, init.jl:111-127 (the GLOBAL mean of dcEdge)
    simulationAlarm, outputAlarm = clock.alarms["simulation_end"], clock.alarms["outputAlarm"]
    i = 0
    while not isRinging(simulationAlarm):                               # ocn_run_loop, run_loop.jl:8-22, in batches between alarms
        n = 0
        while True:
            advance(clock)
            n += 1
            if isRinging(simulationAlarm) or isRinging(outputAlarm):
                break
            if clock.currTime > simulationAlarm.ringTime:
                raise MokaError("ocn_run_decomposed: the clock stepped over the simulation_end alarm without hitting it")
        model.step(dt, n, stepper=stepper)
        i += n
        if isRinging(outputAlarm):
            if series is not None:
                model.finish()
                series.append({"time": clock.currTime, "steps": i, "mass": model.reduce("mass"), "energy": model.reduce("energy"),
                               "ssh2": model.reduce("ssh2")})
            reset(outputAlarm)
    model.finish()
    end, prev = model.gather(fields["nCells"], fields["nEdges"]), model.gather(fields["nCells"], fields["nEdges"], previous=True)
    if rank == 0:
        io_netcdf.write_netcdf(Setup, None, _GatheredProg(end, prev))
        print(f"Moka.jl ran on {world} GPUs")
        print(clock.currTime)
    return Setup, model, i

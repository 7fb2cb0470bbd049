"""GPU: the driver path of the reference (src/driver/mpas_ocean.jl:20-52) end to end -- YAML config -> ocn_init (NetCDF
mesh + initial state) -> ocn_init_alarms -> clocked ocn_run_loop -> write_netcdf -- against the CPU oracle stepping the
same arrays."""
import numpy as np
import pytest
from scipy.io import netcdf_file

import moka_b200 as mb
import moka_oracle_c as OC
from moka_b200 import driver
from conftest import hex_mesh, rel_l2

pytestmark = pytest.mark.gpu

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


def _stage(tmp_path, stepper):
    m = hex_mesh(24, 24, 300.0e3)                      # dc = 300 km -> the reference's dt rule gives 900 s (init.jl:118)
    state = mb.inertialGravityWave(m).initial_state()
    mesh_fp, out_fp, cfg = str(tmp_path / "mesh.nc"), str(tmp_path / f"out_{stepper}.nc"), str(tmp_path / f"cfg_{stepper}.yml")
    mb.write_mesh_netcdf(mesh_fp, m, state)
    with open(cfg, "w") as f:
        f.write(YAML.format(stepper=stepper, mesh=mesh_fp, out=out_fp))
    return m, state, cfg, out_fp


@pytest.mark.parametrize("stepper", ["ForwardEuler", "RK4"])
def test_ocn_run_from_yaml_and_netcdf(backend, tmp_path, stepper):
    m, (ssh, u, h), cfg, out_fp = _stage(tmp_path, stepper)
    Setup, Diag, Tend, Prog, nsteps = mb.ocn_run(cfg, backend=backend)
    assert nsteps == 12 and Setup.timeManager.timeStep.seconds() == 900.0
    om = OC.OracleModel(m, ssh, u, h)
    name = "ForwardEuler" if stepper == "ForwardEuler" else "RungeKutta4"
    om.run_loop(900.0, 12, name)
    if stepper == "ForwardEuler":                       # live reference path: bit-exact
        assert np.array_equal(Prog.normalVelocity, om.normalVelocity[1]) and np.array_equal(Prog.layerThickness, om.layerThickness[1])
    else:
        assert rel_l2(Prog.normalVelocity, om.normalVelocity[1]) <= 1e-12 and rel_l2(Prog.layerThickness, om.layerThickness[1]) <= 1e-12
    with netcdf_file(out_fp, "r", mmap=False) as ds:
        assert float(ds.dt) == 900.0 and float(ds.variables["time"][0]) == 10800.0
        assert ds.dimensions["nCells"] == m["nCells"] and ds.dimensions["TWO"] == 2 and ds.dimensions["time"] == 1
        for k in ("xCell", "yEdge", "xVertex", "dcEdge", "areaCell", "areaTriangle", "nEdgesOnCell", "nEdgesOnEdge"):
            assert np.array_equal(ds.variables[k][:], m[k]), k
        for k in ("angleEdge", "edgeSignOnCell", "cellsOnEdge", "verticesOnCell", "verticesOnEdge"):
            assert k in ds.variables                     # defined, never written (OutPut.jl:173-204)
        # the reference writes time level 1 = the state one step before the last (PrognosticVars.jl:108-113)
        f_ssh, f_u = np.array(ds.variables["ssh"][:]), np.array(ds.variables["normalVelocity"][:]).reshape(-1)
        assert ds.variables["layerThickness"].dimensions == ("nVertLevels", "nCells")
    if stepper == "ForwardEuler":
        assert np.array_equal(f_ssh, om.ssh[0]) and np.array_equal(f_u, om.normalVelocity[0])
    else:
        assert rel_l2(f_u, om.normalVelocity[0]) <= 1e-12


def test_clocked_run_loop_batches_between_alarms(backend, tmp_path):
    m, (ssh, u, h), cfg, out_fp = _stage(tmp_path, "RK4")
    Setup, Diag, Tend, Prog = mb.ocn_init(cfg, backend=backend)
    clock, sim, outp = mb.ocn_init_alarms(Setup)
    rings = []
    l0 = backend.launch_count()
    n, s2 = driver.ocn_run_loop(900.0, Prog, Diag, Tend, Setup, mb.RungeKutta4, clock, sim, outp, sum_ssh2=True,
                                on_output=lambda c: rings.append(c.currTime))
    assert n == 12 and len(rings) == 3 and [t.hour for t in rings] == [1, 2, 3]
    om = OC.OracleModel(m, ssh, u, h)
    om.run_loop(900.0, 12, "RungeKutta4")
    assert abs(s2 - float(np.sum(om.ssh[1] ** 2))) <= 1e-10 * s2
    # 4 stage kernels per step + ssh refresh of both time levels per batch + 2 reduction kernels + the two one-off
    # kernels that build the fused-form mesh arrays on the first RK4 call: nothing per step beyond the stages
    assert backend.launch_count() - l0 <= 12 * 4 + 3 * 2 + 2 + 2
    # error behaviour: a time step that does not land on the simulation_end alarm is detected instead of looping forever
    Setup2, Diag2, Tend2, Prog2 = mb.ocn_init(cfg, backend=backend)
    clock2, sim2, out2 = mb.ocn_init_alarms(Setup2, dt_seconds=7000)
    with pytest.raises(mb.MokaError, match="stepped over"):
        driver.ocn_run_loop(7000.0, Prog2, Diag2, Tend2, Setup2, mb.RungeKutta4, clock2, sim2, out2)
    with pytest.raises(mb.MokaError):
        mb.ocn_init_alarms(Setup2, dt_seconds=0.5)


def test_consistent_diagnostics_and_conservation_series(backend, tmp_path):
    """diagnostics at face value (flux of the same state, vorticity zeroed) against the oracle's
    operators, and a mass / energy time series recorded at the output alarms."""
    import moka_oracle as O
    m, (ssh, u, h), cfg, out_fp = _stage(tmp_path, "RK4")
    Setup, Diag, Tend, Prog = mb.ocn_init(cfg, backend=backend)
    clock, sim, outp = mb.ocn_init_alarms(Setup)
    series = []
    driver.ocn_run_loop(900.0, Prog, Diag, Tend, Setup, mb.RungeKutta4, clock, sim, outp, series=series)
    assert [r["steps"] for r in series] == [4, 8, 12]
    mass0 = float(np.sum(m["areaCell"] * h))
    assert all(abs(r["mass"] - mass0) <= 1e-13 * mass0 for r in series)          # flux form: mass to round-off
    e = [r["energy"] for r in series]
    assert e[0] > e[1] > e[2] and e[0] - e[2] <= 1e-4 * e[0]                     # RK4 damps slightly (2.5e-5 over 8 steps)
    om = OC.OracleModel(m, ssh, u, h)
    om.run_loop(900.0, 12, "RungeKutta4")
    he = O.interpolate_cell2edge(m, om.layerThickness[1])
    want = float(np.sum(m["areaCell"] * 0.5 * 9.80616 * om.ssh[1] ** 2) + np.sum(0.5 * m["dcEdge"] * m["dvEdge"] * he * om.normalVelocity[1] ** 2))
    assert abs(e[2] - want) <= 1e-12 * want
    for _ in range(2):                                                           # twice: nothing accumulates
        mb.diagnostic_compute(Setup.mesh, Diag, Prog, consistent=True)
    un, hn = Prog.normalVelocity, Prog.layerThickness
    hedge = O.interpolate_cell2edge(m, hn)
    assert np.array_equal(Diag.layerThicknessEdge, hedge) and np.array_equal(Diag.thicknessFlux, un * hedge)
    assert np.array_equal(Diag.velocityDivCell, O.divergence_on_cell(m, un)[0])
    assert np.array_equal(Diag.relativeVorticity, O.curl_on_vertex(m, un))

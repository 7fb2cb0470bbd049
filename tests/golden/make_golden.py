"""Generates tests/golden/igw16_rk4_fe.npz from the C oracle (the reference cannot run here: no
Julia).  Inputs are analytic (inertialGravityWave.jl), so the fixture is fully determined by
(nx, ny, dt, nsteps).  Run: python tests/golden/make_golden.py"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [os.path.join(ROOT, "mpas-ocean.jl_b200"), os.path.join(ROOT, "oracle")]
import moka_b200.planar_hex as ph  # noqa: E402
import moka_oracle as O  # noqa: E402
import moka_oracle_c as OC  # noqa: E402

nx, nsteps = 16, 25
m = ph.periodic_hex(nx, nx, 1.0e7 / nx)
OC.sign_index_fields(m)
ssh, u, h = O.InertialGravityWave(m).initial_state()
dt = O.reference_dt(m)
out = {"meta": json.dumps({"nx": nx, "ny": nx, "dc": 1.0e7 / nx, "dt": dt, "nsteps": nsteps})}
for stepper in ("ForwardEuler", "RungeKutta4"):
    om = OC.OracleModel(m, ssh, u, h)
    om.run_loop(dt, nsteps, stepper)
    out[f"{stepper}_ssh"] = om.ssh[1].copy()
    out[f"{stepper}_normalVelocity"] = om.normalVelocity[1].copy()
    out[f"{stepper}_layerThickness"] = om.layerThickness[1].copy()
    om2 = OC.OracleModel(m, ssh, u, h)
    om2.run_loop(dt, 1, stepper)
    om2.diagnostic_compute()
    out[f"{stepper}_tendU_after1"] = om2.compute_normal_velocity_tendency().copy()
    out[f"{stepper}_tendH_after1"] = om2.compute_layer_thickness_tendency().copy()
np.savez_compressed(os.path.join(HERE, "igw16_rk4_fe.npz"), **out)
print("wrote", os.path.join(HERE, "igw16_rk4_fe.npz"), "dt", dt)

"""Generates tests/golden/igw16_adjoint.npz and igw16_adjoint_fe.npz from the adjoint oracle (oracle/adjoint_oracle.py): J = sum ssh^2 after
`nsteps` RK4 steps of the 16x16 inertia-gravity-wave case and dJ/d(normalVelocity_0, layerThickness_0), plus the
finite-difference values of the reference's own acceptance test (test/enzyme/test_Enzyme_end2end.jl:112-180) at the
index it checks (5, 1-based); the second file holds the same for ForwardEuler (the stepper the reference differentiates),
plus dJ/d(ssh_0).  Inputs are analytic, so the fixtures are determined by (nx, dt, nsteps).
Run: python tests/golden/make_golden_adjoint.py"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [os.path.join(ROOT, "mpas-ocean.jl_b200"), os.path.join(ROOT, "oracle")]
import adjoint_oracle as A  # noqa: E402
import moka_b200.planar_hex as ph  # noqa: E402
import moka_oracle as O  # noqa: E402
import moka_oracle_c as OC  # noqa: E402

nx, nsteps = 16, 8
m = ph.periodic_hex(nx, nx, 1.0e7 / nx)
OC.sign_index_fields(m)
ssh, u, h = O.InertialGravityWave(m).initial_state()
dt = 0.5 * m["dc"] / float(np.sqrt(O.GRAVITY * 1000.0))
J, gu, gh = A.gradient_sum_ssh2(m, u, h, dt, nsteps)
k = 4
out = {"meta": json.dumps({"nx": nx, "dt": dt, "nsteps": nsteps, "fd_index": k}), "J": J, "d_normalVelocity": gu, "d_layerThickness": gh,
       "fd_layerThickness": A.finite_difference(m, u, h, dt, nsteps, "h", k, eps=1e-7),
       "fd_normalVelocity": A.finite_difference(m, u, h, dt, nsteps, "u", k, eps=1e-4)}
np.savez_compressed(os.path.join(HERE, "igw16_adjoint.npz"), **out)
print("wrote igw16_adjoint.npz: J", J, "dJ/dh[k]", gh[k], "fd", out["fd_layerThickness"], "dJ/du[k]", gu[k], "fd", out["fd_normalVelocity"])

J, gu, gh, gs, ge = A.gradient_sum_ssh2_fe(m, ssh, u, h, dt, nsteps)
out = {"meta": json.dumps({"nx": nx, "dt": dt, "nsteps": nsteps, "fd_index": k, "stepper": "ForwardEuler"}), "J": J,
       "d_normalVelocity": gu, "d_layerThickness": gh, "d_ssh": gs, "d_layerThicknessEdge": ge,
       "fd_layerThickness": A.finite_difference_fe(m, ssh, u, h, dt, nsteps, "h", k, eps=1e-7),
       "fd_normalVelocity": A.finite_difference_fe(m, ssh, u, h, dt, nsteps, "u", k, eps=1e-4)}
np.savez_compressed(os.path.join(HERE, "igw16_adjoint_fe.npz"), **out)
print("wrote igw16_adjoint_fe.npz: J", J, "dJ/dh[k]", gh[k], "fd", out["fd_layerThickness"], "dJ/du[k]", gu[k], "fd", out["fd_normalVelocity"])

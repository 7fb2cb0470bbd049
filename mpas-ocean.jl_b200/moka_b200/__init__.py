"""moka_b200 -- host-side mirror of MPAS-Ocean.jl's forward-model API over libmoka_b200.so (sm_100a)."""
from ._lib import LIB_PATH, MokaError, SYMBOLS  # noqa: F401
from .api import (B200, CurlOnVertex, ShadowPrognosticVars, autodiff_reverse_run_loop, ocn_init_shadows, DiagnosticVars, DivergenceOnCell, ForwardEuler, GradientOnEdge, Mesh,  # noqa: F401
                  PrognosticVars, RungeKutta4, TendencyVars, cfl_dt, check_eltype_args, check_typeof_args,
                  computeLayerThicknessTendency, computeNormalVelocityTendency, diagnostic_compute,
                  inertialGravityWave, interpolateCell2Edge, kelvinWave, ocn_run_loop, ocn_timestep, reduce_sum, reference_dt)
from .planar_hex import channel_hex, periodic_hex  # noqa: F401

"""moka_b200 -- host-side mirror of MPAS-Ocean.jl's forward-model API over libmoka_b200.so (sm_100a)."""
from ._lib import LIB_PATH, MokaError, SYMBOLS  # noqa: F401
from .api import (B200, CurlOnVertex, DivergenceOnCell_vjp, GradientOnEdge_vjp, ShadowPrognosticVars, autodiff_reverse_run_loop, ocn_init_shadows, DiagnosticVars, DivergenceOnCell, ForwardEuler, GradientOnEdge, Mesh,  # noqa: F401
                  PrognosticVars, RungeKutta4, TendencyVars, cfl_dt, check_eltype_args, check_typeof_args,
                  computeLayerThicknessTendency, computeNormalVelocityTendency, diagnostic_compute,
                  inertialGravityWave, interpolateCell2Edge, kelvinWave, ocn_run_loop, ocn_timestep, reduce_sum, reference_dt)
from .planar_hex import channel_hex, periodic_hex  # noqa: F401
from .planar_voronoi import periodic_voronoi  # noqa: F401
from .spherical_voronoi import geostrophic_zonal_flow, spherical_voronoi  # noqa: F401
from . import config, driver, io_netcdf, time_manager  # noqa: F401,E402
from .config import ConfigAdd, ConfigGet, ConfigRead, ConfigSet, GlobalConfig, yaml_config  # noqa: F401,E402
from .driver import ModelSetup, ocn_init, ocn_init_alarms, ocn_run, ocn_run_decomposed  # noqa: F401,E402
from .io_netcdf import ReadHorzMesh, VerticalMesh, write_mesh_netcdf, write_netcdf  # noqa: F401,E402

"""python tests/sim/run_on_sim.py script.py args...  -- run a repo script with the simulated runtime bound as THE library"""
import ctypes, os, runpy, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(ROOT, "mpas-ocean.jl_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests", "sim"), ROOT]
import simcuda
from moka_b200 import _lib
simcuda.runtime(); _lib.bind(ctypes.CDLL(simcuda._build.LIB))
sys.argv = sys.argv[1:]
runpy.run_path(sys.argv[0], run_name="__main__")

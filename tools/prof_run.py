import sys, ctypes, os
sys.path.insert(0, '/root/repo')
from pyrhe_b200 import _lib
_lib.LIB_PATH = os.path.join(os.path.dirname(_lib.LIB_PATH), 'libpyrhe_b200_prof.so')
sys.argv = ['bench.py', '--workload', 'profile5', '--steps', '1', '--warmup', '0', '--no_cpu_baseline', '--no_e2e']
import runpy
try:
    runpy.run_path('/root/repo/bench.py', run_name='__main__')
finally:
    lib = _lib.load()
    lib.rhe_tc_prof_dump()

"""bench.py against an older build of the library (same-box comparisons): python tools/bench_old_lib.py <path to .so> [bench flags]"""
import ctypes, sys
sys.path.insert(0, ".")
from randnla_b200 import _lib
_lib.LIB_PATH = sys.argv[1]
probe = ctypes.CDLL(_lib.LIB_PATH)
for k in list(_lib.SIGNATURES):
    if not hasattr(probe, k):
        del _lib.SIGNATURES[k]
import bench
sys.argv = ["bench.py"] + sys.argv[2:]
bench.main()

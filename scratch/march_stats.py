import ctypes, os, sys, numpy as np
sys.path.insert(0, os.getcwd())
os.environ["ENRGY_B200_LIB"] = os.path.join(os.getcwd(), "scratch/variants/marchstats.so")
sys.argv = ["x", "--nsteps", sys.argv[1] if len(sys.argv) > 1 else "256"]
exec(open("scripts/measure_shadow.py").read())
from enrgy_b200 import _lib
lib = _lib.load()
buf = (ctypes.c_ulonglong * 32)()
lib.enrgy_debug_march_stats(buf)
v = np.array(list(buf), dtype=np.float64) / 5.0     # 5 launches in measure_shadow (1 + 4)
print("chunks sampled", v[0], "mean active rays/chunk", v[1] / v[0], "inactive rows per chunk (of 8)", v[2] / v[0])
print("chunk index histogram", (v[4:16] / v[0]).round(3))
print("active-ray histogram (bins of 32)", (v[16:25] / v[0]).round(3))

"""Debug: MSM f64 run in one piece vs split at n-1 with a dump in between."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from enrgy_b200.synthetic import make_case
from tests import parity as P
case = make_case(48, 16, w=60, seed=37)
msm = dict(depths=[0.1, 0.1, 0.3, 0.5, 0.5, 0.5, 3.0], temperatures=[-6.9, -6.93, -7.025, -7.31, -6.93, -7.12, -7.0, -5.57], elev=275.0)
def run(mode):
    eng = P.make_engine(case, True, computed=True, shadow=True, msm=msm)
    n = 16
    if mode == 0:
        eng.run(0, n)
    elif mode == 1:
        eng.run(0, n - 1); eng.run(n - 1, n)
    elif mode == 2:
        eng.run(0, n - 1); eng.dump_steps(n - 1, n); eng.run(n - 1, n)
    elif mode == 3:
        eng.defer_snow_total(True); eng.run(0, n - 1); eng.defer_snow_total(False); eng.dump_steps(n - 1, n); eng.run(n - 1, n)
    st = eng.state(np.float64)
    eng.close()
    return st
ref = run(0)
for m in (1, 2, 3):
    st = run(m)
    print("mode", m, [float(np.nanmax(np.abs(a - b))) for a, b in zip(st, ref)])

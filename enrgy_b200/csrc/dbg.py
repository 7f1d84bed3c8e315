import sys
sys.path.insert(0,'.')
from enrgy_b200.synthetic import make_case
from tests import parity as P
import numpy as np
case = make_case(64, 6, w=96)
for f64 in (False, True):
    try:
        eng = P.make_engine(case, f64, computed=True, shadow=True)
        st = eng.run(0, 6)
        print("f64" if f64 else "f32", "ok", st[3, :3])
        eng.close()
    except Exception as e:
        print("f64" if f64 else "f32", "FAIL", str(e)[:200])

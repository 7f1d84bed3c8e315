import sys, os
sys.path.insert(0, os.getcwd())
from enrgy_b200.synthetic import make_case
from tests import parity as P
case = make_case(64, 6, w=96)
for sh in (True, False):
    res = P.compare_run(case, False, computed=True, shadow=sh)
    print("shadow", sh, {k: float("%.3g" % v) for k, v in sorted(res.items(), key=lambda kv: -kv[1])[:6]})
case = make_case(96, 24, w=128, seed=3)
res = P.compare_run(case, False, computed=True, shadow=True)
print("96x128x24", {k: float("%.3g" % v) for k, v in sorted(res.items(), key=lambda kv: -kv[1])[:6]})

#!/usr/bin/env python
"""Host-side breakdown of one end-to-end pass (pinned host rasters -> device -> state back) on C2."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from enrgy_b200 import _lib                                   # noqa: E402
from enrgy_b200._lib import check                             # noqa: E402
from enrgy_b200.engine import Engine                          # noqa: E402
from enrgy_b200.forcing import build_forcing                  # noqa: E402
from enrgy_b200.synthetic import make_band_case               # noqa: E402

n, T = 2048, 2200
case, dem = make_band_case(n, T)
keys = list(case.albedo_maps)


def pin(a):
    t = torch.empty(a.shape, dtype=torch.float32, pin_memory=True)
    t.numpy()[...] = a
    return t


pd, ps = pin(dem), pin(case.swe)
pa = [pin(case.albedo_maps[k]) for k in keys]
out = [torch.empty(case.dem.shape, dtype=torch.float32, pin_memory=True) for _ in range(3)]
stats = torch.empty((T, _lib.S_COUNT), dtype=torch.float64, pin_memory=True)
eng = Engine(n, n, precision=_lib.F32)
eng.set_params(cell_size=10.0, elev_aws=case.elev_aws, aws_row=case.aws_rc[0], aws_col=case.aws_rc[1],
               sensor_z=1.6, zm=1e-3, z_h_or_e=1e-4, emissivity=0.98, lat=case.lat, lon=case.lon,
               insol_mode=_lib.INSOL_COMPUTED, shadow=False)
table = build_forcing(case.aws_rows, keys)
for it in range(4):
    t = [time.perf_counter()]
    eng.set_dem(pd.numpy()); t.append(time.perf_counter())
    eng.set_forcing(table); t.append(time.perf_counter())
    eng.set_albedo_maps([a.numpy() for a in pa]); t.append(time.perf_counter())
    eng.set_swe(ps.numpy()); t.append(time.perf_counter())
    eng.prepass(); t.append(time.perf_counter())
    check(eng.lib.enrgy_run(eng.h, 0, T, stats.numpy().ctypes.data)); t.append(time.perf_counter())
    check(eng.lib.enrgy_get_state(eng.h, 32, *[o.numpy().ctypes.data for o in out])); t.append(time.perf_counter())
    names = ["set_dem", "set_forcing", "set_albedo_maps", "set_swe", "prepass", "run", "get_state"]
    print("  ".join("%s %.2f" % (nm, (b - a) * 1e3) for nm, a, b in zip(names, t, t[1:])), " total %.2f ms  kernel %.2f" % ((t[-1] - t[0]) * 1e3, eng.last_kernel_ms()))
eng.close()

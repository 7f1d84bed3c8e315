#!/usr/bin/env python
"""Kernel time of the shading path (INSOL = 2) on the C2 raster for the library named by
ENRGY_B200_LIB; prints cell-steps/s, kernel ms and the launch configuration.
  python scripts/measure_shadow.py [--size 2048] [--nsteps 256] [--dtype f32]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from enrgy_b200 import _lib                                   # noqa: E402
from enrgy_b200.engine import Engine                          # noqa: E402
from enrgy_b200.forcing import build_forcing                  # noqa: E402
from enrgy_b200.synthetic import make_band_case               # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=2048)
ap.add_argument("--nsteps", type=int, default=256)
ap.add_argument("--dtype", default="f32")
ap.add_argument("--noshadow", action="store_true")
ap.add_argument("--step-s", type=int, default=3600, help="seconds between AWS rows (900 = config C4's time base)")
a = ap.parse_args()
case, dem = make_band_case(a.size, a.nsteps, step_s=a.step_s)
keys = list(case.albedo_maps)
eng = Engine(a.size, a.size, precision=_lib.F32 if a.dtype == "f32" else _lib.F64)
eng.set_params(cell_size=10.0, elev_aws=case.elev_aws, aws_row=case.aws_rc[0], aws_col=case.aws_rc[1],
               sensor_z=1.6, zm=1e-3, z_h_or_e=1e-4, emissivity=0.98, lat=case.lat, lon=case.lon,
               insol_mode=_lib.INSOL_COMPUTED, shadow=not a.noshadow)
eng.set_dem(dem)
eng.set_albedo_maps([case.albedo_maps[k] for k in keys])
eng.set_swe(case.swe)
eng.set_forcing(build_forcing(case.aws_rows, keys))
eng.prepass()
eng.snapshot(save=True)
best = 1e30
for _ in range(4):
    eng.snapshot(save=False)
    st = eng.run(0, a.nsteps)
    best = min(best, eng.last_kernel_ms())
import numpy as np
chk = float(np.nansum(st[:, _lib.S_MELT]))
print("%-28s %.4g cell-steps/s  %.3f ms  %s  check %.10g" % (os.path.basename(_lib.LIB_PATH), float(a.size) ** 2 * a.nsteps / (best * 1e-3), best, eng.kernel_info(), chk))
eng.close()

#!/usr/bin/env python
"""Kernel times of the shading path on a square raster for the library named by ENRGY_B200_LIB: the
line sweep (shade.cu) and the mask-fed fused kernel; prints cell-steps/s, ms and the launch configuration.
Every pass sweeps afresh (the mask budget is set so that the cached masks are never reused).
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
import numpy as np
import torch
best, best_sw = 1e30, 1e30
if a.noshadow:
    for _ in range(4):
        eng.snapshot(save=False)
        st = eng.run(0, a.nsteps)
        best = min(best, eng.last_kernel_ms())
    best_sw = 0.0
else:
    s0, s1 = eng.sub_range(0, a.nsteps)
    masks = torch.empty((s1 - s0) * eng.mask_words(a.size), dtype=torch.int32, device="cuda")
    stats = torch.zeros((a.nsteps, _lib.S_COUNT), dtype=torch.float64, device="cuda")
    for _ in range(4):
        eng.snapshot(save=False)
        eng.shade_scan(s0, s1, [(0, a.size, masks.data_ptr())])
        eng.run_masked(0, a.nsteps, masks.data_ptr(), stats.data_ptr())
        eng.synchronize()
        best = min(best, eng.last_kernel_ms())
        best_sw = min(best_sw, eng.last_sweep_ms())
    st = stats.cpu().numpy()
chk = float(np.nansum(st[:, _lib.S_MELT]))
cs = float(a.size) ** 2 * a.nsteps
print("%-22s total %.4g cell-steps/s | sweep %.3f ms (%.4g cell-sub-steps/s) | fused %.3f ms (%.4g cell-steps/s) | %s check %.10g"
      % (os.path.basename(_lib.LIB_PATH), cs / ((best + best_sw) * 1e-3), best_sw,
         (float(a.size) ** 2 * (0 if a.noshadow else s1 - s0) / (best_sw * 1e-3)) if best_sw else 0.0, best, cs / (best * 1e-3),
         eng.kernel_info(), chk))
eng.close()

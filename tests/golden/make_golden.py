#!/usr/bin/env python
"""Generates tests/golden/*.npz by running the UNMODIFIED reference (tepextepex/ENRGY at
/root/reference) through oracle/ref_harness.py on small seeded synthetic inputs.

Run in the build container only (the reference checkout does not exist on the GPU box):
    python tests/golden/make_golden.py
Each fixture stores the inputs' recipe (so tests regenerate identical inputs from
enrgy_b200.synthetic), the insolation rasters fed to the reference, and the reference's outputs:
final state rasters, the per-step flux rasters of two steps, the heat_fluxes.csv text and the
NumPy version (dtype flow depends on it).
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from enrgy_b200.synthetic import make_case            # noqa: E402
from oracle import insolation_oracle as I             # noqa: E402
from oracle.ref_harness import run_reference          # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = {
    "maps": dict(),
    "const_albedo": dict(const_albedo=(0.35, 0.75)),
    "snow_ageing_andreas": dict(last_snowfall="20220522", max_ice_albedo=0.38, andreas=True),
    "corr_defaults": dict(cloud_corr=0.2, sensible_corr=1.1, latent_corr=0.9, emissivity=None, zm=None,
                          z_h_or_e=None),
    "msm": dict(msm=dict(depths=[0.1, 0.1, 0.3, 0.5, 0.5, 0.5, 3.0],
                         temperatures=[-6.9, -6.93, -7.025, -7.31, -6.93, -7.12, -7.0, -5.57], elev=275.0),
                snow_density=350.0),
}
RECIPE = dict(n=40, n_steps=10, seed=11, w=56, calm_every=4)
KEEP = (1, 7)


def main():
    case = make_case(RECIPE["n"], RECIPE["n_steps"], seed=RECIPE["seed"], w=RECIPE["w"],
                     calm_every=RECIPE["calm_every"])
    pot64 = I.insolation_series(case, shadow=True, dtype=np.float64)
    for f64 in (False, True):
        pot = pot64 if f64 else pot64.astype(np.float32)
        for name, kw in CASES.items():
            full = dict(z=1.6, zm=1e-3, z_h_or_e=1e-4, emissivity=0.98)
            full.update(kw)
            r = run_reference(case, pot, f64=f64, **full)
            out = {
                "recipe": json.dumps(RECIPE), "kwargs": json.dumps(full), "f64": f64,
                "numpy": r["numpy"], "stats_csv": r["stats_csv"], "solar_csv": r["solar_csv"],
                "pot": pot, "swe": r["swe"], "total_snow": r["total_snow"], "total_ice": r["total_ice"],
            }
            for i in KEEP:
                for k in ("lwd", "lwu", "rs", "sens", "lat", "atmo", "g", "mf"):
                    out["step%d_%s" % (i, k)] = r["rows"][i][k]
                out["step%d_snow" % i] = r["melt"][i][0]
                out["step%d_ice" % i] = r["melt"][i][1]
            if r["layer_temperatures"] is not None:
                out["layer_temperatures"] = np.stack(r["layer_temperatures"])
            path = os.path.join(HERE, "%s_%s.npz" % (name, "f64" if f64 else "f32"))
            np.savez_compressed(path, **out)
            print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Generates tests/golden_spec/*.npz from THIS REPO's own specifications (oracle/), not from the
reference: the station blend + cloud attenuation of BASELINE config C4 (oracle/enrgy_oracle.py, "several
weather stations") has no counterpart upstream, so these fixtures pin nothing against the reference --
they freeze the specification, so that a later edit of the oracle cannot move it unnoticed.
    python tests/golden_spec/make_spec_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from enrgy_b200.synthetic import make_case, make_station_rows      # noqa: E402
from tests import parity as P                                      # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
RECIPE = dict(n=40, n_steps=10, seed=11, w=56, calm_every=4)
SPOTS = [(6.0, 40.0, 140.0, 11), (33.4, 9.5, -90.0, 12), (18.0, 50.0, 60.0, 13)]
KEEP = (1, 7)


def stations_of(case):
    return [dict(row=r, col=c, elev=case.elev_aws + dz, rows=make_station_rows(case, case.elev_aws + dz, seed=sd))
            for (r, c, dz, sd) in SPOTS]


def main():
    case = make_case(RECIPE["n"], RECIPE["n_steps"], seed=RECIPE["seed"], w=RECIPE["w"], calm_every=RECIPE["calm_every"])
    pot = P.random_insolation(case, RECIPE["n_steps"])
    for f64 in (False, True):
        r = P.run_oracle(case, pot, f64, stations=stations_of(case), cloud_k=0.8, cloud_corr=0.1, last_snowfall="20220525")
        out = {"recipe": json.dumps(RECIPE), "spots": json.dumps(SPOTS), "f64": f64, "numpy": np.__version__,
               "stats_csv": r["stats_csv"], "swe": r["swe"], "total_snow": r["total_snow"], "total_ice": r["total_ice"]}
        for i in KEEP:
            for k in ("lwd", "rs", "sens", "lat", "atmo", "mf"):
                out["step%d_%s" % (i, k)] = r["rows"][i][k]
        path = os.path.join(HERE, "stations_%s.npz" % ("f64" if f64 else "f32"))
        np.savez_compressed(path, **out)
        print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()

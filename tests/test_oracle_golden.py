"""The NumPy oracle against the golden fixtures made by the UNMODIFIED reference
(tests/golden/make_golden.py): bit-identical rasters and identical CSV text."""
import glob
import json
import os

import numpy as np
import pytest

from enrgy_b200.synthetic import make_case
from oracle import enrgy_oracle as O
from tests.parity import clipped_albedo

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))


def _run_oracle(g):
    recipe = json.loads(str(g["recipe"]))
    kw = json.loads(str(g["kwargs"]))
    f64 = bool(g["f64"])
    dt = np.float64 if f64 else np.float32
    case = make_case(recipe["n"], recipe["n_steps"], seed=recipe["seed"], w=recipe["w"],
                     calm_every=recipe["calm_every"])
    cfg = O.ModelConfig(z=kw["z"], elev_aws=case.elev_aws, xy_aws=case.xy_aws, zm=kw["zm"],
                        z_h_or_e=kw["z_h_or_e"], andreas=kw.get("andreas", False),
                        const_albedo=tuple(kw["const_albedo"]) if kw.get("const_albedo") else None,
                        last_snowfall=kw.get("last_snowfall"), max_ice_albedo=kw.get("max_ice_albedo"),
                        emissivity=kw["emissivity"], cloud_corr=kw.get("cloud_corr"),
                        sensible_corr=kw.get("sensible_corr", 1), latent_corr=kw.get("latent_corr", 1),
                        msm=kw.get("msm"), snow_density=kw.get("snow_density"))
    alb = None if kw.get("const_albedo") else clipped_albedo(case, dt)
    return O.run_model(case.dem.astype(dt), case.geotransform, case.aws_rows, g["pot"], cfg,
                       swe=case.swe.astype(dt), albedo_arrays=alb, state_dtype=dt)


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_oracle_reproduces_reference(path):
    g = np.load(path, allow_pickle=False)
    if str(g["numpy"]).split(".")[0] != np.__version__.split(".")[0]:
        pytest.skip("fixture made with numpy %s: dtype flow differs across major versions (NEP 50)" % g["numpy"])
    out = _run_oracle(g)
    for k in ("swe", "total_snow", "total_ice"):
        assert out[k].dtype == g[k].dtype, k
        assert np.array_equal(out[k], g[k], equal_nan=True), k
    for key in g.files:
        if key.startswith("step"):
            step, name = key.split("_", 1)
            i = int(step[4:])
            if name in ("snow", "ice"):
                got = out["melt"][i][0 if name == "snow" else 1]
            else:
                got = out["rows"][i][name]
            assert np.array_equal(got, g[key], equal_nan=True), key
    assert out["stats_csv"] == str(g["stats_csv"])
    assert out["solar_csv"] == str(g["solar_csv"])
    if "layer_temperatures" in g.files:
        assert np.array_equal(np.stack(out["layer_temperatures"]), g["layer_temperatures"], equal_nan=True)


def test_golden_present():
    assert len(GOLDEN) >= 10

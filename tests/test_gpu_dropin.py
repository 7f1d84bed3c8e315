"""The reference-shaped `Energy` class end to end on files (GDAL-free .npy rasters), against the
oracle's CSV and rasters."""
import os

import numpy as np
import pytest

from enrgy_b200 import Energy
from enrgy_b200.raster_utils import save_npy_raster
from enrgy_b200.synthetic import make_case
from oracle import insolation_oracle as I
from tests import parity as P

pytestmark = pytest.mark.gpu


def _write_case(case, d):
    dem = save_npy_raster(os.path.join(d, "dem.npy"), case.dem, case.geotransform)
    swe = save_npy_raster(os.path.join(d, "swe.npy"), case.swe, case.geotransform)
    alb = {k: save_npy_raster(os.path.join(d, "alb_%s.npy" % k), a, case.geotransform)
           for k, a in case.albedo_maps.items()}
    aws = case.write_aws_csv(os.path.join(d, "aws.csv"))
    return dem, swe, alb, aws


def _csv_numbers(text):
    rows = []
    for line in text.split("\n"):
        if line.startswith("#") or line.startswith("DATE") or not line.strip():
            continue
        parts = line.split(",")
        rows.append((parts[0], [float(x) for x in parts[1:]]))
    return rows


@pytest.mark.parametrize("precision", ["f32", "f64"])
def test_precomputed_insolation_pickles(tmp_path, precision):
    """use_precomputed + add_pickle_dir: the reference's own cache layout (model.py:465-481)."""
    case = make_case(72, 30, w=90, seed=13)
    d = str(tmp_path)
    dem, swe, alb, aws = _write_case(case, d)
    pot = I.insolation_series(case, shadow=True, dtype=np.float32)
    pick = os.path.join(d, "pickle", "10")
    os.makedirs(pick)
    for i, row in enumerate(case.aws_rows):
        np.save(os.path.join(pick, "%s_total.sdat.npy" % row["DATE"]), pot[i])
    e = Energy(dem, None, os.path.join(d, "out"), res=10, precision=precision)
    e.use_precomputed = True
    e.add_pickle_dir(os.path.join(d, "pickle"))
    e.add_snow(swe)
    e.add_checkpoints(["20220601"])
    e.max_resident_insolation_bytes = 8 * 72 * 90 * 4          # force several streamed chunks
    e.model(aws_file=aws, albedo_maps=alb, z=1.6, elev_aws=case.elev_aws, xy_aws=case.xy_aws, zm=1e-3,
            z_h_or_e=1e-4, emissivity=0.98, last_snowfall="20220525", max_ice_albedo=0.4, v=False)
    f64 = precision == "f64"
    ora = P.run_oracle(case, pot, f64, last_snowfall="20220525", max_ice_albedo=0.4)
    tol = 1e-9 if f64 else 1e-4
    assert P.max_rel_err(e.total_ice_melt_array, ora["total_ice"], 1e-6 if f64 else 1e-3) < (1e-6 if f64 else tol)
    assert P.max_rel_err(e.swe_array, ora["swe"], 1e-6 if f64 else 1e-3) < (1e-6 if f64 else tol)
    got = _csv_numbers(open(os.path.join(d, "out", "heat_fluxes.csv")).read())
    want = _csv_numbers(ora["stats_csv"])
    assert [g[0] for g in got] == [w[0] for w in want]
    for (_, a), (_, b) in zip(got, want):
        # printed with %.1f / %.2f / %.4f / %.0f: allow one unit in the last printed digit
        lim = [0.1001] * 8 + [0.0101] + [0.00011] * 3 + [1.001]
        assert all(abs(x - y) <= l for x, y, l in zip(a, b, lim)), (a, b)
    files = os.listdir(os.path.join(d, "out"))
    assert any("20220601 12:00:00 total_melt_ice" in f for f in files)        # checkpoint export
    assert any(case.aws_rows[-1]["DATE"] + " remaining_snow_cover" in f for f in files)
    assert os.path.isfile(os.path.join(d, "out", "solar_output.csv"))


def test_in_kernel_insolation_replaces_saga(tmp_path):
    """Without use_precomputed the reference would call SAGA per step; here the fused kernel
    computes insolation + shading (specification: oracle/insolation_oracle.py)."""
    case = make_case(64, 26, w=80, seed=17)
    d = str(tmp_path)
    dem, swe, alb, aws = _write_case(case, d)
    e = Energy(dem, None, os.path.join(d, "out"), res=10, precision="f64")
    e.lat, e.lon = case.lat, case.lon
    e.add_snow(swe)
    e.model(aws_file=aws, const_albedo=(0.35, 0.75), z=1.6, elev_aws=case.elev_aws, xy_aws=case.xy_aws,
            zm=1e-3, z_h_or_e=1e-4, emissivity=0.98, v=False)
    pot = I.insolation_series(case, shadow=True, dtype=np.float64)
    ora = P.run_oracle(case, pot, True, const_albedo=(0.35, 0.75))
    # state rasters come back as float32 (model.py:76-80)
    assert e.total_ice_melt_array.dtype == np.float32
    assert P.max_rel_err(e.total_ice_melt_array, ora["total_ice"], 1e-3) < 2e-7
    assert P.max_rel_err(e.total_snow_melt_array, ora["total_snow"], 1e-3) < 2e-7


def test_from_config(tmp_path):
    case = make_case(48, 12, w=64, seed=19)
    d = str(tmp_path)
    dem, swe, alb, aws = _write_case(case, d)
    cfg = {
        "input": {"dem": dem, "outlines": None,
                  "aws": {"file": aws, "elev": case.elev_aws, "xy": list(case.xy_aws), "sensor_z": 1.6},
                  "vertical_lapse_rates": {"t_air": -0.0065}},
        "output": {"out_dir": os.path.join(d, "out"), "resolution": 10, "verbose": False, "png_export": 720},
        "albedo": {"use_const": False, "last_snowfall": "20220522", "max_ice_albedo": 0.38, "albedo_maps": alb},
        "solar": {"use_precomputed": False},
        "turbo": {"zm": 0.001, "z_h_or_e": 0.0001, "andreas": False, "sensible_corr_factor": 1, "latent_corr_factor": 1},
        "longwave": {"emissivity": 0.98, "cloud_corr": 0.0},
        "snow": {"use": True, "density": 387.0, "swe_grid": swe},
        "msm": {"use": False},
    }
    e, kw = Energy.from_config(cfg, precision="f32")
    e.lat, e.lon = case.lat, case.lon
    e.model(**kw)
    assert e.stats.shape == (12, _lib_count())
    assert np.isfinite(np.nanmean(e.total_ice_melt_array))


def _lib_count():
    from enrgy_b200 import _lib
    return _lib.S_COUNT


def test_error_behaviour(tmp_path):
    case = make_case(32, 4, seed=23)
    d = str(tmp_path)
    dem, swe, alb, aws = _write_case(case, d)
    e = Energy(dem, None, os.path.join(d, "out"), res=10)
    with pytest.raises(ValueError):
        e.add_cloud_corr(1.5)                                   # model.py:91-92
    with pytest.raises(IOError):
        e.add_pickle_dir(os.path.join(d, "nope"))               # model.py:98-100
    bad = dict(alb)
    rows = case.aws_rows
    rows[0]["HUMID"] = "140"
    case.write_aws_csv(aws)
    with pytest.raises(ValueError):                             # helpers.py:87
        e.model(aws_file=aws, albedo_maps=bad, z=1.6, elev_aws=case.elev_aws, xy_aws=case.xy_aws, v=False)

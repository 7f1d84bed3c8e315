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


def test_insolation_cache_writer_roundtrip(tmp_path):
    """GPU-computed insolation written in the reference's pickle layout (model.py:477-481) and read
    back through the reference-shaped `use_precomputed` path reproduces the in-kernel run."""
    from enrgy_b200.insolation_pickler import pickle_insolation_series
    case = make_case(56, 14, w=72, seed=29)
    d = str(tmp_path)
    dem, swe, alb, aws = _write_case(case, d)
    files = pickle_insolation_series(case.dem, case.geotransform, case.aws_rows, os.path.join(d, "pickle"), 10,
                                     lat=case.lat, lon=case.lon, shadow=True, dtype=np.float64)
    assert len(files) == 14 and all(os.path.isfile(f) for f in files)
    want = I.potential_insolation(case.dem, case.cell, case.lat, case.lon, I.to_unix(case.aws_rows[9]["DATE"]), 3600)
    assert P.max_rel_err(np.load(files[9]), want, 1e-6) < 1e-9
    # the oracle (= the reference's arithmetic) on these files vs the fused in-kernel path
    pot = np.stack([np.load(f) for f in files])
    ora = P.run_oracle(case, pot, True)
    res = P.compare_run(case, True, computed=True, shadow=True)
    assert max(res.values()) < 1e-9
    e = Energy(dem, None, os.path.join(d, "out"), res=10, precision="f64")
    e.use_precomputed = True
    e.add_pickle_dir(os.path.join(d, "pickle"))
    e.add_snow(swe)
    e.model(aws_file=aws, albedo_maps=alb, z=1.6, elev_aws=case.elev_aws, xy_aws=case.xy_aws, zm=1e-3,
            z_h_or_e=1e-4, emissivity=0.98, v=False)
    # streamed rasters are float32 on the device (SAGA .sdat is float32): 1e-7 class agreement
    assert P.max_rel_err(e.total_ice_melt_array, ora["total_ice"], 1e-3) < 5e-7


def test_add_msm_drop_in(tmp_path):
    """add_msm + model(): the sub-surface model through the reference-shaped class."""
    case = make_case(48, 16, w=60, seed=37)
    d = str(tmp_path)
    dem, swe, alb, aws = _write_case(case, d)
    msm = dict(depths=[0.1, 0.1, 0.3, 0.5, 0.5, 0.5, 3.0],
               temperatures=[-6.9, -6.93, -7.025, -7.31, -6.93, -7.12, -7.0, -5.57], elev=275.0)
    e = Energy(dem, None, os.path.join(d, "out"), res=10, precision="f64")
    e.lat, e.lon = case.lat, case.lon
    e.add_snow(swe)
    e.add_msm(msm["depths"], msm["temperatures"], msm["elev"])
    assert len(e.layer_temperatures) == 8 and float(np.nanmax(e.layer_temperatures[0])) <= 0.0
    e.model(aws_file=aws, albedo_maps=alb, z=1.6, elev_aws=case.elev_aws, xy_aws=case.xy_aws, zm=1e-3,
            z_h_or_e=1e-4, emissivity=0.98, v=False)
    pot = I.insolation_series(case, shadow=True, dtype=np.float64)
    ora = P.run_oracle(case, pot, True, msm=msm)
    assert P.max_rel_err(e.total_ice_melt_array, ora["total_ice"], 1e-3) < 2e-7
    assert P.max_rel_err(e.layer_temperatures[0], ora["layer_temperatures"][0], 1e-2) < 1e-5
    rows = _csv_numbers(open(os.path.join(d, "out", "heat_fluxes.csv")).read())
    want = _csv_numbers(ora["stats_csv"])
    for (_, a), (_, b) in zip(rows, want):
        assert abs(a[6] - b[6]) <= 0.1001 and abs(a[8] - b[8]) <= 0.0101      # in-glacier flux, POINT_T_SURF

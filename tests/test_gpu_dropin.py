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


def test_from_config_against_the_oracle(tmp_path):
    """config_template.json layout -> Energy + model() keyword arguments; the run against the oracle
    given the same knobs by hand (GRADIENT column as the lapse rate, snow ageing, ice-albedo cap,
    correction factors, cloud correction, snow density, checkpoints, stakes)."""
    import json
    case = make_case(48, 30, w=64, seed=19, with_gradient=True)
    d = str(tmp_path)
    dem, swe, alb, aws = _write_case(case, d)
    stakes = os.path.join(d, "stakes.csv")
    pts = [(10, 20), (24, 32), (40, 50), (0, 0)]                    # the last one is off-glacier (NaN)
    with open(stakes, "w") as f:
        f.write("name,easting,northing\n")
        for k, (r, c) in enumerate(pts):
            f.write("s%d,%.1f,%.1f\n" % (k, case.geotransform[0] + (c + 0.5) * case.cell,
                                        case.geotransform[3] - (r + 0.5) * case.cell))
    cfg = {
        "input": {"dem": dem, "outlines": None,
                  "aws": {"file": aws, "elev": case.elev_aws, "xy": list(case.xy_aws), "sensor_z": 1.6},
                  "vertical_lapse_rates": {"t_air": "GRADIENT"}},
        "output": {"out_dir": os.path.join(d, "out"), "resolution": 10, "verbose": False, "png_export": 720,
                   "dates": ["20220601"], "stake_coords": stakes, "debug_point_output": "point.csv"},
        "albedo": {"use_const": False, "last_snowfall": "20220522", "max_ice_albedo": 0.38, "albedo_maps": alb},
        "solar": {"use_precomputed": False},
        "turbo": {"zm": 0.001, "z_h_or_e": 0.0001, "andreas": True, "sensible_corr_factor": 1.1, "latent_corr_factor": 0.9},
        "longwave": {"emissivity": 0.97, "cloud_corr": 0.1},
        "snow": {"use": True, "density": 350.0, "swe_grid": swe},
        "msm": {"use": False},
    }
    cfg_path = os.path.join(d, "config.json")
    with open(cfg_path, "w") as f:
        json.dump(cfg, f)
    from enrgy_b200.model import PARAMS
    saved = dict(PARAMS)               # set_density changes the module global, as in the reference (model.py:84-88)
    try:
        e, kw = Energy.from_config(cfg_path, precision="f64")
        e.lat, e.lon = case.lat, case.lon
        e.model(**kw)
    finally:
        PARAMS.clear()
        PARAMS.update(saved)
    pot = I.insolation_series(case, shadow=True, dtype=np.float64)
    okw = dict(temp_lapse_rate="GRADIENT", last_snowfall="20220522", max_ice_albedo=0.38, andreas=True,
               sensible_corr=1.1, latent_corr=0.9, emissivity=0.97, cloud_corr=0.1, snow_density=350.0)
    ora = P.run_oracle(case, pot, True, **okw)
    assert e.stats.shape == (30, _lib_count())
    assert P.max_rel_err(e.total_ice_melt_array, ora["total_ice"], 1e-3) < 2e-7
    assert P.max_rel_err(e.total_snow_melt_array, ora["total_snow"], 1e-3) < 2e-7
    assert P.max_rel_err(e.swe_array, ora["swe"], 1e-3) < 2e-7
    got, want = _csv_numbers(open(os.path.join(d, "out", "heat_fluxes.csv")).read()), _csv_numbers(ora["stats_csv"])
    assert [g[0] for g in got] == [w[0] for w in want]
    lim = [0.1001] * 8 + [0.0101] + [0.00011] * 3 + [1.001]
    for (_, a), (_, b) in zip(got, want):
        assert all(abs(x - y) <= l for x, y, l in zip(a, b, lim)), (a, b)
    # debug_point_output (model.py:170-180, :441-448): header + "DATE,sensible,latent" at the AWS cell per row
    lines = open(os.path.join(d, "out", "point.csv")).read().split("\n")
    assert lines[0] == "SENSIBLE,LATENT" and len(lines) == 31
    r, c = case.aws_rc
    for i, line in enumerate(lines[1:]):
        date, sens, lat = line.split(",")
        assert date == case.aws_rows[i]["DATE"]
        assert abs(float(sens) - float(ora["rows"][i]["sens"][r, c])) <= 0.0501
        assert abs(float(lat) - float(ora["rows"][i]["lat"][r, c])) <= 0.0501
    # stakes sampled at the checkpoint row (model.py:102-120, :279-283): ice melt after the first 13 rows
    k = [row["DATE"] for row in case.aws_rows].index("20220601 12:00:00") + 1
    import dataclasses
    part = dataclasses.replace(case, aws_rows=case.aws_rows[:k + 1])     # (one more row: same time step for row k-1)
    out = open(os.path.join(d, "out", "ice_melt_point.csv")).read().strip().split("\n")
    assert out[0] == "name,20220601 12:00:00" and len(out) == 5
    ora_full = P.run_oracle(part, pot[:k + 1], True, **okw)
    ice_k = np.asarray(ora_full["total_ice"], dtype=np.float64) - np.asarray(ora_full["melt"][k][1], dtype=np.float64)
    for line, (rr, cc) in zip(out[1:], pts):
        name, val = line.split(",")
        if np.isnan(case.dem[rr, cc]):
            assert val == ""                                          # NaN -> empty field, as pandas writes it
        else:
            assert abs(float(val) - ice_k[rr, cc]) <= 0.0011


def test_views_left_behind_and_repeated_calls(tmp_path):
    """After model(): `aws`, `vars` (DistributedVars of the last row), `albedo`, `incoming_shortwave`
    (model.py:229-236, :408); a second model() call continues from the state of the first
    (model.py:258-261 accumulates in place)."""
    import dataclasses
    case = make_case(56, 24, w=72, seed=43)
    d = str(tmp_path)
    dem, swe, alb, aws = _write_case(case, d)
    first, second = dataclasses.replace(case, aws_rows=case.aws_rows[:12]), dataclasses.replace(case, aws_rows=case.aws_rows[12:])
    aws1, aws2 = first.write_aws_csv(os.path.join(d, "aws1.csv")), second.write_aws_csv(os.path.join(d, "aws2.csv"))
    kw = dict(albedo_maps=alb, z=1.6, elev_aws=case.elev_aws, xy_aws=case.xy_aws, zm=1e-3, z_h_or_e=1e-4,
              emissivity=0.98, v=False)
    e = Energy(dem, None, os.path.join(d, "out"), res=10, precision="f64")
    e.lat, e.lon = case.lat, case.lon
    e.add_snow(swe)
    e.model(aws_file=aws1, **kw)
    pot = I.insolation_series(case, shadow=True, dtype=np.float64)
    o1 = P.run_oracle(first, pot[:12], True)
    valid = ~np.isnan(case.dem)
    # views of the last row of the first call
    last = o1["rows"][-1]
    assert np.allclose(e.albedo[valid], np.asarray(last["albedo"])[valid], rtol=1e-9)
    inc = np.asarray(last["rs"])[valid] / (1.0 - np.asarray(last["albedo"])[valid])
    assert np.allclose(e.incoming_shortwave[valid], inc, rtol=1e-9, atol=1e-9)
    row = first.aws_rows[-1]
    assert e.aws.Tz == float(row["T_AIR"]) + 273.15 and e.aws.P == float(row["PRESSURE"]) * 100
    t_air = float(row["T_AIR"]) + (case.dem.astype(np.float64) - case.elev_aws) * -0.006
    assert np.allclose(e.vars.t_air[valid], t_air[valid]) and np.allclose(e.vars.Tz[valid], t_air[valid] + 273.15)
    assert np.allclose(e.vars.pressure[valid], (float(row["PRESSURE"]) + (case.dem - case.elev_aws) * -0.1145)[valid])
    assert e.vars.rel_humidity.shape == case.dem.shape and e.vars.wind_speed.dtype == np.float32
    # second call: starts from the first call's SWE, melt totals keep growing
    e.model(aws_file=aws2, **kw)
    second_case = dataclasses.replace(second, swe=np.asarray(o1["swe"], dtype=np.float32))
    o2 = P.run_oracle(second_case, pot[12:], True)
    tot_ice = np.asarray(o1["total_ice"], dtype=np.float64) + np.asarray(o2["total_ice"], dtype=np.float64)
    tot_snow = np.asarray(o1["total_snow"], dtype=np.float64) + np.asarray(o2["total_snow"], dtype=np.float64)
    assert P.max_rel_err(e.total_ice_melt_array, tot_ice, 1e-3) < 5e-7
    assert P.max_rel_err(e.total_snow_melt_array, tot_snow, 1e-3) < 5e-7
    assert P.max_rel_err(e.swe_array, o2["swe"], 1e-3) < 5e-7


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


def test_extra_stations_through_the_energy_class(tmp_path):
    """add_station / add_cloud_transmissivity (BASELINE config C4) on files: the station positions given in
    real-world coordinates arrive in the kernel as the cell-unit positions the specification uses."""
    import csv as _csv
    from enrgy_b200.synthetic import make_station_rows
    case = make_case(64, 20, w=80, seed=23)
    d = str(tmp_path)
    dem, swe, alb, aws = _write_case(case, d)
    pot = I.insolation_series(case, shadow=False, dtype=np.float64)
    pick = os.path.join(d, "pickle", "10")
    os.makedirs(pick)
    for i, row in enumerate(case.aws_rows):
        np.save(os.path.join(pick, "%s_total.sdat.npy" % row["DATE"]), pot[i])
    spots = [(12.0, 60.0, 110.0, 31), (50.0, 15.0, -70.0, 32)]
    stations = []
    e = Energy(dem, None, os.path.join(d, "out"), res=10, precision="f64")
    e.use_precomputed = True
    e.add_pickle_dir(os.path.join(d, "pickle"))
    e.add_snow(swe)
    e.add_cloud_corr(0.05)
    ul_x, x_dist, _, ul_y, _, y_dist = case.geotransform
    for k, (r, c, dz, sd) in enumerate(spots):
        rows = make_station_rows(case, case.elev_aws + dz, seed=sd)
        path = os.path.join(d, "station%d.csv" % k)
        with open(path, "w", newline="") as f:
            wr = _csv.DictWriter(f, fieldnames=list(rows[0]))
            wr.writeheader()
            wr.writerows(rows)
        e.add_station(path, (ul_x + (c + 0.5) * x_dist, ul_y + (r + 0.5) * y_dist), case.elev_aws + dz)
        stations.append(dict(row=r, col=c, elev=case.elev_aws + dz, rows=rows))
    e.add_cloud_transmissivity(0.6)
    e.model(aws_file=aws, albedo_maps=alb, z=1.6, elev_aws=case.elev_aws, xy_aws=case.xy_aws, zm=1e-3,
            z_h_or_e=1e-4, emissivity=0.98, v=False)
    ora = P.run_oracle(case, pot, True, stations=stations, cloud_k=0.6, cloud_corr=0.05)
    # (the class keeps its state rasters in float32 like the reference, model.py:76-80)
    assert P.max_rel_err(e.total_ice_melt_array, ora["total_ice"], 1e-3) < 1e-6
    assert P.max_rel_err(e.swe_array, ora["swe"], 1e-3) < 1e-6
    plain = P.run_oracle(case, pot, True, cloud_corr=0.05)
    assert P.max_rel_err(e.total_ice_melt_array, plain["total_ice"], 1e-3) > 1e-3     # the stations really act

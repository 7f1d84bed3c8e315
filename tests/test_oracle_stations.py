"""The station-blend / cloud-transmissivity specification (BASELINE config C4; oracle/enrgy_oracle.py,
"several weather stations") against the pinned single-station oracle: with no extra station it must BE
the reference's arithmetic, bit for bit, in both dtype configurations."""
import copy

import numpy as np
import pytest

from enrgy_b200.synthetic import make_case, make_station_rows
from tests import parity as P
from oracle import enrgy_oracle as O


def _run(case, pot, f64, **kw):
    dt = np.float64 if f64 else np.float32
    cfg = P.oracle_config(case, last_snowfall="20220525")
    return O.run_model(case.dem.astype(dt), case.geotransform, case.aws_rows, np.asarray(pot, dtype=dt), cfg,
                       swe=case.swe.astype(dt), albedo_arrays=P.clipped_albedo(case, dt), state_dtype=dt,
                       want_means=True, **kw)


@pytest.mark.parametrize("f64", [False, True])
def test_one_station_is_the_reference_bit_for_bit(f64):
    case = make_case(40, 12, w=56, seed=3)
    pot = P.random_insolation(case, 12)
    ref = _run(case, pot, f64)
    got = _run(case, pot, f64, stations=[], cloud_k=0.8)
    for k in ("swe", "total_snow", "total_ice"):
        assert np.array_equal(ref[k], got[k], equal_nan=True), k
    assert ref["stats_csv"] == got["stats_csv"]
    for a, b in zip(ref["rows"], got["rows"]):
        for fld in P.FLUX_FIELDS:
            assert a[fld].dtype == b[fld].dtype and np.array_equal(a[fld], b[fld], equal_nan=True), fld


def test_station_blend_properties():
    case = make_case(40, 12, w=56, seed=3)
    pot = P.random_insolation(case, 12)
    ref = _run(case, pot, True)
    # a second station that reports exactly what the lapse rates predict from the first changes nothing
    # to temperature and pressure beyond rounding (pressure -0.1145 hPa/m, temperature with the run's lapse
    # rate, same cloudiness); its humidity is not consistent (the reference scales vapour pressure with
    # elevation, the twin repeats the relative humidity), which moves the melt by a few per cent at most
    twin = make_station_rows(case, elev=case.elev_aws + 120.0, seed=None)
    st = [dict(row=5.0, col=50.0, elev=case.elev_aws + 120.0, rows=twin)]
    got = _run(case, pot, True, stations=st, cloud_k=0.8)
    assert P.max_rel_err(got["total_ice"], ref["total_ice"], 1e-3) < 5e-2
    # a warmer, cloudier station next to the glacier tongue: more melt near it, less shortwave under its cloud
    warm = copy.deepcopy(twin)
    for r in warm:
        r["T_AIR"] = "%.3f" % (float(r["T_AIR"]) + 3.0)
        r["CLOUDINESS"] = "1.0"
    st2 = [dict(row=35.0, col=10.0, elev=case.elev_aws + 120.0, rows=warm)]
    a = _run(case, pot, True, stations=st2, cloud_k=None)
    b = _run(case, pot, True, stations=st2, cloud_k=1.2)
    m = ~np.isnan(ref["total_ice"])
    assert np.nansum(a["total_ice"]) > np.nansum(ref["total_ice"])
    assert np.nansum(b["total_ice"]) < np.nansum(a["total_ice"])
    # the influence falls off with distance from the station
    d_near = np.abs(a["total_ice"] - ref["total_ice"])[30:, :20][m[30:, :20]].mean()
    d_far = np.abs(a["total_ice"] - ref["total_ice"])[:10, 36:][m[:10, 36:]].mean()
    assert d_near > 3 * d_far
    # weights: non-negative, sum to one, largest at the station's own cell
    w, D, V = O.station_fields(case.dem.astype(np.float64), P.oracle_config(case), case.geotransform, st2)
    assert np.allclose(w[0] + w[1], 1.0, atol=1e-15) and (w[0] >= 0).all() and (w[1] >= 0).all()
    assert np.unravel_index(np.argmax(w[1]), w[1].shape) == (35, 10)


@pytest.mark.parametrize("f64", [False, True])
def test_station_specification_is_frozen(f64):
    """tests/golden_spec/stations_*.npz were written by the oracle itself (make_spec_golden.py): the
    specification must not move under later edits."""
    import json
    import os
    import sys
    here = os.path.join(os.path.dirname(__file__), "golden_spec")
    g = np.load(os.path.join(here, "stations_%s.npz" % ("f64" if f64 else "f32")), allow_pickle=False)
    if str(g["numpy"]).split(".")[0] != np.__version__.split(".")[0]:
        pytest.skip("fixture made with numpy %s" % g["numpy"])
    sys.path.insert(0, here)
    import make_spec_golden as M
    rec = json.loads(str(g["recipe"]))
    case = make_case(rec["n"], rec["n_steps"], seed=rec["seed"], w=rec["w"], calm_every=rec["calm_every"])
    pot = P.random_insolation(case, rec["n_steps"])
    r = P.run_oracle(case, pot, f64, stations=M.stations_of(case), cloud_k=0.8, cloud_corr=0.1, last_snowfall="20220525")
    for k in ("swe", "total_snow", "total_ice"):
        assert np.array_equal(r[k], g[k], equal_nan=True), k
    for key in g.files:
        if key.startswith("step"):
            step, name = key.split("_", 1)
            assert np.array_equal(r["rows"][int(step[4:])][name], g[key], equal_nan=True), key
    assert r["stats_csv"] == str(g["stats_csv"])

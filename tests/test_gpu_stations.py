"""Several weather stations + cloud attenuation of the shortwave (BASELINE config C4) through the C ABI
against the specification in oracle/enrgy_oracle.py ("several weather stations"; ours -- the reference has
one AWS -- and bit-identical to the reference with one station, tests/test_oracle_stations.py)."""
import numpy as np
import pytest

from enrgy_b200.synthetic import make_case, make_station_rows
from tests import parity as P

pytestmark = pytest.mark.gpu


def _stations(case, n):
    h, w = case.shape
    spots = [(0.18 * h, 0.30 * w, 140.0, 11), (0.80 * h + 0.37, 0.62 * w + 0.5, -90.0, 12), (0.45 * h, 0.85 * w, 60.0, 13)]
    return [dict(row=r, col=c, elev=case.elev_aws + dz, rows=make_station_rows(case, case.elev_aws + dz, seed=sd))
            for (r, c, dz, sd) in spots[:n]]


@pytest.mark.parametrize("f64", [True, False])
@pytest.mark.parametrize("n_extra,cloud_k", [(3, 0.9), (1, None), (2, 0.4)])
def test_station_blend_against_the_oracle(f64, n_extra, cloud_k):
    """Per-step flux rasters, melt, final state and the area means: 1e-9 (float64) / 1e-4 (float32)."""
    case = make_case(56, 30, w=88, seed=17)
    pot = P.random_insolation(case, 30)
    res = P.compare_run(case, f64, pot=pot, stations=_stations(case, n_extra), cloud_k=cloud_k,
                        last_snowfall="20220525", cloud_corr=0.1)
    tol = 1e-9 if f64 else 1e-4
    bad = {k: v for k, v in res.items() if v >= tol and not k.endswith("_l2")}
    assert not bad, bad


@pytest.mark.parametrize("f64", [True, False])
def test_station_blend_with_computed_insolation_and_shading(f64):
    case = make_case(64, 16, w=96, seed=21)
    res = P.compare_run(case, f64, computed=True, shadow=True, stations=_stations(case, 2), cloud_k=0.7)
    tol = 1e-9 if f64 else 1e-4
    bad = {k: v for k, v in res.items() if v >= tol and not k.endswith("_l2")}
    assert not bad, bad


@pytest.mark.parametrize("f64", [True, False])
@pytest.mark.parametrize("mode", ["streamed", "computed", "shadow"])
def test_blend_with_the_primary_station_alone_equals_a_plain_run(f64, mode):
    """No extra station: the blend kernel (weights, folded vapour-pressure factors, exp(0) cloud factor,
    half the cells per thread) must give the plain kernel's rasters: bit for bit in float32, to a few ulp in
    float64."""
    case = make_case(72, 26, w=150, seed=5)
    pot = P.random_insolation(case, 26) if mode == "streamed" else None
    kw = dict(computed=mode != "streamed", shadow=mode == "shadow")
    out = []
    for st in (None, []):
        eng = P.make_engine(case, f64, pot=pot, stations=st, cloud_k=0.8, **kw)
        try:
            stats = eng.run(0, 26)
            out.append((eng.state(np.float64), stats))
        finally:
            eng.close()
    for a, b in zip(out[0][0], out[1][0]):
        if f64:
            # (float64 takes its reciprocals by Newton steps from a hardware seed: not correctly rounded, and
            # the two instantiations are scheduled differently -- agreement to a few ulp, not bit for bit)
            assert np.array_equal(np.isnan(a), np.isnan(b)) and np.allclose(a, b, rtol=1e-13, atol=1e-15, equal_nan=True)
        else:
            assert np.array_equal(a, b, equal_nan=True)
    # (the longwave area sum comes from the kernel instead of the DEM moments; float32 sums in another order)
    assert np.allclose(out[0][1], out[1][1], rtol=1e-11 if f64 else 3e-6, atol=1e-8 if f64 else 1e-2)


@pytest.mark.parametrize("mode", ["streamed", "shadow"])
def test_station_blend_on_row_bands(mode):
    """Station positions are cells of the FULL raster: two row bands give the whole run's rasters bit for bit
    (float32), and their area sums add up."""
    case = make_case(80, 20, w=96, seed=9)
    pot = P.random_insolation(case, 20) if mode == "streamed" else None
    kw = dict(computed=mode != "streamed", shadow=mode == "shadow", stations=_stations(case, 3), cloud_k=0.8)
    whole = P.make_engine(case, False, pot=pot, **kw)
    try:
        s_whole = whole.run(0, 20)
        st_whole = whole.state(np.float32)
    finally:
        whole.close()
    total = np.zeros_like(s_whole)
    for band in ((0, 48), (48, 32)):
        eng = P.make_engine(case, False, pot=pot, band=band, **kw)
        try:
            total += eng.run(0, 20)
            part = eng.state(np.float32)
        finally:
            eng.close()
        for a, b in zip(part, st_whole):
            assert np.array_equal(a, b[band[0]:band[0] + band[1]], equal_nan=True)
    from enrgy_b200 import _lib
    cols = [_lib.S_RS, _lib.S_LWD, _lib.S_SENS, _lib.S_LAT, _lib.S_MELT, _lib.S_SNOW, _lib.S_ICE]
    assert np.allclose(total[1:, cols], s_whole[1:, cols], rtol=3e-6, atol=1e-3)


def test_station_errors():
    from enrgy_b200._lib import EnrgyError
    case = make_case(40, 6, w=56, seed=3)
    eng = P.make_engine(case, True, pot=P.random_insolation(case, 6))
    try:
        ser = np.zeros((4, 6, 4))
        with pytest.raises(EnrgyError):
            eng.set_stations([(1, 1, 300.0)] * 4, ser, None)               # more than 3 extra stations
        with pytest.raises(EnrgyError):
            eng.set_stations([(1, 1, 300.0)], np.full((1, 6, 4), np.nan), None)
        with pytest.raises(EnrgyError):
            eng.set_stations([], None, -1.0)                               # negative cloud_k
        eng.set_stations([], None, 0.5)
        with pytest.raises(EnrgyError):
            eng.run_members([0.0, 0.1])                                    # not with the blend
        eng.set_stations(None, None)
        eng.prepass()
        assert np.isfinite(eng.run(0, 6)).all()
    finally:
        eng.close()

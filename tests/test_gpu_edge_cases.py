"""Edge cases of the C ABI path: row bands vs whole raster, split runs and restarts, empty work,
ragged raster shapes, error codes."""
import numpy as np
import pytest

from enrgy_b200 import _lib
from enrgy_b200.engine import Engine
from enrgy_b200.forcing import build_forcing
from enrgy_b200.parallel import row_bands
from enrgy_b200.synthetic import make_case
from tests import parity as P

pytestmark = pytest.mark.gpu


def _band_engine(case, r0, n, f64, pot=None, computed=True, shadow=True):
    h, w = case.shape
    eng = Engine(h, w, precision=_lib.F64 if f64 else _lib.F32)
    eng.set_params(cell_size=case.cell, elev_aws=case.elev_aws, aws_row=case.aws_rc[0], aws_col=case.aws_rc[1],
                   sensor_z=1.6, zm=1e-3, z_h_or_e=1e-4, emissivity=0.98,
                   insol_mode=_lib.INSOL_COMPUTED if computed else _lib.INSOL_STREAMED, shadow=shadow,
                   lat=case.lat, lon=case.lon, band_row0=r0, band_rows=n)
    eng.set_dem(case.dem)                                  # the FULL DEM on every band
    alb = P.clipped_albedo(case, np.float32)
    keys = list(alb)
    eng.set_albedo_maps([alb[k][r0:r0 + n] for k in keys])
    eng.set_swe(case.swe[r0:r0 + n])
    eng.set_forcing(build_forcing(case.aws_rows, keys))
    eng.prepass()
    return eng


@pytest.mark.parametrize("f64", [True, False])
def test_row_bands_equal_whole_raster(f64):
    """Three bands (unequal, cut by glacier-cell count) with shading rays crossing the cuts give the
    whole-raster result: rasters bit-identical, statistics equal after summing."""
    case = make_case(112, 20, w=90, seed=51)
    n = 20
    whole = _band_engine(case, 0, 112, f64)
    stats_w = whole.run(0, n)
    state_w = whole.state(np.float64)
    whole.close()
    bands = row_bands(112, 3, align=16, valid_per_row=(~np.isnan(case.dem)).sum(axis=1))
    assert len({b[1] for b in bands}) > 1
    total = np.zeros_like(stats_w)
    for (r0, rows) in bands:
        eng = _band_engine(case, r0, rows, f64)
        total += eng.run(0, n)
        part = eng.state(np.float64)
        eng.close()
        for a, b in zip(part, state_w):
            assert np.array_equal(a, b[r0:r0 + rows], equal_nan=True)
    # step 0 uses the first-row SWE quirk per band; all columns are sums/counts
    assert np.allclose(total, stats_w, rtol=1e-12 if f64 else 1e-6, atol=1e-9)


def test_split_runs_snapshot_and_restart():
    case = make_case(64, 18, w=70, seed=53)
    pot = P.random_insolation(case, 18)
    one = P.make_engine(case, True, pot=pot)
    s_all = one.run(0, 18)
    ref = one.state(np.float64)
    one.close()
    eng = P.make_engine(case, True, pot=pot)
    a = eng.run(0, 5)
    eng.snapshot(save=True)
    b = eng.run(5, 11)
    mid = eng.state(np.float64)
    c = eng.run(11, 18)
    got = eng.state(np.float64)
    assert np.array_equal(np.concatenate([a, b, c]), s_all)            # deterministic, split-invariant
    # swe and total_ice are carried step by step: bit-identical; total_snow is added once per run as
    # swe(start) - swe(end), so a split run rounds it differently in the last bit
    assert np.array_equal(got[0], ref[0], equal_nan=True) and np.array_equal(got[2], ref[2], equal_nan=True)
    assert np.allclose(got[1], ref[1], rtol=1e-13, atol=0, equal_nan=True)
    eng.snapshot(save=False)                                            # rewind to step 5
    b2 = eng.run(5, 11)
    assert np.array_equal(b, b2)
    eng.close()
    # restart from exported rasters (the reference's manual restart via add_snow, model.py:122-124)
    eng2 = P.make_engine(case, True, pot=pot)
    eng2.set_state(*mid)
    c2 = eng2.run(11, 18)
    got2 = eng2.state(np.float64)
    eng2.close()
    assert np.array_equal(c2[:, :_lib.S_NSWE], c[:, :_lib.S_NSWE])
    assert np.array_equal(got2[0], ref[0], equal_nan=True) and np.array_equal(got2[2], ref[2], equal_nan=True)
    assert np.allclose(got2[1], ref[1], rtol=1e-13, atol=0, equal_nan=True)


def test_empty_work_and_all_nan():
    case = make_case(32, 6, w=40, seed=55)
    pot = P.random_insolation(case, 6)
    eng = P.make_engine(case, False, pot=pot)
    assert eng.run(3, 3).shape == (0, _lib.S_COUNT)                     # zero steps
    swe, tsn, tic = eng.state()
    assert np.array_equal(np.isnan(swe), np.isnan(case.swe)) and np.all(tic[~np.isnan(case.dem)] == 0)
    eng.close()
    # a band without a single glacier cell: nothing to launch, NaN state, zero counts
    band = Engine(32, 40, precision=_lib.F32)
    band.set_params(cell_size=10.0, elev_aws=case.elev_aws, aws_row=case.aws_rc[0], aws_col=case.aws_rc[1],
                    const_albedo=(0.3, 0.7), insol_mode=_lib.INSOL_COMPUTED, lat=case.lat, lon=case.lon,
                    band_row0=0, band_rows=1)
    dem = case.dem.copy()
    dem[0, :] = np.nan
    band.set_dem(dem)
    band.set_forcing(build_forcing(case.aws_rows, None))
    band.prepass()
    st = band.run(0, 6)
    assert np.all(st[:, _lib.S_NVALID] == 0) and np.all(st[1:, _lib.S_RS] == 0)
    assert np.all(np.isnan(band.state()[0]))
    band.close()


@pytest.mark.parametrize("shape", [(1, 1), (3, 129), (17, 5), (130, 257)])
def test_ragged_shapes(shape):
    h, w = shape
    rng = np.random.default_rng(h * 1000 + w)
    dem = (300 + 40 * rng.random((h, w))).astype(np.float32)
    if h * w > 4:
        dem[rng.random((h, w)) < 0.2] = np.nan
    dem[h // 2, w // 2] = 320.0
    case = make_case(8, 5, seed=1, glacier_mask=False)
    eng = Engine(h, w, precision=_lib.F64)
    eng.set_params(cell_size=10.0, elev_aws=320.0, aws_row=h // 2, aws_col=w // 2, sensor_z=1.6, zm=1e-3,
                   z_h_or_e=1e-4, emissivity=0.98, const_albedo=(0.35, 0.75), insol_mode=_lib.INSOL_COMPUTED,
                   shadow=True, lat=78.0, lon=14.0)
    eng.set_dem(dem)
    eng.set_forcing(build_forcing(case.aws_rows, None))
    eng.prepass()
    st = eng.run(0, 5)
    swe, tsn, tic = eng.state(np.float64)
    eng.close()
    valid = ~np.isnan(dem)
    assert st[0, _lib.S_NVALID] == valid.sum()
    assert np.array_equal(np.isnan(tic), ~valid) and np.all(tic[valid] >= 0)
    assert np.isfinite(st[:, :_lib.S_NSNOW]).all()


def test_error_codes():
    lib = _lib.load()
    with pytest.raises(_lib.EnrgyError) as e:
        Engine(0, 10)
    assert e.value.code == _lib.ERR_ARG
    eng = Engine(16, 16)
    with pytest.raises(_lib.EnrgyError):                                # call order
        eng.set_dem(np.zeros((16, 16), np.float32))
    eng.set_params(cell_size=10.0, elev_aws=0.0, aws_row=8, aws_col=8)
    dem = np.full((16, 16), 100.0, np.float32)
    eng.set_dem(dem)
    alb = np.full((16, 16), 0.4, np.float32)
    alb[3, 3] = np.nan                                                  # NaN on a glacier cell
    with pytest.raises(_lib.EnrgyError) as e:
        eng.set_albedo_maps([alb])
    assert e.value.code == _lib.ERR_MASK
    tab = np.zeros((2, _lib.F_COUNT))
    tab[:, _lib.F_DT] = 3600
    tab[:, _lib.F_RH] = 1.4                                             # not a 0..1 fraction
    with pytest.raises(_lib.EnrgyError) as e:
        eng.set_forcing(tab)
    assert e.value.code == _lib.ERR_RANGE
    tab[:, _lib.F_RH] = 0.8
    tab[:, _lib.F_DT] = 0                                               # the one-row CSV of the reference
    with pytest.raises(_lib.EnrgyError) as e:
        eng.set_forcing(tab)
    assert e.value.code == _lib.ERR_RANGE
    eng.close()
    assert lib.enrgy_destroy(None) == 0


def test_streamed_band_without_the_aws_cell_needs_its_insolation():
    """A row band that does not hold the AWS cell cannot form the shortwave factor from its own rasters: the
    pre-pass fails loudly until the AWS-cell values are handed in (enrgy_set_insolation_aws), and then the
    band gives the whole run's rasters bit for bit."""
    from enrgy_b200._lib import EnrgyError
    case = make_case(64, 10, w=72, seed=7)
    pot = P.random_insolation(case, 10)
    whole = P.make_engine(case, False, pot=pot)
    try:
        whole.run(0, 10)
        ref = whole.state(np.float32)
    finally:
        whole.close()
    band = (0, 16)                                   # the AWS cell sits in row 32
    eng = P.make_engine(case, False, pot=pot, band=band)
    try:
        eng.run(0, 10)
        got = eng.state(np.float32)
        for a, b in zip(got, ref):
            assert np.array_equal(a, b[:16], equal_nan=True)
        eng.set_insolation_aws(0, np.full(10, np.nan))
        with pytest.raises(EnrgyError):
            eng.prepass()
    finally:
        eng.close()

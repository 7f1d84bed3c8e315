"""Shading masks: bit-exact against the NumPy statement of the specification
(oracle/insolation_oracle.py).  Upstream parity of this part is UNPINNED (SAGA GIS is an external
binary); the masks are graded against this repo's own specification."""
import numpy as np
import pytest

from enrgy_b200 import _lib
from enrgy_b200.synthetic import make_case
from oracle import insolation_oracle as I
from oracle.enrgy_oracle import time_step_seconds
from tests import parity as P

pytestmark = pytest.mark.gpu


def _engine(case, f64=False):
    return P.make_engine(case, f64, computed=True, shadow=True)


@pytest.mark.parametrize("shape", [(96, 160), (130, 70), (64, 300)])
def test_masks_and_tables_bit_exact(shape):
    case = make_case(shape[0], 24, w=shape[1], seed=shape[0])
    eng = _engine(case)
    try:
        n_checked = 0
        for step in (0, 5, 11, 17, 23):
            row = case.aws_rows[step]
            dt = time_step_seconds(case.aws_rows, step)
            table = I.substep_table(I.to_unix(row["DATE"]), dt, case.lat, case.lon, case.cell)
            subs = eng.substeps(step)
            assert len(subs) == len(table)
            for got, want in zip(subs, table):
                # same libm, same expression order: the tables agree to the last bit
                assert got[0] == want["E"] and got[1] == want["N"] and got[2] == want["U"]
                assert got[3] == want["B"] and got[4] == want["D"]
                assert int(got[5]) == want["dc_fix"] and int(got[6]) == want["dr_fix"]
                assert np.float32(got[7]) == want["dz"]
            masks = eng.shade_masks(step)
            assert masks.shape[0] == len(table)
            valid = ~np.isnan(case.dem)
            for j, sub in enumerate(table):
                lit = I.shadow_mask(case.dem, sub["dc_fix"], sub["dr_fix"], sub["dz"])
                assert np.array_equal(masks[j][valid], lit[valid]), (step, j)
                n_checked += int(valid.sum())
        assert n_checked > 0
    finally:
        eng.close()


@pytest.mark.parametrize("f64", [False, True])
def test_potential_insolation_raster(f64):
    case = make_case(100, 30, w=140, seed=3)
    eng = _engine(case, f64)
    try:
        for step in (2, 13, 29):
            row = case.aws_rows[step]
            want = I.potential_insolation(case.dem, case.cell, case.lat, case.lon, I.to_unix(row["DATE"]),
                                          time_step_seconds(case.aws_rows, step), shadow=True)
            got = eng.potential_insolation(step)
            err = P.max_rel_err(got, want, 1e-6)
            assert err < (1e-9 if f64 else 1e-4), (step, err)
    finally:
        eng.close()


def test_masks_independent_of_precision():
    case = make_case(80, 12, w=96, seed=8)
    e32, e64 = _engine(case, False), _engine(case, True)
    try:
        for step in (3, 9):
            assert np.array_equal(e32.shade_masks(step), e64.shade_masks(step))
    finally:
        e32.close()
        e64.close()


def test_large_raster_spot_check():
    """2048 x 1024: the full NumPy mask is too slow, so 20000 random cells are ray-traced one by one
    with the same specification (oracle trace_cells)."""
    case = make_case(2048, 14, w=1024, seed=5)
    eng = _engine(case)
    try:
        step = 1                                         # 01:00 UTC: sun ~10 deg above the northern horizon
        table = I.substep_table(I.to_unix(case.aws_rows[step]["DATE"]), 3600, case.lat, case.lon, case.cell)
        masks = eng.shade_masks(step)
        rng = np.random.default_rng(0)
        rr = rng.integers(0, 2048, 20000)
        cc = rng.integers(0, 1024, 20000)
        ok = ~np.isnan(case.dem[rr, cc])
        rr, cc = rr[ok], cc[ok]
        for j, sub in enumerate(table):
            lit = I.trace_cells(case.dem, rr, cc, sub["dc_fix"], sub["dr_fix"], sub["dz"])
            assert np.array_equal(masks[j][rr, cc], lit), j
            assert 0.005 < 1.0 - lit.mean() < 0.995    # the case really has both shade and light
    finally:
        eng.close()


@pytest.mark.parametrize("step_s,n_steps", [(900, 40), (6 * 3600, 10), (24 * 3600, 4)])
@pytest.mark.parametrize("f64", [False, True])
def test_other_time_bases_with_shading(step_s, n_steps, f64):
    """15-minute rows (BASELINE config C4's time base: one sun position per row), 6-hourly rows
    (24 sub-steps per row) and daily rows (96 sub-steps per row, more than the default sub-step
    capacity of a shading time block, which therefore has to grow) -- whole run vs the oracle."""
    case = make_case(72, n_steps, w=100, seed=21, step_s=step_s, start="20220620 00:00:00")
    res = P.compare_run(case, f64, computed=True, shadow=True)
    worst = max(res.values())
    assert worst < (1e-9 if f64 else 1e-4), res


def test_float_sample_variant_gives_the_same_masks():
    """The march samples an integer copy of the DEM when no elevation is negative (one fused
    integer add-min per sample); shadow=2 forces the float-sample variant.  Same masks, same run."""
    case = make_case(120, 12, w=136, seed=31)
    e_int, e_flt = P.make_engine(case, False, computed=True, shadow=1), P.make_engine(case, False, computed=True, shadow=2)
    try:
        for step in (1, 6, 10):
            assert np.array_equal(e_int.shade_masks(step), e_flt.shade_masks(step))
        assert np.array_equal(e_int.run(0, 12), e_flt.run(0, 12))
        for a, b in zip(e_int.state(np.float64), e_flt.state(np.float64)):
            assert np.array_equal(a, b, equal_nan=True)
    finally:
        e_int.close()
        e_flt.close()


def test_negative_elevations_fall_back_to_float_samples():
    """A DEM with negative elevations (bit patterns of negative floats do not order like the values)
    takes the float-sample march on its own; masks bit-exact vs the oracle."""
    case = make_case(96, 8, w=110, seed=33)
    case.dem[...] = case.dem - np.float32(520.0)          # spans about -320 .. +280 m
    case.elev_aws = float(case.dem[case.aws_rc])
    assert np.nanmin(case.dem) < 0 < np.nanmax(case.dem)
    eng = _engine(case)
    try:
        valid = ~np.isnan(case.dem)
        for step in (0, 3, 7):
            table = I.substep_table(I.to_unix(case.aws_rows[step]["DATE"]), time_step_seconds(case.aws_rows, step),
                                    case.lat, case.lon, case.cell)
            masks = eng.shade_masks(step)
            assert masks.shape[0] == len(table)
            for j, sub in enumerate(table):
                lit = I.shadow_mask(case.dem, sub["dc_fix"], sub["dr_fix"], sub["dz"])
                assert np.array_equal(masks[j][valid], lit[valid]), (step, j)
    finally:
        eng.close()


def test_step_rise_skip_on_smooth_terrain():
    """Smooth slopes under a high sun (46 N) with an isolated sharp ridge and a NaN hole: here the
    march skips most chunks through the step-rise pyramids (terrain that rises less per step than
    the rays cannot catch them), so this is where a non-conservative skip would show.  Masks
    bit-exact vs the oracle, both sample variants."""
    case = make_case(160, 16, w=192, seed=41, start="20220621 04:00:00", glacier_mask=False)
    r, c = np.mgrid[0:160, 0:192].astype(np.float64)
    dem = 1000.0 + 0.6 * r + 25.0 * np.sin(c / 40.0) + 4.0 * np.cos(r / 9.0)
    dem += 160.0 * np.exp(-(((r - 70) / 3.0) ** 2 + ((c - 110) / 18.0) ** 2))      # a sharp east-west ridge
    dem += 90.0 * np.exp(-(((r - 120) / 10.0) ** 2 + ((c - 40) / 2.5) ** 2))       # and a north-south one
    dem[30:44, 150:170] = np.nan                                                    # a hole rays cross
    case.dem[...] = dem.astype(np.float32)
    for a in case.albedo_maps.values():
        a[np.isnan(case.dem)] = np.nan
    case.swe[np.isnan(case.dem)] = np.nan
    case.elev_aws = float(case.dem[case.aws_rc])
    case.lat, case.lon = 46.0, 8.0
    valid = ~np.isnan(case.dem)
    shaded_any = 0
    for shadow in (1, 2):
        eng = P.make_engine(case, False, computed=True, shadow=shadow)
        try:
            for step in range(0, 16, 2):
                table = I.substep_table(I.to_unix(case.aws_rows[step]["DATE"]), time_step_seconds(case.aws_rows, step),
                                        case.lat, case.lon, case.cell)
                if not table:
                    continue
                masks = eng.shade_masks(step)
                for j, sub in enumerate(table):
                    lit = I.shadow_mask(case.dem, sub["dc_fix"], sub["dr_fix"], sub["dz"])
                    assert np.array_equal(masks[j][valid], lit[valid]), (shadow, step, j)
                    shaded_any += int((~lit[valid]).sum())
        finally:
            eng.close()
    assert shaded_any > 1000

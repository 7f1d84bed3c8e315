"""BASELINE.json's full sizes.  The NumPy oracle is far too slow for 2048 x 2048 x 2200 (20 minutes),
so the full-size runs are checked (a) against the oracle on a WINDOW of the raster around the AWS
cell for the whole season -- cells are independent given the AWS-cell scalars, so the window's cells
must come out the same as in the full run -- and (b) through size-independent properties: closure of
the melt totals against the per-step area sums, SWE bookkeeping, split runs, row bands."""
import dataclasses

import numpy as np
import pytest

from enrgy_b200 import _lib
from enrgy_b200.engine import Engine
from enrgy_b200.forcing import build_forcing
from enrgy_b200.synthetic import make_case
from oracle import insolation_oracle as I
from tests import parity as P

pytestmark = pytest.mark.gpu

N, T = 2048, 2200            # config C2 of BASELINE.json


@pytest.fixture(scope="module")
def c2_case():
    return make_case(N, T, seed=0)


def _engine(case, f64, shadow=False, band=None):
    h, w = case.shape
    eng = Engine(h, w, precision=_lib.F64 if f64 else _lib.F32)
    r0, rows = band if band else (0, h)
    eng.set_params(cell_size=case.cell, elev_aws=case.elev_aws, aws_row=case.aws_rc[0], aws_col=case.aws_rc[1],
                   sensor_z=1.6, zm=1e-3, z_h_or_e=1e-4, emissivity=0.98, insol_mode=_lib.INSOL_COMPUTED,
                   shadow=shadow, lat=case.lat, lon=case.lon, band_row0=r0, band_rows=rows)
    eng.set_dem(case.dem)
    alb = P.clipped_albedo(case, np.float32)
    keys = list(alb)
    eng.set_forcing(build_forcing(case.aws_rows, keys))
    eng.set_albedo_maps([alb[k][r0:r0 + rows] for k in keys])
    eng.set_swe(case.swe[r0:r0 + rows])
    eng.prepass()
    return eng


def _window(case, half):
    """The case cut to (2 half + 1)^2 cells around the AWS cell, plus a one-cell halo for the terrain
    normals; same forcing, same AWS cell (geotransform shifted)."""
    r, c = case.aws_rc
    r0, r1, c0, c1 = r - half - 1, r + half + 2, c - half - 1, c + half + 2
    gt = list(case.geotransform)
    gt[0] += c0 * case.cell
    gt[3] -= r0 * case.cell
    sl = (slice(r0, r1), slice(c0, c1))
    return dataclasses.replace(case, dem=case.dem[sl].copy(), geotransform=tuple(gt), swe=case.swe[sl].copy(),
                               albedo_maps={k: a[sl].copy() for k, a in case.albedo_maps.items()},
                               aws_rc=(r - r0, c - c0)), sl


@pytest.mark.parametrize("f64", [True, False])
def test_c2_full_season_window_vs_oracle(c2_case, f64):
    case = c2_case
    eng = _engine(case, f64)
    try:
        stats = eng.run(0, T)
        swe, tsn, tic = eng.state(np.float64)
    finally:
        eng.close()
    win, sl = _window(case, 24)
    pot = I.insolation_series(win, shadow=False, dtype=np.float64)
    ora = P.run_oracle(win, pot if f64 else pot.astype(np.float32), f64)
    inner = (slice(1, -1), slice(1, -1))           # the halo cells see other neighbours than in the full raster
    tol, floor = (1e-9, 1e-6) if f64 else (1e-4, 1e-3)
    for got, name in ((swe, "swe"), (tsn, "total_snow"), (tic, "total_ice")):
        err = P.max_rel_err(got[sl][inner], np.asarray(ora[name], dtype=np.float64)[inner], floor)
        assert err < tol, (name, err)
    assert np.nanmax(tic[sl][inner]) > 0.5         # the season really melts ice there (metres w.e.)

    # closure: melt totals of the rasters == per-step area sums added up over the season
    valid = ~np.isnan(case.dem)
    total_rasters = float(np.sum(tsn[valid]) + np.sum(tic[valid]))
    total_steps = float(np.sum(stats[:, _lib.S_SNOW]) + np.sum(stats[:, _lib.S_ICE]))
    assert abs(total_rasters - total_steps) < (1e-9 if f64 else 2e-5) * total_rasters
    # SWE bookkeeping: what left the snow pack is the snow melt
    assert np.allclose(case.swe[valid].astype(np.float64) - swe[valid], tsn[valid], rtol=0, atol=1e-12 if f64 else 1e-6)
    # off-glacier cells are NaN in all three rasters (model.py:258)
    assert np.isnan(swe[~valid]).all() and np.isnan(tsn[~valid]).all() and np.isnan(tic[~valid]).all()
    assert stats[-1, _lib.S_NVALID] == valid.sum()


def test_c2_split_runs_and_row_bands_bit_identical(c2_case):
    """Season in one launch == season in three launches == two row bands side by side (float32)."""
    case = c2_case
    one = _engine(case, False)
    s_one = one.run(0, T)
    st_one = one.state(np.float32)
    one.close()
    three = _engine(case, False)
    s_three = np.concatenate([three.run(0, 700), three.run(700, 1501), three.run(1501, T)])
    st_three = three.state(np.float32)
    three.close()
    assert np.array_equal(s_one, s_three)
    for a, b in zip(st_one, st_three):
        assert np.array_equal(a, b, equal_nan=True)
    total = np.zeros_like(s_one)
    for band in ((0, 1040), (1040, N - 1040)):
        eng = _engine(case, False, band=band)
        total += eng.run(0, T)
        part = eng.state(np.float32)
        eng.close()
        for a, b in zip(part, st_one):
            assert np.array_equal(a, b[band[0]:band[0] + band[1]], equal_nan=True)
    cols = [_lib.S_RS, _lib.S_SENS, _lib.S_LAT, _lib.S_MELT, _lib.S_SNOW, _lib.S_ICE]
    assert np.allclose(total[1:, cols], s_one[1:, cols], rtol=2e-6, atol=1e-6)


@pytest.mark.parametrize("f64", [False, True])
def test_c2_per_step_fluxes_on_a_band_of_the_full_raster(c2_case, f64):
    """Every flux raster of EVERY row of the C2 season (not only the season totals), taken from a 64-row
    band of the full 2048 x 2048 raster -- same patches, same analytic-beam decisions as the full run --
    against the oracle on the window around the AWS cell.  The worst errors per field go to
    gpurun_out/f32_margin_c2_<precision>.json (copied to profiles/ by hand)."""
    import json
    import os
    case = c2_case
    half = 24
    r, c = case.aws_rc
    b0 = (r - half - 8) // 8 * 8                                   # band start on a patch boundary
    eng = _engine(case, f64, band=(b0, 64))
    win, sl = _window(case, half)
    pot = I.insolation_series(win, shadow=False, dtype=np.float64)
    ora = P.run_oracle(win, pot if f64 else pot.astype(np.float32), f64)
    inner = (slice(1, -1), slice(1, -1))
    rows = slice(sl[0].start - b0, sl[0].stop - b0)
    tol, ff, mfl = (1e-9, 1e-3, 1e-7) if f64 else (1e-4, 1.0, 1e-3)
    worst = {name: 0.0 for name in P.FLUX_FIELDS + ("snow", "ice")}
    hist = np.zeros(12, dtype=np.int64)                            # decades of the relative error, 1e-12 .. 1
    try:
        for t0 in range(0, T, 40):
            t1 = min(T, t0 + 40)
            dump = eng.dump_steps(t0, t1)
            eng.run(t0, t1, want_stats=False)
            for i in range(t0, t1):
                for name in worst:
                    if name == "g":
                        continue
                    if name in ("snow", "ice"):
                        ref = np.asarray(ora["melt"][i][0 if name == "snow" else 1], dtype=np.float64)
                        floor = mfl
                    else:
                        ref = np.array(ora["rows"][i][name], dtype=np.float64)
                        floor = ff
                    got = dump[i - t0, _lib.DUMP_NAMES.index(name)][rows, sl[1]][inner]
                    ref = ref[inner]
                    ok = ~np.isnan(ref)
                    e = np.abs(got[ok] - ref[ok]) / np.maximum(np.abs(ref[ok]), floor)
                    worst[name] = max(worst[name], float(e.max()))
                    if name in ("atmo", "mf"):
                        hist += np.histogram(np.log10(np.maximum(e, 1e-12)), bins=np.arange(-12, 1))[0]
    finally:
        eng.close()
    out = {"precision": "f64" if f64 else "f32", "rows": T, "cells": int((2 * half + 1) ** 2), "floors": {"flux_W_m2": ff, "melt_m_we": mfl},
           "worst_relative_error": worst, "atmo_mf_error_decades_1e-12_to_1": hist.tolist(), "bar": tol}
    if os.path.isdir("gpurun_out"):
        with open("gpurun_out/f32_margin_c2_%s.json" % out["precision"], "w") as f:
            json.dump(out, f, indent=1)
    print(out)
    assert max(worst.values()) < tol, worst


def test_c3_full_raster_shading_bands_rays_and_window():
    """Config C3 at its full raster: 8192 x 8192 with shading, 24 hourly rows (one day: the sun goes all
    the way round at 78 N).  (a) Two row bands give bit-identical rasters and statistics sums to the whole
    raster (the lines cross the cut); (b) the full masks of two rows against the oracle's sweep, and
    60000 cells ray-traced one by one along their scan lines; (c) the energy balance of a window around
    the AWS cell against the oracle for all 24 rows, its insolation computed by the oracle with rays
    through the full raster."""
    from oracle.enrgy_oracle import time_step_seconds
    n, t = 8192, 24
    case = make_case(n, t, seed=11, start="20220615 00:00:00")
    valid = ~np.isnan(case.dem)
    whole = _engine(case, False, shadow=True)
    try:
        s_whole = whole.run(0, t)
        st = whole.state(np.float32)
        tables, masks = {}, {}
        for step in (3, 14):
            tables[step] = I.substep_table(I.to_unix(case.aws_rows[step]["DATE"]), time_step_seconds(case.aws_rows, step),
                                           case.lat, case.lon, case.cell)
            masks[step] = whole.shade_masks(step)
    finally:
        whole.close()
    # (a) row bands (the cut is a multiple of the patch height)
    total = np.zeros_like(s_whole)
    for band in ((0, 4112), (4112, n - 4112)):
        eng = _engine(case, False, shadow=True, band=band)
        total += eng.run(0, t)
        part = eng.state(np.float32)
        eng.close()
        for a, b in zip(part, st):
            assert np.array_equal(a, b[band[0]:band[0] + band[1]], equal_nan=True)
    # (the per-CTA statistic rows are float32 sums; bands change which cells a CTA adds up)
    cols = [_lib.S_RS, _lib.S_SENS, _lib.S_LAT, _lib.S_MELT, _lib.S_SNOW]
    assert np.allclose(total[1:, cols], s_whole[1:, cols], rtol=2e-5, atol=1e-6)
    assert np.array_equal(total[1:, _lib.S_NSNOW], s_whole[1:, _lib.S_NSNOW])
    # (b) masks
    rng = np.random.default_rng(1)
    rr, cc = rng.integers(0, n, 80000), rng.integers(0, n, 80000)
    ok = valid[rr, cc]
    rr, cc = rr[ok], cc[ok]
    assert rr.size >= 50000
    n_traced, shade = 0, []
    for step, table in tables.items():
        assert masks[step].shape[0] == len(table) == 4
        for j, sub in enumerate(table):
            full = I.shadow_mask(case.dem, sub["dc_fix"], sub["dr_fix"], sub["dz"])
            assert np.array_equal(masks[step][j][valid], full[valid]), (step, j)
            lit = I.trace_cells(case.dem, rr, cc, sub["dc_fix"], sub["dr_fix"], sub["dz"])
            assert np.array_equal(masks[step][j][rr, cc], lit), (step, j)
            shade.append(1.0 - lit.mean())
            n_traced += rr.size
    assert n_traced >= 400000 and max(shade) > 0.1                 # 03:00 UTC: long shadows
    # (c) window vs oracle, all rows, with shading
    win, sl = _window(case, 24)
    r0, c0 = sl[0].start, sl[1].start
    pot = np.stack([I.potential_insolation_window(case.dem, case.cell, case.lat, case.lon, I.to_unix(row["DATE"]),
                                                  time_step_seconds(case.aws_rows, i),
                                                  (r0, sl[0].stop, c0, sl[1].stop), shadow=True)
                    for i, row in enumerate(case.aws_rows)])
    ora = P.run_oracle(win, pot.astype(np.float32), False)
    inner = (slice(1, -1), slice(1, -1))
    for got, name in zip(st, ("swe", "total_snow", "total_ice")):
        err = P.max_rel_err(got[sl][inner].astype(np.float64), np.asarray(ora[name], dtype=np.float64)[inner], 1e-3)
        assert err < 1e-4, (name, err)

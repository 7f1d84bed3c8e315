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


def test_large_raster_full_masks():
    """2048 x 1024: the oracle's sweep gives the full mask in milliseconds, and 20000 random cells are
    also ray-traced one by one along their scan lines (oracle trace_cells: an independent
    formulation of the same specification)."""
    case = make_case(2048, 14, w=1024, seed=5)
    eng = _engine(case)
    try:
        rng = np.random.default_rng(0)
        rr = rng.integers(0, 2048, 20000)
        cc = rng.integers(0, 1024, 20000)
        ok = ~np.isnan(case.dem[rr, cc])
        rr, cc = rr[ok], cc[ok]
        valid = ~np.isnan(case.dem)
        shade = []
        for step in (1, 7, 12):                          # 01:00 UTC: sun ~10 deg above the northern horizon
            table = I.substep_table(I.to_unix(case.aws_rows[step]["DATE"]), 3600, case.lat, case.lon, case.cell)
            masks = eng.shade_masks(step)
            for j, sub in enumerate(table):
                full = I.shadow_mask(case.dem, sub["dc_fix"], sub["dr_fix"], sub["dz"])
                assert np.array_equal(masks[j][valid], full[valid]), (step, j)
                lit = I.trace_cells(case.dem, rr, cc, sub["dc_fix"], sub["dr_fix"], sub["dz"])
                assert np.array_equal(masks[j][rr, cc], lit), (step, j)
                shade.append(1.0 - lit.mean())
        assert max(shade) > 0.2 and min(shade) < 0.05      # deep shade at night, next to none at noon
    finally:
        eng.close()


def test_all_directions_and_ragged_shapes():
    """At 78 N in June the sun goes all the way round: the 96 sub-steps of a day cover both scan-line
    families (row type / column type), both sweep directions and both signs of the shear.  Shapes
    that are no multiple of anything; every mask bit-exact."""
    seen = set()
    for shape, seed in (((77, 131), 3), ((203, 61), 4)):
        case = make_case(shape[0], 24, w=shape[1], seed=seed)
        eng = _engine(case)
        try:
            valid = ~np.isnan(case.dem)
            for step in range(24):
                table = I.substep_table(I.to_unix(case.aws_rows[step]["DATE"]), 3600, case.lat, case.lon, case.cell)
                masks = eng.shade_masks(step)
                assert masks.shape[0] == len(table)
                for j, sub in enumerate(table):
                    lit = I.shadow_mask(case.dem, sub["dc_fix"], sub["dr_fix"], sub["dz"])
                    assert np.array_equal(masks[j][valid], lit[valid]), (shape, step, j)
                    row_type, sigma, dfix = I.line_geometry(sub["dc_fix"], sub["dr_fix"])
                    seen.add((row_type, sigma, dfix > 0))
        finally:
            eng.close()
    assert len(seen) == 8, seen


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


def test_segments_chunks_and_external_masks():
    """enrgy_shade_scan into several row segments = the whole mask; a run whose masks are swept in
    small chunks (mask budget of a few sub-steps) = the run in one piece; enrgy_run_masked with the
    caller's masks = enrgy_run, all bit for bit."""
    import torch
    case = make_case(152, 30, w=136, seed=31)
    eng = _engine(case)
    ref = _engine(case)
    try:
        n = 30
        want_stats = ref.run(0, n)
        want_state = ref.state(np.float64)
        s0, s1 = eng.sub_range(0, n)
        assert (s0, s1) == (0, int(eng.point_scalars()[:, _lib.P_NSUB].sum()))
        # (a) three segments (row starts are multiples of 8) against the one-piece masks
        cuts = [(0, 48), (48, 64), (112, 40)]
        bufs = [torch.zeros((s1 - s0) * eng.mask_words(r), dtype=torch.int32, device="cuda") for _, r in cuts]
        whole = torch.zeros((s1 - s0) * eng.mask_words(152), dtype=torch.int32, device="cuda")
        eng.shade_scan(s0, s1, [(r0, r, b.data_ptr()) for (r0, r), b in zip(cuts, bufs)])
        eng.shade_scan(s0, s1, [(0, 152, whole.data_ptr())])
        eng.synchronize()
        words = 256 // 32                                            # pitch = round_up(136, 128)
        w_all = whole.cpu().numpy().view(np.uint32).reshape(s1 - s0, -1, words, 8)
        for (r0, r), b in zip(cuts, bufs):
            part = b.cpu().numpy().view(np.uint32).reshape(s1 - s0, -1, words, 8)
            rg = (r + 7) // 8
            got, want = part[:, :rg], w_all[:, r0 // 8:r0 // 8 + rg]
            nw = (136 + 31) // 32
            if r % 8:                                                # rows past the segment are not written
                keep = r % 8
                assert np.array_equal(got[:, -1, :nw, :keep], want[:, -1, :nw, :keep])
                got, want = got[:, :-1], want[:, :-1]
            assert np.array_equal(got[:, :, :nw], want[:, :, :nw]), (r0, r)
        # (b) the caller's masks
        stats = torch.zeros((n, _lib.S_COUNT), dtype=torch.float64, device="cuda")
        eng.run_masked(0, n, whole.data_ptr(), stats.data_ptr())
        eng.synchronize()
        assert np.array_equal(stats.cpu().numpy(), want_stats)
        for a, b in zip(eng.state(np.float64), want_state):
            assert np.array_equal(a, b, equal_nan=True)
    finally:
        eng.close()
        ref.close()
    # (c) chunked sweeps
    eng = _engine(case)
    try:
        eng.set_mask_budget(9 * 4 * (eng.mask_words(152) + 136 * 8))   # nine sub-steps per chunk
        got = eng.run(0, n)
        assert np.array_equal(got, want_stats)
        for a, b in zip(eng.state(np.float64), want_state):
            assert np.array_equal(a, b, equal_nan=True)
    finally:
        eng.close()


def test_uncropped_terrain_shades_the_glacier():
    """enrgy_set_terrain: relief outside the glacier outline casts shadows onto the glacier and shapes
    the slopes at its margin (the reference hands SAGA the uncropped DEM, model.py:469); the oracle is
    given the same uncropped raster."""
    case = make_case(120, 24, w=144, seed=12)
    full = make_case(120, 24, w=144, seed=12, glacier_mask=False)
    terrain = full.dem.copy()
    terrain[np.isnan(case.dem)] += np.float32(120.0)            # valley walls above the glacier surface
    assert np.array_equal(terrain[~np.isnan(case.dem)], case.dem[~np.isnan(case.dem)])
    eng = _engine(case)
    try:
        eng.set_terrain(terrain)
        eng.prepass()
        valid = ~np.isnan(case.dem)
        differs = 0
        for step in (0, 6, 13, 19):
            table = I.substep_table(I.to_unix(case.aws_rows[step]["DATE"]), 3600, case.lat, case.lon, case.cell)
            masks = eng.shade_masks(step)
            for j, sub in enumerate(table):
                lit = I.shadow_mask(terrain, sub["dc_fix"], sub["dr_fix"], sub["dz"])
                assert np.array_equal(masks[j][valid], lit[valid]), (step, j)
                differs += int((lit != I.shadow_mask(case.dem, sub["dc_fix"], sub["dr_fix"], sub["dz"]))[valid].sum())
            want = I.potential_insolation(terrain, case.cell, case.lat, case.lon, I.to_unix(case.aws_rows[step]["DATE"]),
                                          3600, shadow=True)
            got = eng.potential_insolation(step)
            assert P.max_rel_err(got[valid], want[valid], 1e-6) < 1e-4, step
        assert differs > 500                                        # the walls really matter here
    finally:
        eng.close()


def test_negative_elevations():
    """A DEM that straddles sea level; masks bit-exact vs the oracle."""
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


def test_smooth_terrain_with_ridges_and_a_hole():
    """Smooth slopes under a high sun (46 N) with two isolated sharp ridges and a NaN hole the lines
    cross (NaN cells never cast a shadow and do not interrupt a line); masks bit-exact vs the oracle."""
    case = make_case(160, 16, w=192, seed=41, start="20220621 04:00:00", glacier_mask=False)
    r, c = np.mgrid[0:160, 0:192].astype(np.float64)
    dem = 1000.0 + 0.6 * r + 25.0 * np.sin(c / 40.0) + 4.0 * np.cos(r / 9.0)
    dem += 160.0 * np.exp(-(((r - 70) / 3.0) ** 2 + ((c - 110) / 18.0) ** 2))      # a sharp east-west ridge
    dem += 90.0 * np.exp(-(((r - 120) / 10.0) ** 2 + ((c - 40) / 2.5) ** 2))       # and a north-south one
    dem[30:44, 150:170] = np.nan                                                    # a hole the lines cross
    case.dem[...] = dem.astype(np.float32)
    for a in case.albedo_maps.values():
        a[np.isnan(case.dem)] = np.nan
    case.swe[np.isnan(case.dem)] = np.nan
    case.elev_aws = float(case.dem[case.aws_rc])
    case.lat, case.lon = 46.0, 8.0
    valid = ~np.isnan(case.dem)
    shaded_any = 0
    eng = P.make_engine(case, False, computed=True, shadow=True)
    try:
        for step in range(0, 16, 2):
            table = I.substep_table(I.to_unix(case.aws_rows[step]["DATE"]), time_step_seconds(case.aws_rows, step),
                                    case.lat, case.lon, case.cell)
            if not table:
                continue
            masks = eng.shade_masks(step)
            for j, sub in enumerate(table):
                lit = I.shadow_mask(case.dem, sub["dc_fix"], sub["dr_fix"], sub["dz"])
                assert np.array_equal(masks[j][valid], lit[valid]), (step, j)
                shaded_any += int((~lit[valid]).sum())
    finally:
        eng.close()
    assert shaded_any > 1000

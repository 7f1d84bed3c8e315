"""Seeded synthetic inputs shared by the parity tests, the oracle harness and bench.py.

Shapes and value ranges follow SURVEY.md section 8(d): a 10 m DEM with an elevation ramp plus
roughness (so slope, aspect and shading are non-trivial), an optional elliptical glacier outline
(NaN outside, as GDAL's cutline crop leaves it, reference raster_utils.py:36-53), albedo maps at a
few dates, an SWE field increasing with elevation and an hourly (or 15-min) AWS series with the CSV
columns the reference reads (model.py:197-230).

Everything is analytic in (row, col) plus PCG64 draws for phases, so any size is generated in
O(H*W) and is bit-identical wherever it is generated (container, GPU box, oracle, GPU path).
"""
from __future__ import annotations

import csv
import math
import os
from dataclasses import dataclass, field
from datetime import datetime, timedelta

import numpy as np

# Aldegonda glacier (Svalbard) neighbourhood, EPSG:32633 -- the reference's own test site
# (reference model.py:556-557).
DEFAULT_ULX = 470000.0
DEFAULT_ULY = 8660000.0
DEFAULT_LAT = 77.98
DEFAULT_LON = 14.10


@dataclass
class SyntheticCase:
    dem: np.ndarray                    # float32 [H, W], NaN off-glacier
    geotransform: tuple                # GDAL order (ulx, dx, 0, uly, 0, -dy)
    cell: float
    albedo_maps: dict                  # "YYYYmmdd" -> float32 [H, W]
    swe: np.ndarray                    # float32 [H, W]
    aws_rows: list                     # list of dicts with the reference's CSV columns (strings)
    elev_aws: float
    xy_aws: tuple
    aws_rc: tuple                      # (row, col) of the AWS cell
    lat: float = DEFAULT_LAT
    lon: float = DEFAULT_LON
    meta: dict = field(default_factory=dict)

    @property
    def shape(self):
        return self.dem.shape

    def write_aws_csv(self, path):
        cols = list(self.aws_rows[0].keys())
        with open(path, "w", newline="") as f:
            w = csv.DictWriter(f, fieldnames=cols)
            w.writeheader()
            for r in self.aws_rows:
                w.writerow(r)
        return path


def _roughness(rows, cols, rng, n_comp=12, amp=40.0):
    """Sum of separable sinusoids with random wavelengths/phases (cheap fractal-ish relief)."""
    z = np.zeros((rows.size, cols.size), dtype=np.float64)
    for k in range(1, n_comp + 1):
        wl_r = rng.uniform(8.0, 90.0) * (1.0 + 6.0 / k)
        wl_c = rng.uniform(8.0, 90.0) * (1.0 + 6.0 / k)
        ph_r, ph_c = rng.uniform(0.0, 2.0 * math.pi, size=2)
        a = amp / k
        z += a * np.outer(np.sin(2.0 * math.pi * rows / wl_r + ph_r),
                          np.cos(2.0 * math.pi * cols / wl_c + ph_c))
    return z


def make_dem(h, w, seed=0, cell=10.0, glacier_mask=True, row0=0, total_rows=None):
    """float32 DEM [h, w]; rows row0..row0+h of a raster with total_rows rows (for row bands)."""
    total_rows = h if total_rows is None else total_rows
    rng = np.random.default_rng(seed)
    rows = np.arange(row0, row0 + h, dtype=np.float64)
    cols = np.arange(w, dtype=np.float64)
    # north (row 0) is high, as for a north-flowing... no: simply a ramp of 600 m over the raster
    ramp = 200.0 + 600.0 * (1.0 - rows / max(total_rows - 1, 1))
    z = ramp[:, None] + 50.0 * np.outer(np.cos(rows / 23.0), np.sin(cols / 17.0))
    z = z + _roughness(rows, cols, rng)
    dem = z.astype(np.float32)
    if glacier_mask:
        cy, cx = (total_rows - 1) / 2.0, (w - 1) / 2.0
        ry, rx = 0.485 * total_rows, 0.46 * w
        rr = ((rows[:, None] - cy) / ry) ** 2 + ((cols[None, :] - cx) / rx) ** 2
        # wavy outline so the NaN boundary is not tile-aligned
        wob = 0.06 * np.sin(7.0 * np.arctan2(rows[:, None] - cy, cols[None, :] - cx))
        dem[rr > (1.0 + wob)] = np.nan
    return dem


def make_albedo_maps(h, w, dates, seed=1, nan_like=None, row0=0):
    rng = np.random.default_rng(seed)
    rows = np.arange(row0, row0 + h, dtype=np.float64)
    cols = np.arange(w, dtype=np.float64)
    out = {}
    for d in dates:
        f = _roughness(rows, cols, rng, n_comp=5, amp=0.11)
        a = np.clip(0.40 + f, 0.05, 0.95).astype(np.float32)
        if nan_like is not None:
            a[np.isnan(nan_like)] = np.nan
        out[d] = a
    return out


def make_swe(dem, seed=2):
    """0..0.5 m w.e., increasing with elevation, a few snow-free patches at the low end."""
    rng = np.random.default_rng(seed)
    zmin = float(np.nanmin(dem))
    zmax = float(np.nanmax(dem))
    rel = (dem.astype(np.float64) - zmin) / max(zmax - zmin, 1.0)
    swe = 0.5 * rel - 0.05 + 0.01 * rng.standard_normal(1)[0]
    swe = np.where(swe < 0.0, 0.0, swe)
    swe = swe.astype(np.float32)
    swe[np.isnan(dem)] = np.nan
    return swe


def make_aws_rows(n_steps, start="20220601 00:00:00", step_s=3600, seed=3, with_gradient=False,
                  calm_every=0):
    """Synthetic AWS series with the reference's column names (model.py:197-230)."""
    rng = np.random.default_rng(seed)
    t0 = datetime.strptime(start, "%Y%m%d %H:%M:%S")
    rows = []
    for i in range(n_steps):
        t = t0 + timedelta(seconds=i * step_s)
        hod = t.hour + t.minute / 60.0
        doy = (t - datetime(t.year, 1, 1)).days
        diurnal = math.sin(2.0 * math.pi * (hod - 9.0) / 24.0)
        season = 2.5 * math.sin(2.0 * math.pi * (doy - 110) / 365.0)
        t_air = 3.0 + 4.0 * diurnal + season + 0.6 * rng.standard_normal()
        wind = float(np.clip(3.5 + 2.5 * math.sin(i / 37.0) + 0.8 * rng.standard_normal(), 0.0, 12.0))
        if calm_every and i % calm_every == calm_every - 1:
            wind = 0.0                   # exercises the 0 -> 0.1 m/s rule (var_classes.py:81-82)
        pres = 992.5 + 7.5 * math.sin(i / 61.0)
        hum = float(np.clip(78.0 + 17.0 * math.sin(i / 29.0 + 1.0), 30.0, 100.0))
        cld = float(np.clip(0.5 + 0.5 * math.sin(i / 41.0 + 2.0), 0.0, 1.0))
        elev_proxy = 0.45 + 0.55 * math.sin(2.0 * math.pi * (hod - 6.0) / 24.0)
        swd = max(0.0, 600.0 * elev_proxy) * (1.0 - 0.6 * cld)
        row = {
            "DATE": t.strftime("%Y%m%d %H:%M:%S"),
            "T_AIR": "%.2f" % t_air,
            "WIND_SPEED": "%.2f" % wind,
            "PRESSURE": "%.1f" % pres,
            "HUMID": "%.1f" % hum,
            "CLOUDINESS": "%.2f" % cld,
            "SWD": "%.1f" % swd,
        }
        if with_gradient:
            row["GRADIENT"] = "%.4f" % (-0.006 + 0.001 * math.sin(i / 53.0))
        rows.append(row)
    return rows


def make_station_rows(case, elev, seed=11, lapse=-0.006):
    """Series of an extra weather station (BASELINE config C4) on the case's time base: what the lapse
    rates predict at `elev` from the primary station (seed None), plus seeded departures (temperature,
    pressure, humidity, cloudiness) otherwise."""
    rng = None if seed is None else np.random.default_rng(seed)
    dz = elev - case.elev_aws
    rows = []
    for i, r in enumerate(case.aws_rows):
        t = float(r["T_AIR"]) + dz * lapse
        p = float(r["PRESSURE"]) + dz * -0.1145
        h = float(r["HUMID"])
        c = float(r["CLOUDINESS"])
        if rng is not None:
            t += 0.8 * math.sin(i / 19.0) + 0.3 * rng.standard_normal()
            p += 0.6 * math.sin(i / 47.0)
            h = float(np.clip(h + 8.0 * math.sin(i / 23.0 + 0.5), 30.0, 100.0))
            c = float(np.clip(c + 0.35 * math.sin(i / 31.0 + 1.0), 0.0, 1.0))
        rows.append({"DATE": r["DATE"], "T_AIR": "%.3f" % t, "PRESSURE": "%.2f" % p, "HUMID": "%.1f" % h,
                     "CLOUDINESS": "%.2f" % c})
    return rows


def make_case(n=256, n_steps=24, seed=0, cell=10.0, glacier_mask=True, albedo_dates=None,
              step_s=3600, start="20220601 00:00:00", with_gradient=False, calm_every=0,
              w=None):
    h = n
    w = n if w is None else w
    dem = make_dem(h, w, seed=seed, cell=cell, glacier_mask=glacier_mask)
    gt = (DEFAULT_ULX, cell, 0.0, DEFAULT_ULY, 0.0, -cell)
    if albedo_dates is None:
        albedo_dates = ["20220520", "20220615", "20220710", "20220805", "20220915"]
    albedo = make_albedo_maps(h, w, albedo_dates, seed=seed + 1, nan_like=dem)
    swe = make_swe(dem, seed=seed + 2)
    aws = make_aws_rows(n_steps, start=start, step_s=step_s, seed=seed + 3,
                        with_gradient=with_gradient, calm_every=calm_every)
    r, c = h // 2, w // 2
    if np.isnan(dem[r, c]):
        raise ValueError("AWS cell is off-glacier")
    # cell-centre coordinates of the AWS cell; reference get_value_by_real_coords
    # (raster_utils.py:85-89) truncates toward the upper-left corner.
    x = gt[0] + (c + 0.5) * cell
    y = gt[3] - (r + 0.5) * cell
    elev_aws = float(np.float32(dem[r, c]))
    return SyntheticCase(dem=dem, geotransform=gt, cell=cell, albedo_maps=albedo, swe=swe,
                         aws_rows=aws, elev_aws=elev_aws, xy_aws=(x, y), aws_rc=(r, c),
                         meta={"n": n, "n_steps": n_steps, "seed": seed, "step_s": step_s})


def make_band_case(n, n_steps, world=1, rank=0, seed=0, cell=10.0, glacier_mask=True,
                   albedo_dates=None, step_s=3600, start="20220601 00:00:00", balance=True, rows_full=None):
    """Weak-scaling workload: a (world*n) x n raster cut into `world` row bands -- or, with
    `rows_full`, STRONG scaling: a fixed rows_full x n raster cut into `world` row bands.

    With `balance` the band edges are placed so that every rank visits the same number of TILES
    (off-glacier tiles cost nothing, a tile on the glacier margin costs as much as a full one;
    SURVEY.md section 7 "load balance"), aligned to 16 rows;
    otherwise every band has n rows.  Returns (case, dem_full): `case` holds the band-local albedo
    maps / SWE of `rank` and the AWS description of the FULL raster (aws_rc in full-raster
    coordinates); `dem_full` is the whole DEM (replicated on every rank)."""
    from .parallel import row_bands, tile_cost_per_row
    strong = rows_full is not None
    rows_full = n * world if rows_full is None else int(rows_full)
    dem_full = make_dem(rows_full, n, seed=seed, cell=cell, glacier_mask=glacier_mask)
    if balance and world > 1:
        # equal numbers of visited tiles (a tile on the glacier margin costs as much as a full one)
        bands = row_bands(rows_full, world, align=16, valid_per_row=tile_cost_per_row(~np.isnan(dem_full)))
    elif strong:
        bands = row_bands(rows_full, world, align=16)
    else:
        bands = [(r * n, n) for r in range(world)]
    r0, nrows = bands[rank]
    band = dem_full[r0:r0 + nrows]
    gt = (DEFAULT_ULX, cell, 0.0, DEFAULT_ULY, 0.0, -cell)
    if albedo_dates is None:
        albedo_dates = ["20220520", "20220615", "20220710", "20220805", "20220915"]
    albedo = make_albedo_maps(nrows, n, albedo_dates, seed=seed + 1, nan_like=band, row0=r0)
    zmin, zmax = float(np.nanmin(dem_full)), float(np.nanmax(dem_full))
    rel = (band.astype(np.float64) - zmin) / max(zmax - zmin, 1.0)
    swe = np.where(0.5 * rel - 0.05 < 0.0, 0.0, 0.5 * rel - 0.05).astype(np.float32)
    swe[np.isnan(band)] = np.nan
    aws = make_aws_rows(n_steps, start=start, step_s=step_s, seed=seed + 3)
    r, c = rows_full // 2, n // 2
    if np.isnan(dem_full[r, c]):
        raise ValueError("AWS cell is off-glacier")
    x = gt[0] + (c + 0.5) * cell
    y = gt[3] - (r + 0.5) * cell
    case = SyntheticCase(dem=band, geotransform=gt, cell=cell, albedo_maps=albedo, swe=swe,
                         aws_rows=aws, elev_aws=float(np.float32(dem_full[r, c])), xy_aws=(x, y),
                         aws_rc=(r, c), meta={"n": n, "n_steps": n_steps, "world": world, "rank": rank,
                                              "band_row0": r0, "band_rows": nrows, "rows_full": rows_full,
                                              "bands": bands})
    return case, dem_full

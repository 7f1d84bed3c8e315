"""TEST INFRASTRUCTURE -- NumPy statement of this repo's potential-insolation specification.

PARITY UNPINNED UPSTREAM: the reference does not compute slope, aspect, solar geometry, atmospheric
transmittance or topographic shading itself; it shells out to SAGA GIS (`saga_cmd ta_lighting 2`,
reference saga_lighting.py:42-49), whose source is not in the reference checkout, whose version is
not pinned anywhere, and which is not installed here. The only things the reference fixes are the
tool options (saga_lighting.py:42-44): solar constant 1367 W m-2, output kWh m-2 integrated over
[t, t + time_step], sun position every 0.25 h, shadows on, lumped atmospheric transmittance 70 %,
output = direct + diffuse. This file freezes OUR specification of that computation (DESIGN.md
"Insolation specification"); "shading masks bit-exact" is graded against THIS file.

Specification (all angles radians, time base UTC, one sun position per sub-step for the whole grid):
  sun vector   low-precision almanac ephemeris from the Julian date -> (E, N, U) unit vector,
               evaluated with libm (`math`) in float64.
  sub-steps    hour step hs = 0.25 h; n = max(1, ceil(dt_h / hs - 1e-9)); width w_j = min(hs, rest);
               sun position at the sub-step midpoint; sub-steps with U <= 0 contribute nothing.
  terrain      central differences over the 4-neighbourhood; a missing (NaN / outside) neighbour is
               replaced by the mirrored opposite difference, or 0 if that is missing too;
               normal n = (-dz/dx, -dz/dy, 1) / sqrt(1 + |grad|^2), x = east, y = north, row 0 = north.
  direct       S0 * tau^(1/U) * max(0, n . s) * lit          (tau = 0.70; `beer_lambert.py:82-95`
               functional form flux*exp(-k*thickness) with k = -ln tau, thickness = 1/U)
  diffuse      S0 * (0.271 - 0.294 * tau^(1/U)) * U * (1 + n_z) / 2
  shadow       slim ray trace from the cell toward the sun: step k = 1, 2, ... lands on cell
               (r + ((k*dr_fix + 32768) >> 16), c + ((k*dc_fix + 32768) >> 16)) with Q16 fixed-point
               direction (dominant axis = +-65536); ray height z_k = dem + float32(k) * dz in float32
               (one rounded multiply, one rounded add); shaded iff some in-grid, non-NaN sample is
               > z_k; the ray stops at the grid edge or once z_k > max(dem).
  result       sum_j (direct_j + diffuse_j) * w_j / 1000   [kWh m-2]; NaN where the DEM is NaN.
"""
from __future__ import annotations

import math
from datetime import datetime, timezone

import numpy as np

S0 = 1367.0          # saga_lighting.py:42  -SOLARCONST 1367.0
TAU = 0.70           # saga_lighting.py:44  -METHOD 2 -LUMPED 70
HOUR_STEP = 0.25     # saga_lighting.py:43  -HOUR_STEP 0.25
DEG = 0.017453292519943295


def to_unix(date_str):
    """AWS DATE strings are taken as UTC."""
    try:
        d = datetime.strptime(date_str, "%Y%m%d")
    except ValueError:
        d = datetime.strptime(date_str, "%Y%m%d %H:%M:%S")
    return float(d.replace(tzinfo=timezone.utc).timestamp())


def sun_vector(t_unix, lat_deg, lon_deg):
    """(E, N, U) unit vector toward the sun; almanac low-precision formulae, float64 libm."""
    n = t_unix / 86400.0 + 2440587.5 - 2451545.0
    mean_lon = math.fmod(280.460 + 0.9856474 * n, 360.0)
    g = math.fmod(357.528 + 0.9856003 * n, 360.0) * DEG
    lam = (mean_lon + 1.915 * math.sin(g) + 0.020 * math.sin(2.0 * g)) * DEG
    eps = (23.439 - 0.0000004 * n) * DEG
    ra = math.atan2(math.cos(eps) * math.sin(lam), math.cos(lam))
    dec = math.asin(math.sin(eps) * math.sin(lam))
    gmst = math.fmod(280.46061837 + 360.98564736629 * n, 360.0)
    ha = (gmst + lon_deg) * DEG - ra
    phi = lat_deg * DEG
    sd, cd = math.sin(dec), math.cos(dec)
    sp, cp = math.sin(phi), math.cos(phi)
    sh, ch = math.sin(ha), math.cos(ha)
    up = sp * sd + cp * cd * ch
    east = -cd * sh
    north = sd * cp - cd * sp * ch
    return east, north, up


def substeps(t_unix, dt_s, hour_step=HOUR_STEP):
    """[(t_mid_unix, width_h)] covering [t, t + dt]."""
    dt_h = dt_s / 3600.0
    n = max(1, int(math.ceil(dt_h / hour_step - 1e-9)))
    out = []
    for j in range(n):
        w = min(hour_step, dt_h - j * hour_step)
        out.append((t_unix + (j * hour_step + 0.5 * w) * 3600.0, w))
    return out


def substep_table(t_unix, dt_s, lat, lon, cell, s0=S0, tau=TAU, hour_step=HOUR_STEP):
    """Per sub-step with the sun up: dict(E, N, U, B, D, dc_fix, dr_fix, dz) -- the scalars both the
    oracle and the CUDA path consume.  B, D already hold width/1000 (kWh) and D the 1/2 of the
    sky-view factor (1 + n_z)/2."""
    rows = []
    for t_mid, w in substeps(t_unix, dt_s, hour_step):
        e, n, u = sun_vector(t_mid, lat, lon)
        if not (u > 0.0):
            continue
        tb = math.pow(tau, 1.0 / u)
        b = s0 * tb * w / 1000.0
        d = s0 * (0.271 - 0.294 * tb) * u * w / 1000.0 * 0.5
        m = max(abs(e), abs(n))
        if m > 0.0:
            dc_fix = int(math.floor(e / m * 65536.0 + 0.5))
            dr_fix = int(math.floor(-n / m * 65536.0 + 0.5))
            dz = np.float32(cell * u / m)
        else:                       # sun at the zenith: nothing can shade
            dc_fix, dr_fix, dz = 0, 0, np.float32(np.inf)
        rows.append(dict(E=e, N=n, U=u, B=b, D=d, dc_fix=dc_fix, dr_fix=dr_fix, dz=dz, w=w))
    return rows


def terrain_normals(dem, cell):
    """(nx, ny, nz) float64 [H, W]; NaN where the DEM is NaN."""
    z = np.asarray(dem, dtype=np.float64)
    h, w = z.shape
    pad = np.full((h + 2, w + 2), np.nan)
    pad[1:-1, 1:-1] = z
    zn, zs = pad[:-2, 1:-1], pad[2:, 1:-1]
    zw, ze = pad[1:-1, :-2], pad[1:-1, 2:]

    def one_sided(a, b):
        # difference toward neighbour a; mirrored from b when a is missing; else 0
        va, vb = ~np.isnan(a), ~np.isnan(b)
        with np.errstate(invalid="ignore"):
            return np.where(va, a - z, np.where(vb, z - b, 0.0))

    d_n, d_s = one_sided(zn, zs), one_sided(zs, zn)
    d_e, d_w = one_sided(ze, zw), one_sided(zw, ze)
    gy = (d_n - d_s) / (2.0 * cell)
    gx = (d_e - d_w) / (2.0 * cell)
    inv = 1.0 / np.sqrt(1.0 + gx * gx + gy * gy)
    nx, ny, nz = -gx * inv, -gy * inv, inv
    bad = np.isnan(z)
    nx[bad] = np.nan
    ny[bad] = np.nan
    nz[bad] = np.nan
    return nx, ny, nz


def shadow_mask(dem32, dc_fix, dr_fix, dz32, zmax32=None):
    """lit [H, W] bool (True = sunlit) by the slim ray trace of the specification."""
    dem32 = np.asarray(dem32, dtype=np.float32)
    h, w = dem32.shape
    if zmax32 is None:
        zmax32 = np.float32(np.nanmax(dem32))
    valid = ~np.isnan(dem32)
    lit = np.ones((h, w), dtype=bool)
    if not np.isfinite(dz32):
        return lit
    active = valid.copy()
    dz32 = np.float32(dz32)
    k = 1
    while active.any():
        ro = (k * dr_fix + 32768) >> 16
        co = (k * dc_fix + 32768) >> 16
        if abs(ro) >= h or abs(co) >= w:
            break
        zk = dem32 + np.float32(k) * dz32                     # float32 multiply, float32 add
        # sample[r, c] = dem32[r + ro, c + co] where inside, else NaN (ray left the grid)
        sample = np.full((h, w), np.nan, dtype=np.float32)
        r0, r1 = max(0, -ro), min(h, h - ro)
        c0, c1 = max(0, -co), min(w, w - co)
        sample[r0:r1, c0:c1] = dem32[r0 + ro:r1 + ro, c0 + co:c1 + co]
        inside = np.zeros((h, w), dtype=bool)
        inside[r0:r1, c0:c1] = True
        with np.errstate(invalid="ignore"):
            active &= inside & ~(zk > zmax32)
            hit = active & (sample > zk)
        lit[hit] = False
        active &= ~hit
        k += 1
    return lit


def trace_cells(dem32, rows, cols, dc_fix, dr_fix, dz32, zmax32=None):
    """Same ray trace for a list of cells only (spot checks at sizes the full mask is too slow for)."""
    dem32 = np.asarray(dem32, dtype=np.float32)
    h, w = dem32.shape
    if zmax32 is None:
        zmax32 = np.float32(np.nanmax(dem32))
    rows = np.asarray(rows)
    cols = np.asarray(cols)
    z0 = dem32[rows, cols]
    lit = np.ones(rows.shape, dtype=bool)
    if not np.isfinite(dz32):
        return lit
    active = ~np.isnan(z0)
    dz32 = np.float32(dz32)
    k = 1
    while active.any():
        ro = (k * dr_fix + 32768) >> 16
        co = (k * dc_fix + 32768) >> 16
        r2, c2 = rows + ro, cols + co
        inside = (r2 >= 0) & (r2 < h) & (c2 >= 0) & (c2 < w)
        zk = z0 + np.float32(k) * dz32
        with np.errstate(invalid="ignore"):
            active &= inside & ~(zk > zmax32)
            sample = dem32[np.clip(r2, 0, h - 1), np.clip(c2, 0, w - 1)]
            hit = active & (sample > zk)
        lit[hit] = False
        active &= ~hit
        k += 1
    return lit


def potential_insolation(dem32, cell, lat, lon, t_unix, dt_s, shadow=True, normals=None,
                         s0=S0, tau=TAU, hour_step=HOUR_STEP, return_masks=False):
    """kWh m-2 over [t, t + dt], float64 [H, W]; optionally the per-sub-step lit masks."""
    dem32 = np.asarray(dem32, dtype=np.float32)
    if normals is None:
        normals = terrain_normals(dem32, cell)
    nx, ny, nz = normals
    table = substep_table(t_unix, dt_s, lat, lon, cell, s0, tau, hour_step)
    direct = np.zeros(dem32.shape, dtype=np.float64)
    dsum = 0.0
    masks = []
    zmax = np.float32(np.nanmax(dem32))
    for sub in table:
        cosi = nx * sub["E"] + ny * sub["N"] + nz * sub["U"]
        term = sub["B"] * np.maximum(cosi, 0.0)
        if shadow:
            lit = shadow_mask(dem32, sub["dc_fix"], sub["dr_fix"], sub["dz"], zmax)
            term = np.where(lit, term, 0.0)
            masks.append(lit)
        direct = direct + term
        dsum = dsum + sub["D"]
    pot = direct + dsum * (1.0 + nz)
    pot[np.isnan(dem32)] = np.nan
    if return_masks:
        return pot, masks, table
    return pot


def insolation_series(case, shadow=True, dtype=np.float64):
    """[T, H, W] potential insolation for a SyntheticCase (small cases only)."""
    from oracle.enrgy_oracle import time_step_seconds
    normals = terrain_normals(case.dem, case.cell)
    out = np.empty((len(case.aws_rows),) + case.dem.shape, dtype=dtype)
    for i, row in enumerate(case.aws_rows):
        dt = time_step_seconds(case.aws_rows, i)
        out[i] = potential_insolation(case.dem, case.cell, case.lat, case.lon, to_unix(row["DATE"]), dt,
                                      shadow=shadow, normals=normals)
    return out

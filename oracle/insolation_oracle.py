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
  shadow       line-sweep ray trace.  All rays of one sub-step share one direction, so the raster is
               cut into sheared scan LINES and every ray runs along its cell's line:
                 direction in Q16 fixed point as before (dominant axis = +-65536);
                 |dr_fix| >= |dc_fix| ("row type"):  u = sign(dr_fix) * row   (u grows toward the sun)
                     line l holds the cells (row, l + sh(u)),  sh(u) = (u * dc_fix + 32768) >> 16
                 else ("column type"):               u = sign(dc_fix) * col
                     line l holds the cells (l + sh(u), col),  sh(u) = (u * dr_fix + 32768) >> 16
               so the ray of a cell visits, k = 1, 2, ... steps toward the sun, the cell of its own line
               at u + k (one cell per step along the dominant axis; the other coordinate deviates from
               the straight line by less than one cell).  Heights are compared in float64:
                 g(cell) = float64(dem) - float64(u) * float64(dz)      (dz: float32 rise per step)
               (product and difference are exact for any real DEM: 15 + 24 and < 53 significant bits)
               and a cell is SHADED iff some in-grid, non-NaN cell further along its line toward the sun
               has g > g(cell)  <=>  dem_k - dem_0 > k * dz.  Rays end at the grid edge.  Because g does
               not depend on the ray, the mask of a whole sub-step is ONE running maximum per line,
               swept from the sunward edge: lit = not (M > g); M = max(M, g)  -- O(H*W) whatever the
               terrain, which is what the CUDA path does.
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


def line_geometry(dc_fix, dr_fix):
    """(row_type, sigma, dfix): scan-line family of a Q16 direction (module docstring, "shadow")."""
    row_type = abs(dr_fix) >= abs(dc_fix)
    if row_type:
        return True, (1 if dr_fix > 0 else -1), int(dc_fix)
    return False, (1 if dc_fix > 0 else -1), int(dr_fix)


def shear(u, dfix):
    """sh(u) = (u * dfix + 32768) >> 16 (floor), for Python ints or int64 arrays."""
    return (u * dfix + 32768) >> 16


def shadow_mask(dem32, dc_fix, dr_fix, dz32, zmax32=None):
    """lit [H, W] bool (True = sunlit) by the line sweep of the specification.  `zmax32` is accepted
    for call compatibility and ignored (rays end at the grid edge)."""
    dem32 = np.asarray(dem32, dtype=np.float32)
    h, w = dem32.shape
    lit = np.ones((h, w), dtype=bool)
    if not np.isfinite(dz32) or (dc_fix == 0 and dr_fix == 0):
        return lit
    row_type, sigma, dfix = line_geometry(dc_fix, dr_fix)
    z = dem32.astype(np.float64)
    if not row_type:
        z = z.T                                         # sweep axis first
    na, nb = z.shape
    dz = float(np.float32(dz32))
    out = np.ones((na, nb), dtype=bool)
    us = sigma * np.arange(na, dtype=np.int64)
    sh = shear(us, dfix)
    l_off = int(sh.max())                               # line l lives at index l + l_off
    m = np.full(nb + int(sh.max() - sh.min()), -np.inf)
    b = np.arange(nb, dtype=np.int64)
    for a in np.argsort(-us, kind="stable"):            # from the sunward edge (largest u) away from the sun
        u = int(us[a])
        idx = b - int(sh[a]) + l_off
        hrow = z[a]
        ok = ~np.isnan(hrow)
        g = hrow - float(u) * dz
        mm = m[idx]
        with np.errstate(invalid="ignore"):
            out[a] = ~(mm > g)
        m[idx] = np.where(ok, np.maximum(mm, np.where(ok, g, -np.inf)), mm)
    lit = out if row_type else out.T
    lit = lit.copy()
    lit[np.isnan(dem32)] = True                         # off-glacier cells read "sunlit"
    return lit


def trace_cells(dem32, rows, cols, dc_fix, dr_fix, dz32, zmax32=None):
    """The same specification cell by cell (the ray of each listed cell marched along its own line):
    an independent formulation of shadow_mask, and the spot check at sizes where full masks are slow."""
    dem32 = np.asarray(dem32, dtype=np.float32)
    h, w = dem32.shape
    rows = np.asarray(rows, dtype=np.int64)
    cols = np.asarray(cols, dtype=np.int64)
    lit = np.ones(rows.shape, dtype=bool)
    if not np.isfinite(dz32) or (dc_fix == 0 and dr_fix == 0):
        return lit
    row_type, sigma, dfix = line_geometry(dc_fix, dr_fix)
    dz = float(np.float32(dz32))
    z0 = dem32[rows, cols].astype(np.float64)
    u0 = sigma * (rows if row_type else cols)
    line = (cols if row_type else rows) - shear(u0, dfix)
    g0 = z0 - u0.astype(np.float64) * dz
    active = ~np.isnan(z0)
    k = 1
    while active.any():
        u = u0 + k
        major = sigma * u
        minor = line + shear(u, dfix)
        r2, c2 = (major, minor) if row_type else (minor, major)
        inside = (r2 >= 0) & (r2 < h) & (c2 >= 0) & (c2 < w)
        active &= inside                                # both coordinates are monotone along a ray
        smp = dem32[np.clip(r2, 0, h - 1), np.clip(c2, 0, w - 1)].astype(np.float64)
        with np.errstate(invalid="ignore"):
            hit = active & ((smp - u.astype(np.float64) * dz) > g0)
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


def potential_insolation_window(dem32, cell, lat, lon, t_unix, dt_s, window, shadow=True,
                                s0=S0, tau=TAU, hour_step=HOUR_STEP):
    """potential_insolation for the cells of window = (r0, r1, c0, c1) of a LARGE raster: terrain normals
    from the window plus a one-cell halo, sunlit masks by marching the rays of the window's cells
    through the full raster (trace_cells).  Same values as the full computation, at a cost that does
    not depend on the size of the raster."""
    dem32 = np.asarray(dem32, dtype=np.float32)
    h, w = dem32.shape
    r0, r1, c0, c1 = window
    ra, rb, ca, cb = max(r0 - 1, 0), min(r1 + 1, h), max(c0 - 1, 0), min(c1 + 1, w)
    nx, ny, nz = terrain_normals(dem32[ra:rb, ca:cb], cell)
    sl = (slice(r0 - ra, r0 - ra + (r1 - r0)), slice(c0 - ca, c0 - ca + (c1 - c0)))
    nx, ny, nz = nx[sl], ny[sl], nz[sl]
    rr, cc = np.mgrid[r0:r1, c0:c1]
    direct = np.zeros(nz.shape, dtype=np.float64)
    dsum = 0.0
    for sub in substep_table(t_unix, dt_s, lat, lon, cell, s0, tau, hour_step):
        cosi = nx * sub["E"] + ny * sub["N"] + nz * sub["U"]
        term = sub["B"] * np.maximum(cosi, 0.0)
        if shadow:
            lit = trace_cells(dem32, rr.ravel(), cc.ravel(), sub["dc_fix"], sub["dr_fix"], sub["dz"]).reshape(nz.shape)
            term = np.where(lit, term, 0.0)
        direct = direct + term
        dsum = dsum + sub["D"]
    pot = direct + dsum * (1.0 + nz)
    pot[np.isnan(dem32[r0:r1, c0:c1])] = np.nan
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

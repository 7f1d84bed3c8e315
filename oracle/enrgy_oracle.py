"""TEST INFRASTRUCTURE -- CPU (NumPy) restatement of ENRGY's per-cell, per-timestep energy balance.

This file is the parity ORACLE for the CUDA path in `enrgy_b200/csrc/`. It is NOT part of the
product: only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import it, and only as the checker / the timed CPU baseline.

What it restates (file:line into the reference checkout, tepextepex/ENRGY):
  * the time loop and per-step sequencing               model.py:183-283
  * AWS scalars and the lapse-rate distribution         var_classes.py:80-85, :113-125, :144-183
  * Magnus saturation vapour pressure                   turbo.py:368-379
  * bulk-aerodynamic turbulent fluxes, psi functions,
    Monin-Obukhov fixed-count iteration, Andreas z0     turbo.py:43-365
  * longwave, shortwave, SW scaling factor              model.py:464-545
  * albedo (date-linear blend, snow ageing, ice cap)    model.py:298-337, interpolator.py:5-39
  * flux sum, clamp, melt partition, state update       model.py:411-438, msm.py:193-203, model.py:245-261
  * sub-surface model `tick`                            msm.py:18-107
  * area statistics row                                 var_classes.py:45-56, model.py:246-269
  * CSV/time helpers                                    helpers.py:27-87, raster_utils.py:85-89

Pinning: `tests/test_oracle_vs_reference.py` runs the unmodified reference through
`oracle/ref_harness.py` (where /root/reference exists) and requires BIT-IDENTICAL rasters and CSV
text; `tests/golden/*.npz` (made by `tests/golden/make_golden.py` from the reference itself) pin it
where the reference is absent; the scalar known answers of SURVEY.md section 4 are checked in
`tests/test_oracle_known_answers.py`. Because the reference is NumPy, the restatement keeps the
reference's expression order and its Python-scalar / np.float64-scalar / array operand kinds, so
NumPy's promotion rules (NEP 50, numpy >= 2) reproduce the reference's dtype flow exactly.

The insolation raster is an INPUT here (the reference gets it from SAGA GIS, an external binary,
saga_lighting.py:42-49); `oracle/insolation_oracle.py` holds this repo's own specification of it.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from datetime import datetime
from math import exp as _math_exp, pi as _PI

import numpy as np

# ---- constants -------------------------------------------------------------------------------
# turbo.py:30-40
R_AIR = 287.058
KARMAN = 0.4
GRAV = 9.81
CP_AIR = 1010
TS_MELT = 0 + 273.15
ES_MELT = 611
LV = 2.514 * 10 ** 6
LS = 2.849 * 10 ** 6
ZM_DEFAULT = 0.001


def default_params():
    """var_classes.py:7-15 (a fresh dict: the reference mutates its module global, model.py:84-88)."""
    return {
        "ice_density": 900.0,
        "snow_density": 387.0,
        "latent_heat_of_fusion": 3.34 * 10 ** 5,
        "specific_heat_capacity_ice": 2097.0,
        "thermal_diffusivity_ice": 1.16 * 10 ** -6,
        "thermal_diffusivity_snow": 0.40 * 10 ** -6,
        "g": 9.81,
    }


# ---- helpers.py ------------------------------------------------------------------------------
def time_step_seconds(rows, i):
    """helpers.py:63-71 + model.py:190-193: forward difference, last row uses the previous one."""
    def parse(s):
        try:
            return datetime.strptime(s, "%Y%m%d")
        except ValueError:
            return datetime.strptime(s, "%Y%m%d %H:%M:%S")
    if i < len(rows) - 1:
        d = parse(rows[i + 1]["DATE"]) - parse(rows[i]["DATE"])
    else:
        d = parse(rows[i]["DATE"]) - parse(rows[i - 1]["DATE"])
    return int(d.total_seconds())


def unit_guess(value, scale=10):
    """helpers.py:74-87."""
    if 1 < value <= scale:
        return value / scale
    elif value <= 1:
        return value
    raise ValueError("Wrong value encountered")


def kwh_to_w(insol, dt):
    """helpers.py:54-60 then :27-36 -- kWh -> J -> W, in the reference's operation order."""
    return insol * 3.6 * 10 ** 6 / dt


def sample_at(array, gt, easting, northing):
    """raster_utils.py:85-89."""
    ul_x, x_dist, _, ul_y, _, y_dist = gt
    pixel = int((easting - ul_x) / x_dist)
    line = -int((ul_y - northing) / y_dist)
    return array[line][pixel]


# ---- turbo.py --------------------------------------------------------------------------------
def e_max(t_air, air_pressure):
    """turbo.py:368-379 (Kelvin, Pascal -> Pascal)."""
    t_air = t_air - 273.15
    air_pressure = air_pressure / 100
    ew_t = 611.2 * np.exp((17.62 * t_air) / (243.12 + t_air))
    f_p = 1.0016 + 3.15 * 10 ** -6 * air_pressure - 0.074 / air_pressure
    return f_p * ew_t


def dry_air_density(t_air, p_air):
    """turbo.py:83-85."""
    return p_air / (R_AIR * t_air)


def dyer_x(zeta):
    """turbo.py:364-365."""
    return (1 - 16 * zeta) ** (1 / 4)


_A, _B, _C, _D = 0.7, 0.75, 5, 0.35


def minus_psi_m(z, L):
    """turbo.py:308-333."""
    zeta = z / L
    if isinstance(zeta, np.ndarray):
        x = dyer_x(zeta)
        return np.where(zeta >= 0,
                        _A * zeta + _B * (zeta - _C / _D) * np.exp(-_D * zeta) + _B * _C / _D,
                        -(2 * np.log((1 + x) / 2) + np.log((1 + x ** 2) / 2) - 2 * np.arctan(x) + _PI / 2))
    if zeta >= 0:
        return _A * zeta + _B * (zeta - _C / _D) * np.exp(-_D * zeta) + _B * _C / _D
    x = dyer_x(zeta)
    return -(2 * np.log((1 + x) / 2) + np.log((1 + x ** 2) / 2) - 2 * np.arctan(x) + _PI / 2)


def minus_psi_h(z, L):
    """turbo.py:336-361."""
    zeta = z / L
    if isinstance(zeta, np.ndarray):
        x = dyer_x(zeta)
        return np.where(zeta >= 0,
                        (1 + 2 * _A * zeta / 3) ** 1.5 + _B * (zeta - _C / _D) * np.exp(-_D * zeta) + _B * _C / _D - 1,
                        -(2 * np.log((1 + x ** 2) / 2)))
    if zeta >= 0:
        return (1 + 2 * _A * zeta / 3) ** 1.5 + _B * (zeta - _C / _D) * np.exp(-_D * zeta) + _B * _C / _D - 1
    x = dyer_x(zeta)
    return -(2 * np.log((1 + x ** 2) / 2))


def friction_velocity(uz, z, L=None, zm=None):
    """turbo.py:293-305 (note: psi_m is NOT multiplied by z/L here)."""
    if zm is None:
        zm = ZM_DEFAULT
    num = KARMAN * uz
    if L is not None:
        den = np.log(z / zm) + minus_psi_m(z, L)
    else:
        den = np.log(z / zm)
    return num / den


def andreas_bi(Re):
    """turbo.py:199-225."""
    if isinstance(Re, np.ndarray):
        b0 = np.full(Re.shape, 1.25)
        b1 = np.full(Re.shape, 0.00)
        b2 = np.full(Re.shape, 0.00)
        b0 = np.where(Re > 0.135, 0.149, b0)
        b1 = np.where(Re > 0.135, -0.55, b1)
        b2 = np.where(Re > 0.135, 0.0, b2)
        b0 = np.where(Re > 2.5, 0.317, b0)
        b1 = np.where(Re > 2.5, -0.565, b1)
        b2 = np.where(Re > 2.5, -0.183, b2)
        return b0, b1, b2
    if Re <= 0.135:
        return 1.25, 0, 0
    if Re <= 2.5:
        return 0.149, -0.55, 0
    return 0.317, -0.565, -0.183


def andreas_z0(uz, z, zm, L):
    """turbo.py:228-261."""
    u_star = friction_velocity(uz, z, zm=zm, L=L)
    Re = u_star * zm / 1.5e-5
    b0, b1, b2 = andreas_bi(Re)
    ln_re = np.log(Re)
    return zm * np.exp(b0 + b1 * ln_re + b2 * ln_re ** 2)


def exchange_coef(z, L=None, zm=None, z_h_or_e=None, andreas=False, uz=None):
    """turbo.py:264-290."""
    if zm is None:
        zm = ZM_DEFAULT
    if z_h_or_e is None:
        z_h_or_e = zm / 10
    if andreas:
        if uz is None:
            raise ValueError("You must specify Uz parameter to use 'andreas=True' option")
        z_h_or_e = andreas_z0(uz, z, zm, L)
    num = KARMAN ** 2
    if L is not None:
        pm = minus_psi_m(z, L)
        ph = minus_psi_h(z, L)
        den = (np.log(z / zm) + pm * (z / L)) * (np.log(z / z_h_or_e) + ph * (z / L))
    else:
        den = np.log(z / zm) * np.log(z / z_h_or_e)
    return num / den


def sensible(z, uz, Tz, P, Ts=None, zm=None, z_h_or_e=None, L=None, andreas=False):
    """turbo.py:140-156."""
    if Ts is None:
        Ts = TS_MELT
    rho = dry_air_density(Tz, P)
    CH = exchange_coef(z, zm=zm, z_h_or_e=z_h_or_e, L=L, andreas=andreas, uz=uz)
    return CH * CP_AIR * rho * uz * (Tz - Ts)


def latent(z, uz, Tz, P, rh, Ts=None, zm=None, z_h_or_e=None, L=None, andreas=False):
    """turbo.py:159-196, including the Kelvin-vs-zero comparison that makes Ls dead code (F8)."""
    es = ES_MELT if Ts is None else e_max(Ts, P)
    emax = e_max(Tz, P)
    ez = emax * rh
    rho = dry_air_density(Tz, P)
    CE = exchange_coef(z, zm=zm, z_h_or_e=z_h_or_e, L=L, andreas=andreas, uz=uz)
    flux = CE * rho * uz * 0.622 / P * (ez - es)
    if Ts is None:
        return flux * LV
    if type(flux) == np.ndarray:
        return np.where(Ts >= 0, flux * LV, flux * LS)
    return flux * LV if Ts >= 0 else flux * LS


def monin_obukhov_length(Tz, P, u_star, Qh):
    """turbo.py:122-137."""
    rho = dry_air_density(Tz, P)
    num = rho * CP_AIR * u_star ** 3 * Tz
    den = KARMAN * GRAV * Qh
    return num / den


def sensible_iteratively(z, uz, Tz, P, Ts, zm=None, z_h_or_e=None, max_iter=5, andreas=False):
    """turbo.py:88-119: neutral first guess then exactly max_iter updates, no convergence test."""
    if not (isinstance(max_iter, int) and max_iter < 10):
        max_iter = 5
    u_star = friction_velocity(uz, z, zm=zm, L=None)
    Qh = sensible(z, uz, Tz, P, Ts=Ts, zm=zm, z_h_or_e=z_h_or_e, andreas=andreas, L=None)
    L = monin_obukhov_length(Tz, P, u_star, Qh)
    for _ in range(max_iter):
        u_star = friction_velocity(uz, z, zm=zm, L=L)
        Qh = sensible(z, uz, Tz, P, Ts, zm=zm, z_h_or_e=z_h_or_e, andreas=andreas, L=L)
        L = monin_obukhov_length(Tz, P, u_star, Qh)
    return Qh, L


def turbulent_fluxes(z, uz, Tz, P, rh, surface_temp=None, L=None, zm=None, z_h_or_e=None,
                     max_iter=5, andreas=False):
    """turbo.py:43-80 (without the bare except)."""
    if L is None:
        qh, L = sensible_iteratively(z, uz, Tz, P, surface_temp, zm=zm, z_h_or_e=z_h_or_e,
                                     andreas=andreas, max_iter=max_iter)
    else:
        qh = sensible(z, uz, Tz, P, Ts=surface_temp, L=L, zm=zm, z_h_or_e=z_h_or_e, andreas=andreas)
    qe = latent(z, uz, Tz, P, rh, Ts=surface_temp, L=L, zm=zm, z_h_or_e=z_h_or_e, andreas=andreas)
    return qh, qe, L


# ---- msm.py ----------------------------------------------------------------------------------
def melt_partition(melt_flux, swe, dt, params):
    """msm.py:193-203: snow melts first, the rest is ice."""
    q = melt_flux * dt
    kg = q / params["latent_heat_of_fusion"]
    we = kg / 1000
    if isinstance(we, np.ndarray):
        snow = np.where(we > swe, swe, we)
    else:
        snow = swe if we > swe else we
    ice = we - snow
    return snow, ice


def subsurface_tick(depths, temps, dt, flux=None, snow_depth=None, params=None):
    """msm.py:31-107: explicit conduction through the layer stack; returns (temps, qm, ground flux)."""
    params = default_params() if params is None else params
    c = params["specific_heat_capacity_ice"]
    k_ice = params["thermal_diffusivity_ice"]
    k_snow = params["thermal_diffusivity_snow"]
    rho_ice = params["ice_density"]
    rho_snow = params["snow_density"]
    if flux is None:
        flux = 0
    grads = []
    for t0, t1, d in zip(temps, temps[1:], depths):      # msm.py:18-28
        grads.append(np.nan if d == 0 else (t1 - t0) / d)
    new_temps = []
    surf = True
    qm = np.nan
    ground_flux = None
    for i in range(0, len(temps) - 1):
        if snow_depth is None:
            k, rho = k_ice, rho_ice
        else:
            if isinstance(snow_depth, np.ndarray):
                ratio = np.where(snow_depth > depths[i], 1, snow_depth / depths[i])
            else:
                ratio = 1 if snow_depth >= depths[i] else snow_depth / depths[i]
            k = ratio * k_snow + (1 - ratio) * k_ice
            rho = ratio * rho_snow + (1 - ratio) * rho_ice
            snow_depth -= depths[i]
            if isinstance(snow_depth, np.ndarray):
                snow_depth[snow_depth < 0] = 0
            else:
                snow_depth = 0 if snow_depth < 0 else snow_depth
        if depths[i] == 0:
            new_temps.append(temps[i])
            continue
        if surf:
            ground_flux = k * grads[i] * c * rho
            full = flux + ground_flux
            q0 = -temps[i] * c * rho * depths[i] / dt
            qm = full - q0
            if isinstance(qm, np.ndarray):
                qm[qm < 0] = 0
            else:
                qm = 0 if qm < 0 else qm
            delta_t = (full - qm) / (c * rho * depths[i])
            surf = False
        else:
            delta_t = k * (grads[i] - grads[i - 1]) / depths[i]
        new_temps.append(temps[i] + delta_t * dt)
    new_temps.append(temps[-1])
    return new_temps, qm, ground_flux


# ---- interpolator.py -------------------------------------------------------------------------
def albedo_bracket(keys, date_str):
    """interpolator.py:23-39 + :5-10: closest map dates before/after; ValueError outside range."""
    try:
        d = datetime.strptime(date_str, "%Y%m%d")
    except ValueError:
        d = datetime.strptime(date_str, "%Y%m%d %H:%M:%S")
    dates = [datetime.strptime(k, "%Y%m%d") for k in keys]
    lo = [x for x in dates if x <= d]
    hi = [x for x in dates if x >= d]
    if len(lo) == 0 or len(hi) == 0:
        raise ValueError("Passed date is outside of the possible interpolation range!")
    return d, max(lo), min(hi)


def albedo_blend(arrays, date_str):
    """interpolator.py:5-20: linear in WHOLE days between the two bracketing maps."""
    d, before, after = albedo_bracket(list(arrays), date_str)
    ar0 = arrays[before.strftime("%Y%m%d")]
    ar1 = arrays[after.strftime("%Y%m%d")]
    if (after - before).days == 0:
        return ar0
    return ar0 + (d - before).days * (ar1 - ar0) / (after - before).days


# ---- model.py --------------------------------------------------------------------------------
@dataclass
class ModelConfig:
    z: float = 2.0
    elev_aws: float = 0.0
    xy_aws: tuple = None
    zm: float = None
    z_h_or_e: float = None
    andreas: bool = False
    const_albedo: tuple = None            # (ice, snow)
    temp_lapse_rate: object = -0.006      # float or a CSV column name
    last_snowfall: str = None
    max_ice_albedo: float = None
    emissivity: float = None
    cloud_corr: float = None
    sensible_corr: float = 1
    latent_corr: float = 1
    msm: dict = None                      # dict(depths, temperatures, elev)
    snow_density: float = None


def surface_albedo(albedo_arrays, swe, date_str, cfg):
    """model.py:298-337."""
    if cfg.const_albedo is None:
        a = albedo_blend(albedo_arrays, date_str)
        if cfg.last_snowfall is not None:
            delta = (datetime.strptime(date_str, "%Y%m%d %H:%M:%S")
                     - datetime.strptime(cfg.last_snowfall, "%Y%m%d"))
            if delta.days > 0:
                snow_albedo = 0.40 + 0.44 * _math_exp(-0.12 * delta.days)
                a = np.where(swe > 0, snow_albedo, a)
        cap = 0.45 if cfg.max_ice_albedo is None else cfg.max_ice_albedo
        return np.where((swe <= 0) & (a > cap), cap, a)
    return np.where(swe > 0, cfg.const_albedo[1], cfg.const_albedo[0])


def longwave(Tz_surf, Tz, cloudiness, eps=None):
    """model.py:533-545 (sigma is 5.70e-8 in the reference, F11)."""
    sigma = 5.70 * 10 ** -8
    if eps is None:
        eps = 0.98
    lwu = eps * sigma * Tz_surf ** 4
    lwd = (0.765 + 0.22 * cloudiness ** 3) * sigma * Tz ** 4
    return lwd, lwu


def nanmean_f(a):
    return float(np.nanmean(a))


def stats_row(date_str, lwd, lwu, rs, sens, lat, atmo, g, mf, point_t_surf):
    """var_classes.py:42-56: the first ten CSV fields."""
    rl = lwd - lwu
    return "%s,%.1f,%.1f,%.1f,%.1f,%.1f,%.1f,%.1f,%.1f,%.2f" % (
        date_str, nanmean_f(rs), nanmean_f(rl), nanmean_f(lwd), nanmean_f(sens), nanmean_f(lat),
        nanmean_f(atmo), nanmean_f(g), nanmean_f(mf), point_t_surf)


CSV_HEADER = ("# DATE format is %Y%m%d, HEAT FLUXES are in W m-2"
              "# ICE and SNOW_MELT are in m w.e."
              "\n# POINT_T_SURF (degree Celsius) is near the point of glacier body temperature measurements"
              "\nDATE,RS_BALANCE,RL_BALANCE,LWD_FLUX,SENSIBLE,LATENT,ATMO_BALANCE,INSIDE_GLACIER_FLUX,"
              "MELT_FLUX,POINT_T_SURF,SNOW_MELT,ICE_MELT,SNOW_COVER,SNOW_COVER_PERCENT_FROM_SURFACE")
# ^ helpers.py:39-45 (the first two comment strings share one line)


# ---- several weather stations + cloud transmissivity (BASELINE config C4) -----------------------
# PARITY UNPINNED UPSTREAM: the reference supports exactly one AWS (model.py:155, var_classes.py:94-125)
# and has no cloud term in the shortwave.  This is THIS REPO's specification; it is
# built so that with no extra station it IS the reference's arithmetic, operation for operation
# (tests/test_oracle_stations.py: bit-identical rasters), and the CUDA path is graded against it.
#
#   stations   k = 0 is the reference's AWS (cfg.elev_aws, cfg.xy_aws, aws_rows): wind, the Monin-Obukhov
#              solve, the exchange coefficients, the lapse rate, the longwave cloudiness and the observed
#              shortwave factor stay its own.  Extra stations k = 1.. carry (row, col) in cell units
#              (cell centres; fractions allowed), an elevation and a series of T_AIR, PRESSURE, HUMID,
#              CLOUDINESS on the same time base.
#   weights    inverse squared distance in cell units, softened by half a cell:
#                  q_k = 1 / ((r - r_k)^2 + (c - c_k)^2 + 0.25),   w_k = q_k / sum_j q_j
#              (array arithmetic in the DEM's dtype; sums run k = 0, 1, ... left to right)
#   reduction  every station value is brought to the cell's elevation with the reference's own
#              formulas (var_classes.py:144-162) and the results are blended:
#                  D      = sum_k w_k (dem - z_k)                  (replaces dem - elev_aws)
#                  t_air  = sum_k w_k T_k + D * lapse
#                  p      = sum_k w_k P_k + D * -0.1145
#                  e      = sum_k e_k * V_k,   V_k = w_k * 10 ** (-(dem - z_k) / 6300)
#              With one station w_0 = 1 exactly (q / q), so D = dem - elev_aws, t_air = T + D * lapse,
#              e = e_0 * 10 ** (-D / 6300): the reference's expressions.
#   cloud SW   Beer-Lambert attenuation by the cloud field relative to the primary station, where the
#              shortwave factor was observed (model.py:500-530):
#                  incoming = potential * factor * exp(-cloud_k * sum_k w_k (N_k - N_0))
#              (N = cloudiness after cloud_corr and clamping; the k = 0 term is zero, so one station or
#              equal cloudiness gives exp(0) = 1 exactly).
def station_fields(dem, cfg, geotransform, stations):
    """Time-invariant rasters of the station blend: (w [n], D, V [n]) in the DEM's dtype."""
    dt = dem.dtype
    h, w_ = dem.shape
    ul_x, x_dist, _, ul_y, _, y_dist = geotransform
    pixel = int((cfg.xy_aws[0] - ul_x) / x_dist)          # raster_utils.py:85-89
    line = -int((ul_y - cfg.xy_aws[1]) / y_dist)
    pos = [(float(line), float(pixel), cfg.elev_aws)] + [(float(s["row"]), float(s["col"]), float(s["elev"])) for s in stations]
    rr = np.arange(h, dtype=dt)[:, None]
    cc = np.arange(w_, dtype=dt)[None, :]
    q = []
    for (r_k, c_k, _) in pos:
        dr = rr - dt.type(r_k)
        dc = cc - dt.type(c_k)
        q.append(dt.type(1) / (dr * dr + dc * dc + dt.type(0.25)))
    tot = q[0]
    for qk in q[1:]:
        tot = tot + qk
    wts = [qk / tot for qk in q]
    D = None
    V = []
    for wk, (_, _, z_k) in zip(wts, pos):
        dz = dem - z_k
        term = wk * dz
        D = term if D is None else D + term
        V.append(wk * 10 ** (-dz / 6300))
    return wts, D, V


def run_model(dem, geotransform, aws_rows, insolation, cfg, *, swe=None, albedo_arrays=None,
              state_dtype=np.float32, keep_steps=None, want_means=False, stations=None, cloud_k=None):
    """The reference's Energy.__init__ + model() on arrays (model.py:19-82, :155-286).

    dem            [H, W] float32 (as shipped) or float64 ("float64-injected")
    insolation     [T, H, W] array or callable(i) -> [H, W]; kWh m-2 per step, the dtype the
                   reference would np.load (model.py:481)
    swe            initial SWE raster or None (zeros, model.py:79)
    albedo_arrays  dict "YYYYmmdd" -> [H, W], ALREADY clipped as load_raster(remove_outliers=True)
    state_dtype    float32 as shipped (model.py:76-80) or float64 (injected)
    Returns the same dict layout as oracle.ref_harness.run_reference.
    """
    params = default_params()
    if cfg.snow_density is not None:
        params["snow_density"] = cfg.snow_density
    multi = stations is not None                          # (an empty list runs the blend with one station)
    if multi:
        st_w, st_D, st_V = station_fields(dem, cfg, geotransform, stations)
    total_snow = np.zeros_like(dem, dtype=state_dtype)
    total_ice = np.zeros_like(dem, dtype=state_dtype)
    if swe is None:
        swe_arr = np.zeros_like(dem, dtype=state_dtype)
    else:
        swe_arr = np.array(swe, copy=True)

    use_msm = cfg.msm is not None
    if use_msm:                                           # model.py:126-143
        depths = list(cfg.msm["depths"])
        layer_t = []
        for t_point in cfg.msm["temperatures"]:
            td = t_point + (dem - cfg.msm["elev"]) * -0.006
            td[td > 0] = 0.0
            layer_t.append(td)
    else:
        depths = []
        layer_t = [np.zeros_like(dem)]                    # the by-hand workaround of SURVEY F9

    csv_lines = [CSV_HEADER]
    solar_lines = []
    rows_out, melt_out, means_out = [], [], []
    n = len(aws_rows)
    for i in range(n):
        row = aws_rows[i]
        date_str = row["DATE"]
        dt = time_step_seconds(aws_rows, i)
        r_hum = unit_guess(float(row["HUMID"]), 100)
        cld = float(row["CLOUDINESS"])
        if cfg.cloud_corr is not None:                    # model.py:200-204
            cld += cfg.cloud_corr
            cld = 1.0 if cld > 1.0 else cld
            cld = 0.0 if cld < 0.0 else cld
        t_surf = layer_t[0]
        try:                                              # model.py:213-221
            grad_temp = float(cfg.temp_lapse_rate)
        except ValueError:
            try:
                grad_temp = float(row["GRADIENT"])
            except KeyError:
                grad_temp = cfg.temp_lapse_rate

        # AwsVars.__post_init__, var_classes.py:80-85
        t_air = float(row["T_AIR"])
        wind = float(row["WIND_SPEED"])
        pressure = float(row["PRESSURE"])
        swd = float(row["SWD"])
        if wind == 0:
            wind = 0.1
        aws_Tz = t_air + 273.15
        aws_P = pressure * 100
        aws_e = r_hum * e_max(aws_Tz, aws_P)

        # DistributedVars.__post_init__, var_classes.py:113-125
        delta_dem = dem - cfg.elev_aws
        if multi:                                         # station blend (specification above)
            st_t, st_p, st_e, st_n = [t_air], [pressure], [aws_e], [cld]
            for st in stations:
                srow = st["rows"][i]
                n_k = float(srow["CLOUDINESS"])
                if cfg.cloud_corr is not None:
                    n_k += cfg.cloud_corr
                    n_k = 1.0 if n_k > 1.0 else n_k
                    n_k = 0.0 if n_k < 0.0 else n_k
                t_k, p_k = float(srow["T_AIR"]), float(srow["PRESSURE"])
                st_t.append(t_k)
                st_p.append(p_k)
                st_e.append(unit_guess(float(srow["HUMID"]), 100) * e_max(t_k + 273.15, p_k * 100))
                st_n.append(n_k)
            mix_t = st_w[0] * st_t[0]
            mix_p = st_w[0] * st_p[0]
            d_e = st_e[0] * st_V[0]
            mix_n = st_w[0] * (st_n[0] - st_n[0])
            for k in range(1, len(st_t)):
                mix_t = mix_t + st_w[k] * st_t[k]
                mix_p = mix_p + st_w[k] * st_p[k]
                d_e = d_e + st_e[k] * st_V[k]
                mix_n = mix_n + st_w[k] * (st_n[k] - st_n[0])
            delta_dem = st_D
            d_t_air = mix_t + delta_dem * grad_temp
            d_pressure = mix_p + delta_dem * -0.1145
        else:
            d_t_air = t_air + delta_dem * grad_temp
            d_pressure = pressure + delta_dem * -0.1145
            d_e = aws_e * 10 ** (-delta_dem / 6300)
        d_Tz = d_t_air + 273.15
        d_Tz_surf = t_surf + 273.15
        d_wind = np.zeros_like(dem, dtype=np.float32)     # :164-173, float32 on purpose (F10)
        d_wind[~np.isnan(dem)] = wind
        d_wind[np.isnan(dem)] = np.nan
        d_P = d_pressure * 100
        d_emax = e_max(d_Tz, d_P)
        d_rh = np.divide(d_e, d_emax)

        albedo = surface_albedo(albedo_arrays, swe_arr, date_str, cfg)

        # calc_energy_fluxes, model.py:340-461
        point_t_surf = sample_at(layer_t[0], geotransform, *cfg.xy_aws)
        point_t_surf += 273.15
        _, _, L = turbulent_fluxes(cfg.z, wind, aws_Tz, aws_P, r_hum, zm=cfg.zm, z_h_or_e=cfg.z_h_or_e,
                                   andreas=cfg.andreas, surface_temp=point_t_surf)
        sens, lat, L = turbulent_fluxes(cfg.z, d_wind, d_Tz, d_P, d_rh, L=L, zm=cfg.zm,
                                        z_h_or_e=cfg.z_h_or_e, andreas=cfg.andreas,
                                        surface_temp=layer_t[0] + 273.15)
        sens = sens * cfg.sensible_corr
        lat = lat * cfg.latent_corr
        lwd, lwu = longwave(d_Tz_surf, d_Tz, cld, eps=cfg.emissivity)

        # calc_shortwave + potential_to_real_insolation_factor, model.py:464-530
        pot = insolation(i) if callable(insolation) else insolation[i]
        pot = np.asarray(pot)
        incoming = kwh_to_w(pot, dt)
        pot_aws = sample_at(pot, geotransform, *cfg.xy_aws)
        pot_aws = kwh_to_w(pot_aws, dt)
        solar_lines.append("\n%s,%s,%s" % (date_str, pot_aws, swd))
        factor = 1 if pot_aws == 0 else swd / pot_aws
        incoming *= factor                    # in place, keeps the raster dtype (model.py:489)
        if multi and cloud_k is not None:     # Beer-Lambert cloud attenuation relative to the primary station
            incoming = incoming * np.exp(-cloud_k * mix_n)
        rs = incoming * (1 - albedo)

        atmo = rs + lwd - lwu + sens + lat                 # model.py:411
        if use_msm:                                        # model.py:428-431
            snow_depth = swe_arr / params["snow_density"]
            layer_t, mf, g = subsurface_tick(depths, layer_t, dt, flux=atmo, snow_depth=snow_depth,
                                             params=params)
        else:                                              # model.py:434-438
            g = np.zeros(atmo.shape)
            mf = atmo + g
            mf[mf < 0] = 0
        line = stats_row(date_str, lwd, lwu, rs, sens, lat, atmo, g, mf, point_t_surf - 273.15)

        snow_melt, ice_melt = melt_partition(mf, swe_arr, dt, params)   # model.py:245
        mean_snow = nanmean_f(snow_melt)
        mean_ice = nanmean_f(ice_melt)
        mean_swe = nanmean_f(swe_arr)
        snow_px = np.sum(swe_arr > 0)
        total_px = np.count_nonzero(~np.isnan(swe_arr))
        cover = round(snow_px / total_px * 100)
        keep = keep_steps is None or i in keep_steps
        if keep:
            rows_out.append(dict(date=date_str, lwd=np.array(lwd), lwu=np.array(lwu), rs=np.array(rs),
                                 sens=np.array(sens), lat=np.array(lat), atmo=np.array(atmo),
                                 g=np.array(g), mf=np.array(mf), point_t_surf=float(point_t_surf - 273.15),
                                 albedo=np.array(albedo), L=float(L), factor=float(factor)))
            melt_out.append((np.array(snow_melt), np.array(ice_melt), np.array(swe_arr)))
        else:
            rows_out.append(None)
            melt_out.append(None)
        if want_means:
            means_out.append([nanmean_f(rs), nanmean_f(lwd - lwu), nanmean_f(lwd), nanmean_f(sens),
                              nanmean_f(lat), nanmean_f(atmo), nanmean_f(g), nanmean_f(mf),
                              mean_snow, mean_ice, mean_swe, float(snow_px), float(total_px)])
        swe_arr -= snow_melt                               # model.py:258-261 (in place, state dtype)
        total_snow += snow_melt
        total_ice += ice_melt
        csv_lines.append("\n%s,%.4f,%.4f,%.4f,%.0f" % (line, mean_snow, mean_ice, mean_swe, cover))

    return dict(stats_csv="".join(csv_lines), solar_csv="".join(solar_lines), rows=rows_out,
                melt=melt_out, swe=swe_arr, total_snow=total_snow, total_ice=total_ice,
                layer_temperatures=layer_t if use_msm else None, means=np.array(means_out),
                numpy=np.__version__, n_steps=n)

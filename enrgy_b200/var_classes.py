"""Host-side views with the reference's names (tepextepex/ENRGY var_classes.py:7-190).

The reference builds these objects for every AWS row and computes with them (eleven full-raster
temporaries per row, var_classes.py:113-125); here the per-cell arithmetic lives in the fused CUDA
kernel, and these classes only exist so that code written against the reference finds what it
expects after `Energy.model()`: `energy.aws`, `energy.vars` (the distributed meteo fields of the LAST
processed row), `energy.output_row`.  The rasters of `DistributedVars` are evaluated lazily, in
NumPy, from the row's scalars and the DEM the first time they are read -- debug views, not the
product path.
"""
from __future__ import annotations

import numpy as np

# var_classes.py:7-15 -- kept as a mutable module global because reference code reads it that way
# (set_density changes it, model.py:84-88)
PARAMS = {
    "ice_density": 900.0,
    "snow_density": 387.0,
    "latent_heat_of_fusion": 3.34 * 10 ** 5,
    "specific_heat_capacity_ice": 2097.0,
    "thermal_diffusivity_ice": 1.16 * 10 ** -6,
    "thermal_diffusivity_snow": 0.40 * 10 ** -6,
    "g": 9.81,
}


def calc_e_max(t_kelvin, p_pa):
    """Magnus formula with the pressure enhancement factor, turbo.py:368-379."""
    t = t_kelvin - 273.15
    p = p_pa / 100
    return 611.2 * np.exp(17.62 * t / (243.12 + t)) * (1.0016 + 3.15 * 10 ** -6 * p - 0.074 / p)


class OutputRow:
    """Area means of one row, printed like var_classes.py:45-56.  Built from the device statistics
    (sums over the glacier cells) instead of from eleven rasters."""

    def __init__(self, date_time_str, means, point_t_surf):
        self.date_time_str = date_time_str
        (self.mean_rs, self.mean_rl, self.mean_lwd, self.mean_sensible, self.mean_latent, self.mean_atmo,
         self.mean_g, self.mean_melt) = [float(x) for x in means]
        self.point_t_surf = point_t_surf

    def __repr__(self):
        return "%s,%.1f,%.1f,%.1f,%.1f,%.1f,%.1f,%.1f,%.1f,%.2f" % (
            self.date_time_str, self.mean_rs, self.mean_rl, self.mean_lwd, self.mean_sensible,
            self.mean_latent, self.mean_atmo, self.mean_g, self.mean_melt, self.point_t_surf)


class AwsVars:
    """The scalars of one AWS row, var_classes.py:59-85 (wind 0 -> 0.1 m/s, Tz, P, e)."""

    def __init__(self, t_air, wind_speed, pressure, rel_humidity, cloudiness, incoming_shortwave, t_surf, grad_temp,
                 elev, x, y, z):
        self.t_air, self.pressure, self.rel_humidity = t_air, pressure, rel_humidity
        self.cloudiness, self.incoming_shortwave = cloudiness, incoming_shortwave
        self.t_surf, self.grad_temp, self.elev, self.x, self.y, self.z = t_surf, grad_temp, elev, x, y, z
        self.wind_speed = 0.1 if wind_speed == 0 else wind_speed
        self.Tz = t_air + 273.15
        self.P = pressure * 100
        self.e = rel_humidity * calc_e_max(self.Tz, self.P)


class DistributedVars:
    """The lapse-rate distribution of one row over the DEM, var_classes.py:94-183: attributes
    `delta_dem, t_air, Tz, t_surf, Tz_surf, wind_speed, pressure, P, e, e_max, rel_humidity`, each
    computed on first access."""

    _FIELDS = ("delta_dem", "t_air", "Tz", "t_surf", "Tz_surf", "wind_speed", "pressure", "P", "e", "e_max",
               "rel_humidity")

    def __init__(self, aws, dem, date_str, export_png=False):
        self.aws, self.dem, self.date_str, self.export_png = aws, dem, date_str, export_png
        self._cache = {}

    def _compute(self, name):
        a = self.aws
        if name == "delta_dem":
            return self.dem - a.elev
        if name == "t_air":
            return a.t_air + self.delta_dem * a.grad_temp
        if name == "Tz":
            return self.t_air + 273.15
        if name == "t_surf":
            return a.t_surf
        if name == "Tz_surf":
            return np.asarray(self.t_surf) + 273.15
        if name == "wind_speed":
            w = np.full(self.dem.shape, a.wind_speed, dtype=np.float32)      # var_classes.py:164-173
            w[np.isnan(self.dem)] = np.nan
            return w
        if name == "pressure":
            return a.pressure + self.delta_dem * -0.1145                      # hPa per metre, var_classes.py:152
        if name == "P":
            return self.pressure * 100
        if name == "e":
            return a.e * 10 ** (-self.delta_dem / 6300)                       # var_classes.py:162
        if name == "e_max":
            return calc_e_max(self.Tz, self.P)
        if name == "rel_humidity":
            return np.divide(self.e, self.e_max)
        raise AttributeError(name)

    def __getattr__(self, name):
        if name in DistributedVars._FIELDS:
            cache = self.__dict__["_cache"]
            if name not in cache:
                cache[name] = self._compute(name)
            return cache[name]
        raise AttributeError(name)

"""Drop-in for the reference's `model.Energy` (tepextepex/ENRGY model.py:18-545).

Same constructor, `add_*` / `set_*` methods, hand-set attributes and `model()` signature; results
are delivered the same way (attributes `total_ice_melt_array`, `total_snow_melt_array`,
`swe_array`, files `heat_fluxes.csv`, `solar_output.csv`, `"<DATE> total_melt_ice|total_melt_snow|
remaining_snow_cover.tiff"`).  The time loop itself (model.py:183-283) runs as fused CUDA kernels
behind the C ABI of include/enrgy_b200.h; this file only does what the reference does on the host
around it: reading files, parsing rows, formatting CSV lines, exporting rasters.

Differences, all additive:
  * `Energy(..., precision="f32"|"f64", device=0)`: arithmetic of the device path.  "f32" mirrors
    the as-shipped float32 rasters/state, "f64" the float64 evaluation.
  * potential insolation: with `use_precomputed` the per-step rasters are read exactly where the
    reference reads them (pickle dir `.npy`, else `<dem_dir>/<DATE>_total.sdat`) and streamed to
    the GPU; without it the reference shells out to SAGA GIS per step (saga_lighting.py:7-53) --
    here the fused kernel computes insolation and the shading ray march itself.
  * no PNG previews (matplotlib, raster_utils.py:9-32) and no CPU fallback.
"""
from __future__ import annotations

import os
from datetime import datetime

import numpy as np

from . import _lib
from .engine import Engine
from .forcing import build_forcing, read_input_file
from .geo import coords_to_index, get_value_by_real_coords, grid_centre_latlon
from .helpers import fill_header
from .raster_utils import export_array_as_geotiff, load_raster, show_me  # noqa: F401

# var_classes.py:7-15 -- kept as a mutable module global because reference code reads it that way
PARAMS = {
    "ice_density": 900.0,
    "snow_density": 387.0,
    "latent_heat_of_fusion": 3.34 * 10 ** 5,
    "specific_heat_capacity_ice": 2097.0,
    "thermal_diffusivity_ice": 1.16 * 10 ** -6,
    "thermal_diffusivity_snow": 0.40 * 10 ** -6,
    "g": 9.81,
}


class OutputRow:
    """Area means of one step, printed like reference var_classes.py:45-56."""

    def __init__(self, date_time_str, means, point_t_surf):
        self.date_time_str = date_time_str
        (self.mean_rs, self.mean_rl, self.mean_lwd, self.mean_sensible, self.mean_latent, self.mean_atmo,
         self.mean_g, self.mean_melt) = [float(x) for x in means]
        self.point_t_surf = point_t_surf

    def __repr__(self):
        return "%s,%.1f,%.1f,%.1f,%.1f,%.1f,%.1f,%.1f,%.1f,%.2f" % (
            self.date_time_str, self.mean_rs, self.mean_rl, self.mean_lwd, self.mean_sensible,
            self.mean_latent, self.mean_atmo, self.mean_g, self.mean_melt, self.point_t_surf)


def _div(a, b):
    return float(a) / float(b) if b else float("nan")


class Energy:
    def __init__(self, base_dem_path, glacier_outlines_path, out_dir, res=None, precision="f32", device=0):
        self.params = PARAMS
        self.current_date_str = None
        self.input_list = []
        self.output_row = None
        self.debug_point_output = None
        self.res = 100 if res is None else res                      # model.py:30-33
        if not os.path.isdir(out_dir):
            os.mkdir(out_dir)
        self.out_dir = out_dir
        self.png_export = 1
        self.result_export_dates = None
        self.export_potential = False
        self.use_precomputed = False
        self.aws = None
        self.albedo_arrays = None
        self.cloud_corr = None
        self.sensible_corr_factor = 1
        self.latent_corr_factor = 1
        self.stake_df = None
        self.out_stake_df = None
        self.use_msm = False
        self.msm_xy = None
        self.layer_temperatures = None
        self.layer_depths = []
        self.potential_incoming_sr_path = None
        self.pickle_dir = None
        self.base_dem_path = base_dem_path
        self.outlines_path = glacier_outlines_path
        # device path options (additive)
        self.precision = precision
        self.device = device
        self.shadow = True              # SAGA -SHADOW 1 (saga_lighting.py:43)
        self.lat = None                 # grid reference latitude / longitude; None = grid centre, UTM 33N
        self.lon = None
        self.max_resident_insolation_bytes = 4 << 30
        self.stats = None               # [T, S_COUNT] sums of the last model() call
        self.point_scalars = None
        self._engine = None
        self._last = None

        print("Loading base DEM...")
        self.base_dem_array, self.geotransform, self.projection = load_raster(
            base_dem_path, self.outlines_path, self.res, v=False)
        self.total_snow_melt_array = np.zeros_like(self.base_dem_array, dtype=np.float32)
        self.total_ice_melt_array = np.zeros_like(self.base_dem_array, dtype=np.float32)
        self.swe_array = np.zeros_like(self.base_dem_array, dtype=np.float32)
        self._swe_given = False

    # ---- configuration, same semantics as model.py:84-153 -----------------------------------------
    def set_density(self, snow=None, ice=None):
        if snow is not None:
            self.params["snow_density"] = snow
        if ice is not None:
            self.params["ice_density"] = ice

    def add_cloud_corr(self, cloud_corr):
        if (float(cloud_corr) < -1.0) or (float(cloud_corr) > 1.0):
            raise ValueError("cloud_corr value should be a float between [-1.0..+1.0]")
        self.cloud_corr = cloud_corr

    def add_pickle_dir(self, pickle_dir):
        self.pickle_dir = os.path.join(pickle_dir, str(self.res))
        if not os.path.exists(self.pickle_dir):
            raise IOError(f"Cannot find pickled insolation for {self.res} m resolution inside {pickle_dir}!"
                          f"Please choose directory containing pickles or change the spatial resolution.")

    def add_stakes(self, file_path):
        import pandas as pd
        self.stake_df = pd.read_csv(file_path)
        self.out_stake_df = pd.DataFrame({"name": self.stake_df["name"]})

    def write_stakes(self, out_file_path):
        out_path = os.path.join(os.path.dirname(out_file_path), "ice_melt_point.csv")
        self.out_stake_df.to_csv(out_path, index=False, float_format="%.3f")

    def sample_stakes(self):
        vals = []
        for row in self.stake_df.itertuples():
            try:
                value = get_value_by_real_coords(self.total_ice_melt_array, self.geotransform,
                                                 row.easting, row.northing)
                value = round(value, 4)
            except Exception:
                value = None
            vals.append(value)
        self.out_stake_df[self.current_date_str] = vals

    def add_snow(self, swe_map_path):
        print("Initialized snow cover state (SWE) from %s" % swe_map_path)
        self.swe_array = load_raster(swe_map_path, self.outlines_path, self.res, v=False)[0]
        self._swe_given = True

    def add_msm(self, depths, temperatures, elev_aws):
        """Sub-surface model set-up, reference model.py:126-149: `depths` are layer THICKNESSES,
        `temperatures` the boundary temperatures at elevation `elev_aws`; they are distributed with
        -0.006 K/m and capped at 0 degC on the device."""
        print("Initializing subsurface model...")
        self.use_msm = True
        self.layer_depths = list(depths)
        self._msm_point_temps = list(temperatures)
        self._msm_elev = elev_aws
        if len(self._msm_point_temps) != len(self.layer_depths) + 1:
            raise ValueError("temperatures are layer BOUNDARIES: one more than depths")
        delta = self.base_dem_array - elev_aws                     # host copy for callers that read it
        self.layer_temperatures = []
        for t_point in temperatures:
            t = t_point + delta * -0.006
            t[t > 0] = 0.0
            self.layer_temperatures.append(t)

    def add_checkpoints(self, date_str_list):
        self.result_export_dates = [s + " 12:00:00" for s in date_str_list]

    # ---- the model run ----------------------------------------------------------------------------
    def model(self, aws_file=None, albedo_maps=None, z=2.0, elev_aws=0.0, xy_aws=None,
              zm=None, z_h_or_e=None, andreas=False,
              solar_only=False, const_albedo=None, temp_lapse_rate=-0.006, last_snowfall=None,
              max_ice_albedo=None, emissivity=None, v=True):
        if aws_file is None:
            return
        if solar_only:
            raise NotImplementedError("solar_only is a debugging mode of the reference (model.py:400-405); "
                                      "not part of the accelerated path")
        if albedo_maps is not None:
            self.albedo_arrays = {}
            for key in albedo_maps:
                self.albedo_arrays[key] = load_raster(albedo_maps[key], self.outlines_path, self.res,
                                                      remove_outliers=True, v=v)[0]
        out_file = os.path.join(self.out_dir, "heat_fluxes.csv")
        fill_header(out_file)
        if self.debug_point_output is not None:
            with open(os.path.join(self.out_dir, self.debug_point_output), "a") as f:
                f.write("SENSIBLE,LATENT")

        self.input_list = read_input_file(aws_file)
        rows = self.input_list
        n_steps = len(rows)
        keys = list(self.albedo_arrays) if (const_albedo is None and self.albedo_arrays) else None
        if const_albedo is None and keys is None:
            raise ValueError("either albedo_maps or const_albedo is needed")
        table = build_forcing(rows, keys, temp_lapse_rate=temp_lapse_rate, cloud_corr=self.cloud_corr,
                              last_snowfall=last_snowfall if const_albedo is None else None)

        h, w = self.base_dem_array.shape
        aws_row, aws_col = coords_to_index(self.geotransform, xy_aws[0], xy_aws[1])
        if not (0 <= aws_row < h and 0 <= aws_col < w):
            raise IndexError("AWS coordinates fall outside the model grid")
        streamed = bool(self.use_precomputed)
        lat, lon = self.lat, self.lon
        if not streamed and (lat is None or lon is None):
            lat, lon = grid_centre_latlon(self.geotransform, h, w)
        eng = Engine(h, w, precision=_lib.F64 if self.precision == "f64" else _lib.F32, device=self.device)
        self._engine = eng
        try:
            eng.set_params(cell_size=abs(self.geotransform[1]), elev_aws=elev_aws, aws_row=aws_row,
                           aws_col=aws_col, sensor_z=z, zm=zm, z_h_or_e=z_h_or_e, andreas=andreas,
                           sensible_corr=self.sensible_corr_factor, latent_corr=self.latent_corr_factor,
                           emissivity=emissivity, const_albedo=const_albedo, max_ice_albedo=max_ice_albedo,
                           snow_density=self.params["snow_density"], ice_density=self.params["ice_density"],
                           insol_mode=_lib.INSOL_STREAMED if streamed else _lib.INSOL_COMPUTED,
                           shadow=self.shadow, lat=lat or 0.0, lon=lon or 0.0,
                           msm_depths=self.layer_depths if self.use_msm else None)
            eng.set_dem(self.base_dem_array)
            if self.use_msm:
                eng.set_msm(self._msm_point_temps, self._msm_elev)
            eng.set_forcing(table)    # the library starts its host pre-pass here, under the raster uploads
            if keys is not None:
                eng.set_albedo_maps([self.albedo_arrays[k] for k in keys])
            if self._swe_given:
                eng.set_swe(self.swe_array)

            # step ranges: cut at the checkpoint rows (model.py:279-283) and, for streamed
            # insolation, at the residency limit
            cuts = {n_steps}
            if self.result_export_dates is not None:
                for i, row in enumerate(rows):
                    if row["DATE"] in self.result_export_dates:
                        cuts.add(i + 1)
            max_chunk = n_steps
            if streamed:
                max_chunk = max(1, int(self.max_resident_insolation_bytes // (h * w * 4)))
            ranges, t = [], 0
            for c in sorted(cuts):
                while t < c:
                    e = min(c, t + max_chunk)
                    ranges.append((t, e))
                    t = e
            stats = np.zeros((n_steps, _lib.S_COUNT), dtype=np.float64)
            point = None
            if not streamed:
                eng.prepass()
                point = eng.point_scalars()
            solar_file = os.path.join(self.out_dir, "solar_output.csv")
            for (t0, t1) in ranges:
                if streamed:
                    eng.set_insolation(t0, self._read_insolation(rows, t0, t1, v))
                    eng.prepass()
                    point = eng.point_scalars()
                stats[t0:t1] = eng.run(t0, t1)
                self._write_rows(rows, t0, t1, stats, point, out_file, solar_file, table)
                self.current_date_str = rows[t1 - 1]["DATE"]
                if self.result_export_dates is not None and self.current_date_str in self.result_export_dates:
                    self._pull_state(eng)
                    self.export_result()
                    if self.stake_df is not None:
                        self.sample_stakes()
                        self.write_stakes(out_file)
            self._pull_state(eng)
            if self.use_msm:
                self.layer_temperatures = [t.astype(np.float32) for t in eng.layer_temps()]
            self.stats = stats
            self.point_scalars = point
        finally:
            eng.close()
            self._engine = None
        self.export_result()                                        # model.py:285-286

    # ---- helpers ------------------------------------------------------------------------------------
    def _read_insolation(self, rows, t0, t1, v):
        """Per-step potential insolation rasters [kWh m-2], read where model.py:465-481 reads them."""
        h, w = self.base_dem_array.shape
        out = np.empty((t1 - t0, h, w), dtype=np.float32)
        for i in range(t0, t1):
            path = os.path.join(os.path.dirname(self.base_dem_path), "%s_total.sdat" % rows[i]["DATE"])
            self.potential_incoming_sr_path = path
            if self.pickle_dir is None:
                out[i - t0] = load_raster(path, self.outlines_path, self.res, v=v)[0]
            else:
                out[i - t0] = np.load(os.path.join(self.pickle_dir, f"{os.path.basename(path)}.npy"))
        return out

    def _write_rows(self, rows, t0, t1, stats, point, out_file, solar_file, table):
        f32 = self.precision != "f64"
        with open(out_file, "a") as output, open(solar_file, "a") as solar:
            for i in range(t0, t1):
                s = stats[i]
                nv = s[_lib.S_NVALID]
                means = [_div(s[_lib.S_RS], nv), _div(s[_lib.S_LWD] - s[_lib.S_LWU], nv), _div(s[_lib.S_LWD], nv),
                         _div(s[_lib.S_SENS], nv), _div(s[_lib.S_LAT], nv), _div(s[_lib.S_ATMO], nv),
                         _div(s[_lib.S_G], nv), _div(s[_lib.S_MELT], nv)]
                self.output_row = OutputRow(rows[i]["DATE"], means, point[i, _lib.P_TSURF_AWS])
                mean_snow = _div(s[_lib.S_SNOW], nv)
                mean_ice = _div(s[_lib.S_ICE], nv)
                mean_swe = _div(s[_lib.S_SWE], s[_lib.S_NSWE])
                cover = round(_div(s[_lib.S_NSNOW], s[_lib.S_NSWE]) * 100) if s[_lib.S_NSWE] else 0
                output.write("\n%s,%.4f,%.4f,%.4f,%.0f" % (str(self.output_row), mean_snow, mean_ice,
                                                          mean_swe, cover))
                pot = point[i, _lib.P_POT_AWS]
                pot = np.float32(pot) if f32 and self.use_precomputed else pot
                solar.write("\n%s,%s,%s" % (rows[i]["DATE"], pot, table[i, _lib.F_SWD]))   # model.py:519-520

    def _pull_state(self, eng):
        swe, tsn, tic = eng.state(np.float32)
        self.swe_array, self.total_snow_melt_array, self.total_ice_melt_array = swe, tsn, tic

    def export_result(self):
        """model.py:288-295."""
        arrays = (self.total_ice_melt_array, self.total_snow_melt_array, self.swe_array)
        titles = ("total_melt_ice", "total_melt_snow", "remaining_snow_cover")
        for arr, title in zip(arrays, titles):
            export_array_as_geotiff(arr, self.geotransform, self.projection,
                                    os.path.join(self.out_dir, "%s %s.tiff" % (self.current_date_str, title)))
        print("Result saved as GeoTIFF")

    # ---- config_template.json -------------------------------------------------------------------------
    @classmethod
    def from_config(cls, config, precision="f32", device=0):
        """Builds an Energy object and the keyword arguments of model() from a dict / JSON file laid
        out like the reference's config_template.json (which the reference itself never loads,
        SURVEY.md F6).  Returns (energy, model_kwargs)."""
        import json
        if isinstance(config, (str, os.PathLike)):
            with open(config) as f:
                config = json.load(f)
        inp, out = config["input"], config["output"]
        e = cls(inp["dem"], inp.get("outlines"), out["out_dir"], res=out.get("resolution"),
                precision=precision, device=device)
        e.debug_point_output = out.get("debug_point_output")
        if out.get("png_export") is not None:
            e.png_export = out["png_export"]
        if out.get("stake_coords"):
            e.add_stakes(out["stake_coords"])
        if out.get("dates"):
            e.add_checkpoints(out["dates"])
        alb = config.get("albedo", {})
        solar = config.get("solar", {})
        e.use_precomputed = bool(solar.get("use_precomputed", False))
        if solar.get("pickles"):
            e.add_pickle_dir(solar["pickles"])
        turbo = config.get("turbo", {})
        e.sensible_corr_factor = turbo.get("sensible_corr_factor", 1)
        e.latent_corr_factor = turbo.get("latent_corr_factor", 1)
        lw = config.get("longwave", {})
        if lw.get("cloud_corr") is not None:
            e.add_cloud_corr(lw["cloud_corr"])
        snow = config.get("snow", {})
        if snow.get("use"):
            if snow.get("density") is not None:
                e.set_density(snow=snow["density"])
            if snow.get("swe_grid"):
                e.add_snow(snow["swe_grid"])
        msm = config.get("msm", {})
        if msm.get("use"):
            e.add_msm(msm["depths"], msm["temperatures"], msm["elev"])
            e.msm_xy = tuple(msm["xy"]) if msm.get("xy") else None
        aws = inp["aws"]
        lapse = inp.get("vertical_lapse_rates", {}).get("t_air", -0.006)
        kwargs = dict(aws_file=aws["file"], z=aws.get("sensor_z", 2.0), elev_aws=aws.get("elev", 0.0),
                      xy_aws=tuple(aws["xy"]), zm=turbo.get("zm"), z_h_or_e=turbo.get("z_h_or_e"),
                      andreas=bool(turbo.get("andreas", False)), temp_lapse_rate=lapse,
                      emissivity=lw.get("emissivity"), v=bool(out.get("verbose", False)))
        if alb.get("use_const"):
            kwargs["const_albedo"] = tuple(alb["const"]) if alb.get("const") else (0.35, 0.75)
        else:
            kwargs["albedo_maps"] = alb.get("albedo_maps")
            kwargs["last_snowfall"] = alb.get("last_snowfall")
            kwargs["max_ice_albedo"] = alb.get("max_ice_albedo")
        return e, kwargs

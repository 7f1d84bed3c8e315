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
    here the fused kernel computes the insolation and a line sweep (csrc/shade.cu) the shading masks.
  * no PNG previews (matplotlib, raster_utils.py:9-32) and no CPU fallback.
  * several GPUs: under `torchrun` (torch.distributed initialised, one process per GPU) every rank
    builds the same Energy object and calls model(); the raster is cut into row bands balanced by
    visited tiles, every rank runs its band, the per-row area sums are all-reduced, rank 0 writes the
    CSV files and exports, and every rank ends up with the full state rasters.  With shading the sweep
    is sharded by sub-step and the mask rows are exchanged (parallel.ShardedShading).
  * terrain outside the glacier outline: the reference hands SAGA the UNCROPPED DEM (model.py:469 ->
    saga_lighting.py:42), so valley walls shade the glacier.  With GDAL the uncropped raster is read
    on the model grid automatically; for `.npy` rasters call `add_terrain(path_or_array)` -- without
    it only the glacier's own relief casts shadows.
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

from .var_classes import PARAMS, AwsVars, DistributedVars, OutputRow  # noqa: F401  (same names as var_classes.py)


def _div(a, b):
    return float(a) / float(b) if b else float("nan")


class Energy:
    def __init__(self, base_dem_path, glacier_outlines_path, out_dir, res=None, precision="f32", device=None):
        self.params = PARAMS
        self.current_date_str = None
        self.input_list = []
        self.output_row = None
        self.debug_point_output = None
        self.res = 100 if res is None else res                      # model.py:30-33
        os.makedirs(out_dir, exist_ok=True)                          # (every rank of a torchrun job gets here at once)
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
        self.stats = None               # [T, S_COUNT] sums of the last model() call (all ranks: reduced)
        self.point_scalars = None
        self.vars = None                # DistributedVars of the last processed row (model.py:232), lazy NumPy
        self.albedo = None              # albedo / incoming shortwave rasters of the last row (model.py:235, :408)
        self.incoming_shortwave = None  #   ... filled when debug_views is on
        self.debug_views = None         # None: on for rasters up to 4 Mi cells; True / False to force
        self.bands = None               # row bands of the last model() call [(row0, rows)] (one per rank)
        self.terrain_array = None       # uncropped terrain on the model grid (add_terrain)
        self._engine = None
        self._engine_factory = Engine   # (tests inject a stand-in here)
        self._last = None

        print("Loading base DEM...")
        self.base_dem_array, self.geotransform, self.projection = load_raster(
            base_dem_path, self.outlines_path, self.res, v=False)
        self.total_snow_melt_array = np.zeros_like(self.base_dem_array, dtype=np.float32)
        self.total_ice_melt_array = np.zeros_like(self.base_dem_array, dtype=np.float32)
        self.swe_array = np.zeros_like(self.base_dem_array, dtype=np.float32)
        # SAGA sees the uncropped DEM (model.py:469): read it on the model grid where GDAL can
        if not isinstance(base_dem_path, np.ndarray) and not str(base_dem_path).endswith(".npy"):
            try:
                self.add_terrain(None)
            except Exception as e:                                   # pragma: no cover  (needs GDAL)
                print("uncropped terrain not loaded (%s): only the glacier's own relief will cast shadows" % e)

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

    # ---- beyond the reference (BASELINE config C4): several weather stations, cloud attenuation ------
    def add_station(self, aws_file, xy, elev):
        """An extra weather station: a CSV on the time base of the main AWS file with T_AIR, PRESSURE,
        HUMID, CLOUDINESS columns, its real-world coordinates and elevation.  Temperature, pressure and
        vapour pressure are then blended per cell over the main AWS and the extra stations (inverse
        squared distance, each reduced to the cell's elevation with the reference's lapse formulas);
        wind, exchange coefficients, longwave cloudiness and the observed shortwave factor stay the main
        station's.  Specification: oracle/enrgy_oracle.py "several weather stations" (the reference has
        one AWS, model.py:155).  At most three."""
        if not hasattr(self, "stations"):
            self.stations = []
        if len(self.stations) >= 3:
            raise ValueError("at most three extra stations")
        self.stations.append(dict(aws_file=aws_file, xy=(float(xy[0]), float(xy[1])), elev=float(elev)))

    def add_cloud_transmissivity(self, k):
        """Beer-Lambert attenuation of the incoming shortwave by the blended cloud field relative to the
        main station's cloudiness: x exp(-k (N_cell - N_aws)).  Needs at least one add_station to differ
        from the reference."""
        if float(k) < 0:
            raise ValueError("cloud transmissivity coefficient must be >= 0")
        self.cloud_k = float(k)

    def _station_setup(self, n_steps):
        """[(row, col, elev)] in cell units of the model grid (cell centres: the fractional position of the
        station minus one half) and the series of every extra station."""
        from .forcing import build_station_series
        ul_x, x_dist, _, ul_y, _, y_dist = self.geotransform
        pos, series = [], []
        for st in getattr(self, "stations", []):
            rows = read_input_file(st["aws_file"])
            if len(rows) != n_steps:
                raise ValueError("station file %s has %d rows, the main AWS file %d" % (st["aws_file"], len(rows), n_steps))
            col = (st["xy"][0] - ul_x) / x_dist - 0.5
            row = (st["xy"][1] - ul_y) / y_dist - 0.5
            pos.append((row, col, st["elev"]))
            series.append(build_station_series(rows, cloud_corr=self.cloud_corr))
        return pos, series

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

    def add_terrain(self, terrain=None):
        """The UNCROPPED terrain on the model grid: shadow casters and slope neighbours outside the
        glacier outline (the reference gives SAGA the uncropped DEM file, model.py:469 ->
        saga_lighting.py:42, and crops the result).  `terrain`: an array or a `.npy` path on the model
        grid; None re-reads `base_dem_path` without the cutline onto the cropped raster's grid (GDAL)."""
        if terrain is None:
            from .raster_utils import load_uncropped_like
            arr = load_uncropped_like(self.base_dem_path, self.geotransform, self.base_dem_array.shape)
        else:
            arr = load_raster(terrain, None, self.res, v=False)[0]
        if arr.shape != self.base_dem_array.shape:
            raise ValueError("terrain raster %s does not match the model grid %s" % (arr.shape, self.base_dem_array.shape))
        self.terrain_array = np.ascontiguousarray(arr, dtype=np.float32)

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
    @staticmethod
    def _dist():
        """(world, rank, torch.distributed or None): several ranks only under an initialised process group."""
        try:
            import torch.distributed as dist
        except Exception:                                            # pragma: no cover
            return 1, 0, None
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            return dist.get_world_size(), dist.get_rank(), dist
        return 1, 0, None

    def _collective_device(self, dist):
        """Tensors of a collective live on the GPU with NCCL, on the host with gloo."""
        import torch
        return torch.device("cuda", self._device_index()) if dist.get_backend() == "nccl" else torch.device("cpu")

    def _device_index(self):
        if self.device is not None:
            return int(self.device)
        return int(os.environ.get("LOCAL_RANK", "0"))

    def model(self, aws_file=None, albedo_maps=None, z=2.0, elev_aws=0.0, xy_aws=None,
              zm=None, z_h_or_e=None, andreas=False,
              solar_only=False, const_albedo=None, temp_lapse_rate=-0.006, last_snowfall=None,
              max_ice_albedo=None, emissivity=None, v=True):
        if aws_file is None:
            return
        if solar_only:
            raise NotImplementedError("solar_only is a debugging mode of the reference (model.py:400-405); "
                                      "not part of the accelerated path")
        world, rank, dist = self._dist()
        writer = rank == 0                                          # rank 0 writes every file
        if albedo_maps is not None:
            self.albedo_arrays = {}
            for key in albedo_maps:
                self.albedo_arrays[key] = load_raster(albedo_maps[key], self.outlines_path, self.res,
                                                      remove_outliers=True, v=v)[0]
        out_file = os.path.join(self.out_dir, "heat_fluxes.csv")
        if writer:
            fill_header(out_file)
        if self.use_msm and self.msm_xy is not None and self.debug_point_output is not None:
            if coords_to_index(self.geotransform, *self.msm_xy) != coords_to_index(self.geotransform, *xy_aws):
                raise NotImplementedError("debug_point_output prints the layer temperatures at msm_xy (model.py:421-426); "
                                          "the accelerated path keeps them per step for the AWS cell only")
        if self.debug_point_output is not None and writer:           # model.py:170-180
            header = ""
            with open(os.path.join(self.out_dir, self.debug_point_output), "a") as f:
                if self.use_msm:
                    cur_depth = 0.0
                    header += f"{cur_depth},"
                    for layer_thickness in self.layer_depths:
                        cur_depth += layer_thickness
                        header += f"{cur_depth},"
                header += "SENSIBLE,LATENT"
                f.write(header)
        if not self.use_msm and self.layer_temperatures is not None:
            if any(np.nanmax(np.abs(t)) > 0 for t in self.layer_temperatures[:1]):
                raise NotImplementedError("a hand-set surface temperature without add_msm (model.py:207-210) is not "
                                          "supported: without the sub-surface model the surface is at 0 degC")

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
        shading = bool(self.shadow) and not streamed
        lat, lon = self.lat, self.lon
        if not streamed and (lat is None or lon is None):
            lat, lon = grid_centre_latlon(self.geotransform, h, w)
        # row bands, one per rank: equal numbers of visited tiles
        if world > 1:
            from .parallel import row_bands, tile_cost_per_row
            bands = row_bands(h, world, align=16, valid_per_row=tile_cost_per_row(~np.isnan(self.base_dem_array)))
        else:
            bands = [(0, h)]
        self.bands = bands
        r0, nr = bands[rank]
        band = slice(r0, r0 + nr)
        eng = self._engine_factory(h, w, precision=_lib.F64 if self.precision == "f64" else _lib.F32,
                                   device=self._device_index())
        self._engine = eng
        try:
            eng.set_params(cell_size=abs(self.geotransform[1]), elev_aws=elev_aws, aws_row=aws_row,
                           aws_col=aws_col, sensor_z=z, zm=zm, z_h_or_e=z_h_or_e, andreas=andreas,
                           sensible_corr=self.sensible_corr_factor, latent_corr=self.latent_corr_factor,
                           emissivity=emissivity, const_albedo=const_albedo, max_ice_albedo=max_ice_albedo,
                           snow_density=self.params["snow_density"], ice_density=self.params["ice_density"],
                           insol_mode=_lib.INSOL_STREAMED if streamed else _lib.INSOL_COMPUTED,
                           shadow=shading, lat=lat or 0.0, lon=lon or 0.0,
                           msm_depths=self.layer_depths if self.use_msm else None,
                           band_row0=r0 if world > 1 else 0, band_rows=nr if world > 1 else 0)
            eng.set_dem(self.base_dem_array)
            if self.terrain_array is not None and not streamed:
                eng.set_terrain(self.terrain_array)
            if self.use_msm:
                eng.set_msm(self._msm_point_temps, self._msm_elev)
            eng.set_forcing(table)    # the library starts its host pre-pass here, under the raster uploads
            if getattr(self, "stations", None) or getattr(self, "cloud_k", None) is not None:
                pos, series = self._station_setup(n_steps)
                eng.set_stations(pos, series, cloud_k=getattr(self, "cloud_k", None))
            if keys is not None:
                eng.set_albedo_maps([self.albedo_arrays[k][band] for k in keys])
            # the SWE raster as it stands (model.py:245-258 reads self.swe_array; zeros by default) and the
            # melt totals of earlier model() calls (model.py:260-261 accumulates across calls)
            eng.set_swe(self.swe_array[band])
            if np.any(np.nan_to_num(self.total_snow_melt_array) != 0) or np.any(np.nan_to_num(self.total_ice_melt_array) != 0):
                eng.set_state(total_snow=self.total_snow_melt_array[band], total_ice=self.total_ice_melt_array[band])
            if self.use_msm and world > 1:
                # every rank integrates the AWS cell in its pre-pass: hand it that cell's own values
                eng.set_aws_cell([float(self.albedo_arrays[k][aws_row, aws_col]) for k in keys] if keys else None,
                                 float(self.swe_array[aws_row, aws_col]))

            # step ranges: cut at the checkpoint rows (model.py:279-283) and, for streamed
            # insolation, at the residency limit
            cuts = {n_steps}
            if self.result_export_dates is not None:
                for i, row in enumerate(rows):
                    if row["DATE"] in self.result_export_dates:
                        cuts.add(i + 1)
            max_chunk = n_steps
            if streamed:
                max_chunk = max(1, int(self.max_resident_insolation_bytes // (nr * w * 4)))
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
            sharded = None
            if shading and world > 1:
                import torch
                from .parallel import ShardedShading
                stream = torch.cuda.Stream(self._device_index())
                eng.set_stream(stream.cuda_stream)
                sharded = (ShardedShading(eng, bands, rank, world), stream,
                           point[:, _lib.P_NSUB].astype(int))
            want_views = self.debug_views if self.debug_views is not None else (h * w <= (1 << 22))
            solar_file = os.path.join(self.out_dir, "solar_output.csv")
            layers_pt = eng.point_layers() if (self.use_msm and self.debug_point_output is not None and not streamed) else None
            for (t0, t1) in ranges:
                if streamed:
                    pot_rows = self._read_insolation(rows, t0, t1, v)
                    eng.set_insolation(t0, pot_rows[:, band])
                    if world > 1:                # the shortwave factor needs the AWS cell, whichever band holds it
                        eng.set_insolation_aws(t0, pot_rows[:, aws_row, aws_col])
                    eng.prepass()
                    point = eng.point_scalars()
                    if self.use_msm and self.debug_point_output is not None:
                        layers_pt = eng.point_layers()
                last = t1 == n_steps and want_views and sharded is None
                if last and t1 - 1 > t0:
                    eng.defer_snow_total(True)       # two launches, rasters as from one
                    stats[t0:t1 - 1] = eng.run(t0, t1 - 1)
                    eng.defer_snow_total(False)
                if last:
                    self._last_views(eng, t1 - 1, band, world, dist)
                    stats[t1 - 1:t1] = eng.run(t1 - 1, t1)
                elif sharded is not None:
                    stats[t0:t1] = self._run_sharded(eng, sharded, t0, t1)
                else:
                    stats[t0:t1] = eng.run(t0, t1)
                if world > 1:
                    stats[t0:t1] = self._allreduce(stats[t0:t1], dist)
                if writer:
                    self._write_rows(rows, t0, t1, stats, point, out_file, solar_file, table, layers_pt)
                self.current_date_str = rows[t1 - 1]["DATE"]
                if self.result_export_dates is not None and self.current_date_str in self.result_export_dates:
                    self._pull_state(eng, bands, rank, dist)
                    if writer:
                        self.export_result()
                        if self.stake_df is not None:
                            self.sample_stakes()
                            self.write_stakes(out_file)
            self._pull_state(eng, bands, rank, dist)
            if self.use_msm:
                lt = eng.layer_temps()
                self.layer_temperatures = [self._gather_rows(t.astype(np.float32), bands, rank, dist) for t in lt]
            self.stats = stats
            self.point_scalars = point
            # the AwsVars / DistributedVars of the last row, as the reference leaves them behind (model.py:229-232)
            f = table[-1]
            t_surf = self.layer_temperatures[0] if self.use_msm else np.zeros(self.base_dem_array.shape)
            self.aws = AwsVars(f[_lib.F_T_AIR], f[_lib.F_WIND], f[_lib.F_PRESSURE], f[_lib.F_RH], f[_lib.F_CLOUD],
                               f[_lib.F_SWD], t_surf, f[_lib.F_LAPSE], elev_aws, xy_aws[0], xy_aws[1], z)
            self.vars = DistributedVars(self.aws, self.base_dem_array, self.current_date_str, False)
        finally:
            eng.close()
            self._engine = None
        if writer:
            self.export_result()                                    # model.py:285-286

    # ---- helpers ------------------------------------------------------------------------------------
    def _run_sharded(self, eng, sharded, t0, t1):
        """Rows [t0, t1) with shading over several GPUs: this band's statistics sums (not yet reduced)."""
        import torch
        sh, stream, sub_counts = sharded
        d_stats = torch.zeros((t1 - t0, _lib.S_COUNT), dtype=torch.float64, device="cuda:%d" % self._device_index())
        stream.wait_stream(torch.cuda.current_stream(self._device_index()))
        sh.run(t0, t1, d_stats.data_ptr(), stream, sub_counts)
        stream.synchronize()
        return d_stats.cpu().numpy()

    def _allreduce(self, arr, dist):
        import torch
        t = torch.from_numpy(np.ascontiguousarray(arr)).to(self._collective_device(dist))
        dist.all_reduce(t)
        return t.cpu().numpy()

    def _gather_rows(self, part, bands, rank, dist):
        """The full raster from every rank's band (all ranks get it)."""
        if dist is None:
            return part
        import torch
        dev = self._collective_device(dist)
        mx = max(n for _, n in bands)
        buf = torch.full((mx, part.shape[1]), float("nan"), dtype=torch.from_numpy(part).dtype, device=dev)
        buf[:part.shape[0]] = torch.from_numpy(np.ascontiguousarray(part)).to(dev)
        out = [torch.empty_like(buf) for _ in bands]
        dist.all_gather(out, buf)
        return np.concatenate([o[:n].cpu().numpy() for o, (_, n) in zip(out, bands)], axis=0)

    def _last_views(self, eng, step, band, world, dist):
        """Albedo and incoming shortwave rasters of the last row (model.py:235-236, :408), from the
        debug view of the kernel's own arithmetic."""
        d = eng.dump_steps(step, step + 1)[0]
        with np.errstate(invalid="ignore", divide="ignore"):
            alb = d[_lib.D_ALBEDO]
            inc = d[_lib.D_RS] / (1.0 - alb)                        # rs = incoming * (1 - albedo), model.py:497
        if world > 1:
            alb = self._gather_rows(alb, self.bands, dist.get_rank(), dist)
            inc = self._gather_rows(inc, self.bands, dist.get_rank(), dist)
        self.albedo, self.incoming_shortwave = alb, inc

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

    def _write_rows(self, rows, t0, t1, stats, point, out_file, solar_file, table, layers_pt=None):
        f32 = self.precision != "f64"
        dbg = None
        if self.debug_point_output is not None:
            dbg = open(os.path.join(self.out_dir, self.debug_point_output), "a")
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
                if dbg is not None:                                  # model.py:413, :421-426, :441-448
                    line = "\n%s" % rows[i]["DATE"]
                    if self.use_msm and self.msm_xy is not None and layers_pt is not None:
                        for tl in layers_pt[i]:
                            line += ",%.2f" % tl
                    line += ",%.1f,%.1f" % (point[i, _lib.P_SENS_AWS], point[i, _lib.P_LAT_AWS])
                    dbg.write(line)
        if dbg is not None:
            dbg.close()

    def _pull_state(self, eng, bands=None, rank=0, dist=None):
        swe, tsn, tic = eng.state(np.float32)
        if dist is not None:
            swe, tsn, tic = (self._gather_rows(a, bands, rank, dist) for a in (swe, tsn, tic))
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
    def from_config(cls, config, precision="f32", device=None):
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

"""Thin object wrapper over the C ABI: one Engine = one enrgy_ctx = one GPU (row band)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import Params, check


def _f32c(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a


def make_params(*, cell_size, elev_aws, aws_row, aws_col, sensor_z=2.0, zm=None, z_h_or_e=None,
                andreas=False, sensible_corr=1.0, latent_corr=1.0, emissivity=None,
                const_albedo=None, max_ice_albedo=None, snow_density=None, ice_density=None,
                insol_mode=_lib.INSOL_STREAMED, shadow=False, lat=0.0, lon=0.0, solar_const=None,
                transmittance=None, hour_step=None, band_row0=0, band_rows=0, msm_depths=None):
    """struct enrgy_params from keyword arguments; None = the reference's default (NaN in the struct)."""
    nan = float("nan")
    p = Params()
    p.cell_size = cell_size
    p.elev_aws = elev_aws
    p.aws_row, p.aws_col = int(aws_row), int(aws_col)
    p.sensor_z = sensor_z
    p.zm = nan if zm is None else zm
    p.z_h_or_e = nan if z_h_or_e is None else z_h_or_e
    p.andreas = 1 if andreas else 0
    p.sensible_corr, p.latent_corr = float(sensible_corr), float(latent_corr)
    p.emissivity = nan if emissivity is None else emissivity
    if const_albedo is not None:
        p.albedo_const = 1
        p.albedo_ice, p.albedo_snow = float(const_albedo[0]), float(const_albedo[1])
    p.max_ice_albedo = nan if max_ice_albedo is None else max_ice_albedo
    p.snow_density = nan if snow_density is None else snow_density
    p.ice_density = nan if ice_density is None else ice_density
    p.insol_mode = int(insol_mode)
    p.shadow = 1 if shadow else 0
    p.lat_deg, p.lon_deg = float(lat), float(lon)
    p.solar_const = nan if solar_const is None else solar_const
    p.transmittance = nan if transmittance is None else transmittance
    p.hour_step = nan if hour_step is None else hour_step
    p.band_row0, p.band_rows = int(band_row0), int(band_rows)
    if msm_depths is not None:
        if len(msm_depths) > _lib.MAX_LAYERS - 1:
            raise ValueError("at most %d sub-surface layers" % (_lib.MAX_LAYERS - 1))
        p.msm_layers = len(msm_depths)
        for i, d in enumerate(msm_depths):
            p.msm_depths[i] = float(d)
    return p


def host_prepass(dem, forcing, precision=_lib.F64, pot_aws=None, **params):
    """The library's pre-pass on the host alone (enrgy_host_prepass: no device, no handle): per-row
    scalars [T, P_COUNT] -- Monin-Obukhov length, CH, potential insolation and shortwave factor at the
    AWS cell, sunlit sub-step count -- for a full DEM and a forcing table (forcing.build_forcing)."""
    lib = _lib.load()
    dem = np.ascontiguousarray(dem, dtype=np.float32)
    forcing = np.ascontiguousarray(forcing, dtype=np.float64)
    t = forcing.shape[0]
    out = np.zeros((t, _lib.P_COUNT), dtype=np.float64)
    pa = None if pot_aws is None else np.ascontiguousarray(pot_aws, dtype=np.float64)
    p = make_params(**params)
    check(lib.enrgy_host_prepass(C.byref(p), int(precision), dem.shape[0], dem.shape[1], dem.ctypes.data, t,
                                 forcing.ctypes.data, None if pa is None else pa.ctypes.data, out.ctypes.data))
    return out


class Engine:
    def __init__(self, rows, cols, precision=_lib.F32, device=0):
        self.lib = _lib.load()
        self.rows, self.cols, self.precision, self.device = int(rows), int(cols), int(precision), int(device)
        h = C.c_void_p()
        check(self.lib.enrgy_create(self.device, self.rows, self.cols, self.precision, C.byref(h)))
        self.h = h
        self.band_row0, self.band_rows = 0, self.rows
        self.n_steps = 0
        self._keep = []

    # ---- configuration ------------------------------------------------------------------------
    def set_params(self, **kw):
        """Keyword arguments: see make_params()."""
        p = make_params(**kw)
        self.msm_layers = int(p.msm_layers)
        check(self.lib.enrgy_set_params(self.h, C.byref(p)))
        self.band_row0 = int(p.band_row0)
        self.band_rows = int(p.band_rows) if p.band_rows else self.rows
        self.params = p

    def set_dem(self, dem):
        dem = _f32c(dem)
        assert dem.shape == (self.rows, self.cols), dem.shape
        check(self.lib.enrgy_set_dem(self.h, dem.ctypes.data))

    def set_terrain(self, terrain):
        """The uncropped terrain on the model grid (shadow casters and slopes outside the glacier
        outline); see enrgy_set_terrain."""
        terrain = _f32c(terrain)
        assert terrain.shape == (self.rows, self.cols), terrain.shape
        check(self.lib.enrgy_set_terrain(self.h, terrain.ctypes.data))

    def set_albedo_maps(self, maps):
        maps = [_f32c(m) for m in maps]
        for m in maps:
            assert m.shape == (self.band_rows, self.cols), m.shape
        arr = (C.c_void_p * len(maps))(*[m.ctypes.data for m in maps])
        check(self.lib.enrgy_set_albedo_maps(self.h, len(maps), arr))

    def set_swe(self, swe):
        if swe is None:
            check(self.lib.enrgy_set_swe(self.h, None))
            return
        swe = _f32c(swe)
        assert swe.shape == (self.band_rows, self.cols), swe.shape
        check(self.lib.enrgy_set_swe(self.h, swe.ctypes.data))

    def set_msm(self, temperatures, elev):
        """Initial boundary temperatures at the reference elevation (reference model.py:126-143)."""
        t = np.ascontiguousarray(temperatures, dtype=np.float64)
        assert t.shape == (self.msm_layers + 1,), t.shape
        check(self.lib.enrgy_set_msm(self.h, t.ctypes.data, float(elev)))

    def layer_temps(self):
        out = np.empty((self.msm_layers + 1, self.band_rows, self.cols), dtype=np.float64)
        check(self.lib.enrgy_get_layer_temps(self.h, out.ctypes.data))
        return out

    def set_member(self, albedo_offset=0.0, zm=None, z_h_or_e=None):
        """Ensemble member on the resident handle (BASELINE config C5), see enrgy_set_member."""
        nan = float("nan")
        check(self.lib.enrgy_set_member(self.h, float(albedo_offset), nan if zm is None else float(zm),
                                        nan if z_h_or_e is None else float(z_h_or_e)))

    def run_members(self, albedo_offsets, zm=None, z_h_or_e=None, t0=0, t1=None, want_stats=True):
        """Ensemble members in fused passes of the kernel (enrgy_run_members): every member starts from
        the handle's current state; returns (per-step statistics [members, t1 - t0, S_COUNT] or None -- the
        passes skip them when they are not wanted --, totals [members, 4]: glacier-wide means of the final
        swe, total_snow, total_ice rasters and the glacier cell count).
        The handle needs prepass() again before a plain run()."""
        t1 = self.n_steps if t1 is None else t1
        off = np.ascontiguousarray(albedo_offsets, dtype=np.float64)
        n = off.size

        def opt(a):
            if a is None:
                return None
            a = np.ascontiguousarray([np.nan if v is None else v for v in a], dtype=np.float64)
            assert a.size == n
            return a
        zm_a, zh_a = opt(zm), opt(z_h_or_e)
        stats = np.empty((n, t1 - t0, _lib.S_COUNT), dtype=np.float64) if want_stats else None
        totals = np.empty((n, 4), dtype=np.float64)
        check(self.lib.enrgy_run_members(self.h, n, off.ctypes.data, None if zm_a is None else zm_a.ctypes.data,
                                         None if zh_a is None else zh_a.ctypes.data, int(t0), int(t1),
                                         None if stats is None else stats.ctypes.data, totals.ctypes.data))
        return stats, totals

    def member_state(self, member, dtype=np.float32):
        dt = np.dtype(dtype)
        code = 32 if dt == np.float32 else 64
        outs = [np.empty((self.band_rows, self.cols), dtype=dt) for _ in range(3)]
        check(self.lib.enrgy_get_member_state(self.h, int(member), code, *[o.ctypes.data for o in outs]))
        return tuple(outs)

    def set_stations(self, stations, series, cloud_k=None):
        """Extra weather stations (BASELINE config C4; enrgy_set_stations): stations = [(row, col, elev)] in
        cell units of the full raster, series [n_extra, n_steps, ST_COUNT] (forcing.build_station_series),
        cloud_k = Beer-Lambert coefficient of the cloud attenuation of the shortwave (None: off).
        stations = [] runs the blend with the primary AWS alone; stations = None switches it off.
        Call after set_forcing."""
        if stations is None:
            check(self.lib.enrgy_set_stations(self.h, -1, None, None, None, None, float("nan")))
            return
        n = len(stations)
        ck = float("nan") if cloud_k is None else float(cloud_k)
        if n == 0:
            check(self.lib.enrgy_set_stations(self.h, 0, None, None, None, None, ck))
            return
        pos = np.ascontiguousarray(stations, dtype=np.float64)
        rows, cols, elevs = [np.ascontiguousarray(pos[:, k]) for k in range(3)]
        ser = np.ascontiguousarray(series, dtype=np.float64)
        assert ser.shape == (n, self.n_steps, _lib.ST_COUNT), ser.shape
        check(self.lib.enrgy_set_stations(self.h, n, rows.ctypes.data, cols.ctypes.data, elevs.ctypes.data,
                                          ser.ctypes.data, ck))

    def set_forcing(self, table):
        table = np.ascontiguousarray(table, dtype=np.float64)
        assert table.ndim == 2 and table.shape[1] == _lib.F_COUNT
        self.n_steps = table.shape[0]
        check(self.lib.enrgy_set_forcing(self.h, self.n_steps, table.ctypes.data))

    def set_insolation(self, t0, pot):
        pot = _f32c(pot)
        assert pot.ndim == 3 and pot.shape[1:] == (self.band_rows, self.cols), pot.shape
        check(self.lib.enrgy_set_insolation(self.h, int(t0), pot.shape[0], pot.ctypes.data))

    def set_insolation_aws(self, t0, pot_aws):
        """Potential insolation at the AWS cell for steps t0.. (streamed mode on a row band that does not hold
        the cell); see enrgy_set_insolation_aws."""
        a = np.ascontiguousarray(pot_aws, dtype=np.float64)
        check(self.lib.enrgy_set_insolation_aws(self.h, int(t0), int(a.size), a.ctypes.data))

    def prepass(self):
        check(self.lib.enrgy_prepass(self.h))

    def point_scalars(self):
        out = np.empty((self.n_steps, _lib.P_COUNT), dtype=np.float64)
        check(self.lib.enrgy_get_point_scalars(self.h, out.ctypes.data))
        return out

    def point_layers(self):
        out = np.empty((self.n_steps, self.msm_layers + 1), dtype=np.float64)
        check(self.lib.enrgy_get_point_layers(self.h, out.ctypes.data))
        return out

    def set_aws_cell(self, albedo_at_aws, swe_at_aws):
        a = np.ascontiguousarray(albedo_at_aws if albedo_at_aws is not None else [], dtype=np.float64)
        check(self.lib.enrgy_set_aws_cell(self.h, int(a.size), a.ctypes.data if a.size else None, float(swe_at_aws)))

    # ---- the hot path -------------------------------------------------------------------------
    def run(self, t0=0, t1=None, want_stats=True):
        t1 = self.n_steps if t1 is None else t1
        if want_stats:
            stats = np.empty((t1 - t0, _lib.S_COUNT), dtype=np.float64)
            check(self.lib.enrgy_run(self.h, int(t0), int(t1), stats.ctypes.data))
            return stats
        check(self.lib.enrgy_run(self.h, int(t0), int(t1), None))
        return None

    def run_async(self, t0, t1, d_stats_ptr=None, stream_ptr=None):
        check(self.lib.enrgy_run_async(self.h, int(t0), int(t1), d_stats_ptr, stream_ptr))

    # ---- shading as separate steps (parallel.ShardedRun) ---------------------------------------
    def sub_range(self, t0, t1):
        a, b = C.c_int(0), C.c_int(0)
        check(self.lib.enrgy_sub_range(self.h, int(t0), int(t1), C.byref(a), C.byref(b)))
        return a.value, b.value

    def mask_words(self, rows):
        """uint32 words per sub-step of a mask array for a band of `rows` rows."""
        return int(self.lib.enrgy_mask_words(self.h, int(rows)))

    def shade_scan(self, sub0, sub1, segments, stream_ptr=None):
        """Sweeps the sub-steps [sub0, sub1) over the full raster; segments = [(row0, rows, device_ptr)]."""
        n = len(segments)
        r0 = (C.c_int * n)(*[int(s[0]) for s in segments])
        nr = (C.c_int * n)(*[int(s[1]) for s in segments])
        ptr = (C.c_void_p * n)(*[int(s[2]) for s in segments])
        check(self.lib.enrgy_shade_scan(self.h, int(sub0), int(sub1), n, r0, nr, ptr, stream_ptr))

    def run_masked(self, t0, t1, d_masks_ptr, d_stats_ptr=None, stream_ptr=None):
        check(self.lib.enrgy_run_masked(self.h, int(t0), int(t1), d_masks_ptr, d_stats_ptr, stream_ptr))

    def defer_snow_total(self, on):
        check(self.lib.enrgy_defer_snow_total(self.h, 1 if on else 0))

    def set_mask_budget(self, n_bytes):
        check(self.lib.enrgy_set_mask_budget(self.h, int(n_bytes)))

    def last_sweep_ms(self):
        return float(self.lib.enrgy_last_sweep_ms(self.h))

    def synchronize(self):
        check(self.lib.enrgy_synchronize(self.h))

    # ---- debug views --------------------------------------------------------------------------
    def dump_steps(self, t0, t1):
        out = np.empty((t1 - t0, _lib.D_COUNT, self.band_rows, self.cols), dtype=np.float64)
        check(self.lib.enrgy_dump_steps(self.h, int(t0), int(t1), out.ctypes.data))
        return out

    def substeps(self, step, max_sub=256):
        buf = np.zeros((max_sub, 8), dtype=np.float64)
        n = C.c_int(0)
        check(self.lib.enrgy_get_substeps(self.h, int(step), max_sub, buf.ctypes.data, C.byref(n)))
        return buf[:n.value].copy()

    def shade_masks(self, step, max_sub=256):
        words = (self.cols + 31) // 32
        buf = np.zeros((max_sub, self.band_rows, words), dtype=np.uint32)
        n = C.c_int(0)
        check(self.lib.enrgy_shade_masks(self.h, int(step), max_sub, buf.ctypes.data, C.byref(n)))
        bits = np.unpackbits(buf[:n.value].view(np.uint8), axis=-1, bitorder="little")
        return bits[:, :, :self.cols].astype(bool)

    def potential_insolation(self, step):
        out = np.empty((self.band_rows, self.cols), dtype=np.float64)
        check(self.lib.enrgy_potential_insolation(self.h, int(step), out.ctypes.data))
        return out

    # ---- state --------------------------------------------------------------------------------
    def state(self, dtype=np.float32):
        dt = np.dtype(dtype)
        code = 32 if dt == np.float32 else 64
        outs = [np.empty((self.band_rows, self.cols), dtype=dt) for _ in range(3)]
        check(self.lib.enrgy_get_state(self.h, code, *[o.ctypes.data for o in outs]))
        return tuple(outs)

    def set_state(self, swe=None, total_snow=None, total_ice=None):
        arrs = [None if a is None else np.ascontiguousarray(a, dtype=np.float64)
                for a in (swe, total_snow, total_ice)]
        check(self.lib.enrgy_set_state(self.h, 64, *[None if a is None else a.ctypes.data for a in arrs]))

    def set_stream(self, stream_ptr):
        check(self.lib.enrgy_set_stream(self.h, stream_ptr))

    def snapshot(self, save=True):
        check(self.lib.enrgy_snapshot(self.h, 1 if save else 0))

    def microbench(self, kind):
        v = C.c_double(0.0)
        check(self.lib.enrgy_microbench(self.h, int(kind), C.byref(v)))
        return v.value

    # ---- introspection ------------------------------------------------------------------------
    def launch_count(self):
        return int(self.lib.enrgy_launch_count(self.h))

    def last_kernel_ms(self):
        return float(self.lib.enrgy_last_kernel_ms(self.h))

    def kernel_info(self):
        v = [C.c_int(0) for _ in range(4)]
        check(self.lib.enrgy_kernel_info(self.h, *[C.byref(x) for x in v]))
        return dict(regs=v[0].value, smem_bytes=v[1].value, ctas_per_sm=v[2].value, grid=v[3].value)

    def close(self):
        if getattr(self, "h", None):
            self.lib.enrgy_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

"""TEST INFRASTRUCTURE -- an oracle-backed stand-in for enrgy_b200.engine.Engine, so that the N > 1
host logic of `Energy.model` (row bands, per-band uploads, all-reduce of the statistics, rank 0
writing the files, gathering the state rasters) can run on the CPU under gloo.  The per-band numbers
come from the NumPy oracle; the product never imports this."""
from __future__ import annotations

import numpy as np

from enrgy_b200 import _lib
from oracle import enrgy_oracle as O


class FakeEngine:
    def __init__(self, rows, cols, precision=_lib.F32, device=0):
        self.rows, self.cols = rows, cols
        self.f64 = precision == _lib.F64
        self.dt = np.float64 if self.f64 else np.float32
        self.kw = None
        self.dem = self.swe = None
        self.alb = None
        self.pot = {}
        self.table = None
        self.tsn = self.tic = None
        self.n_steps = 0
        self.aws_rows = None             # set by the test: the CSV rows (the oracle parses strings)
        self.albedo_keys = None
        self.geotransform = None
        self.xy_aws = None

    # ---- configuration (what Energy.model calls) -------------------------------------------------
    def set_params(self, **kw):
        self.kw = kw
        self.r0 = kw.get("band_row0", 0)
        self.nr = kw.get("band_rows", 0) or self.rows

    def set_dem(self, dem):
        assert dem.shape == (self.rows, self.cols)
        self.dem_full = np.asarray(dem, dtype=self.dt)

    def set_forcing(self, table):
        self.table = np.asarray(table)
        self.n_steps = table.shape[0]

    def set_albedo_maps(self, maps):
        for m in maps:
            assert m.shape == (self.nr, self.cols), (m.shape, self.nr)
        self.alb = [np.asarray(m, dtype=self.dt) for m in maps]

    def set_swe(self, swe):
        assert swe.shape == (self.nr, self.cols)
        self.swe = np.array(swe, dtype=self.dt)
        self.tsn = np.zeros_like(self.swe)
        self.tic = np.zeros_like(self.swe)

    def set_state(self, swe=None, total_snow=None, total_ice=None):
        if total_snow is not None:
            self.tsn = np.array(total_snow, dtype=self.dt)
        if total_ice is not None:
            self.tic = np.array(total_ice, dtype=self.dt)

    def set_insolation(self, t0, pot):
        assert pot.shape[1:] == (self.nr, self.cols)
        for i in range(pot.shape[0]):
            self.pot[t0 + i] = np.asarray(pot[i], dtype=self.dt)

    def set_insolation_aws(self, t0, pot_aws):
        if not isinstance(getattr(self, "pot_aws", None), dict):
            self.pot_aws = {}
        for i, v in enumerate(np.asarray(pot_aws, dtype=np.float64)):
            self.pot_aws[t0 + i] = float(v)

    def prepass(self):
        pass

    def point_scalars(self):
        return np.zeros((self.n_steps, _lib.P_COUNT))

    def defer_snow_total(self, on):
        pass

    def set_stream(self, s):
        pass

    def close(self):
        pass

    # ---- the run ---------------------------------------------------------------------------------
    def run(self, t0, t1):
        """Oracle on the band (the AWS row in front so the point sampling sees the AWS cell), rows
        [t0, t1) continuing from the current state; returns the band's statistics SUMS."""
        kw = self.kw
        ar, ac = kw["aws_row"], kw["aws_col"]
        sl = slice(self.r0, self.r0 + self.nr)
        dem = np.concatenate([self.dem_full[ar:ar + 1], self.dem_full[sl]], axis=0)
        aws_valid = np.zeros(self.cols, dtype=bool)
        aws_valid[ac] = True

        def with_aws(band, fill):
            first = np.full((1, self.cols), np.nan, dtype=self.dt)
            first[0, ac] = fill
            return np.concatenate([first, band], axis=0)
        gt = list(self.geotransform)
        gt[3] = self.xy_aws[1] + 0.5 * abs(gt[5])
        # only the AWS cell of the extra row is a glacier cell
        dem[0, ~aws_valid] = np.nan
        rows = self.aws_rows[t0:t1 + 1] if t1 < len(self.aws_rows) else self.aws_rows[t0:t1]
        n = t1 - t0
        pot = np.stack([with_aws(self.pot[t0 + i], self.pot_aws[t0 + i]) for i in range(n)]
                       + ([with_aws(self.pot[t0 + n - 1], 0.0)] if len(rows) > n else []))
        cfg = O.ModelConfig(z=kw["sensor_z"], elev_aws=kw["elev_aws"], xy_aws=self.xy_aws, zm=kw["zm"],
                            z_h_or_e=kw["z_h_or_e"], emissivity=kw["emissivity"], const_albedo=kw.get("const_albedo"))
        alb = None
        if self.alb is not None:
            alb = {k: with_aws(a, 0.5) for k, a in zip(self.albedo_keys, self.alb)}
        out = O.run_model(dem, tuple(gt), rows, pot, cfg, swe=with_aws(self.swe, 0.0), albedo_arrays=alb,
                          state_dtype=self.dt, keep_steps=None)
        s = np.zeros((n, _lib.S_COUNT))
        for i in range(n):
            row, melt = out["rows"][i], out["melt"][i]
            cut = lambda a: np.asarray(a)[1:]          # noqa: E731  (drop the duplicated AWS row)
            s[i, _lib.S_RS] = np.nansum(cut(row["rs"]))
            s[i, _lib.S_LWD] = np.nansum(cut(row["lwd"]))
            s[i, _lib.S_LWU] = np.nansum(np.where(np.isnan(cut(row["lwd"])), np.nan, cut(row["lwu"])))
            s[i, _lib.S_SENS] = np.nansum(cut(row["sens"]))
            s[i, _lib.S_LAT] = np.nansum(cut(row["lat"]))
            s[i, _lib.S_ATMO] = np.nansum(cut(row["atmo"]))
            s[i, _lib.S_MELT] = np.nansum(cut(row["mf"]))
            s[i, _lib.S_SNOW] = np.nansum(cut(melt[0]))
            s[i, _lib.S_ICE] = np.nansum(cut(melt[1]))
            swe = cut(melt[2])
            s[i, _lib.S_SWE] = np.nansum(swe)
            s[i, _lib.S_NSNOW] = np.sum(swe > 0)
            s[i, _lib.S_NSWE] = np.count_nonzero(~np.isnan(swe))
            s[i, _lib.S_NVALID] = np.count_nonzero(~np.isnan(cut(row["atmo"])))
            # the state after this row (the oracle ran one row further for the time step of the last one)
            if i == n - 1:
                snow_cum = sum(np.nan_to_num(cut(out["melt"][j][0])) for j in range(n))
                ice_cum = sum(np.nan_to_num(cut(out["melt"][j][1])) for j in range(n))
                valid = ~np.isnan(self.dem_full[sl])
                new_swe = np.where(valid, np.nan_to_num(self.swe) - snow_cum, np.nan)
                self.tsn = np.where(valid, np.nan_to_num(self.tsn) + snow_cum, np.nan).astype(self.dt)
                self.tic = np.where(valid, np.nan_to_num(self.tic) + ice_cum, np.nan).astype(self.dt)
                self.swe = new_swe.astype(self.dt)
        return s

    def state(self, dtype=np.float32):
        return tuple(np.asarray(a, dtype=dtype) for a in (self.swe, self.tsn, self.tic))

"""Insolation cache writer in the reference's own layout.

The reference caches SAGA's per-step "total potential insolation" rasters as
`<pickle_dir>/<res>/<DATE>_total.sdat.npy` (insolation_pickler.py:12-25) and reads them back in
`Energy.calc_shortwave` (model.py:477-481).  This module writes the same files from the GPU
computation (specification: DESIGN.md section 5), so the UNMODIFIED reference can run on
GPU-computed insolation via `use_precomputed = True; add_pickle_dir(pickle_dir)`.
File names embed the AWS DATE string verbatim, as model.py:466-467 builds them (the author's own
files use a non-zero-padded hour, renamer.py:16-17; pass `date_format` to reproduce that).
"""
from __future__ import annotations

import os
from datetime import datetime

import numpy as np

from . import _lib
from .engine import Engine
from .forcing import build_forcing, read_input_file
from .geo import grid_centre_latlon


def pickle_path(pickle_dir, res, date_str):
    return os.path.join(pickle_dir, str(res), "%s_total.sdat.npy" % date_str)


def pickle_insolation_series(dem, geotransform, aws_rows_or_file, pickle_dir, res, lat=None, lon=None,
                             shadow=True, device=0, date_format=None, dtype=np.float32):
    """Computes the potential insolation [kWh m-2] of every AWS row on the GPU and saves it where the
    reference looks for it.  Returns the list of files written."""
    rows = read_input_file(aws_rows_or_file) if isinstance(aws_rows_or_file, (str, os.PathLike)) else aws_rows_or_file
    dem = np.ascontiguousarray(dem, dtype=np.float32)
    h, w = dem.shape
    if lat is None or lon is None:
        lat, lon = grid_centre_latlon(geotransform, h, w)
    out_dir = os.path.join(pickle_dir, str(res))
    os.makedirs(out_dir, exist_ok=True)
    valid = np.argwhere(~np.isnan(dem))
    if valid.size == 0:
        raise ValueError("the DEM has no valid cell")
    r0, c0 = (int(v) for v in valid[len(valid) // 2])     # any glacier cell serves as the AWS cell here
    eng = Engine(h, w, precision=_lib.F64, device=device)
    files = []
    try:
        eng.set_params(cell_size=abs(geotransform[1]), elev_aws=float(dem[r0, c0]), aws_row=r0, aws_col=c0,
                       const_albedo=(0.35, 0.75), insol_mode=_lib.INSOL_COMPUTED, shadow=shadow, lat=lat, lon=lon)
        eng.set_dem(dem)
        eng.set_forcing(build_forcing(rows, None))
        eng.prepass()
        for i, row in enumerate(rows):
            pot = eng.potential_insolation(i).astype(dtype)
            name = row["DATE"]
            if date_format is not None:
                name = datetime.strptime(row["DATE"], "%Y%m%d %H:%M:%S").strftime(date_format)
            path = pickle_path(pickle_dir, res, name)
            np.save(path, pot)
            files.append(path)
    finally:
        eng.close()
    return files

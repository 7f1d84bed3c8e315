"""Raster I/O of the model path: same entry points as reference raster_utils.py:36-89.

GDAL (`osgeo`) is imported lazily: with it, rasters are read exactly the way the reference reads
them (warp to UTM 33N, cutline crop, resample, Float32, nodata -> NaN, raster_utils.py:36-53) and
results are written as GeoTIFF (raster_utils.py:56-82).  Without it (this build image has no GDAL)
`.npy` rasters are accepted instead: `<name>.npy` plus an optional side-car `<name>.npy.json`
{"geotransform": [...], "projection": "..."}; they are taken as already on the model grid (the
same contract as the reference's own pickle cache, insolation_pickler.py:12-25, model.py:477-481).
This module is I/O only -- it is not on the hot path and has no CUDA dependency.
"""
from __future__ import annotations

import json
import os

import numpy as np

from .geo import get_value_by_real_coords  # noqa: F401  (re-export, raster_utils.py:85-89)

UTM33N = "+proj=utm +zone=33 +datum=WGS84 +units=m +no_defs"


def _gdal():
    try:
        from osgeo import gdal
        return gdal
    except Exception:
        return None


def have_gdal():
    return _gdal() is not None


def show_me(array, out_dir=None, title=None, units=None, show=False, dir=None, verbose=False):
    """PNG previews (raster_utils.py:9-32) are out of scope: accepted and ignored."""
    return None


def _clean(array, remove_negatives, remove_outliers):
    if remove_negatives:
        array[array < 0] = np.nan          # raster_utils.py:46-47
    if remove_outliers:
        array[array < 0] = 0.001           # raster_utils.py:48-50
        array[array > 1] = 1
    return array


def load_raster(raster_path, crop_path, res, remove_negatives=False, remove_outliers=False, v=True):
    """-> (float32 array [H, W] with NaN nodata, geotransform, projection)."""
    if isinstance(raster_path, np.ndarray):                      # in-memory raster
        arr = np.array(raster_path, dtype=np.float32, copy=True)
        return _clean(arr, remove_negatives, remove_outliers), None, None
    if str(raster_path).endswith(".npy"):
        arr = np.load(raster_path).astype(np.float32)
        gt, proj = (0.0, float(res or 1), 0.0, 0.0, 0.0, -float(res or 1)), UTM33N
        side = str(raster_path) + ".json"
        if os.path.isfile(side):
            with open(side) as f:
                meta = json.load(f)
            gt = tuple(meta.get("geotransform", gt))
            proj = meta.get("projection", proj)
        arr = _clean(arr, remove_negatives, remove_outliers)
        if v:
            print("Raster size is %dx%d" % arr.shape)
        return arr, gt, proj
    gdal = _gdal()
    if gdal is None:
        raise ImportError("GDAL (osgeo) is needed to read %r; .npy rasters work without it" % (raster_path,))
    ds = gdal.Open(raster_path)
    crop_ds = gdal.Warp("", ds, dstSRS=UTM33N, format="VRT", cutlineDSName=crop_path,
                        cropToCutline=True, outputType=gdal.GDT_Float32, xRes=res, yRes=res)
    gt = crop_ds.GetGeoTransform()
    proj = crop_ds.GetProjection()
    band = crop_ds.GetRasterBand(1)
    nodata = band.GetNoDataValue()
    array = band.ReadAsArray()
    array[array == nodata] = np.nan
    array = _clean(array, remove_negatives, remove_outliers)
    if v:
        print("Raster size is %dx%d" % array.shape)
    return array, gt, proj


def load_uncropped_like(raster_path, geotransform, shape):
    """`raster_path` warped to UTM 33N onto the grid (geotransform, shape) of an already loaded cropped
    raster, WITHOUT the cutline: the terrain SAGA sees around the glacier (the reference passes it the
    uncropped DEM file, model.py:469 -> saga_lighting.py:42).  Needs GDAL."""
    gdal = _gdal()
    if gdal is None:
        raise ImportError("GDAL (osgeo) is needed to re-read %r without the cutline" % (raster_path,))
    rows, cols = shape
    ulx, dx, _, uly, _, dy = geotransform
    bounds = (ulx, uly + rows * dy, ulx + cols * dx, uly)           # (minX, minY, maxX, maxY); dy < 0
    ds = gdal.Warp("", gdal.Open(raster_path), dstSRS=UTM33N, format="VRT", outputType=gdal.GDT_Float32,
                   xRes=abs(dx), yRes=abs(dy), outputBounds=bounds)
    band = ds.GetRasterBand(1)
    nodata = band.GetNoDataValue()
    array = band.ReadAsArray()
    array[array == nodata] = np.nan
    return array


def export_array_as_geotiff(array_to_export, geotransform, projection, path, scale_mult=None):
    """Float32 GeoTIFF with nodata -9999 (Int16 / -32768 when scale_mult is given),
    raster_utils.py:56-82.  Without GDAL the same array is written to `<path>.npy`."""
    array = np.copy(array_to_export)
    if scale_mult is not None:
        array = np.rint(array * scale_mult)
        nodata = -32768
    else:
        nodata = -9999
    array[np.isnan(array)] = nodata
    gdal = _gdal()
    if gdal is None:
        out = path + ".npy"
        np.save(out, array.astype(np.int16 if scale_mult is not None else np.float32))
        with open(out + ".json", "w") as f:
            json.dump({"geotransform": list(geotransform) if geotransform is not None else None,
                       "projection": projection, "nodata": nodata}, f)
        return out
    gdt = gdal.GDT_Int16 if scale_mult is not None else gdal.GDT_Float32
    ds = gdal.GetDriverByName("GTiff").Create(path, array.shape[1], array.shape[0], 1, gdt)
    ds.SetGeoTransform(geotransform)
    ds.SetProjection(projection)
    band = ds.GetRasterBand(1)
    band.SetNoDataValue(nodata)
    band.WriteArray(array)
    band.FlushCache()
    ds = None
    return path


def save_npy_raster(path, array, geotransform=None, projection=UTM33N):
    """Writes `<path>` (.npy) + side-car json: the GDAL-free raster format load_raster accepts."""
    np.save(path, np.asarray(array, dtype=np.float32))
    real = path if path.endswith(".npy") else path + ".npy"
    with open(real + ".json", "w") as f:
        json.dump({"geotransform": list(geotransform) if geotransform is not None else None,
                   "projection": projection}, f)
    return real

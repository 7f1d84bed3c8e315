"""Grid georeference helpers: raster index of a coordinate and UTM -> geographic conversion.

The reference warps every raster to UTM zone 33N (raster_utils.py:38) and lets SAGA derive the
sun position from the grid's georeference (saga_lighting.py:42, -LOCATION 1); the in-kernel
insolation needs one latitude/longitude, taken at the grid centre.
"""
from __future__ import annotations

import math


def coords_to_index(gt, easting, northing):
    """(line, pixel) exactly as reference raster_utils.py:85-89 computes them."""
    ul_x, x_dist, _, ul_y, _, y_dist = gt
    pixel = int((easting - ul_x) / x_dist)
    line = -int((ul_y - northing) / y_dist)
    return line, pixel


def get_value_by_real_coords(array, gt, easting, northing):
    """raster_utils.py:85-89."""
    line, pixel = coords_to_index(gt, easting, northing)
    return array[line][pixel]


def utm_to_latlon(easting, northing, zone=33, northern=True):
    """WGS84 transverse Mercator inverse (Krueger series, n^4): degrees (lat, lon)."""
    a, f = 6378137.0, 1.0 / 298.257223563
    k0, e0 = 0.9996, 500000.0
    n = f / (2.0 - f)
    big_a = a / (1.0 + n) * (1.0 + n ** 2 / 4.0 + n ** 4 / 64.0)
    b1 = n / 2.0 - 2.0 * n ** 2 / 3.0 + 37.0 * n ** 3 / 96.0
    b2 = n ** 2 / 48.0 + n ** 3 / 15.0
    b3 = 17.0 * n ** 3 / 480.0
    d1 = 2.0 * n - 2.0 * n ** 2 / 3.0 - 2.0 * n ** 3
    d2 = 7.0 * n ** 2 / 3.0 - 8.0 * n ** 3 / 5.0
    d3 = 56.0 * n ** 3 / 15.0
    n0 = 0.0 if northern else 10000000.0
    xi = (northing - n0) / (k0 * big_a)
    eta = (easting - e0) / (k0 * big_a)
    xi_p, eta_p = xi, eta
    for j, b in ((1, b1), (2, b2), (3, b3)):
        xi_p -= b * math.sin(2 * j * xi) * math.cosh(2 * j * eta)
        eta_p -= b * math.cos(2 * j * xi) * math.sinh(2 * j * eta)
    chi = math.asin(math.sin(xi_p) / math.cosh(eta_p))
    lat = chi + d1 * math.sin(2 * chi) + d2 * math.sin(4 * chi) + d3 * math.sin(6 * chi)
    lon0 = math.radians(zone * 6 - 183)
    lon = lon0 + math.atan2(math.sinh(eta_p), math.cos(xi_p))
    return math.degrees(lat), math.degrees(lon)


def grid_centre_latlon(gt, rows, cols, zone=33, northern=True):
    cx = gt[0] + 0.5 * cols * gt[1]
    cy = gt[3] + 0.5 * rows * gt[5]
    return utm_to_latlon(cx, cy, zone, northern)

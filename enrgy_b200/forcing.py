"""Host-side row logic of `Energy.model`: CSV rows -> the forcing table of the C ABI.

Mirrors, with the same names where the reference has them:
  helpers.py:48-51   read_input_file        helpers.py:63-71   get_time_step
  helpers.py:74-87   heuristic_unit_guesser model.py:197-226   humidity / cloud / lapse handling
  interpolator.py:23-39  _get_closest_dates (bracketing albedo maps, integer days)
  model.py:311-320   snow-albedo ageing since `last_snowfall`
"""
from __future__ import annotations

import csv
from datetime import datetime, timezone

import numpy as np

from . import _lib


def read_input_file(input_file):
    """helpers.py:48-51."""
    with open(input_file) as f:
        return list(csv.DictReader(f))


def _parse(s, pattern=None):
    if pattern is not None:
        return datetime.strptime(s, pattern)
    try:
        return datetime.strptime(s, "%Y%m%d")
    except ValueError:
        return datetime.strptime(s, "%Y%m%d %H:%M:%S")


def get_time_step(time_list, i, pattern):
    """helpers.py:63-71: forward difference; the last row repeats the previous step."""
    if i < len(time_list) - 1:
        d = _parse(time_list[i + 1]["DATE"], pattern) - _parse(time_list[i]["DATE"], pattern)
    else:
        d = _parse(time_list[i]["DATE"], pattern) - _parse(time_list[i - 1]["DATE"], pattern)
    return int(d.total_seconds())


def heuristic_unit_guesser(value, scale=10):
    """helpers.py:74-87: per cent -> fraction when 1 < value <= scale."""
    if 1 < value <= scale:
        return value / scale
    elif value <= 1:
        return value
    else:
        raise ValueError("Wrong value encountered")


def get_closest_dates(keys, needed_date):
    """interpolator.py:23-39."""
    dates = [datetime.strptime(k, "%Y%m%d") for k in keys]
    before = [d for d in dates if d <= needed_date]
    after = [d for d in dates if d >= needed_date]
    if len(before) == 0 or len(after) == 0:
        raise ValueError("Passed date is outside of the possible interpolation range!")
    return max(before), min(after)


def to_unix(dt):
    """AWS timestamps are taken as UTC."""
    return float(dt.replace(tzinfo=timezone.utc).timestamp())


def build_forcing(rows, albedo_keys=None, temp_lapse_rate=-0.006, cloud_corr=None, last_snowfall=None):
    """[n_rows, F_COUNT] float64 table for enrgy_set_forcing, row semantics of model.py:186-230."""
    n = len(rows)
    out = np.zeros((n, _lib.F_COUNT), dtype=np.float64)
    keys = list(albedo_keys) if albedo_keys is not None else None
    for i, row in enumerate(rows):
        try:                                   # model.py:190-193
            dt = get_time_step(rows, i, "%Y%m%d")
        except ValueError:
            dt = get_time_step(rows, i, "%Y%m%d %H:%M:%S")
        date = _parse(row["DATE"])
        r_hum = heuristic_unit_guesser(float(row["HUMID"]), 100)
        cld = float(row["CLOUDINESS"])
        if cloud_corr is not None:             # model.py:200-204
            cld += cloud_corr
            cld = 1.0 if cld > 1.0 else cld
            cld = 0.0 if cld < 0.0 else cld
        try:                                   # model.py:213-221
            grad = float(temp_lapse_rate)
        except ValueError:
            try:
                grad = float(row["GRADIENT"])
            except KeyError:
                raise ValueError("lapse-rate column %r is missing from the AWS file" % temp_lapse_rate)
        o = out[i]
        o[_lib.F_TIME] = to_unix(date)
        o[_lib.F_DT] = dt
        o[_lib.F_T_AIR] = float(row["T_AIR"])
        o[_lib.F_WIND] = float(row["WIND_SPEED"])
        o[_lib.F_PRESSURE] = float(row["PRESSURE"])
        o[_lib.F_RH] = r_hum
        o[_lib.F_CLOUD] = cld
        o[_lib.F_SWD] = float(row["SWD"])
        o[_lib.F_LAPSE] = grad
        if keys is not None:
            before, after = get_closest_dates(keys, date)
            o[_lib.F_ALB_I0] = keys.index(before.strftime("%Y%m%d"))
            o[_lib.F_ALB_I1] = keys.index(after.strftime("%Y%m%d"))
            o[_lib.F_ALB_DAYS] = (date - before).days
            o[_lib.F_ALB_SPAN] = (after - before).days
            if last_snowfall is not None:      # model.py:311-320 (needs the full timestamp format)
                delta = (datetime.strptime(row["DATE"], "%Y%m%d %H:%M:%S")
                         - datetime.strptime(last_snowfall, "%Y%m%d"))
                o[_lib.F_SNOW_DAYS] = delta.days if delta.days > 0 else 0
    return out


def build_station_series(rows, cloud_corr=None):
    """[n_rows, ST_COUNT] float64 series of an extra weather station for enrgy_set_stations (BASELINE
    config C4): T_AIR, PRESSURE, HUMID as a fraction, CLOUDINESS after cloud_corr -- the row semantics of
    model.py:197-204 applied to the station's own file."""
    out = np.zeros((len(rows), _lib.ST_COUNT), dtype=np.float64)
    for i, row in enumerate(rows):
        cld = float(row["CLOUDINESS"])
        if cloud_corr is not None:
            cld += cloud_corr
            cld = 1.0 if cld > 1.0 else cld
            cld = 0.0 if cld < 0.0 else cld
        out[i] = (float(row["T_AIR"]), float(row["PRESSURE"]), heuristic_unit_guesser(float(row["HUMID"]), 100), cld)
    return out

"""ctypes binding of the C ABI declared in include/enrgy_b200.h.

The shared library is built in-tree (enrgy_b200/csrc/build.sh, called by __graft_entry__.build()).
There is no fallback: if the library is missing, or no CUDA device is visible when a context is
created, the caller gets an exception -- never a CPU code path.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ENRGY_B200_LIB") or os.path.join(_HERE, "csrc", "libenrgy_b200.so")

F32, F64 = 32, 64
INSOL_STREAMED, INSOL_COMPUTED = 0, 1
MAX_LAYERS = 8

# forcing columns (enum ENRGY_F_*)
(F_TIME, F_DT, F_T_AIR, F_WIND, F_PRESSURE, F_RH, F_CLOUD, F_SWD, F_LAPSE, F_ALB_I0, F_ALB_I1,
 F_ALB_DAYS, F_ALB_SPAN, F_SNOW_DAYS, F_COUNT) = range(15)
# enrgy_station_col
(ST_T_AIR, ST_PRESSURE, ST_RH, ST_CLOUD, ST_COUNT) = range(5)
# statistics columns (enum ENRGY_S_*)
(S_RS, S_LWD, S_LWU, S_SENS, S_LAT, S_ATMO, S_G, S_MELT, S_SNOW, S_ICE, S_SWE, S_NSNOW, S_NSWE,
 S_NVALID, S_COUNT) = range(15)
# point scalars (enum ENRGY_P_*)
(P_L, P_CH, P_POT_AWS, P_SW_FACTOR, P_TSURF_AWS, P_QH_AWS, P_NSUB, P_SENS_AWS, P_LAT_AWS, P_COUNT) = range(10)
# dump fields (enum ENRGY_D_*)
(D_RS, D_LWD, D_LWU, D_SENS, D_LAT, D_ATMO, D_MELT, D_SNOW, D_ICE, D_ALBEDO, D_POT, D_G,
 D_COUNT) = range(13)
DUMP_NAMES = ("rs", "lwd", "lwu", "sens", "lat", "atmo", "mf", "snow", "ice", "albedo", "pot", "g")

ERR_ARG, ERR_CUDA, ERR_NODEVICE, ERR_MASK, ERR_RANGE = -1, -2, -3, -4, -5


class Params(C.Structure):
    """struct enrgy_params (include/enrgy_b200.h)."""
    _fields_ = [
        ("cell_size", C.c_double), ("elev_aws", C.c_double),
        ("aws_row", C.c_int32), ("aws_col", C.c_int32),
        ("sensor_z", C.c_double),
        ("zm", C.c_double), ("z_h_or_e", C.c_double),
        ("andreas", C.c_int32), ("_pad0", C.c_int32),
        ("sensible_corr", C.c_double), ("latent_corr", C.c_double),
        ("emissivity", C.c_double),
        ("albedo_const", C.c_int32), ("_pad1", C.c_int32),
        ("albedo_ice", C.c_double), ("albedo_snow", C.c_double), ("max_ice_albedo", C.c_double),
        ("snow_density", C.c_double), ("ice_density", C.c_double),
        ("insol_mode", C.c_int32), ("shadow", C.c_int32),
        ("lat_deg", C.c_double), ("lon_deg", C.c_double),
        ("solar_const", C.c_double), ("transmittance", C.c_double), ("hour_step", C.c_double),
        ("msm_layers", C.c_int32), ("_pad2", C.c_int32),
        ("msm_depths", C.c_double * MAX_LAYERS),
        ("band_row0", C.c_int32), ("band_rows", C.c_int32),
    ]


# every symbol include/enrgy_b200.h declares: name -> (restype, argtypes)
_P = C.c_void_p
_PROTOTYPES = {
    "enrgy_abi_version": (C.c_int, []),
    "enrgy_last_error": (C.c_char_p, []),
    "enrgy_device_count": (C.c_int, []),
    "enrgy_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(_P)]),
    "enrgy_destroy": (C.c_int, [_P]),
    "enrgy_set_params": (C.c_int, [_P, C.POINTER(Params)]),
    "enrgy_set_dem": (C.c_int, [_P, _P]),
    "enrgy_set_terrain": (C.c_int, [_P, _P]),
    "enrgy_set_albedo_maps": (C.c_int, [_P, C.c_int, C.POINTER(_P)]),
    "enrgy_set_swe": (C.c_int, [_P, _P]),
    "enrgy_set_msm": (C.c_int, [_P, _P, C.c_double]),
    "enrgy_set_member": (C.c_int, [_P, C.c_double, C.c_double, C.c_double]),
    "enrgy_set_insolation_aws": (C.c_int, [_P, C.c_int, C.c_int, _P]),
    "enrgy_set_stations": (C.c_int, [_P, C.c_int, _P, _P, _P, _P, C.c_double]),
    "enrgy_run_members": (C.c_int, [_P, C.c_int, _P, _P, _P, C.c_int, C.c_int, _P, _P]),
    "enrgy_get_member_state": (C.c_int, [_P, C.c_int, C.c_int, _P, _P, _P]),
    "enrgy_set_forcing": (C.c_int, [_P, C.c_int, _P]),
    "enrgy_set_insolation": (C.c_int, [_P, C.c_int, C.c_int, _P]),
    "enrgy_prepass": (C.c_int, [_P]),
    "enrgy_get_point_scalars": (C.c_int, [_P, _P]),
    "enrgy_get_point_layers": (C.c_int, [_P, _P]),
    "enrgy_set_aws_cell": (C.c_int, [_P, C.c_int, _P, C.c_double]),
    "enrgy_host_prepass": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, _P, C.c_int, _P, _P, _P]),
    "enrgy_run": (C.c_int, [_P, C.c_int, C.c_int, _P]),
    "enrgy_run_async": (C.c_int, [_P, C.c_int, C.c_int, _P, _P]),
    "enrgy_synchronize": (C.c_int, [_P]),
    "enrgy_dump_steps": (C.c_int, [_P, C.c_int, C.c_int, _P]),
    "enrgy_get_substeps": (C.c_int, [_P, C.c_int, C.c_int, _P, C.POINTER(C.c_int)]),
    "enrgy_shade_masks": (C.c_int, [_P, C.c_int, C.c_int, _P, C.POINTER(C.c_int)]),
    "enrgy_potential_insolation": (C.c_int, [_P, C.c_int, _P]),
    "enrgy_sub_range": (C.c_int, [_P, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "enrgy_mask_words": (C.c_int64, [_P, C.c_int]),
    "enrgy_shade_scan": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                   C.POINTER(_P), _P]),
    "enrgy_run_masked": (C.c_int, [_P, C.c_int, C.c_int, _P, _P, _P]),
    "enrgy_set_mask_budget": (C.c_int, [_P, C.c_int64]),
    "enrgy_defer_snow_total": (C.c_int, [_P, C.c_int]),
    "enrgy_get_state": (C.c_int, [_P, C.c_int, _P, _P, _P]),
    "enrgy_set_state": (C.c_int, [_P, C.c_int, _P, _P, _P]),
    "enrgy_get_layer_temps": (C.c_int, [_P, _P]),
    "enrgy_set_stream": (C.c_int, [_P, _P]),
    "enrgy_snapshot": (C.c_int, [_P, C.c_int]),
    "enrgy_microbench": (C.c_int, [_P, C.c_int, C.POINTER(C.c_double)]),
    "enrgy_launch_count": (C.c_int64, [_P]),
    "enrgy_last_kernel_ms": (C.c_double, [_P]),
    "enrgy_last_sweep_ms": (C.c_double, [_P]),
    "enrgy_kernel_info": (C.c_int, [_P, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                    C.POINTER(C.c_int)]),
}
EXPORTED_SYMBOLS = tuple(_PROTOTYPES)

_lib = None


class EnrgyError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("enrgy_b200 error %d: %s" % (code, message))
        self.code = code


def load():
    """Load libenrgy_b200.so (once).  Raises if it has not been built -- no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise ImportError(
            "%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(enrgy_b200 has no CPU fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _PROTOTYPES.items():
        fn = getattr(lib, name)          # AttributeError if the library lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise EnrgyError(rc, load().enrgy_last_error().decode("utf-8", "replace"))

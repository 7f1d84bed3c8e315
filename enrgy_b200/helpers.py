"""Unit/CSV helpers of the model path, same names and semantics as reference helpers.py."""
from __future__ import annotations

from .forcing import get_time_step, heuristic_unit_guesser, read_input_file  # noqa: F401 (re-export)

CSV_HEADER_LINES = (
    "# DATE format is %Y%m%d, HEAT FLUXES are in W m-2",
    "# ICE and SNOW_MELT are in m w.e.",
    "\n# POINT_T_SURF (degree Celsius) is near the point of glacier body temperature measurements",
    "\nDATE,RS_BALANCE,RL_BALANCE,LWD_FLUX,SENSIBLE,LATENT,ATMO_BALANCE,INSIDE_GLACIER_FLUX,MELT_FLUX,"
    "POINT_T_SURF,SNOW_MELT,ICE_MELT,SNOW_COVER,SNOW_COVER_PERCENT_FROM_SURFACE",
)


def J_to_W(insol, time_step=None):
    """helpers.py:27-36: energy per period [J] -> mean flux [W]; one day by default."""
    if time_step is None:
        time_step = 86400
    return insol / time_step


def kWh_to_J(insol):
    """helpers.py:54-60."""
    return insol * 3.6 * 10 ** 6


def fill_header(out_file):
    """helpers.py:39-45 (the first two comment strings share a line there as well)."""
    with open(out_file, "w") as output:
        for part in CSV_HEADER_LINES:
            output.write(part)

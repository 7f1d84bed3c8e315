"""GPU parity: the CUDA path (through the C ABI) against the NumPy oracle on the same inputs.

Bars (BASELINE.json north_star): float64 mode 1e-9 relative, float32 mode 1e-4 relative, relative
error taken against max(|ref|, floor) because fluxes cross zero:
floor = 1e-3 W m-2 (f64) / 1 W m-2 (f32) for fluxes, 1e-7 / 1e-3 m w.e. for per-step melt (the
last snow of a cell melts as `swe` itself, a float32 state of magnitude 0.5 m whose spacing is
3e-8 m), 1e-6 / 1e-3 m w.e. for season totals.
"""
import numpy as np
import pytest

from enrgy_b200.synthetic import make_case
from tests import parity as P

pytestmark = pytest.mark.gpu

VARIANTS = {
    "maps": dict(),
    "const_albedo": dict(const_albedo=(0.35, 0.75)),
    "snow_ageing": dict(last_snowfall="20220522", max_ice_albedo=0.38),
    "andreas": dict(andreas=True),
    "corr": dict(cloud_corr=0.2, sensible_corr=1.1, latent_corr=0.9, emissivity=None, zm=None,
                 z_h_or_e=None),
    "no_swe": dict(use_swe=False, const_albedo=(0.35, 0.75)),
    "msm": dict(msm=dict(depths=[0.1, 0.1, 0.3, 0.5, 0.5, 0.5, 3.0],
                         temperatures=[-6.9, -6.93, -7.025, -7.31, -6.93, -7.12, -7.0, -5.57], elev=275.0),
                snow_density=350.0, last_snowfall="20220522"),
    "msm_const": dict(msm=dict(depths=[0.25, 0.5, 1.0], temperatures=[-2.0, -3.0, -4.0, -5.0], elev=400.0),
                      const_albedo=(0.35, 0.75)),
}


@pytest.mark.parametrize("variant", list(VARIANTS))
def test_streamed_f64(variant):
    case = make_case(96, 12, calm_every=5, w=200)
    pot = P.random_insolation(case, 12)
    res = P.compare_run(case, True, pot=pot, **VARIANTS[variant])
    print(variant, res)
    for k, v in res.items():
        assert v < 1e-9, (k, v)


@pytest.mark.parametrize("variant", list(VARIANTS))
def test_streamed_f32(variant):
    case = make_case(96, 12, calm_every=5, w=200)
    pot = P.random_insolation(case, 12)
    res = P.compare_run(case, False, pot=pot, **VARIANTS[variant])
    print(variant, res)
    for k, v in res.items():
        assert v < 1e-4, (k, v)


@pytest.mark.parametrize("f64", [True, False])
def test_msm_with_in_kernel_insolation(f64):
    case = make_case(72, 30, w=100, seed=41)
    res = P.compare_run(case, f64, computed=True, shadow=True, **VARIANTS["msm"])
    print(f64, res)
    tol = 1e-9 if f64 else 1e-4
    for k, v in res.items():
        assert v < tol, (k, v)


@pytest.mark.parametrize("f64", [True, False])
@pytest.mark.parametrize("shadow", [False, True])
def test_computed_insolation(f64, shadow):
    case = make_case(128, 26, w=160)
    res = P.compare_run(case, f64, computed=True, shadow=shadow)
    print(f64, shadow, res)
    tol = 1e-9 if f64 else 1e-4
    for k, v in res.items():
        assert v < tol, (k, v)


@pytest.mark.parametrize("f64", [True, False])
def test_computed_insolation_high_sun_gentle_slopes(f64):
    """Mid-latitude summer days on gentle slopes: for most rows the sun stands above every slope of a
    patch in all sub-steps, where the kernel replaces the per-sub-step max(cos i, 0) sum by the row's
    analytic sums; steeper patches, dawn and dusk take the sub-step path in the same run."""
    case = make_case(128, 40, w=160, seed=43, start="20220620 03:00:00")
    r, c = np.mgrid[0:128, 0:160].astype(np.float64)
    nan = np.isnan(case.dem)
    dem = 1500.0 + 1.2 * r + 30.0 * np.sin(c / 25.0) + 6.0 * np.cos(r / 7.0) + 0.15 * (case.dem.astype(np.float64) - np.nanmean(case.dem))
    dem[:, 100:] += 3.5 * (c[:, 100:] - 100)          # a steep flank (35 % grade): its patches keep the sub-step path
    dem[nan] = np.nan
    case.dem[...] = dem.astype(np.float32)
    case.elev_aws = float(case.dem[case.aws_rc])
    case.lat, case.lon = 46.0, 8.0
    res = P.compare_run(case, f64, computed=True, shadow=False)
    print(f64, res)
    tol = 1e-9 if f64 else 1e-4
    for k, v in res.items():
        assert v < tol, (k, v)
    # both paths really occur: tan(sun elevation) runs from 0.02 at dawn to 2.3 at noon, the patches'
    # steepest slopes from ~0.2 to ~0.5
    from oracle import insolation_oracle as I
    tan2 = []
    for row in case.aws_rows:
        tab = I.substep_table(I.to_unix(row["DATE"]), 3600, 46.0, 8.0, case.cell)
        if tab:
            tan2.append(min(s["U"] ** 2 / (s["E"] ** 2 + s["N"] ** 2) for s in tab))
    assert max(tan2) > 1.0 and min(tan2) < 0.01

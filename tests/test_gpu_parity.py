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

"""Scalar known answers captured from the reference (numpy 2.3.5, float64)."""
import numpy as np
import pytest

from oracle import enrgy_oracle as O

RTOL = 1e-12


def test_turbulent_fluxes_known_answers():
    f = O.turbulent_fluxes
    cases = [
        (dict(z=1.6, uz=2.5, Tz=276.15, P=99000, rh=0.85, surface_temp=273.15, zm=1e-3, z_h_or_e=1e-4),
         (20.022365228122418, 3.437569577804156, 7.279896018496393)),
        (dict(z=1.6, uz=2.5, Tz=270.15, P=99000, rh=0.85, surface_temp=273.15, zm=1e-3, z_h_or_e=1e-4),
         (-21.336263310909324, -21.714868609907597, -11.987122239228613)),
        (dict(z=1.6, uz=0.1, Tz=276.15, P=99000, rh=0.85, surface_temp=273.15, zm=1e-3, z_h_or_e=1e-4),
         (1.8928122048674906e-07, 0.14528279191991347, 40.26157070077273)),
        (dict(z=1.6, uz=2.5, Tz=276.15, P=99000, rh=0.85, surface_temp=273.15, zm=1e-3, andreas=True),
         (21.26456371260018, 3.6505425891167613, 6.599655787397531)),
        (dict(z=2, uz=4, Tz=278.15, P=95000, rh=.6, surface_temp=None),
         (49.361715414038976, -13.770591743912142, 11.644416150922229)),
        # the reference's own __main__ demo, turbo.py:382-398
        (dict(z=1.6, uz=2.5, Tz=276.15, P=99000, rh=0.85, surface_temp=None, zm=0.01),
         (39.21029634109856, 7.353104473501779, 11.846208732123374)),
    ]
    for kw, want in cases:
        rh = kw.pop("rh")
        got = f(kw.pop("z"), kw.pop("uz"), kw.pop("Tz"), kw.pop("P"), rh, **kw)
        assert np.allclose(got, want, rtol=RTOL, atol=0), (got, want)


def test_psi_and_coefficients():
    assert O.minus_psi_m(1.6, 11.8) == pytest.approx(0.688494401082556, rel=RTOL)
    assert O.minus_psi_h(1.6, 11.8) == pytest.approx(0.6899804147445412, rel=RTOL)
    assert O.minus_psi_m(1.6, -20.0) == pytest.approx(-0.23915979174700652, rel=RTOL)
    assert O.minus_psi_h(1.6, -20.0) == pytest.approx(-0.4542447601760426, rel=RTOL)
    assert O.exchange_coef(1.6, zm=1e-3, z_h_or_e=1e-4) == pytest.approx(0.0022402925404105363, rel=RTOL)
    assert O.exchange_coef(1.6, L=11.8, zm=1e-3, z_h_or_e=1e-4) == pytest.approx(0.0021911226774993724, rel=RTOL)
    assert O.friction_velocity(2.5, 1.6, L=11.8, zm=1e-3) == pytest.approx(0.12397329486860466, rel=RTOL)
    assert O.e_max(276.15, 99000) == pytest.approx(761.1500572718732, rel=RTOL)
    assert O.e_max(273.15, 99000) == pytest.approx(614.0382615434344, rel=RTOL)
    assert O.andreas_z0(2.5, 1.6, 1e-3, 11.8) == pytest.approx(0.00018404390612247393, rel=RTOL)


def test_melt_and_tick():
    snow, ice = O.melt_partition(300.0, 0.001, 3600, O.default_params())
    assert snow == 0.001 and ice == pytest.approx(0.0022335329341317367, rel=RTOL)
    temps, qm, g = O.subsurface_tick([.1, .1, .3, .5, .5, .5, 3.], [-6.9, -6.93, -7.025, -7.31, -6.93, -7.12, -7.0, -5.57],
                                     3600, flux=200., snow_depth=.25)
    want = [0.0, -6.93936, -7.025, -7.295718079999999, -6.93952128, -7.11482176, -6.99967056, -5.57]
    assert np.allclose(temps, want, rtol=1e-12, atol=1e-12)
    assert qm == pytest.approx(44.35764031999997, rel=1e-11)
    assert g == pytest.approx(-0.09738467999999793, rel=1e-10)


def test_albedo_blend_integer_days():
    a0 = np.arange(4.).reshape(2, 2)
    got = O.albedo_blend({"20190727": a0, "20190803": a0 + 4}, "20190731 13:00:00")
    assert np.array_equal(got, a0 + 4 * 4 / 7)
    assert O.albedo_blend({"20190727": a0, "20190803": a0 + 4}, "20190727") is a0
    with pytest.raises(ValueError):
        O.albedo_blend({"20190727": a0, "20190803": a0 + 4}, "20190710")


def test_unit_guess_and_time_step():
    assert O.unit_guess(85.0, 100) == 0.85 and O.unit_guess(0.7, 100) == 0.7
    with pytest.raises(ValueError):
        O.unit_guess(120.0, 100)
    rows = [{"DATE": "20220601 00:00:00"}, {"DATE": "20220601 01:00:00"}, {"DATE": "20220601 03:00:00"}]
    assert [O.time_step_seconds(rows, i) for i in range(3)] == [3600, 7200, 7200]

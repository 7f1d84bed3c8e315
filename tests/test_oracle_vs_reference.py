"""Live pin: the oracle against the unmodified reference run in this container (skipped where the
reference checkout is absent, e.g. on the GPU box)."""
import numpy as np
import pytest

from enrgy_b200.synthetic import make_case
from oracle import ref_harness
from tests import parity as P

pytestmark = pytest.mark.skipif(not ref_harness.reference_available(), reason="reference checkout absent")


@pytest.mark.parametrize("f64", [False, True])
@pytest.mark.parametrize("variant", ["maps", "const_albedo", "snow_ageing", "andreas", "corr"])
def test_bit_identical(f64, variant):
    from tests.test_gpu_parity import VARIANTS
    kw = dict(VARIANTS[variant])
    case = make_case(48, 7, seed=21, w=40, calm_every=3)
    pot = P.random_insolation(case, 7, seed=9)
    ref = ref_harness.run_reference(case, pot, f64=f64, **{**dict(z=1.6, zm=1e-3, z_h_or_e=1e-4, emissivity=0.98), **kw})
    ora = P.run_oracle(case, pot, f64, **kw)
    for k in ("swe", "total_snow", "total_ice"):
        assert ref[k].dtype == ora[k].dtype
        assert np.array_equal(ref[k], ora[k], equal_nan=True), k
    for i in range(7):
        for k in ("lwd", "lwu", "rs", "sens", "lat", "atmo", "g", "mf"):
            assert np.array_equal(ref["rows"][i][k], ora["rows"][i][k], equal_nan=True), (i, k)
    assert ref["stats_csv"] == ora["stats_csv"]
    assert ref["solar_csv"] == ora["solar_csv"]

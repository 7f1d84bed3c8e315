"""Ensemble members (config C5) against the oracle run on the perturbed inputs."""
import numpy as np
import pytest

from enrgy_b200.ensemble import make_members, run_members, shard
from enrgy_b200.synthetic import make_case
from tests import parity as P

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("const", [False, True])
def test_members_match_reference_on_perturbed_inputs(const):
    case = make_case(64, 14, w=80, seed=61)
    pot = P.random_insolation(case, 14)
    kw = dict(const_albedo=(0.35, 0.75)) if const else dict(last_snowfall="20220525")
    members = make_members(3, seed=7, albedo_sigma=0.08)
    members.append(dict(albedo_offset=0.0, zm=1e-3, z_h_or_e=1e-4))   # the unperturbed member (None = keep!)
    eng = P.make_engine(case, True, pot=pot, **kw)
    try:
        got = run_members(eng, members, keep_rasters=True)
    finally:
        eng.close()
    for i, m in enumerate(members):
        off = m["albedo_offset"]
        c2 = make_case(64, 14, w=80, seed=61)
        okw = dict(kw)
        if const:
            okw["const_albedo"] = tuple(float(np.clip(a + off, 0.001, 1.0)) if off else a for a in kw["const_albedo"])
        else:
            for k in c2.albedo_maps:
                a = c2.albedo_maps[k].astype(np.float64)
                a[a < 0] = 0.001
                a[a > 1] = 1
                c2.albedo_maps[k] = np.clip(a + off, np.float64(np.float32(0.001)), 1.0)
        if m.get("zm") is not None:
            okw["zm"], okw["z_h_or_e"] = m["zm"], m["z_h_or_e"]
        ora = P.run_oracle_arrays(c2, pot, True, **okw)
        assert P.max_rel_err(got[i]["total_ice"], ora["total_ice"], 1e-3) < 2e-7, i
        assert P.max_rel_err(got[i]["swe"], ora["swe"], 1e-3) < 2e-7, i
        # the per-member totals come from the per-step area sums: same number as the raster mean
        assert abs(got[i]["mean_ice"] - np.nanmean(got[i]["total_ice"])) < 1e-6 * max(np.nanmean(got[i]["total_ice"]), 1e-3), i
    assert abs(got[0]["mean_ice"] - got[3]["mean_ice"]) > 1e-6    # the perturbation really acts
    assert shard(members, 3, 1) == [1]

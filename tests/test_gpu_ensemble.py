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


@pytest.mark.parametrize("f64", [False, True])
@pytest.mark.parametrize("mode", ["streamed", "computed", "shadow", "const"])
def test_fused_members_equal_one_run_per_member(f64, mode):
    """enrgy_run_members (four members per pass, a remainder in pairs, an odd one padded) gives the state
    rasters of one enrgy_run per member bit for bit and its per-step statistics to rounding -- 5 members,
    so that a group of four, a pair and the padding copy are all exercised; glacier margin included;
    with and without the statistics."""
    case = make_case(72, 30, w=150, seed=5)
    pot = P.random_insolation(case, 30) if mode in ("streamed", "const") else None
    kw = dict(const_albedo=(0.35, 0.75)) if mode == "const" else dict(last_snowfall="20220525") if mode == "computed" else {}
    members = make_members(5, seed=3, albedo_sigma=0.08)
    members[2] = dict(albedo_offset=0.9, zm=2e-3, z_h_or_e=1e-4)          # clipped at 1
    eng = P.make_engine(case, f64, pot=pot, computed=mode in ("computed", "shadow"), shadow=mode == "shadow", **kw)
    try:
        one = run_members(eng, members, keep_rasters=True, fused=False)
        eng.snapshot(save=False)
        fused = run_members(eng, members, keep_rasters=True, fused=True)
        nostat = run_members(eng, members, keep_rasters=True, fused=True, want_stats=False)
        # a sub-range and a subset of the members, after the fused run: the handle's own state was left alone
        eng.prepass()
        stats_own = eng.run(0, 30)
    finally:
        eng.close()
    for i in range(len(members)):
        for k in ("swe", "total_snow", "total_ice"):
            assert np.array_equal(one[i][k], fused[i][k], equal_nan=True), (i, k)
            assert np.array_equal(one[i][k], nostat[i][k], equal_nan=True), (i, k)
        # (the passes walk a patch in two halves: float32 sums in another order)
        assert np.allclose(one[i]["stats"], fused[i]["stats"], rtol=1e-12 if f64 else 2e-6, atol=1e-9 if f64 else 1e-3), i
        assert abs(fused[i]["mean_ice_raster"] - np.nanmean(fused[i]["total_ice"].astype(np.float64))) < 1e-6
    assert np.isfinite(stats_own).all()
    assert abs(fused[0]["mean_ice"] - fused[1]["mean_ice"]) > 1e-7


def test_fused_members_on_row_bands():
    """Fused members on two row bands (one engine per band, as one rank per GPU would hold them) give the
    whole raster's member states bit for bit; with shading, so the bands see terrain outside themselves."""
    case = make_case(80, 20, w=96, seed=9)
    members = make_members(4, seed=2, albedo_sigma=0.05)
    kw = dict(computed=True, shadow=True)
    whole = P.make_engine(case, False, **kw)
    try:
        ref = run_members(whole, members, keep_rasters=True, want_stats=False)
    finally:
        whole.close()
    for band in ((0, 32), (32, 48)):
        eng = P.make_engine(case, False, band=band, **kw)
        try:
            got = run_members(eng, members, keep_rasters=True, want_stats=False)
        finally:
            eng.close()
        for i in range(len(members)):
            for k in ("swe", "total_snow", "total_ice"):
                assert np.array_equal(got[i][k], ref[i][k][band[0]:band[0] + band[1]], equal_nan=True), (i, k, band)

"""Parameter ensembles (BASELINE config C5): members share the DEM, terrain, albedo maps and forcing
on the device and differ in an albedo offset and the roughness lengths.  Members are independent, so
the ensemble axis shards over ranks with no collective except gathering the per-member totals.

The reference has no ensemble code; a member is defined as the reference run on perturbed inputs
(albedo maps / constants shifted by the offset and clipped to [0.001, 1] like raster_utils.py:48-50,
`zm` / `z_h_or_e` replaced), which is what tests/test_gpu_ensemble.py checks against the oracle.
"""
from __future__ import annotations

import numpy as np

from . import _lib


def make_members(n, seed=0, albedo_sigma=0.03, zm_range=(3e-4, 3e-3)):
    """Seeded perturbations of SURVEY.md 8(d) C5: albedo offset ~ N(0, sigma), zm log-uniform;
    z_h_or_e = zm / 10 (the reference's default ratio, turbo.py:275-277)."""
    rng = np.random.default_rng(seed)
    off = albedo_sigma * rng.standard_normal(n)
    zm = np.exp(rng.uniform(np.log(zm_range[0]), np.log(zm_range[1]), n))
    return [dict(albedo_offset=float(o), zm=float(z), z_h_or_e=float(z) / 10.0) for o, z in zip(off, zm)]


def shard(members, world, rank):
    """Member indices of `rank` (round-robin)."""
    return list(range(rank, len(members), world))


def _result(stats, n_steps):
    # glacier-wide melt totals from the per-step area sums (the melt rasters add up to exactly
    # these, tests/test_gpu_full_size.py): no raster leaves the device unless it is asked for
    n_valid = float(stats[-1, _lib.S_NVALID]) if n_steps else 0.0
    return dict(stats=stats,
                mean_ice=float(stats[:, _lib.S_ICE].sum() / n_valid) if n_valid else float("nan"),
                mean_snow=float(stats[:, _lib.S_SNOW].sum() / n_valid) if n_valid else float("nan"))


def run_members(engine, members, indices=None, n_steps=None, keep_rasters=False, fused=True, want_stats=True):
    """Runs the given members on a loaded Engine (DEM, albedo maps, SWE, forcing set; state = initial);
    every member starts from the engine's current state.  Returns {index: dict(stats=[T, S_COUNT],
    mean_ice, mean_snow[, rasters])}.

    fused (default): enrgy_run_members -- four members per pass of the kernel share everything a member
    does not change (terrain, insolation and sunlit masks, lapse-rate meteorology, flux factors, net
    longwave); the same state rasters as one run per member, bit for bit (float32 statistics to rounding:
    the passes sum in another order).  want_stats=False skips the per-step statistics in the fused passes
    (a third of a member's share of a step); mean_ice / mean_snow then come from the final rasters.
    The sub-surface model (8 temperatures per cell and member) runs the members one after the other."""
    n_steps = engine.n_steps if n_steps is None else n_steps
    indices = list(range(len(members)) if indices is None else indices)
    out = {}
    if fused and getattr(engine, "msm_layers", 0) == 0 and indices:
        sel = [members[i] for i in indices]
        stats, totals = engine.run_members([m.get("albedo_offset", 0.0) for m in sel], [m.get("zm") for m in sel],
                                           [m.get("z_h_or_e") for m in sel], 0, n_steps, want_stats=want_stats)
        for k, i in enumerate(indices):
            res = _result(stats[k], n_steps) if want_stats else dict(stats=None)
            # (the season totals of the rasters; equal to the per-step sums added up, to rounding)
            res.update(mean_ice_raster=float(totals[k, 2]), mean_snow_raster=float(totals[k, 1]))
            if not want_stats:
                res.update(mean_ice=float(totals[k, 2]), mean_snow=float(totals[k, 1]))
            if keep_rasters:
                swe, tsn, tic = engine.member_state(k, np.float32)
                res.update(swe=swe, total_snow=tsn, total_ice=tic)
            out[i] = res
        return out
    engine.snapshot(save=True)
    for i in indices:
        m = members[i]
        engine.snapshot(save=False)
        engine.synchronize()
        engine.set_member(m.get("albedo_offset", 0.0), m.get("zm"), m.get("z_h_or_e"))
        engine.prepass()
        res = _result(engine.run(0, n_steps), n_steps)
        if keep_rasters:
            swe, tsn, tic = engine.state(np.float32)
            res.update(swe=swe, total_snow=tsn, total_ice=tic)
        out[i] = res
    engine.snapshot(save=False)          # like the fused passes: the engine's own state is left alone
    return out

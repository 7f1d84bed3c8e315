"""Multi-GPU plumbing: one process per GPU, the raster cut into row bands.

Cells are independent once the AWS-cell pre-pass has produced the per-step scalars, so the data
path needs NO collective: every rank gets the full DEM (replicated, for the shading rays) and its
own band of the albedo / SWE / state rasters.  The only exchange is the sum of the per-step area
statistics (`[T, S_COUNT]` float64 sums and counts) -- one all-reduce per pass, NCCL for CUDA
tensors (NVLink/NVSwitch), gloo on the CPU in the tests.
"""
from __future__ import annotations

import numpy as np

from . import _lib


def row_bands(rows, world, align=16, valid_per_row=None):
    """[(row0, n_rows)] for `world` ranks.  Band edges are multiples of `align` (the tile height of
    the fused kernel).  With `valid_per_row` (glacier cells per raster row) the bands are balanced
    by glacier cells instead of rows -- off-glacier tiles cost nothing (load balance)."""
    if world < 1:
        raise ValueError("world must be >= 1")
    if valid_per_row is None:
        weights = np.ones(rows, dtype=np.float64)
    else:
        weights = np.asarray(valid_per_row, dtype=np.float64)
        if weights.shape != (rows,):
            raise ValueError("valid_per_row must have one entry per raster row")
    cum = np.concatenate([[0.0], np.cumsum(weights)])
    total = cum[-1]
    edges = [0]
    for r in range(1, world):
        target = total * r / world
        e = int(np.searchsorted(cum, target))
        e = int(round(e / align)) * align
        e = min(max(e, edges[-1]), rows)
        edges.append(e)
    edges.append(rows)
    return [(edges[i], edges[i + 1] - edges[i]) for i in range(world)]


def tile_cost_per_row(valid, tile_h=8, tile_w=128):
    """Cost weights for row_bands(valid_per_row=...) in units of VISITED TILES: the fused kernel walks
    tiles of tile_h x tile_w cells and a tile with one glacier cell costs as much as a full one, so a
    band along the glacier margin is more expensive than its glacier-cell count says.  `valid` is the
    boolean glacier mask [rows, cols]; every row gets 1 / tile_h of the tiles its row group touches."""
    valid = np.asarray(valid, dtype=bool)
    rows, cols = valid.shape
    hp, wp = -(-rows // tile_h) * tile_h, -(-cols // tile_w) * tile_w
    v = np.zeros((hp, wp), dtype=bool)
    v[:rows, :cols] = valid
    tiles = v.reshape(hp // tile_h, tile_h, wp // tile_w, tile_w).any(axis=(1, 3)).sum(axis=1)
    return np.repeat(tiles / float(tile_h), tile_h)[:rows]


def rebalance_bands(bands, seconds, valid_per_row, align=16):
    """New band edges from MEASURED per-band times: with the shading ray march the cost of a glacier
    cell depends on the terrain around it, so equal glacier-cell counts are not equal times.  The
    measured time of every band is spread over its rows in proportion to their glacier cells
    (piecewise-constant cost per glacier cell), and the cuts are placed at equal shares of that
    estimate.  One or two rounds settle (scripts/measure_modes.py strong)."""
    w = np.asarray(valid_per_row, dtype=np.float64)
    rows = w.shape[0]
    if len(bands) != len(seconds):
        raise ValueError("one measured time per band")
    cost = np.zeros(rows, dtype=np.float64)
    for (r0, n), t in zip(bands, seconds):
        tot = w[r0:r0 + n].sum()
        if n > 0:
            cost[r0:r0 + n] = (w[r0:r0 + n] * (float(t) / tot)) if tot > 0 else 0.0
    return row_bands(rows, len(bands), align=align, valid_per_row=cost)


def allreduce_stats(stats):
    """In-place SUM of a [T, S_COUNT] statistics tensor over all ranks (torch.distributed)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
    return stats


def means_from_sums(stats):
    """Area means laid out like the reference's CSV columns (var_classes.py:45-56, model.py:246-252)
    from the summed statistics: [T, 12] = rs, rl, lwd, sens, lat, atmo, g, melt, snow, ice, swe,
    snow cover per cent."""
    s = np.asarray(stats, dtype=np.float64)
    nv = s[:, _lib.S_NVALID]
    with np.errstate(invalid="ignore", divide="ignore"):
        cover = np.round(s[:, _lib.S_NSNOW] / s[:, _lib.S_NSWE] * 100)
        cols = [s[:, _lib.S_RS] / nv, (s[:, _lib.S_LWD] - s[:, _lib.S_LWU]) / nv, s[:, _lib.S_LWD] / nv,
                s[:, _lib.S_SENS] / nv, s[:, _lib.S_LAT] / nv, s[:, _lib.S_ATMO] / nv, s[:, _lib.S_G] / nv,
                s[:, _lib.S_MELT] / nv, s[:, _lib.S_SNOW] / nv, s[:, _lib.S_ICE] / nv,
                s[:, _lib.S_SWE] / s[:, _lib.S_NSWE], cover]
    return np.stack(cols, axis=1)

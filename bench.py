#!/usr/bin/env python
"""bench.py -- cell-timesteps/s of the fused surface-energy-balance path on B200.

Workload (BASELINE.json configs[1], "C2"): 2048 x 2048 synthetic 10 m DEM + 5 albedo maps, one
ablation season of hourly AWS rows (2200 steps), potential insolation computed in the kernel
without shadows.  A "step" of this benchmark is ONE PASS OF THE WHOLE SEASON over the raster.
With N > 1 GPUs the raster grows to (N*2048) x 2048 and is cut into N row bands, one per rank
(weak scaling); the DEM is replicated, the only exchange is one NCCL all-reduce of the per-step
area statistics [T x 15] float64 per pass.

  value  whole-job cell-timesteps/s, inputs resident in HBM, CUDA events on the launching stream
  e2e    the same through the public API from pinned HOST buffers: upload of every raster + forcing,
         pre-pass, kernels, download of the three state rasters + statistics, every pass
  --impl reference   the CPU path (oracle/enrgy_oracle.py, a bit-exact NumPy restatement of the
         reference's Energy.model pinned against it) on the host cores, one process per core on
         row bands, on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_CELLS = 2048
N_STEPS = 2200
# Algorithmic work per glacier cell-step.  SURVEY.md 8(d): energy-balance core 93 FLOP (+ 2 exp, not
# counted) + 12 FLOP per insolation sub-step; C2 has 4 sub-steps per step -> 141 FLOP.  That is the
# REFERENCE's arithmetic after hoisting per-cell invariants; the kernel executes less (DESIGN.md 4.1:
# ez = e, exp(0) = 1, one reciprocal per quantity, analytic longwave sum, daily albedo blend, flux
# scalars folded into the balance FMA chain): 77 FLOP = 46.5 FP32-pipe operations (FMA/ADD/MUL, two
# cells per packed instruction) + 14 min/max/select/compare/MUFU/SHFL operations on the other pipes,
# counted in the SASS of the hot basic block (346 instructions per 256 cell-steps; rows that take the
# analytic direct-beam path execute 40 packed instructions fewer, not credited here).  Both are
# reported; `roofline.frac` uses the SURVEY figure as the contract asks.
FLOP_PER_CELL_STEP = 141.0
FLOP_EXECUTED_PER_CELL_STEP = 77.0
FP32_PIPE_OPS_PER_CELL_STEP = 46.5       # lane-operations on the FMA pipe (a packed FFMA2 is two)
ISSUE_SLOTS_PER_CELL_STEP = 39.5         # 42.5 packed / 2 + 4 scalar FMA-pipe + 14 ALU/XU/SHFL instructions
METRIC = "cell-timesteps/s"
SHADOW = False                   # --shadow: C3-style run with the per-sub-step shading ray march


def log(*a):
    print(*a, file=sys.stderr, flush=True)


_JSON_FD = None


def claim_stdout():
    """Keeps fd 1 for the ONE JSON line: everything else written to stdout by this process or by
    native libraries (NCCL prints its version banner there) is sent to stderr."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(obj):
    line = (json.dumps(obj) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(line.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, line)


# ---------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML during the timed region."""

    def __init__(self, index, period=0.05):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception as e:          # pragma: no cover
            log("clock sampler unavailable:", e)

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ---------------------------------------------------------------------------------------------
def build_engine(case, dem_full, precision, device, pinned=None):
    from enrgy_b200 import _lib
    from enrgy_b200.engine import Engine
    from enrgy_b200.forcing import build_forcing
    m = case.meta
    eng = Engine(m["rows_full"], case.dem.shape[1], precision=precision, device=device)
    eng.set_params(cell_size=case.cell, elev_aws=case.elev_aws, aws_row=case.aws_rc[0],
                   aws_col=case.aws_rc[1], sensor_z=1.6, zm=1e-3, z_h_or_e=1e-4, emissivity=0.98,
                   insol_mode=_lib.INSOL_COMPUTED, shadow=SHADOW, lat=case.lat, lon=case.lon,
                   band_row0=m["band_row0"], band_rows=case.dem.shape[0])
    keys = list(case.albedo_maps)
    table = build_forcing(case.aws_rows, keys)
    upload(eng, case, dem_full, table, pinned)
    return eng, table


def upload(eng, case, dem_full, table, pinned=None):
    """Everything a user's Energy.model() call moves to the device, from (pinned) host memory."""
    src = pinned if pinned is not None else {
        "dem": dem_full, "swe": case.swe, "alb": [case.albedo_maps[k] for k in case.albedo_maps]}
    eng.set_dem(src["dem"])
    eng.set_forcing(table)            # the host pre-pass starts here and overlaps the raster uploads
    eng.set_albedo_maps(src["alb"])
    eng.set_swe(src["swe"])
    eng.prepass()


def pin(arr):
    import torch
    t = torch.empty(arr.shape, dtype=torch.float32, pin_memory=True)
    t.numpy()[...] = arr
    return t


def run_ours(args):
    import torch
    import torch.distributed as dist
    from enrgy_b200 import _lib

    claim_stdout()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        log("warning: WORLD_SIZE=%d but --gpus=%d; using WORLD_SIZE" % (world, args.gpus))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    n, T = args.n, args.t
    case, dem_full = make_workload(n, T, world, rank)
    precision = _lib.F32 if args.dtype == "f32" else _lib.F64
    stream = torch.cuda.Stream()
    t_setup = time.time()
    eng, table = build_engine(case, dem_full, precision, local_rank)
    eng.set_stream(stream.cuda_stream)
    log("rank %d: setup %.1f s, kernel %s" % (rank, time.time() - t_setup, eng.kernel_info()))
    stats = torch.zeros((T, _lib.S_COUNT), dtype=torch.float64, device="cuda")
    eng.snapshot(save=True)

    def one_pass():
        eng.snapshot(save=False)                       # rewind the season (device-to-device)
        eng.run_async(0, T, stats.data_ptr(), None)    # fused kernel + statistics finalize
        if world > 1:
            with torch.cuda.stream(stream):
                dist.all_reduce(stats)                 # glacier-wide sums (NCCL)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        one_pass()
    barrier()
    sampler = ClockSampler(local_rank, period=0.02)
    sampler.start()
    l0 = eng.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    kernel_ms = []
    for _ in range(args.steps):
        one_pass()
    e1.record(stream)
    barrier()
    clocks = sampler.stop()
    launches = eng.launch_count() - l0 + (args.steps if world > 1 else 0)
    ms_total = e0.elapsed_time(e1)
    kernel_ms = eng.last_kernel_ms()                   # fused kernel alone, last pass
    t_ms = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_step = float(t_ms.item()) / args.steps
    cell_steps = float(n) * n * world * T
    value = cell_steps / (ms_step * 1e-3)
    stats_host = stats.cpu().numpy()

    # ---- end to end through the public API from pinned host buffers ---------------------------
    pinned_t = {"dem": pin(dem_full), "swe": pin(case.swe), "alb": [pin(case.albedo_maps[k]) for k in case.albedo_maps]}
    pinned = {"dem": pinned_t["dem"].numpy(), "swe": pinned_t["swe"].numpy(), "alb": [t.numpy() for t in pinned_t["alb"]]}
    out_state = [torch.empty(case.dem.shape, dtype=torch.float32, pin_memory=True) for _ in range(3)]
    stats_h = torch.empty((T, _lib.S_COUNT), dtype=torch.float64, pin_memory=True)
    from enrgy_b200._lib import check

    def e2e_pass():
        upload(eng, case, dem_full, table, pinned)
        check(eng.lib.enrgy_run(eng.h, 0, T, stats_h.numpy().ctypes.data))
        check(eng.lib.enrgy_get_state(eng.h, 32, *[o.numpy().ctypes.data for o in out_state]))
        if world > 1:
            g = stats_h.cuda(non_blocking=True)
            dist.all_reduce(g)
            stats_h.copy_(g)
    eng.set_stream(None)
    e2e_steps = max(1, min(args.steps, 5))
    e2e_pass()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_pass()
    barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    t_e = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
    e2e_value = cell_steps / float(t_e.item())
    # the DEM goes up whole only with shading (the rays leave the band); else the band + one row either side
    m = case.meta
    dem_rows_up = dem_full.shape[0] if SHADOW else (min(dem_full.shape[0], m["band_row0"] + case.dem.shape[0] + 1) - max(0, m["band_row0"] - 1))
    h2d = dem_rows_up * dem_full.shape[1] * 4 + case.swe.nbytes + sum(a.nbytes for a in case.albedo_maps.values()) + table.nbytes
    d2h = 3 * case.dem.size * 4 + stats_h.numel() * 8

    result = None
    if rank == 0:
        peak32 = eng.microbench(0)
        peak64 = eng.microbench(1)
        peak = peak32 if args.dtype == "f32" else peak64
        # algorithmic FLOPs are counted on GLACIER cells only (off-glacier cells are skipped, they
        # count toward the metric's H*W*T but do no arithmetic)
        # (this rank's band: the kernel time below is this rank's too; the statistics are global)
        n_valid = float(np.count_nonzero(~np.isnan(case.dem)))
        achieved = FLOP_PER_CELL_STEP * n_valid * T / (kernel_ms * 1e-3) / 1e12
        peaks_file = os.path.join(ROOT, "MEASURED_PEAKS.json")
        hbm = json.load(open(peaks_file))["hbm_gbs"] if os.path.isfile(peaks_file) else 6650.0
        bytes_per_launch = algorithmic_bytes(case, precision)
        result = {
            "metric": METRIC, "value": value, "unit": "cell-timesteps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": "C2: %dx%d 10 m synthetic DEM + 5 albedo maps per GPU, %d hourly steps, "
                                   "in-kernel insolation (4 sub-steps/step), %s" % (n, n, T, "shading ray march" if SHADOW else "no shading"),
                       "raster": [n * world, n], "steps_per_pass": T, "parallelism": "row bands x%d, balanced by visited tiles" % world,
                       "band_rows": [b[1] for b in case.meta["bands"]],
                       "l2": "per-pass inputs ~%.0f MB > 126 MB L2, no flush" % (bytes_per_launch / 1e6)},
            "roofline": {"bound": "fp32" if args.dtype == "f32" else "fp64",
                         "bound_note": "no dense contraction and 0.03 B of HBM traffic per cell-step: neither 'tensor' nor 'hbm' binds; the FP32 (FP64) pipe does" + ("; with --shadow the ray march dominates (integer/LDS work, see profiles/r01_summary.md) and this FLOP roofline covers the energy balance only" if SHADOW else ""), "achieved": achieved, "peak": peak,
                         "unit": "TFLOP/s", "frac": achieved / peak if peak else None,
                         "traffic": measured_traffic(args.dtype, SHADOW),
                         "peak_source": "enrgy_microbench FMA loop on this GPU (MEASURED_PEAKS.json has no FP32/FP64 pipe peak)",
                         "flop_per_cell_step": FLOP_PER_CELL_STEP, "kernel_ms": kernel_ms,
                         "flop_source": "SURVEY.md 8(d): 93 (core) + 4 x 12 (insolation sub-steps), glacier cells only",
                         "as_executed": {"flop_per_cell_step": FLOP_EXECUTED_PER_CELL_STEP,
                                         "achieved": FLOP_EXECUTED_PER_CELL_STEP * n_valid * T / (kernel_ms * 1e-3) / 1e12,
                                         "frac": (FLOP_EXECUTED_PER_CELL_STEP * n_valid * T / (kernel_ms * 1e-3) / 1e12 / peak) if peak else None},
                         "glacier_cell_fraction": n_valid / float(case.dem.size),
                         "hbm": {"achieved": bytes_per_launch / (kernel_ms * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                                 "frac": bytes_per_launch / (kernel_ms * 1e-3) / 1e9 / hbm},
                         # the binding resources (ncu: DRAM < 1 %): FP32-pipe operations and issue slots
                         "pipes": issue_roofline(n_valid * T, kernel_ms, clocks, args.dtype)},
            "e2e": {"value": e2e_value, "unit": "cell-timesteps/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": float(t_e.item()) * 1e3, "steps": e2e_steps},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "kernel": eng.kernel_info(),
            "check": {"mean_melt_flux_last_step": float(stats_host[-1, _lib.S_MELT] / stats_host[-1, _lib.S_NVALID])},
        }
        if args.cpu_baseline and world == 1:
            result["cpu_baseline"] = cpu_baseline(n, args.cpu_sample_steps, cores=1)
    eng.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        emit(result)


def issue_roofline(cell_steps, kernel_ms, clocks, dtype):
    """The binding resources of the fused kernel (ncu: DRAM < 1 %): FP32-pipe lane-operations per
    second against 128 lanes x SMs x clock, and instruction issue slots against 4 schedulers x 32
    lanes x SMs x clock, both at the SM clock sampled during the run.  float32 figures (the packed
    path); float64 runs the same algorithm on the FP64 pipe (64 lanes per SM)."""
    mhz = clocks.get("sm_mhz") or clocks.get("sm_max_mhz") or 1965.0
    lanes = 128 if dtype == "f32" else 64
    peak = 148 * lanes * mhz * 1e6 / 1e12
    pipe = FP32_PIPE_OPS_PER_CELL_STEP * cell_steps / (kernel_ms * 1e-3) / 1e12
    slots = ISSUE_SLOTS_PER_CELL_STEP * cell_steps / (kernel_ms * 1e-3) / 1e12
    issue_peak = 148 * 128 * mhz * 1e6 / 1e12
    return {"fp_pipe": {"achieved": pipe, "peak": peak, "unit": "Tera lane-ops/s", "frac": pipe / peak,
                        "ops_per_cell_step": FP32_PIPE_OPS_PER_CELL_STEP},
            "issue_slots": {"achieved": slots, "peak": issue_peak, "unit": "Tera thread-instr/s", "frac": slots / issue_peak,
                            "slots_per_cell_step": ISSUE_SLOTS_PER_CELL_STEP,
                            "note": "float32: two cells per packed instruction (FFMA2/FADD2/FMUL2)"},
            "sm_mhz": mhz}


def measured_traffic(dtype, shadow):
    """dram__bytes_read.sum + dram__bytes_write.sum of the fused kernel for this workload, from the
    committed ncu --set full capture (profiles/r01_traffic.json), or None."""
    path = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if not os.path.isfile(path):
        return None
    key = ("shadow_" if shadow else "") + dtype
    return json.load(open(path)).get(key)


def algorithmic_bytes(case, precision):
    """HBM bytes one pass has to move: DEM + 3 normals + 5 albedo maps + 3 state rasters in and 3 out."""
    cells = case.dem.size
    r = 4 if precision == 32 else 8
    return cells * (4 + 3 * r + 5 * 4 + 3 * r + 3 * r)


def make_workload(n, T, world, rank):
    from enrgy_b200.synthetic import make_band_case
    return make_band_case(n, T, world=world, rank=rank)


# ---------------------------------------------------------------------------------------------
def _oracle_band(job):
    """One process: the NumPy oracle on a row band of the sample (cells are independent given the
    AWS-cell scalars; every band carries the AWS cell's row so the point solve is identical)."""
    import numpy as np
    from oracle import enrgy_oracle as O
    dem, gt, rows, pot, cfg_kw, swe, alb = job
    cfg = O.ModelConfig(**cfg_kw)
    t0 = time.perf_counter()
    O.run_model(dem, gt, rows, pot, cfg, swe=swe, albedo_arrays=alb, state_dtype=np.float32)
    return time.perf_counter() - t0


def cpu_sample(n, steps):
    from enrgy_b200.synthetic import make_band_case
    from oracle import insolation_oracle as I
    from oracle.enrgy_oracle import time_step_seconds
    case, dem_full = make_band_case(n, steps, world=1, rank=0)
    normals = I.terrain_normals(case.dem, case.cell)
    pot = np.empty((steps,) + case.dem.shape, dtype=np.float32)
    for i, row in enumerate(case.aws_rows):
        pot[i] = I.potential_insolation(case.dem, case.cell, case.lat, case.lon, I.to_unix(row["DATE"]),
                                        time_step_seconds(case.aws_rows, i), shadow=False, normals=normals)
    return case, pot


def cpu_baseline(n, steps, cores):
    """Times the oracle (kind "port": bit-identical NumPy restatement of reference model.py:155-286)
    on a bounded sample: the C2 raster for `steps` hourly steps, insolation precomputed in memory
    (the reference np.loads it per step, model.py:481)."""
    import multiprocessing as mp
    case, pot = cpu_sample(n, steps)
    alb = {}
    for k, a in case.albedo_maps.items():
        a = a.copy(); a[a < 0] = 0.001; a[a > 1] = 1
        alb[k] = a
    cfg_kw = dict(z=1.6, elev_aws=case.elev_aws, xy_aws=case.xy_aws, zm=1e-3, z_h_or_e=1e-4, emissivity=0.98)
    rows_total = case.dem.shape[0]
    ar = case.aws_rc[0]
    jobs = []
    if cores <= 1:
        jobs.append((case.dem, case.geotransform, case.aws_rows, pot, cfg_kw, case.swe, alb))
    else:
        # row bands; each band is given the AWS row as its first row (duplicated) so that the
        # reference's point sampling (raster_utils.py:85-89) sees the same cell in every process
        edges = np.linspace(0, rows_total, cores + 1).astype(int)
        for b in range(cores):
            sl = slice(edges[b], edges[b + 1])
            def band(a):
                return np.concatenate([a[ar:ar + 1], a[sl]], axis=0)
            gt = list(case.geotransform)
            gt[3] = case.xy_aws[1] + 0.5 * case.cell          # AWS row becomes row 0 of the band
            jobs.append((band(case.dem), tuple(gt), case.aws_rows,
                         np.concatenate([pot[:, ar:ar + 1], pot[:, sl]], axis=1), cfg_kw, band(case.swe),
                         {k: band(a) for k, a in alb.items()}))
    t0 = time.perf_counter()
    if cores <= 1:
        _oracle_band(jobs[0])
    else:
        with mp.get_context("fork").Pool(cores) as pool:
            pool.map(_oracle_band, jobs)
    wall = time.perf_counter() - t0
    cell_steps = float(rows_total) * case.dem.shape[1] * steps
    return {"value": cell_steps / wall, "unit": "cell-timesteps/s", "cores": cores, "kind": "port",
            "sample": "%dx%d raster, first %d hourly steps of the season, float32 as shipped, insolation "
                      "precomputed in memory; %.1f s wall" % (rows_total, case.dem.shape[1], steps, wall)}


def run_reference(args):
    """--impl reference: the CPU path on all host cores (rank 0 only)."""
    claim_stdout()
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n, T = args.n, args.t
    vals = []
    t_all = time.perf_counter()
    last = None
    for _ in range(args.warmup + args.steps):
        last = cpu_baseline(n, args.cpu_sample_steps, cores)
        vals.append(last["value"])
        if time.perf_counter() - t_all > 240:
            break
    timed = vals[args.warmup:] if len(vals) > args.warmup else vals
    value = float(np.mean(timed))
    cb = dict(last)
    cb["value"] = value
    emit({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "cell-timesteps/s",
        "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": len(timed), "warmup": min(args.warmup, len(vals) - len(timed)),
        "ms_per_step": float(n) * n * T / value * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "C2: %dx%d 10 m synthetic DEM + 5 albedo maps, %d hourly steps (bounded sample: first %d steps)"
                               % (n, n, T, args.cpu_sample_steps)},
        "cpu_baseline": cb,
        "e2e": {"value": value, "unit": "cell-timesteps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--dtype", default="f32", choices=["f32", "f64"])
    ap.add_argument("--n", type=int, default=N_CELLS, help="raster edge per GPU")
    ap.add_argument("--t", type=int, default=N_STEPS, help="AWS rows per pass")
    ap.add_argument("--cpu-sample-steps", type=int, default=24)
    ap.add_argument("--no-cpu-baseline", dest="cpu_baseline", action="store_false")
    ap.add_argument("--shadow", action="store_true", help="topographic shading ray march on (config C3)")
    args = ap.parse_args()
    global SHADOW
    SHADOW = bool(args.shadow)
    if args.warmup < 3 and args.impl == "ours":
        log("note: the timing rules ask for >= 3 warm-up passes")
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

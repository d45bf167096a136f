#!/usr/bin/env python3
"""bench.py -- throughput of the fused Qingdai loop step on N B200s (one process per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload hires|full181|default181|config3|ensemble64] [--impl b200|reference]

A "step" is one pass of the per-timestep loop (scripts/run_simulation.py:1760-2344: precipitation / cloud diagnosis,
dual-star forcing, P019 snow, albedo, SpectralModel.time_step, slab ocean, hydrology bucket) over every ensemble member
resident on the GPU.  Prints ONE JSON line (rank 0).

Workloads (BASELINE.json configs; the reference's own harness takes the grid on the command line the same way,
scripts/benchmark_jax.py:43,122-172):
  hires       configs[4], the configuration the north-star roofline target is stated on: 1441x2880 full physics, dt=37 s.
              N=1: one domain on one GPU.  N>1: the SAME domain split into latitude bands with peer-to-peer halo
              exchange (strong scaling); the line then carries "band_parity" = banded run vs un-split run.  DEFAULT.
  full181     configs[1]: 181x360 full physics (topography + orography, energy branch + sea ice, cloud coupling,
              dynamic ocean, hydrology), dt=300 s, one member per GPU.
  default181  configs[0]: the script's default path at 181x360 (built-in 0.29 land mask seed 42, time_step(Teq, dt)
              without the albedo argument, combo filter, no routing / ecology).
  config3     configs[2]: full181 + D8 river routing (6 h events) + sub-daily ecology albedo feedback.
  ensemble64  configs[3]: 64 independent 181x360 members (planet rotated per member + a QD_* parameter sweep) split
              across the GPUs, no communication.
The default run also records short sub-records of the other workloads under "also" (N=1: default181, full181, config3,
ensemble64; N>1: ensemble64 split over the ranks), so one driver invocation measures every BASELINE config.

Inputs: the 181x360 topography written by the reference's scripts/generate_topography.py defaults (seed 42, land 0.40;
tests/golden/topography_qingdai_181x360_seed42.nc, made by tests/golden/make_golden.py topo), read through the
QD_TOPO_NC loader, which regrids it bilinearly for 1441x2880 exactly like the reference's loader does for a coarser file.
"""
from __future__ import annotations

import argparse
import gc
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DAY = 72000.0            # planet day, s (2*pi/Omega, constants.py:33)
TOPO_NC = os.path.join(ROOT, "tests", "golden", "topography_qingdai_181x360_seed42.nc")
MASK0 = os.path.join(ROOT, "tests", "golden", "default_mask_181x360.npz")
BAND_FIELDS = ("u", "v", "h", "ts", "q", "cloud", "hice", "uo", "vo", "eta", "sst", "precip", "albedo", "wland")

# Algorithmic bytes per CELL per launch for each kernel in the default (loop, full-physics)
# configuration: every distinct field the kernel must read or write, counted once, 8 B (1 B masks).
ALG_BYTES = {
    "k_column": (12 + 17) * 8 + 2, "k_energy": (13 + 5) * 8 + 1, "k_advect_momentum": (6 + 4) * 8,
    "k_column_energy": (14 + 19) * 8 + 2,
    "k_laplacian": 5 * 16, "k_hyper_update": 5 * 24, "k_tail": (11 + 7) * 8 + 2, "k_advect": 4 * 8,
    "k_shapiro_lon": 3 * 16, "k_shapiro_lat": 3 * 16, "k_gauss_lat": 2 * 16, "k_gauss_lon": 16,
    "k_precip_a": (5 + 2) * 8, "k_precip_b": 4 * 8, "k_precip_c": 3 * 8, "k_precip_d": 2 * 8,
    "k_cloud_a": (4 + 2) * 8, "k_cloud_b": 4 * 8, "k_cloud_c": 3 * 8, "k_select_coop": 8, "k_select_cluster": 8,
    "k_ocean_prep": 6 * 8, "k_ocean_momentum": 7 * 8 + 1, "k_ocean_lap": 3 * 16, "k_ocean_hyper": 3 * 24,
    "k_ocean_continuity": 6 * 8 + 1, "k_ocean_continuity2": 6 * 8 + 1, "k_ocean_sst_finish": 9 * 8 + 2, "k_ocean_sst_finish2": 11 * 8 + 2,
    "k_ocean_fused": (6 + 4) * 8 + 1, "k_ocean_close": (4 + 2) * 8 + 2,
    "k_gauss2d_tile<plain>": 16, "k_gauss2d_tile<precip>": 3 * 8, "k_gauss2d_tile<cloud_b>": 4 * 8, "k_gauss2d_tile<cloud_c>": 3 * 8,
}


def alg_bytes_per_cell(name):
    """Algorithmic bytes per cell of one launch; the fused del^4 tile kernel reads and writes each of its
    n fields once (its profile name carries n: ``k_hyper4_stream[3]``, ``k_hyper4_tile<8>[3]``)."""
    if name.startswith("k_hyper4_") and name.endswith("]"):
        return 16 * int(name[name.rindex("[") + 1:-1])
    return ALG_BYTES.get(name, 16)


def workload(name):
    from qingdai_b200.params import QDParams
    full = dict(orog_enabled=True, energy_w=1.0, cloud_couple=True)
    if name == "hires":
        return dict(name=name, nlat=1441, nlon=2880, dt=37, members_total=None, members_per_gpu=1, params=QDParams(**full), with_albedo=True,
                    label="configs[4] 1441x2880 full physics, dt=37 s (QD_TOPO_NC topography regridded + orography, energy branch + sea ice, cloud coupling, dynamic ocean, hydrology)")
    if name == "full181":
        return dict(name=name, nlat=181, nlon=360, dt=300, members_total=None, members_per_gpu=1, params=QDParams(**full), with_albedo=True,
                    label="configs[1] 181x360 full physics (QD_TOPO_NC topography + orography, energy branch + sea ice, cloud coupling, dynamic ocean, hydrology)")
    if name == "default181":
        return dict(name=name, nlat=181, nlon=360, dt=300, members_total=None, members_per_gpu=1, params=QDParams(), with_albedo=False, mask0=True,
                    label="configs[0] 181x360 default script path (built-in land mask 0.29 seed 42, time_step(Teq, dt), combo filter, no routing / ecology)")
    if name == "config3":
        return dict(name=name, nlat=181, nlon=360, dt=300, members_total=None, members_per_gpu=1, params=QDParams(**full), with_albedo=True, config3=True,
                    label="configs[2] 181x360 full physics + P014 D8 routing (C++-built network, 6 h events) + P015 sub-daily ecology albedo feedback (NB=16)")
    if name == "ensemble64":
        return dict(name=name, nlat=181, nlon=360, dt=300, members_total=64, members_per_gpu=None, params=QDParams(**full), with_albedo=True,
                    label="configs[3] 64-member 181x360 full-physics ensemble (planet rotated per member + QD_* parameter sweep) split across GPUs")
    raise SystemExit(f"unknown workload {name}")


def member_inputs(spec, m):
    """(topography dict, QDParams) of global ensemble member m -- the same for the B200 arm and the CPU arms."""
    from qingdai_b200.synthetic import load_reference_topography
    nlat, nlon = spec["nlat"], spec["nlon"]
    if spec.get("mask0"):
        d = np.load(MASK0)
        topo = dict(land_mask=d["land_mask"], base_albedo=d["base_albedo"], friction=d["friction"], elevation=None)
    else:
        topo = load_reference_topography(TOPO_NC, nlat, nlon, roll_columns=(m * (nlon - 1)) // 64 if spec["members_total"] else 0)
    p = spec["params"]
    if spec["members_total"]:
        # QD_* sweep (continuous parameters only: members of one batch share the launch structure)
        p = p.replace(gh_newton=0.36 + 0.08 * (m % 8) / 7.0, sw_a0=0.05 + 0.02 * ((m // 8) % 8) / 7.0)
    return topo, p


def config_of(spec, world, band, halo, members, total_members, flush, state_bytes):
    nlat, nlon = spec["nlat"], spec["nlon"]
    par = "single GPU" if world == 1 else (f"latitude bands x{world}, halo {halo} rows over NVLink peer stores" if band else f"independent members x{world}")
    return {"workload": spec["label"], "grid": [nlat, nlon], "dt_s": spec["dt"], "members_per_gpu": members, "members_total": total_members,
            "parallelism": par,
            "l2": (f"256 MiB L2 flush before every timed step (state ~{state_bytes / 1e6:.0f} MB per GPU)" if flush else
                   f"no flush: the state (~{state_bytes / 1e6:.0f} MB per GPU) is larger than the 126 MB L2"),
            "loop_with_albedo": bool(spec["with_albedo"]),
            "inputs": ("built-in land mask (create_land_sea_mask 0.29 seed 42)" if spec.get("mask0") else
                       "reference generate_topography.py 181x360 seed 42 land 0.40 via the QD_TOPO_NC loader" + (" (bilinear regrid)" if nlat != 181 else ""))}


class ClockSampler:
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc, self.marks = gpu_index, [], None, {}

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(self.idx)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def mark(self, name):
        self.marks[name] = time.perf_counter()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.1)
        self.proc.terminate()
        t0, t1 = self.marks.get("t0", 0.0), self.marks.get("t1", float("inf"))

        def parse(rows):
            sm, mx, reasons = [], [], set()
            for _, r in rows:
                try:
                    sm.append(float(r[1])); mx.append(float(r[2]))
                except (ValueError, IndexError):
                    continue
                for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                    if len(r) > col and r[col].lower().startswith("active"):
                        reasons.add(name)
            return sm, mx, reasons
        timed = [x for x in self.rows if t0 <= x[0] <= t1]
        sm, mx, reasons = parse(timed)
        sm_all, mx_all, reasons_all = parse(self.rows)
        return {"sm_mhz": statistics.median(sm) if sm else (statistics.median(sm_all) if sm_all else None),
                "sm_max_mhz": max(mx_all) if mx_all else None, "reasons": sorted(reasons | reasons_all),
                "samples": len(sm), "samples_total": len(sm_all),
                "note": "nvidia-smi every 20 ms from before the warm-up; sm_mhz = median over the timed + end-to-end regions (whole run when those are shorter than one sample)"}


def oracle_run(spec, nsteps, warmup=1):
    """The CPU arm: the oracle port of the reference loop (NumPy, one core) on the same workload and inputs."""
    from oracle import model
    nlat, nlon, dt = spec["nlat"], spec["nlon"], spec["dt"]
    topo, p = member_inputs(spec, 0)
    g = model.make_grid(nlat, nlon)
    st = model.new_atmos_state(g, p, topo["land_mask"], topo["friction"], base_albedo=topo["base_albedo"], elevation=topo["elevation"])
    oc = model.new_ocean_state(g, topo["land_mask"], init_Ts=np.where(topo["land_mask"] == 0, st.T_s, 288.0))
    k = 0
    for _ in range(max(1, warmup)):                                            # first-touch / allocator warm-up
        model.loop_step(st, oc, g, p, t=k * dt, dt=dt, with_albedo_arg=spec["with_albedo"]); k += 1
    t0 = time.perf_counter()
    for _ in range(nsteps):
        model.loop_step(st, oc, g, p, t=k * dt, dt=dt, with_albedo_arg=spec["with_albedo"]); k += 1
    return (time.perf_counter() - t0) / nsteps


def load_peaks():
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        peaks = {}
    if "hbm_gbs" in peaks:
        return float(peaks["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class MultiSim:
    """The ensemble members of this rank as several batched Simulations on their own CUDA streams
    (qingdai_b200.ensemble.EnsembleRunner); the few engine calls the bench makes, fanned out / aggregated."""

    def __init__(self, runner):
        self.r = runner
        self.sims = runner.sims
        self.engine = self

    def step(self, n=1):
        self.r.step(n)

    def sync(self):
        self.r.synchronize()

    def launches(self):
        return sum(s.engine.launches() for s in self.sims)

    def graph_status(self):
        g = [s.engine.graph_status() for s in self.sims]
        return {"live": sum(x["live"] for x in g), "failed": sum(x["failed"] for x in g), "streams": len(self.sims)}

    def diag(self):
        return [d for s in self.sims for d in s.engine.diag()]

    def last_nsub(self):
        return [int(x) for s in self.sims for x in s.engine.last_nsub()]


def run_b200(name, args, rank, world, local, short=False, sampler=None):
    """Build the workload, time it, return the JSON record (rank 0) or None."""
    import torch
    import torch.distributed as dist
    from qingdai_b200.simulation import Simulation

    spec = workload(name)
    if spec["members_total"] and args.members:
        spec["members_total"] = args.members
        spec["label"] = spec["label"].replace("64-member", f"{args.members}-member")
    nlat, nlon, dt = spec["nlat"], spec["nlon"], spec["dt"]
    ncell = nlat * nlon
    dev = f"cuda:{local}"
    steps = max(1, min(args.steps, 40) if short else args.steps)
    warmup = max(args.warmup, 3)
    band = None
    if spec["members_total"]:
        assert spec["members_total"] % world == 0
        members = spec["members_total"] // world
        scaling, ids = "strong", [rank * members + m for m in range(members)]
    elif name == "hires" and world > 1 and not args.replicas:
        members, scaling, ids, band = 1, "strong", [0], (rank, world, args.halo)
        spec["label"] += "; ONE domain in latitude bands over the GPUs"
    else:
        members, scaling, ids = spec["members_per_gpu"], "weak", [0] * spec["members_per_gpu"]
    ins = [member_inputs(spec, m) for m in ids]
    topos, plist = [t for t, _ in ins], [p for _, p in ins]
    extra = {}
    if spec.get("config3"):
        from qingdai_b200.grid import SphericalGrid
        from qingdai_b200.hydrology_network import build_network
        t0 = time.perf_counter()
        net = build_network(SphericalGrid(nlat, nlon), topos[0]["elevation"], topos[0]["land_mask"])
        print(f"[bench] routing network built in {time.perf_counter() - t0:.2f} s (n_lakes={net['n_lakes']}, pit sweeps={net['pit_sweeps']})", file=sys.stderr)
        extra = dict(with_eco=True, eco_env={}, routing_network=net, dt_hydro_hours=6.0)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def make_sim(b):
        if spec["members_total"]:
            from qingdai_b200.ensemble import EnsembleRunner
            by_id = dict(zip(ids, ins))
            return MultiSim(EnsembleRunner(nlat, nlon, spec["members_total"], lambda m: by_id[m][0], lambda m: by_id[m][1], dt=dt, rank=rank,
                                           world=world, device=dev, streams=args.streams or None, with_ocean=True, with_hydrology=True,
                                           loop_with_albedo=spec["with_albedo"]))
        return Simulation(nlat, nlon, topos, plist, dt=dt, batch=members, with_ocean=True, with_hydrology=True,
                          loop_with_albedo=spec["with_albedo"], device=dev, band=b, **extra)

    # -------- latitude bands: the banded run against an un-split run of the same library (every rank holds an un-split
    # copy and checks its OWN rows; max over ranks) -- the correctness evidence a single-GPU test box cannot give
    band_parity = None
    if band:
        ref_sim, bsim = make_sim(None), make_sim(band)
        nchk = 4
        for _ in range(nchk):
            ref_sim.step(1); bsim.step(1)
        r0, r1, _, err = bsim.engine.band_info()
        worst, worst_f = 0.0, None
        for f in BAND_FIELDS:
            a, b_ = ref_sim.engine.get(f), bsim.engine.get(f)
            e = float(np.max(np.abs(a[r0:r1] - b_[r0:r1]))) / max(float(np.max(np.abs(a))), 1e-300)
            if e > worst:
                worst, worst_f = e, f
        worst_all = allmax(worst)
        err_all = allmax(float(err))
        band_parity = {"max_rel": worst_all, "fields": len(BAND_FIELDS), "steps": nchk, "tolerance": 1e-10, "ok": bool(worst_all <= 1e-10 and err_all == 0.0),
                       "exchange_error_word": int(err_all), "worst_field_rank0": worst_f,
                       "how": "every rank steps an un-split copy next to the banded run and compares its own rows (max-norm relative to the field max, max over ranks)"}
        del ref_sim, bsim
        gc.collect(); torch.cuda.empty_cache()

    sim = make_sim(band)
    eng = sim.engine
    state_bytes = members * ncell * 8 * 45
    flush_on = state_bytes < (512 << 20)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if flush_on else None

    for _ in range(warmup):
        sim.step(1)
    barrier()
    gstat = eng.graph_status()
    # -------- device-timed region: K steps.  Small states: CUDA events around every step with an L2 flush before each;
    # states larger than L2: one event pair around the K back-to-back steps
    l0 = eng.launches()
    if sampler:
        sampler.mark("t0")
    barrier()
    multi = isinstance(sim, MultiSim) and len(sim.sims) > 1
    cur = torch.cuda.current_stream()

    def fork():          # the member groups' streams start after everything enqueued on the timing stream ...
        if multi:
            for st in sim.r._streams:
                st.wait_stream(cur)

    def join():          # ... and the closing event waits for all of them
        if multi:
            for st in sim.r._streams:
                cur.wait_stream(st)
    if flush_on:
        evs = []
        for _ in range(steps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fork(); sim.step(1); join(); e1.record()
            evs.append((e0, e1))
        barrier()
        dev_ms = sum(a.elapsed_time(b) for a, b in evs)
    else:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fork()
        for _ in range(steps):
            sim.step(1)
        join(); e1.record()
        barrier()
        dev_ms = e0.elapsed_time(e1)
    launches = eng.launches() - l0
    dev_ms = allmax(dev_ms)
    ms_per_step = dev_ms / steps
    total_members = 1 if band else members * world
    steps_per_s = 1e3 / ms_per_step
    value = total_members * steps_per_s * dt / DAY

    # -------- end-to-end through the public API: per step H2D forcing + D2H of the step's metrics, host sync every step
    barrier()
    w0 = time.perf_counter()
    metric_bytes = 0
    for _ in range(steps):
        sim.step(1)
        if band:
            s = eng.scalars()                   # latitude bands: per-rank step scalars (n_sub, global sums); sync D2H
            metric_bytes = s.size * 8
        else:
            d = eng.diag()                      # the step's metrics: one reduction launch + D2H of the global means
            metric_bytes = len(d) * len(d[0]) * 8
    barrier()
    e2e_s = allmax(time.perf_counter() - w0)
    if sampler:
        sampler.mark("t1")
    e2e_val = total_members * (steps / e2e_s) * dt / DAY
    nsub = [int(x) for x in eng.last_nsub()]
    prof_eng = sim.sims[0].engine if isinstance(sim, MultiSim) else eng          # per-kernel times: the first member group
    # -------- the reference's object-level operator interface with HOST arrays every step (rank 0, one member):
    # SpectralModel.time_step(Teq, dt, albedo) + WindDrivenSlabOcean.step(dt, u, v, Q_net, ice_mask) + reads of T_s
    dropin = None
    if rank == 0 and world == 1 and members == 1 and not spec.get("config3") and not (short and ncell > 200000):
        from qingdai_b200.dynamics import SpectralModel
        from qingdai_b200.grid import SphericalGrid
        from qingdai_b200.ocean import WindDrivenSlabOcean
        grid = SphericalGrid(nlat, nlon)
        tp = topos[0]
        gcm = SpectralModel(grid, tp["friction"], land_mask=tp["land_mask"], greenhouse_factor=0.40, tau_rad=864000.0)
        oc = WindDrivenSlabOcean(grid, tp["land_mask"], 50.0)
        rng = np.random.default_rng(0)
        Teq = 250.0 + 40.0 * np.cos(np.deg2rad(grid.lat_mesh)) + rng.standard_normal((nlat, nlon))
        alb = np.clip(0.3 + 0.05 * rng.standard_normal((nlat, nlon)), 0.0, 1.0)
        qn = 50.0 * rng.standard_normal((nlat, nlon))
        nd = max(5, min(steps, 50 if ncell <= 200000 else 10))
        for k in range(3 + nd):
            if k == 3:
                torch.cuda.synchronize(); wd = time.perf_counter()
            gcm.time_step(Teq, dt, albedo=alb)
            u_h, v_h = gcm.u, gcm.v
            oc.step(dt, u_h, v_h, Q_net=qn, ice_mask=gcm.h_ice > 0.0)
            gcm.T_s = np.where(tp["land_mask"] == 0, oc.Ts, gcm.T_s)          # run_simulation.py:2252-2253
        torch.cuda.synchronize()
        sec = (time.perf_counter() - wd) / nd
        fb = ncell * 8
        dropin = {"value": (1.0 / sec) * dt / DAY, "unit": "planet-days/s", "ms_per_step": sec * 1e3,
                  "h2d_bytes_per_step": 6 * fb + ncell, "d2h_bytes_per_step": 5 * fb,
                  "note": "drop-in SpectralModel.time_step + WindDrivenSlabOcean.step with host NumPy arrays in and out every step (cores only, no loop physics)"}
        del gcm, oc, grid

    # -------- per-kernel device time (CUDA events around every launch, stream mode) -> roofline of the dominant kernel
    roof = None
    buf = None
    if not args.no_profile and (rank == 0 or band):       # band mode: every rank must run the same (stream-mode) steps
        import ctypes
        prof_eng.lib.qd_profile(prof_eng.ctx, 1)
        nprof = min(steps, 20)
        psim = sim.sims[0] if isinstance(sim, MultiSim) else sim
        for _ in range(nprof):
            psim.step(1)
        buf = ctypes.create_string_buffer(1 << 16)
        prof_eng.lib.qd_profile_report(prof_eng.ctx, buf, len(buf))
        prof_eng.lib.qd_profile(prof_eng.ctx, 0)
    if rank == 0 and buf is not None:
        rows = [ln.split() for ln in buf.value.decode().strip().splitlines()]
        rows = [(r[0], int(r[1]), float(r[2])) for r in rows if len(r) == 3]
        tot = sum(r[2] for r in rows) or 1.0
        peak, peak_src = load_peaks()
        kname, cnt, ms = rows[0]
        per_launch_s = ms / cnt * 1e-3
        pmembers = prof_eng.batch if isinstance(sim, MultiSim) else members
        alg = alg_bytes_per_cell(kname) * ncell * pmembers
        achieved = alg / per_launch_s / 1e9
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(name, {}).get(kname)
        except (OSError, ValueError):
            pass
        step_alg = 233 * ncell * members
        roof = {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "frac_of_nominal_8TBps": achieved / 8000.0, "traffic": traffic, "peak_source": peak_src, "alg_bytes_per_launch": alg,
                "us_per_launch": per_launch_s * 1e6, "share_of_step": ms / tot,
                "how": "CUDA events around every launch on the launching stream (qd_profile, stream mode) over %d steps after the timed region" % nprof,
                "whole_step": {"alg_bytes_per_step": step_alg, "achieved": step_alg / (ms_per_step * 1e-3) / 1e9,
                               "frac": step_alg / (ms_per_step * 1e-3) / 1e9 / peak, "note": "233 B/cell-step (SURVEY 8d) over the whole fused step (graph mode, the timed region)"},
                "top_kernels": [{"kernel": r[0], "launches_per_step": r[1] / nprof, "us_per_launch": r[2] / r[1] * 1e3, "share": r[2] / tot,
                                 "frac": alg_bytes_per_cell(r[0]) * ncell * pmembers / (r[2] / r[1] * 1e-3) / 1e9 / peak}
                                for r in rows[:(40 if (band or args.all_kernels) else 10)]]}

    # -------- CPU baseline (oracle port, rank 0, N=1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and not spec.get("config3") and not spec["members_total"]:
        per_cell_cpu = 0.3 / 65160          # s per cell-step of the oracle, rough (for sizing the CPU sample)
        budget = 4.0 if short else 15.0
        n = args.cpu_steps or int(max(2 if ncell > 200000 else 1, min(60, budget / max(per_cell_cpu * ncell, 1e-9))))
        sec = oracle_run(spec, n)
        cpu = {"value": (1.0 / sec) * dt / DAY, "unit": "planet-days/s", "cores": 1, "kind": "port",
               "sample": f"{n} loop steps (after 1 warm-up step) of one {nlat}x{nlon} member, NumPy oracle port of the reference loop (single-threaded like the reference's NumPy path)",
               "ms_per_step": sec * 1e3, "host_cores_available": os.cpu_count(),
               "jax_cpu": "unavailable (no jax on the box; the reference's QD_USE_JAX=1 path, pygcm/jax_compat.py:28,52, cannot run)"}

    rec = None
    if rank == 0:
        rec = {"metric": "simulated planet-days per wall-second", "value": value, "unit": "planet-days/s", "n_gpus": world,
               "steps": steps, "warmup": warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
               "scaling": scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
               "cell_steps_per_s": total_members * ncell * steps_per_s,
               "config": config_of(spec, world, band, args.halo, members, total_members, flush_on, state_bytes),
               "e2e": {"value": e2e_val, "unit": "planet-days/s", "h2d_bytes_per_step": 80, "d2h_bytes_per_step": metric_bytes,
                       "note": "Simulation.step(1) per step through the C ABI: forcing scalars H2D (the loop has no other per-step host input), then the step's global diagnostics (one reduction launch) D2H; host sync every step",
                       "host_array_dropin": dropin},
               "gpu_launches": launches, "ocean_substeps_last_step": nsub[:4], "graphs": gstat, "roofline": roof, "cpu_baseline": cpu}
        if band_parity is not None:
            rec["band_parity"] = band_parity
    del sim, eng, flush
    gc.collect(); torch.cuda.empty_cache()
    return rec


def slim(rec):
    """Sub-record under "also": the main numbers of a secondary workload."""
    if rec is None:
        return None
    r = rec.get("roofline") or {}
    out = {k: rec[k] for k in ("value", "unit", "ms_per_step", "steps", "warmup", "scaling", "cell_steps_per_s", "gpu_launches", "ocean_substeps_last_step", "graphs")}
    out["config"] = {k: rec["config"][k] for k in ("workload", "grid", "dt_s", "members_per_gpu", "members_total", "parallelism", "l2")}
    out["e2e"] = {k: rec["e2e"][k] for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step")}
    if rec["e2e"].get("host_array_dropin"):
        out["e2e"]["host_array_dropin"] = {k: rec["e2e"]["host_array_dropin"][k] for k in ("value", "ms_per_step")}
    if r:
        out["roofline"] = {"kernel": r["kernel"], "frac": r["frac"], "achieved": r["achieved"], "us_per_launch": r["us_per_launch"], "share_of_step": r["share_of_step"],
                           "whole_step_frac": r["whole_step"]["frac"],
                           "top_kernels": [{k: (round(v, 4) if isinstance(v, float) else v) for k, v in t.items()} for t in r["top_kernels"][:8]]}
    if rec.get("cpu_baseline"):
        out["cpu_baseline"] = {k: rec["cpu_baseline"][k] for k in ("value", "unit", "cores", "kind", "ms_per_step")}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="hires")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-steps", type=int, default=0, help="oracle steps for cpu_baseline (0 = auto, about 15 s)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-profile", action="store_true")
    ap.add_argument("--no-also", action="store_true", help="skip the short sub-records of the other workloads")
    ap.add_argument("--all-kernels", action="store_true", help="list every kernel in roofline.top_kernels")
    ap.add_argument("--halo", type=int, default=0, help="latitude bands: halo rows per exchange (0 = 8 rows on 2 GPUs, 16 on more: measured, profiles/README.md)")
    ap.add_argument("--streams", type=int, default=0, help="ensemble workloads: member groups (CUDA streams) per GPU, 0 = automatic")
    ap.add_argument("--members", type=int, default=0, help="ensemble workloads: total members instead of 64 (tuning runs)")
    ap.add_argument("--replicas", action="store_true", help="hires at N>1: independent replicas instead of latitude bands")
    ap.add_argument("--no-all-cores", action="store_true", help="reference arm: skip the all-cores (independent copies) figure")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.halo <= 0:
        # fewer halo rows = less redundant compute next to a cut but more exchanges: with two ranks (720 own rows each at
        # 1441x2880) 8 rows win (1.358 vs 1.384 ms/step); with 4 / 8 ranks the per-rank compute is too small to pay for
        # the extra exchanges
        args.halo = 8 if max(world, args.gpus) == 2 else 16

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return
        spec = workload(args.workload)
        nlat, nlon, dt = spec["nlat"], spec["nlon"], spec["dt"]
        ncell = nlat * nlon
        per_cell_cpu = 0.3 / 65160
        est = max(per_cell_cpu * ncell, 1e-9)
        # --steps / --warmup are honoured as given as long as the whole run stays within ~4 minutes on one host core
        steps, warm = max(1, args.steps), max(1, args.warmup)
        budget = 240.0
        if (steps + warm) * est > budget:
            warm = max(1, min(warm, int(0.2 * budget / est)))
            steps = int(max(1, min(steps, (budget - warm * est) / est)))
        sec = oracle_run(spec, steps, warmup=warm)
        val = (1.0 / sec) * dt / DAY
        band = args.workload == "hires" and args.gpus > 1 and not args.replicas
        members = (spec["members_total"] // args.gpus) if spec["members_total"] else 1
        state_bytes = members * ncell * 8 * 45
        if band:
            spec["label"] += "; ONE domain in latitude bands over the GPUs"
        line = {"impl": "reference", "metric": "simulated planet-days per wall-second", "value": val, "unit": "planet-days/s",
                "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": sec * 1e3, "higher_is_better": True,
                "scaling": "strong" if (band or spec["members_total"]) else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "cell_steps_per_s": ncell / sec,
                "config": config_of(spec, args.gpus, band, args.halo, members, 1 if band else members * args.gpus, state_bytes < (512 << 20), state_bytes),
                "cpu_baseline": {"value": val, "unit": "planet-days/s", "cores": 1, "kind": "port",
                                 "sample": f"{steps} loop steps (after {warm} warm-up) of ONE {nlat}x{nlon} domain with the NumPy oracle port of the reference loop (single-threaded like the reference's NumPy path; /root/reference is Python and cannot travel)",
                                 "host_cores_available": os.cpu_count(),
                                 "jax_cpu": "unavailable (no jax on the box; the reference's QD_USE_JAX=1 path, pygcm/jax_compat.py:28,52, cannot run)"},
                "e2e": {"value": val, "unit": "planet-days/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        if not args.no_all_cores and ncell <= 200000:
            # One domain cannot use more than one core in the reference (NumPy, no threaded kernels on this path).  What
            # ALL host cores can do is run independent copies (the ensemble use case): P concurrent processes of
            # the same sample, aggregate throughput reported next to -- not instead of -- the single-domain value.
            P = max(1, min(os.cpu_count() or 1, 16))
            k = max(2, min(steps, 20))
            cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--workload", args.workload, "--steps", str(k), "--warmup", "1", "--no-all-cores"]
            env = dict(os.environ, RANK="0", WORLD_SIZE="1", OMP_NUM_THREADS="1", OPENBLAS_NUM_THREADS="1", MKL_NUM_THREADS="1")
            try:
                procs = [subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, env=env, text=True) for _ in range(P)]
                vals = [json.loads(pr.communicate(timeout=240)[0].strip().splitlines()[-1])["value"] for pr in procs]
                line["cpu_baseline"]["all_cores"] = {"processes": P, "value": float(sum(vals)), "unit": "planet-days/s",
                                                     "per_process": float(sum(vals) / P),
                                                     "note": f"{P} independent copies of the sample running concurrently ({k} steps each): aggregate over copies, an ensemble figure"}
            except Exception as exc:          # noqa: BLE001  (the extra figure must never break the reference line)
                line["cpu_baseline"]["all_cores"] = {"unavailable": str(exc)[:120]}
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------ B200 arm
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()                         # before the warm-up: the timed region can be shorter than one sample
    line = run_b200(args.workload, args, rank, world, local, short=False, sampler=sampler)
    clocks = sampler.stop() if sampler else None
    if args.workload == "hires" and not args.no_also:
        also = {}
        names = ("default181", "full181", "config3", "ensemble64") if world == 1 else ("ensemble64",)
        for nm in names:
            try:
                also[nm] = slim(run_b200(nm, args, rank, world, local, short=True))
            except Exception as exc:          # noqa: BLE001  (a secondary record must never cost the headline line)
                if world > 1:
                    raise
                also[nm] = {"error": f"{type(exc).__name__}: {exc}"[:300]}
        if line is not None:
            line["also"] = also
    if rank == 0:
        line["clocks"] = clocks
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

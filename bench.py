#!/usr/bin/env python3
"""bench.py -- throughput of the fused Qingdai loop step on N B200s (one process per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload full181|config3|ensemble64|hires] [--impl b200|reference]

A "step" is one pass of the per-timestep loop (scripts/run_simulation.py:1760-2344: precipitation /
cloud diagnosis, dual-star forcing, P019 snow, albedo, SpectralModel.time_step, slab ocean,
hydrology bucket) over every ensemble member resident on the GPU.  Prints ONE JSON line (rank 0).

Workloads (BASELINE.json configs):
  full181     configs[1]: 181x360 full physics (topography + orography, energy branch + sea ice,
              cloud coupling, dynamic ocean, hydrology), dt=300 s, one member per GPU.  DEFAULT.
  config3     configs[2]: full181 + D8 river routing (6 h events) + sub-daily ecology albedo feedback.
  ensemble64  configs[3]: 64 independent 181x360 members (topography seeds 42..105) split across GPUs.
  hires       configs[4] at one GPU: 1441x2880 full physics, dt=37 s.
Multi-GPU: ensemble members are independent -> no data-path collective; `--workload hires` at N>1 splits ONE
domain into latitude bands with peer-to-peer halo exchange.
"""
from __future__ import annotations

import argparse
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

# Algorithmic bytes per CELL per launch for each kernel in the default (loop, full-physics)
# configuration: every distinct field the kernel must read or write, counted once, 8 B (1 B masks).
ALG_BYTES = {
    "k_column": (12 + 17) * 8 + 2, "k_energy": (13 + 5) * 8 + 1, "k_advect_momentum": (6 + 4) * 8,
    "k_laplacian": 5 * 16, "k_hyper_update": 5 * 24, "k_tail": (11 + 7) * 8 + 2, "k_advect": 4 * 8,
    "k_shapiro_lon": 3 * 16, "k_shapiro_lat": 3 * 16, "k_gauss_lat": 2 * 16, "k_gauss_lon": 16,
    "k_precip_a": (5 + 2) * 8, "k_precip_b": 4 * 8, "k_precip_c": 3 * 8, "k_precip_d": 2 * 8,
    "k_cloud_a": (4 + 2) * 8, "k_cloud_b": 4 * 8, "k_cloud_c": 3 * 8, "k_select_coop": 8, "k_select_cluster": 8,
    "k_ocean_prep": 6 * 8, "k_ocean_momentum": 7 * 8 + 1, "k_ocean_lap": 3 * 16, "k_ocean_hyper": 3 * 24,
    "k_ocean_continuity": 6 * 8 + 1, "k_ocean_sst_finish": 9 * 8 + 2,
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
    if name == "full181":
        return dict(nlat=181, nlon=360, dt=300, members_total=None, members_per_gpu=1, params=QDParams(**full),
                    label="configs[1] 181x360 full physics (topography+orography, energy branch+sea ice, cloud coupling, dynamic ocean, hydrology)")
    if name == "config3":
        return dict(nlat=181, nlon=360, dt=300, members_total=None, members_per_gpu=1, params=QDParams(**full), config3=True,
                    label="configs[2] 181x360 full physics + P014 D8 routing (C++-built network, 6 h events) + P015 sub-daily ecology albedo feedback (NB=16)")
    if name == "ensemble64":
        return dict(nlat=181, nlon=360, dt=300, members_total=64, members_per_gpu=None, params=QDParams(**full),
                    label="configs[3] 64-member 181x360 full-physics ensemble (topography seeds 42..105) split across GPUs")
    if name == "hires":
        return dict(nlat=1441, nlon=2880, dt=37, members_total=None, members_per_gpu=1, params=QDParams(**full),
                    label="configs[4] 1441x2880 full physics, dt=37 s, one domain per GPU (replicas)")
    raise SystemExit(f"unknown workload {name}")


class ClockSampler:
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(self.idx)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def oracle_run(spec, nsteps, seed=42):
    """The CPU arm: the oracle port of the reference loop (NumPy, one core) on the same workload."""
    from oracle import model
    from qingdai_b200.synthetic import make_topography
    nlat, nlon, dt, p = spec["nlat"], spec["nlon"], spec["dt"], spec["params"]
    topo = make_topography(nlat, nlon, seed=seed, land_frac=0.40)
    g = model.make_grid(nlat, nlon)
    st = model.new_atmos_state(g, p, topo["land_mask"], topo["friction"], base_albedo=topo["base_albedo"], elevation=topo["elevation"])
    oc = model.new_ocean_state(g, topo["land_mask"], init_Ts=np.where(topo["land_mask"] == 0, st.T_s, 288.0))
    model.loop_step(st, oc, g, p, t=0.0, dt=dt, with_albedo_arg=True)          # warm-up (first-touch)
    t0 = time.perf_counter()
    for i in range(nsteps):
        model.loop_step(st, oc, g, p, t=(i + 1) * dt, dt=dt, with_albedo_arg=True)
    return (time.perf_counter() - t0) / nsteps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--workload", default="full181")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-steps", type=int, default=0, help="oracle steps for cpu_baseline (0 = auto, about 15 s)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-profile", action="store_true")
    ap.add_argument("--all-kernels", action="store_true", help="list every kernel in roofline.top_kernels")
    ap.add_argument("--halo", type=int, default=16, help="latitude bands: halo rows per exchange")
    ap.add_argument("--replicas", action="store_true", help="hires at N>1: independent replicas instead of latitude bands")
    ap.add_argument("--no-all-cores", action="store_true", help="reference arm: skip the all-cores (independent copies) figure")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    spec = workload(args.workload)
    nlat, nlon, dt = spec["nlat"], spec["nlon"], spec["dt"]
    ncell = nlat * nlon
    per_cell_cpu = 0.3 / 65160          # s per cell-step of the oracle, rough (for sizing the CPU sample)

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return
        steps = max(1, args.steps)
        budget = 150.0
        steps = int(max(1, min(steps, budget / max(per_cell_cpu * ncell, 1e-9))))
        sec = oracle_run(spec, steps)
        val = (1.0 / sec) * dt / DAY
        line = {"impl": "reference", "metric": "simulated planet-days per wall-second", "value": val, "unit": "planet-days/s",
                "n_gpus": args.gpus, "steps": steps, "warmup": 1, "ms_per_step": sec * 1e3, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "cell_steps_per_s": ncell / sec,
                "config": {"workload": spec["label"], "grid": [nlat, nlon], "dt_s": dt, "members": 1},
                "cpu_baseline": {"value": val, "unit": "planet-days/s", "cores": 1, "kind": "port",
                                 "sample": f"{steps} loop steps of one {nlat}x{nlon} member with the NumPy oracle port of the reference (single-threaded like the reference's NumPy path; /root/reference is Python and cannot travel)",
                                 "host_cores_available": os.cpu_count()},
                "e2e": {"value": val, "unit": "planet-days/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        if not args.no_all_cores and ncell <= 200000:
            # One domain cannot use more than one core in the reference (NumPy, no threaded kernels on this path).  What
            # ALL host cores can do is run independent copies (the ensemble use case): P concurrent processes of
            # the same sample, aggregate throughput reported next to -- not instead of -- the single-domain value.
            import subprocess
            P = max(1, min(os.cpu_count() or 1, 16))
            k = max(2, min(steps, 20))
            cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--workload", args.workload, "--steps", str(k), "--no-all-cores"]
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
    from qingdai_b200.simulation import Simulation
    from qingdai_b200.synthetic import make_topography
    from qingdai_b200.engine import S as SC

    band = None
    if spec["members_total"]:
        assert spec["members_total"] % world == 0
        members = spec["members_total"] // world
        scaling, seeds = "strong", [42 + rank * members + m for m in range(members)]
    elif args.workload == "hires" and world > 1 and not args.replicas:
        # configs[4]: ONE 1441x2880 domain split into latitude bands over the GPUs (halo rows over NVLink)
        members, scaling, seeds, band = 1, "strong", [42], (rank, world, args.halo)
        spec["label"] = "configs[4] 1441x2880 full physics, dt=37 s, ONE domain in latitude bands over the GPUs"
    else:
        members = spec["members_per_gpu"]
        scaling, seeds = "weak", [42 + rank * members + m for m in range(members)]
    topos = [make_topography(nlat, nlon, seed=s, land_frac=0.40) for s in seeds]
    extra = {}
    if spec.get("config3"):
        from qingdai_b200.grid import SphericalGrid
        from qingdai_b200.hydrology_network import build_network
        t0 = time.perf_counter()
        net = build_network(SphericalGrid(nlat, nlon), topos[0]["elevation"], topos[0]["land_mask"])
        print(f"[bench] routing network built in {time.perf_counter() - t0:.2f} s (n_lakes={net['n_lakes']}, pit sweeps={net['pit_sweeps']})", file=sys.stderr)
        extra = dict(with_eco=True, eco_env={}, routing_network=net, dt_hydro_hours=6.0)
    sim = Simulation(nlat, nlon, topos, spec["params"], dt=dt, batch=members, with_ocean=True, with_hydrology=True,
                     loop_with_albedo=True, device=f"cuda:{local}", band=band, **extra)
    eng = sim.engine
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{local}")
    state_bytes = members * ncell * 8 * 45

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        sim.step(1)
    barrier()
    # -------- device-timed region: K steps, CUDA events around every step, L2 flushed before each
    sampler = ClockSampler(local)
    sampler.start()
    l0 = eng.launches()
    evs = []
    barrier()
    for _ in range(args.steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        sim.step(1)
        e1.record()
        evs.append((e0, e1))
    barrier()
    launches = eng.launches() - l0
    dev_ms = sum(a.elapsed_time(b) for a, b in evs)
    t = torch.tensor([dev_ms], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms = float(t.item())
    ms_per_step = dev_ms / args.steps
    total_members = 1 if band else members * world
    steps_per_s = 1e3 / ms_per_step
    value = total_members * steps_per_s * dt / DAY

    # -------- end-to-end through the public API: per step H2D forcing + D2H of a step metric (mean Ts)
    barrier()
    w0 = time.perf_counter()
    metric_bytes = 0
    for _ in range(args.steps):
        sim.step(1)
        if band:
            eng.sync()
            _ = eng.scalars()                   # latitude bands: per-rank step scalars (n_sub, global sums)
            metric_bytes = eng.scalars().size * 8
        else:
            d = eng.diag()                      # the step's metrics: one reduction launch + D2H of the global means
            metric_bytes = len(d) * len(d[0]) * 8
    barrier()
    e2e_s = time.perf_counter() - w0
    clocks = sampler.stop()                     # sampled over the device-timed and the end-to-end region
    t = torch.tensor([e2e_s], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    e2e_val = total_members * (args.steps / e2e_s) * dt / DAY
    # -------- the reference's object-level operator interface with HOST arrays every step (rank 0, small grids):
    # SpectralModel.time_step(Teq, dt, albedo) + WindDrivenSlabOcean.step(dt, u, v, Q_net, ice_mask) + reads of T_s
    dropin = None
    if rank == 0 and world == 1 and members == 1 and ncell <= 200000:
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
        nd = max(10, min(args.steps, 50))
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

    # -------- per-kernel device time (CUDA events around every launch) -> roofline of the dominant kernel
    roof = None
    if not args.no_profile and (rank == 0 or band):       # band mode: every rank must run the same (stream-mode) steps
        import ctypes
        eng.lib.qd_profile(eng.ctx, 1)
        nprof = min(args.steps, 20)
        for _ in range(nprof):
            sim.step(1)
        buf = ctypes.create_string_buffer(1 << 16)
        eng.lib.qd_profile_report(eng.ctx, buf, len(buf))
        eng.lib.qd_profile(eng.ctx, 0)
    if roof is None and rank == 0 and not args.no_profile:
        rows = [ln.split() for ln in buf.value.decode().strip().splitlines()]
        rows = [(r[0], int(r[1]), float(r[2])) for r in rows if len(r) == 3]
        tot = sum(r[2] for r in rows) or 1.0
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except (OSError, ValueError):
            pass
        peak, peak_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json)") if "hbm_gbs" in peaks else (6650.0, "fallback (B200_PROFILING.md)")
        name, cnt, ms = rows[0]
        per_launch_s = ms / cnt * 1e-3
        alg = alg_bytes_per_cell(name) * ncell * members
        achieved = alg / per_launch_s / 1e9
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(args.workload, {}).get(name)
        except (OSError, ValueError):
            pass
        step_alg = 233 * ncell * members
        roof = {"bound": "hbm", "kernel": name, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "frac_of_nominal_8TBps": achieved / 8000.0, "traffic": traffic, "peak_source": peak_src, "alg_bytes_per_launch": alg, "us_per_launch": per_launch_s * 1e6,
                "share_of_step": ms / tot,
                "whole_step": {"alg_bytes_per_step": step_alg, "achieved": step_alg / (ms_per_step * 1e-3) / 1e9,
                               "frac": step_alg / (ms_per_step * 1e-3) / 1e9 / peak, "note": "233 B/cell-step (SURVEY 8d) over the whole fused step"},
                "top_kernels": [{"kernel": r[0], "launches_per_step": r[1] / nprof, "us_per_launch": r[2] / r[1] * 1e3, "share": r[2] / tot} for r in rows[:(40 if (band or args.all_kernels) else 8)]]}

    # -------- CPU baseline (oracle port, rank 0, N=1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        n = args.cpu_steps or int(max(1, min(60, 15.0 / max(per_cell_cpu * ncell, 1e-9))))
        sec = oracle_run(spec, n)
        cpu = {"value": (1.0 / sec) * dt / DAY, "unit": "planet-days/s", "cores": 1, "kind": "port",
               "sample": f"{n} loop steps of one {nlat}x{nlon} member, NumPy oracle port (single-threaded like the reference)",
               "ms_per_step": sec * 1e3, "host_cores_available": os.cpu_count()}

    if rank == 0:
        line = {"metric": "simulated planet-days per wall-second", "value": value, "unit": "planet-days/s", "n_gpus": world,
                "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
                "scaling": scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "cell_steps_per_s": total_members * ncell * steps_per_s,
                "config": {"workload": spec["label"], "grid": [nlat, nlon], "dt_s": dt, "members_per_gpu": members,
                           "members_total": total_members, "parallelism": (f"latitude bands x{world}, halo {args.halo} rows over NVLink peer stores" if band else f"independent members x{world}") if world > 1 else "single GPU",
                           "l2": f"256 MiB L2 flush before every timed step (state ~{state_bytes / 1e6:.0f} MB per GPU)",
                           "loop_with_albedo": True},
                "clocks": clocks,
                "e2e": {"value": e2e_val, "unit": "planet-days/s", "h2d_bytes_per_step": 80, "d2h_bytes_per_step": metric_bytes,
                        "note": "Simulation.step(1) per step through the C ABI: forcing scalars H2D (the loop has no other per-step host input), then the step's global diagnostics (one reduction launch) D2H; host sync every step",
                        "host_array_dropin": dropin},
                "gpu_launches": launches, "roofline": roof, "cpu_baseline": cpu}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

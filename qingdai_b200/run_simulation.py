"""Environment-driven entry point of the fused loop: the subset of ``scripts/run_simulation.py:main()`` that is the hot
path (no plots, no daily ecology), configured by the same ``QD_*`` variables.

    python -m qingdai_b200.run_simulation

    QD_N_LAT / QD_N_LON     grid (the reference hard-codes 121x240 at run_simulation.py:1195)   default 121 x 240
    QD_DT_SECONDS           time step (run_simulation.py:1593)                                              default 300
    QD_TOTAL_YEARS / QD_SIM_DAYS   duration, same priority as run_simulation.py:1595-1601                   default 5 planet years
    QD_TOPO_NC              topography file (run_simulation.py:1198-1203); otherwise a mask file given by QD_MASK_NPZ is
                            required -- the procedural generator is init-time host code of the reference and stays there
    QD_HYDRO_NETCDF, QD_HYDRO_ENABLE   offline routing network (run_simulation.py:1297-1311)
    QD_ECO_ENABLE           sub-daily ecology albedo feedback (run_simulation.py:1335-1337)
    QD_INIT_BANDED, QD_INIT_T_EQ, QD_INIT_T_POLE   banded initial surface temperature (run_simulation.py:310-328)
    QD_RESTART_IN / QD_RESTART_OUT   warm restart files in the reference's format (run_simulation.py:1435-1470, 2493-2506)
    QD_LOOP_WITH_ALBEDO     1: pass the albedo into SpectralModel.time_step (energy branch live, BASELINE configs[1-4])
    every physics parameter: qingdai_b200.params.QDParams.from_env

Prints the periodic global diagnostics the reference prints (area-weighted means from one device reduction) and returns
the Simulation.  For the full script -- plots, daily ecology, autosave -- run the reference's ``main()`` with the four
class imports swapped; tests/test_reference_main_dropin.py does exactly that.
"""
from __future__ import annotations

import os
import sys

import numpy as np

from .forcing import OrbitalSystem
from .params import QDParams
from .simulation import DAY_SECONDS, Simulation


def _topography(nlat, nlon, env):
    path = env.get("QD_TOPO_NC")
    if path and os.path.exists(path):
        from .synthetic import load_reference_topography
        return load_reference_topography(path, nlat, nlon)
    mask_npz = env.get("QD_MASK_NPZ")
    if mask_npz and os.path.exists(mask_npz):
        d = np.load(mask_npz)
        if d["land_mask"].shape != (nlat, nlon):
            raise ValueError(f"QD_MASK_NPZ holds a {d['land_mask'].shape} mask, the grid is {(nlat, nlon)}")
        return dict(land_mask=d["land_mask"], base_albedo=d["base_albedo"], friction=d["friction"], elevation=None)
    raise SystemExit("qingdai_b200.run_simulation needs QD_TOPO_NC (a topography NetCDF, e.g. from the reference's "
                     "scripts/generate_topography.py) or QD_MASK_NPZ; the procedural generator is not part of the hot path")


def main(env=None, log=print):
    env = dict(os.environ if env is None else env)
    nlat, nlon = int(env.get("QD_N_LAT", "121")), int(env.get("QD_N_LON", "240"))
    dt = int(env.get("QD_DT_SECONDS", "300"))
    p = QDParams.from_env(env)
    topo = _topography(nlat, nlon, env)
    extra = {}
    hydro_nc = env.get("QD_HYDRO_NETCDF", "data/hydrology.nc")
    if int(env.get("QD_HYDRO_ENABLE", "1")) == 1 and os.path.exists(hydro_nc):
        from .routing import load_network
        extra.update(routing_network=load_network(hydro_nc), dt_hydro_hours=float(env.get("QD_HYDRO_DT_HOURS", "6")))
    if int(env.get("QD_ECO_ENABLE", "0")) == 1:
        extra.update(with_eco=True, eco_env=env)
    sim = Simulation(nlat, nlon, topo, p, dt=dt, loop_with_albedo=int(env.get("QD_LOOP_WITH_ALBEDO", "0")) == 1, **extra)
    e = sim.engine
    if int(env.get("QD_INIT_BANDED", "0")) == 1:                       # run_simulation.py:310-328
        t_eq, t_pole = float(env.get("QD_INIT_T_EQ", "295.0")), float(env.get("QD_INIT_T_POLE", "265.0"))
        ts0 = t_pole + (t_eq - t_pole) * (np.cos(np.deg2rad(sim.grid.lat_mesh)) ** 2)
        e.set("ts", ts0)
        e.set("sst", np.where(np.asarray(topo["land_mask"]) == 0, ts0, e.get("sst")))
    if env.get("QD_RESTART_IN") and os.path.exists(env["QD_RESTART_IN"]):
        sim.load_restart(env["QD_RESTART_IN"])
        log(f"[Restart] Loaded '{env['QD_RESTART_IN']}' (t = {sim.t:.0f} s)")
    if env.get("QD_TOTAL_YEARS"):
        duration = float(env["QD_TOTAL_YEARS"]) * OrbitalSystem().T_planet
    elif env.get("QD_SIM_DAYS"):
        duration = float(env["QD_SIM_DAYS"]) * DAY_SECONDS
    else:
        duration = 5 * OrbitalSystem().T_planet
    nsteps = int(np.ceil(duration / dt))                               # time_steps = np.arange(0, duration, dt)
    every = max(1, int(float(env.get("QD_DIAG_EVERY_DAYS", "1")) * DAY_SECONDS / dt))
    done = 0
    while done < nsteps:
        n = min(every, nsteps - done)
        sim.step(n)
        done += n
        d = sim.diagnostics()
        log(f"[day {sim.t / DAY_SECONDS:8.3f}] <Ts> {d['ts_mean']:.3f} K  <SST> {d['sst_mean']:.3f} K  <q> {d['q_mean']:.3e}  "
            f"<cloud> {d['cloud_mean']:.3f}  <P> {d['precip_mean']:.3e}  <albedo> {d['albedo_mean']:.3f}  max|u| {d['u_absmax']:.2f} m/s")
    if env.get("QD_RESTART_OUT"):
        sim.save_restart(env["QD_RESTART_OUT"])
        log(f"[Restart] Saved final state to '{env['QD_RESTART_OUT']}'.")
    return sim


if __name__ == "__main__":
    main()
    sys.exit(0)

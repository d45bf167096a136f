"""Fused script loop: the per-step body of scripts/run_simulation.py:1760-2344 on the device.

``Simulation`` is what ``qingdai_b200.run_simulation.main`` (the env-driven entry point) drives and what bench.py times: state
lives in HBM, each step consumes one ``qd_forcing_t`` (10 doubles) and nothing comes back unless
asked for (diagnostics / plots / autosave pull fields through the C ABI on demand).
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np

from . import constants as const
from ._binding import Forcing
from .engine import Engine
from .forcing import OrbitalSystem, ThermalForcing
from .grid import SphericalGrid
from .params import QDParams

DAY_SECONDS = 2 * np.pi / const.PLANET_OMEGA          # run_simulation.py:1594


def q_sat_host(T, p0):
    """humidity.py:85-101, host evaluation for the initial q only (dynamics.py:84)."""
    Tc = np.clip(np.asarray(T, dtype=float) - 273.15, -80.0, 60.0)
    es = 610.94 * np.exp(17.625 * Tc / (Tc + 243.04))
    den = np.maximum(p0 - (1.0 - 0.622) * es, 1.0)
    return np.clip(0.622 * es / den, 0.0, 0.5)


class Simulation:
    CHUNK = 64          # steps enqueued per library call; the host evaluates the next chunk's orbital scalars meanwhile

    def __init__(self, nlat, nlon, topo: dict | Sequence[dict], params: Optional[QDParams | Sequence[QDParams]] = None,
                 dt=300, batch=1, with_ocean=True, with_hydrology=True, with_eco=False, loop_with_albedo=False,
                 device=None, lib=None, t0=0.0, eco_env=None, band=None, routing_network=None, dt_hydro_hours=6.0):
        self.grid = SphericalGrid(nlat, nlon)
        plist = list(params) if isinstance(params, (list, tuple)) else [params or QDParams.from_env()] * batch
        self.engine = Engine(nlat, nlon, batch=batch, params=plist, dt=dt, device=device, lib=lib, band=band)
        self.dt = dt
        self.t = float(t0)
        self.step_index = 0
        self.forcing = ThermalForcing(self.grid, OrbitalSystem())
        self.cfg = dict(with_ocean=with_ocean, with_hydrology=with_hydrology, with_eco=with_eco,
                        loop_with_albedo=loop_with_albedo, with_routing=False)
        topos = list(topo) if isinstance(topo, (list, tuple)) else [topo] * batch
        e = self.engine
        for b, tp in enumerate(topos):
            e.set_mask("land", tp["land_mask"], member=b)
            e.set("friction", tp["friction"], member=b)
            e.set("base_albedo", tp["base_albedo"], member=b)
            land = np.asarray(tp["land_mask"])
            e.set("cs_map", np.where(land == 1, plist[b].Cs_land, plist[b].Cs_ocean).astype(float), member=b)
        has_elev = [tp.get("elevation") is not None for tp in topos]
        if any(has_elev) != all(has_elev):
            raise ValueError("ensemble members of one batch must all come with an elevation map or all without (one launch structure)")
        if has_elev[0]:
            for b, tp in enumerate(topos):
                e.set_elevation(tp["elevation"], member=b)
        self.reset_state()
        self.routing = None
        if routing_network is not None:
            # run_simulation.py:1297-1311: P014 river routing on the land runoff of the bucket, events every dt_hydro
            from .engine import _ENGINES
            from .routing import RiverRouting
            if batch != 1:
                raise ValueError("river routing drives one ensemble member per Simulation")
            _ENGINES[id(self.grid)] = e                      # the drop-in binds to the engine of its grid object
            self.routing = RiverRouting(self.grid, routing_network, dt_hydro_hours=dt_hydro_hours, diag=False)
            self.cfg["with_routing"] = True
        self.eco = None
        if with_eco:
            # run_simulation.py:1335-1337 builds the adapter; :1716-1723 calls it once at t=0 before the loop
            from .ecology import EcologyAdapter
            if batch != 1:
                raise ValueError("with_eco drives one ensemble member per Simulation")
            self.eco = EcologyAdapter(self.grid, topos[0]["land_mask"], engine=e, env=eco_env)
            self.eco.step_subdaily(self.forcing.calculate_insolation(float(t0)), 0.0, dt)

    def reset_state(self):
        """SpectralModel.__init__ / WindDrivenSlabOcean.__init__ initial fields (dynamics.py:56-88, ocean.py:85-94)."""
        e, g = self.engine, self.grid
        zeros = np.zeros(e.shape)
        lat_rad = np.deg2rad(g.lat_mesh)
        for b, p in enumerate(e.params):
            Ts = np.full(e.shape, 288.0)
            e.set("u", zeros, b); e.set("v", zeros, b)
            e.set("h", np.full(e.shape, float(p.H)) + 300 * (np.sin(lat_rad) ** 2), b)
            e.set("ts", Ts, b); e.set("cloud", zeros, b); e.set("hice", zeros, b)
            e.set("q", float(np.clip(p.q_init_rh, 0.0, 1.0)) * q_sat_host(Ts, p.p0), b)
            for k in ("uo", "vo", "eta", "wland", "ssnow", "eflux", "pcond", "lh", "lhrel", "eday"):
                e.set(k, zeros, b)
            land = e._land[b] if e._land[b] is not None else np.zeros(e.shape, dtype=np.uint8)
            e.set("sst", np.where(land == 0, Ts, 288.0), b)
        e.set_counters(0, 0, 0)

    # reference restart variable -> device field (run_simulation.py:63-123, 1436-1470)
    _RESTART_MAP = {"u": "u", "v": "v", "h": "h", "T_s": "ts", "cloud_cover": "cloud", "q": "q", "h_ice": "hice",
                    "uo": "uo", "vo": "vo", "eta": "eta", "Ts": "sst", "W_land": "wland", "S_snow": "ssnow", "C_snow": "csnow"}

    def save_restart(self, path, member=0, dtype="f4"):
        """Reference-format warm-restart file (run_simulation.save_restart: minimal prognostic state, float32 unless
        dtype="f8") of one member, readable by the reference's load_restart.  For a bit-exact resume use save_checkpoint."""
        from types import SimpleNamespace
        from .restart import save_restart
        e = self.engine
        g = {k: e.get(f, member) for k, f in self._RESTART_MAP.items()}
        gcm = SimpleNamespace(**{k: g[k] for k in ("u", "v", "h", "T_s", "cloud_cover", "q", "h_ice")})
        oc = SimpleNamespace(**{k: g[k] for k in ("uo", "vo", "eta", "Ts")}) if self.cfg["with_ocean"] else None
        save_restart(path, self.grid, gcm, oc, e.get_mask("land", member), W_land=g["W_land"], S_snow=g["S_snow"],
                     C_snow=g["C_snow"], t_seconds=self.t, dtype=dtype)

    def load_restart(self, path, member=None):
        """Warm restart from a reference-format file (run_simulation.py:1436-1470): the variables the file holds
        replace the device fields (all members unless ``member`` is given), t_seconds becomes the clock.  Grid shapes
        must match.  Lagged diagnostics (precipitation, cloud_eff, latent heat ...) restart from their current values,
        as in the reference."""
        from .restart import load_restart
        d = load_restart(path)
        for k, f in self._RESTART_MAP.items():
            a = d.get(k)
            if a is None:
                continue
            a = np.asarray(a, dtype=np.float64)
            if a.shape != self.engine.shape:
                raise ValueError(f"restart variable {k!r} has shape {a.shape}, the simulation is {self.engine.shape}")
            if k == "cloud_cover":
                a = np.clip(a, 0.0, 1.0)                   # run_simulation.py:1444
            elif k == "h_ice":
                a = np.maximum(a, 0.0)                     # run_simulation.py:1446
            self.engine.set(f, a, member)
        if d.get("t_seconds") is not None:
            self.t = float(d["t_seconds"])
        self._forcing_ahead = None
        return d

    def save_checkpoint(self, path):
        """Every field, mask, counter and the clock as float64 NetCDF-3 (restart.py); resume is bit-exact."""
        from .restart import save_checkpoint
        save_checkpoint(path, self)

    def load_checkpoint(self, path):
        from .restart import load_checkpoint
        load_checkpoint(path, self)

    def forcing_for(self, t):
        ((fa, sa, ca, aa), (fb, sb, cb, ab)), theta = self.forcing.star_geometry(t)
        return Forcing(t, fa, sa, ca, aa, fb, sb, cb, ab, theta)

    def step(self, nsteps=1):
        done = 0
        while done < nsteps:
            n = nsteps - done
            rr = self.routing
            if rr is not None:      # stop at the next routing event (routing.py:238: t_accum >= dt_hydro)
                left = rr.dt_hydro_seconds - rr.t_accum
                n = max(1, min(n, int(np.ceil((left - 1e-9) / self.dt))))
            n = min(n, self.CHUNK)
            # {t: Forcing} evaluated while the previous chunk was running on the device (host NumPy, ~30 us per step)
            pre = getattr(self, "_forcing_ahead", None) or {}
            fl = []
            for k in range(n):
                tk = self.t + k * self.dt
                f = pre.get(tk)
                fl.append(f if f is not None else self.forcing_for(tk))
            self.engine.loop_steps(fl, self.dt, **self.cfg)    # asynchronous: the step graphs are only enqueued here
            self.t += n * self.dt
            self.step_index += n
            done += n
            # orbital scalars of what comes next -- the rest of this call, else the first step of the next call --
            # overlap with the device work just enqueued
            ahead = max(1, min(nsteps - done, self.CHUNK))
            self._forcing_ahead = {self.t + k * self.dt: self.forcing_for(self.t + k * self.dt) for k in range(ahead)}
            if rr is not None:      # the fused step accumulated R * area * dt on the device (k_route_accumulate)
                rr.t_accum += n * self.dt
                if rr.t_accum + 1e-9 >= rr.dt_hydro_seconds:
                    rr.maybe_route(self.engine.get("precip"), self.engine.get("eflux"))

    def diagnostics(self, member=0):
        """Area-weighted means used by the reference's periodic prints, from one device reduction kernel."""
        d = self.engine.diag()[member]
        out = {name + "_mean": d[name] for name in ("ts", "h", "q", "cloud", "precip", "albedo", "sst")}
        out["u_absmax"] = d["uabs_max"]
        return out

    def energy_diagnostics(self, member=0):
        """energy.compute_energy_diagnostics (energy.py:494-538) of the current state."""
        d = self.engine.diag()[member]
        toa = d["I"] - d["R"] - d["OLR"]
        sfc = d["SW_sfc"] - d["LW_sfc"] - d["SH"] - d["LH"]
        return {"TOA_net": toa, "SFC_net": sfc, "ATM_net": toa - sfc, "I_mean": d["I"], "R_mean": d["R"], "OLR_mean": d["OLR"],
                "SW_sfc_mean": d["SW_sfc"], "LW_sfc_mean": d["LW_sfc"], "SH_mean": d["SH"], "LH_mean": d["LH"]}

    def water_closure(self, member=0, dt_since_prev=None, prev_total=None):
        """hydrology.diagnose_water_closure (hydrology.py:270-340) of the current state."""
        d = self.engine.diag()[member]
        p = self.engine.params[member]
        res = {"CWV_mean": float(p.rho_a) * float(p.h_mbl) * d["q"], "ICE_mean": float(p.rho_i) * d["hice"], "W_land_mean": d["wland"],
               "S_snow_mean": d["ssnow"], "E_mean": d["eflux"], "P_mean": d["precip"], "R_mean": d["rland"]}
        res["total_reservoir_mean"] = res["CWV_mean"] + res["ICE_mean"] + res["W_land_mean"] + res["S_snow_mean"]
        if dt_since_prev is not None and prev_total is not None and dt_since_prev > 0:
            res["d/dt_total_mean"] = (res["total_reservoir_mean"] - prev_total) / float(dt_since_prev)
            res["closure_residual"] = res["d/dt_total_mean"] - (res["E_mean"] - res["P_mean"] - res["R_mean"])
        return res

    def ocean_diagnostics(self, member=0):
        """WindDrivenSlabOcean.diagnostics (ocean.py:535-561)."""
        d = self.engine.diag()[member]
        p = self.engine.params[member]
        a = const.PLANET_RADIUS
        dlat, dlon = self.engine.dlat, self.engine.dlon
        min_cos = float(np.min(np.maximum(np.cos(np.deg2rad(self.grid.lat)), 0.5)))
        dx_min = min(a * dlat, a * dlon * max(1e-3, min_cos))
        return {"KE_mean": d["KE_ocean"], "U_max": d["uocean_max"], "eta_min": d["eta_min"], "eta_max": d["eta_max"],
                "cfl_per_s": float(np.sqrt(p.oc_g * p.oc_H) / max(1e-12, dx_min))}

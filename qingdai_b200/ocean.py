"""Drop-in for ``pygcm.ocean.WindDrivenSlabOcean`` (ocean.py:27-561) backed by libqd_b200."""
from __future__ import annotations

import os

import numpy as np

from . import constants as const
from .engine import engine_for_grid


def _prop(fld):
    def get(self):
        return self._engine.get(fld)

    def set_(self, value):
        self._engine.set(fld, np.asarray(value, dtype=np.float64))
    return property(get, set_)


class WindDrivenSlabOcean:
    uo = _prop("uo")
    vo = _prop("vo")
    eta = _prop("eta")
    Ts = _prop("sst")

    def __init__(self, grid, land_mask, H_m, init_Ts=None, rho_w=None, cp_w=None):
        self.grid = grid
        self.land_mask = np.asarray(land_mask, dtype=int)
        self.H = float(H_m)
        self._engine = engine_for_grid(grid)
        e = self._engine
        over = dict(oc_H=self.H)
        over["oc_rho_w"] = float(os.getenv("QD_RHO_W", str(rho_w if rho_w is not None else 1000.0)))   # ocean.py:49-50
        over["oc_cp_w"] = float(os.getenv("QD_CP_W", str(cp_w if cp_w is not None else 4200.0)))
        e.set_params([p.replace(**over) for p in e.params])
        self.rho_w, self.cp_w, self.g = over["oc_rho_w"], over["oc_cp_w"], 9.81
        self.a = const.PLANET_RADIUS
        e.set_mask("land", self.land_mask)
        shape = grid.lat_mesh.shape
        zeros = np.zeros(shape)
        e.set("uo", zeros); e.set("vo", zeros); e.set("eta", zeros)
        e.set("sst", np.full(shape, 288.0) if init_Ts is None else np.array(init_Ts, dtype=float))
        a, _, c = e.counters()
        e.set_counters(a, 0, c)

    @property
    def _step(self):
        return self._engine.counters()[1]

    def step(self, dt, u_atm, v_atm, Q_net=None, ice_mask=None):
        """ocean.py:265: one ocean step driven by host wind / heat-flux / ice arrays."""
        e = self._engine
        # the winds are staged in tensors owned by this object: the atmosphere's u / v on the shared engine stay untouched
        import torch
        if getattr(self, "_wind", None) is None:
            self._wind = torch.empty((2, e.batch, e.nlat, e.nlon), dtype=torch.float64, device=e.device)
        for k, a in enumerate((u_atm, v_atm)):
            self._wind[k].copy_(torch.from_numpy(np.array(np.broadcast_to(np.asarray(a, dtype=np.float64), (e.batch,) + e.shape))))
        if Q_net is not None:
            e.set("qnet", np.asarray(Q_net, dtype=np.float64))
        if ice_mask is not None:
            e.set_mask("ice", np.asarray(ice_mask).astype(np.uint8))
        e.ocean_step(dt, has_q=Q_net is not None, has_ice=ice_mask is not None, winds=(self._wind[0], self._wind[1]))

    def diagnostics(self):
        """ocean.py:535-561."""
        p = self._engine.params[0]
        lat_rad = np.deg2rad(self.grid.lat_mesh)
        w = np.maximum(np.cos(lat_rad), 0.0)
        uo, vo, eta = self.uo, self.vo, self.eta
        KE = 0.5 * (uo ** 2 + vo ** 2)
        coslat = np.maximum(np.cos(lat_rad), 0.5)
        dx_min = min(self.a * self.grid.dlat_rad, self.a * self.grid.dlon_rad * max(1e-3, float(np.min(coslat))))
        return {"KE_mean": float(np.sum(KE * w) / (np.sum(w) + 1e-15)),
                "U_max": float(np.max(np.sqrt(uo ** 2 + vo ** 2))),
                "eta_min": float(np.min(eta)), "eta_max": float(np.max(eta)),
                "cfl_per_s": float(np.sqrt(p.oc_g * self.H) / max(1e-12, dx_min))}

"""Drop-in for ``pygcm.dynamics.SpectralModel`` (dynamics.py:17-667) backed by libqd_b200.

Same constructor, same ``time_step(Teq_field, dt, albedo=None)``, same attribute names.  State
lives on the GPU; attribute reads pull a host copy through the C ABI, attribute writes upload
(``gcm.T_s = ...``, ``gcm.cloud_cover = ...`` as scripts/run_simulation.py:1441-1447,1900,2253 do).
There is no NumPy fallback: without the CUDA library / a CUDA device construction raises.
"""
from __future__ import annotations

import os

import numpy as np

from . import constants as const
from .engine import engine_for_grid
from .params import QDParams

# reference attribute -> device field
_FIELDS = {"u": "u", "v": "v", "h": "h", "T_s": "ts", "q": "q", "cloud_cover": "cloud", "h_ice": "hice",
           "isr": "isr", "isr_A": "isr_a", "isr_B": "isr_b", "olr": "olr", "E_flux_last": "eflux",
           "P_cond_flux_last": "pcond", "LH_last": "lh", "LH_release_last": "lhrel", "C_snow_map_last": "csnow"}


def _field_property(attr, fld):
    def get(self):
        return self._engine.get(fld)

    def set_(self, value):
        self._engine.set(fld, np.asarray(value, dtype=np.float64))
    return property(get, set_, doc=f"host view of device field '{fld}' (reference attribute SpectralModel.{attr})")


class SpectralModel:
    def __init__(self, grid, friction_map, initial_state=None, g=9.81, H=8000, tau_rad=1e6, greenhouse_factor=0.15,
                 C_s_map=None, land_mask=None, Cs_ocean=None, Cs_land=None, Cs_ice=None, seaice_enabled=None,
                 t_freeze=None, rho_i=None, L_f=None):
        self.grid = grid
        env = QDParams.from_env()
        over = dict(g=float(g), H=float(H), tau_rad=float(tau_rad), gh_newton=float(greenhouse_factor))
        if Cs_ocean is not None:
            over["Cs_ocean"] = float(Cs_ocean)
        if Cs_land is not None:
            over["Cs_land"] = float(Cs_land)
        if Cs_ice is not None:
            over["Cs_ice"] = float(Cs_ice)
        if seaice_enabled is not None:
            over["seaice_enabled"] = bool(seaice_enabled)
        if t_freeze is not None:
            over["t_freeze"] = float(t_freeze)
        if rho_i is not None:
            over["rho_i"] = float(rho_i)
        if L_f is not None:
            over["L_f"] = float(L_f)
        if land_mask is None:
            over["seaice_enabled"] = False            # dynamics.py:392: sea-ice path needs a land mask
        self._params = env.replace(**over)
        self._engine = engine_for_grid(grid, params=self._params)
        e = self._engine
        e.set_params(self._params)
        self.friction_map = np.asarray(friction_map, dtype=np.float64)
        self.C_s_map = C_s_map
        self.land_mask = land_mask
        self.Cs_ocean, self.Cs_land, self.Cs_ice = Cs_ocean, Cs_land, Cs_ice
        self.seaice_enabled = self._params.seaice_enabled
        self.t_freeze, self.rho_i, self.L_f = self._params.t_freeze, self._params.rho_i, self._params.L_f
        self.g, self.H, self.tau_rad, self.greenhouse_factor = g, H, tau_rad, greenhouse_factor
        self.a = const.PLANET_RADIUS
        self.dlat_rad = np.deg2rad(grid.lat[1] - grid.lat[0])
        self.dlon_rad = np.deg2rad(grid.lon[1] - grid.lon[0])
        self.energy_w = self._params.energy_w
        self.hum_params = self._params
        self.energy_params = self._params
        shape = grid.lat_mesh.shape
        e.set("friction", self.friction_map)
        e.set_mask("land", np.zeros(shape, dtype=np.uint8) if land_mask is None else land_mask)
        if C_s_map is not None:
            e.set("cs_map", C_s_map)
        # initial state (dynamics.py:56-88)
        zeros = np.zeros(shape)
        lat_rad = np.deg2rad(grid.lat_mesh)
        Ts = np.full(shape, 288.0)
        e.set("u", zeros); e.set("v", zeros)
        e.set("h", np.full(shape, float(H)) + 300 * (np.sin(lat_rad) ** 2))
        e.set("ts", Ts); e.set("cloud", zeros); e.set("hice", zeros)
        from .simulation import q_sat_host
        RH0 = float(os.getenv("QD_Q_INIT_RH", "0.5"))
        e.set("q", float(np.clip(RH0, 0.0, 1.0)) * q_sat_host(Ts, self._params.p0))
        for k in ("isr", "isr_a", "isr_b", "olr", "eflux", "pcond", "lh", "lhrel"):
            e.set(k, zeros)
        a, o, _ = e.counters()
        e.set_counters(0, o, 0)
        self.glacier_mask_last = np.zeros(shape, dtype=bool)

    @property
    def _step_counter(self):
        return self._engine.counters()[0]

    @property
    def cloud_eff_last(self):
        if not self._engine.counters()[2]:
            raise AttributeError("cloud_eff_last")        # set only by the energy branch (dynamics.py:353)
        return self._engine.get("cloud_eff")

    def time_step(self, Teq_field, dt, albedo=None):
        """dynamics.py:260: advance one step; ``albedo`` given -> explicit energy branch live."""
        e = self._engine
        e.set("teq", np.asarray(Teq_field, dtype=np.float64))
        if albedo is not None:
            e.set("albedo", np.asarray(albedo, dtype=np.float64))
        e.atmos_step(dt, has_albedo=albedo is not None)

    # numerics helpers kept for API parity (dynamics.py:90-258)
    def _advect(self, field, dt):
        cos = np.maximum(1e-6, np.cos(np.deg2rad(self.grid.lat)))
        return self._engine.op_advect(field, self.u, self.v, dt, cos)

    def _laplacian_sphere(self, F):
        return self._engine.op_laplacian(F, np.maximum(np.cos(np.deg2rad(self.grid.lat)), 0.2))

    def _hyperdiffuse(self, F, k4, dt, n_substeps=1):
        return self._engine.op_hyperdiffuse(F, k4, dt, n_substeps, np.maximum(np.cos(np.deg2rad(self.grid.lat)), 0.2))

    def _shapiro_filter(self, F, n=2, lon_wrap=True):
        if not lon_wrap:
            raise NotImplementedError("lon_wrap=False is never used by the reference's step")
        return self._engine.op_shapiro(F, n)

    def _spectral_zonal_filter(self, F, cutoff=0.75, damp=0.5):
        return self._engine.op_bandstop(F, cutoff, damp)


for _a, _f in _FIELDS.items():
    setattr(SpectralModel, _a, _field_property(_a, _f))

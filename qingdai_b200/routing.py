"""Drop-in for ``pygcm.routing.RiverRouting`` (routing.py:74-354) backed by libqd_b200.

Per model step the land runoff is accumulated on the device (``qd_route_accumulate``); every
``dt_hydro`` the event is routed by the level-synchronous gather of ``csrc/qd_route.cuh`` whose
floating-point additions happen in exactly the order of the reference's serial loop, so the flow
accumulation map and the ocean inflow are bit-identical.  Lake (P-E) bookkeeping and the closure
sums are host NumPy on the arrays the event returns (event cadence: every 6 model hours).
The network can be given as a dict of arrays or as a NetCDF path (netCDF4 when installed, else the NetCDF-3 reader
of qingdai_b200.ncio).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import numpy as np

from . import constants
from .engine import engine_for_grid, _ptr


def load_network(path_or_dict):
    if isinstance(path_or_dict, dict):
        return {k: np.asarray(v) for k, v in path_or_dict.items()}
    try:
        from netCDF4 import Dataset
    except Exception:           # no netCDF4 on this box: NetCDF-3 files (hydrology_network.save_network) still open
        from .ncio import Dataset
    out = {}
    with Dataset(path_or_dict, "r") as ds:
        for k in ("land_mask", "flow_to_index", "flow_order", "lake_mask", "lake_id", "lake_outlet_index", "lake_outlet_i", "lake_outlet_j"):
            if k in ds.variables:
                out[k] = np.array(ds.variables[k][:])
    return out


class RiverRouting:
    def __init__(self, grid, network_nc_path, dt_hydro_hours: float = 6.0, treat_lake_as_water: bool = True,
                 alpha_lake: Optional[float] = None, diag: bool = True, member: int = 0):
        self.grid = grid
        self.dt_hydro_seconds = float(dt_hydro_hours) * 3600.0
        self.treat_lake_as_water = bool(treat_lake_as_water)
        self.alpha_lake = alpha_lake
        self.diag_enabled = bool(diag)
        self.n_lat, self.n_lon = int(grid.n_lat), int(grid.n_lon)
        self.shape = (self.n_lat, self.n_lon)
        self.n_cells = self.n_lat * self.n_lon
        self.member = member
        net = load_network(network_nc_path)
        self.land_mask = (net["land_mask"] > 0).astype(np.uint8)
        self.flow_to_index = net["flow_to_index"].astype(np.int64)
        if self.flow_to_index.shape != self.shape:
            raise RuntimeError(f"flow_to_index shape {self.flow_to_index.shape} != grid shape {self.shape}")
        if "flow_order" in net:
            self.flow_order = net["flow_order"].astype(np.int64)
        else:
            self.flow_order = np.where(self.land_mask.flatten(order="C") == 1)[0].astype(np.int64)   # routing.py:126-130
        self.lake_mask = net.get("lake_mask")
        self.lake_id = net.get("lake_id")
        out_idx = net.get("lake_outlet_index")
        if out_idx is None and "lake_outlet_i" in net and "lake_outlet_j" in net:
            out_idx = net["lake_outlet_j"].astype(np.int64) * self.n_lon + net["lake_outlet_i"].astype(np.int64)
        self.lake_outlet_index = out_idx
        self.n_lakes = int(np.max(self.lake_id)) if self.lake_id is not None else 0
        if self.n_lakes > 0 and self.lake_outlet_index is not None and self.lake_outlet_index.shape[0] != self.n_lakes:
            self.n_lakes = min(self.n_lakes, self.lake_outlet_index.shape[0])
            self.lake_outlet_index = self.lake_outlet_index[: self.n_lakes]
        self.cell_area = self._compute_cell_areas()
        self.t_accum = 0.0
        self._diag_cache = None
        self.lake_volume_kg = np.zeros(self.n_lakes) if self.n_lakes > 0 else None
        self._engine = engine_for_grid(grid)
        e = self._engine
        has_lakes = self.lake_mask is not None and self.lake_id is not None and self.n_lakes > 0 and self.lake_outlet_index is not None
        fo = np.ascontiguousarray(self.flow_order, dtype=np.int64)
        ft = np.ascontiguousarray(self.flow_to_index.reshape(-1), dtype=np.int64)
        land = np.ascontiguousarray(self.land_mask.reshape(-1), dtype=np.uint8)
        lk = np.ascontiguousarray((self.lake_mask.reshape(-1) > 0).astype(np.uint8)) if has_lakes else None
        lid = np.ascontiguousarray(self.lake_id.reshape(-1).astype(np.int32)) if has_lakes else None
        lo = np.ascontiguousarray(self.lake_outlet_index.astype(np.int64)) if has_lakes else None
        e._chk(e.lib.qd_route_setup(e.ctx, int(fo.size), _ptr(fo), _ptr(ft), _ptr(land),
                                    _ptr(lk) if has_lakes else None, _ptr(lid) if has_lakes else None,
                                    self.n_lakes if has_lakes else 0, _ptr(lo) if has_lakes else None), "qd_route_setup")
        self.levels = int(e.lib.qd_route_levels(e.ctx))
        self._has_lakes = has_lakes
        if self.diag_enabled:
            print(f"[Routing] Loaded network: land={int(self.land_mask.sum())} cells, n_lakes={self.n_lakes}, "
                  f"dt_hydro={self.dt_hydro_seconds / 3600.0:.1f} h, levels={self.levels}")

    def _compute_cell_areas(self):
        """routing.py:176-200."""
        R = float(constants.PLANET_RADIUS)
        lats = np.asarray(self.grid.lat, dtype=float)
        lons = np.asarray(self.grid.lon, dtype=float)
        dphi = np.deg2rad(abs(lats[1] - lats[0]))
        dlam = np.deg2rad(abs(lons[1] - lons[0]))
        pc = np.deg2rad(self.grid.lat_mesh[:, 0])
        band = np.sin(np.clip(pc + 0.5 * dphi, -0.5 * np.pi, 0.5 * np.pi)) - np.sin(np.clip(pc - 0.5 * dphi, -0.5 * np.pi, 0.5 * np.pi))
        return np.repeat(((R * R) * dlam * band)[:, None], self.n_lon, axis=1)

    def reset(self):
        e = self._engine
        z = np.zeros(self.n_cells)
        e._chk(e.lib.qd_route_buffer(e.ctx, self.member, _ptr(z), 1), "qd_route_buffer")
        self.t_accum = 0.0
        if self.lake_volume_kg is not None:
            self.lake_volume_kg.fill(0.0)
        self._diag_cache = None

    @property
    def buffer_kg(self):
        e = self._engine
        out = np.empty(self.n_cells)
        e._chk(e.lib.qd_route_buffer(e.ctx, self.member, _ptr(out), 0), "qd_route_buffer")
        return out

    def accumulate_device(self, dt_seconds):
        """Accumulate the runoff already resident in the device field ``rland`` (fused-loop path)."""
        e = self._engine
        e._chk(e.lib.qd_route_accumulate(e.ctx, float(dt_seconds)), "qd_route_accumulate")
        self.t_accum += float(dt_seconds)

    def step(self, R_land_flux, dt_seconds, precip_flux=None, evap_flux=None):
        """routing.py:211-335."""
        R = np.asarray(R_land_flux, dtype=float)
        if R.shape != self.shape:
            raise ValueError(f"R_land_flux shape {R.shape} != grid shape {self.shape}")
        e = self._engine
        e.set("rland", R, self.member)
        self.accumulate_device(dt_seconds)
        self.maybe_route(precip_flux, evap_flux)

    def maybe_route(self, precip_flux=None, evap_flux=None):
        if self.t_accum + 1e-9 < self.dt_hydro_seconds:
            return False
        event_dt = self.t_accum
        self.t_accum = 0.0
        e = self._engine
        flow = np.empty(self.n_cells)
        after = np.empty(self.n_cells)
        inp = np.empty(self.n_cells)
        ocean = C.c_double(0.0)
        lake_store = np.zeros(max(self.n_lakes, 1))
        e._chk(e.lib.qd_route_event(e.ctx, self.member, _ptr(flow), C.byref(ocean), _ptr(after), _ptr(inp), _ptr(lake_store)), "qd_route_event")
        mass_input = float(np.sum(inp))                         # routing.py:252
        residual = float(np.sum(after))                         # routing.py:301
        ocean_kg = float(ocean.value)
        if self.lake_volume_kg is not None:
            self.lake_volume_kg += lake_store[: self.n_lakes]
        lake_delta = 0.0
        if self._has_lakes and self.lake_volume_kg is not None and precip_flux is not None and evap_flux is not None:
            P = np.asarray(precip_flux, dtype=float)                # routing.py:303-318
            E = np.asarray(evap_flux, dtype=float)
            lm = self.lake_mask.astype(bool)
            net = (P - E) * self.cell_area * event_dt
            lake_add = float(np.sum(np.where(lm, net, 0.0)))
            if lake_add != 0.0 and self.n_lakes > 0:
                ids = self.lake_id
                tot = np.sum(np.where(lm, self.cell_area, 0.0))
                for k in range(1, self.n_lakes + 1):
                    la = np.sum(np.where(ids == k, self.cell_area, 0.0))
                    self.lake_volume_kg[k - 1] += (0.0 if la <= 0 else la / tot) * lake_add
                lake_delta = lake_add
        closure = mass_input - (ocean_kg + lake_delta + residual)
        self._diag_cache = {
            "flow_accum_kgps": (flow / max(event_dt, 1e-9)).reshape(self.shape, order="C"),
            "ocean_inflow_kgps": float(ocean_kg / max(event_dt, 1e-9)),
            "mass_closure_error_kg": float(closure),
            "lake_volume_kg": (self.lake_volume_kg.copy() if self.lake_volume_kg is not None else None),
        }
        if self.diag_enabled:
            print(f"[HydroRouting] ocean_inflow={self._diag_cache['ocean_inflow_kgps']:.3e} kg/s | "
                  f"mass_error={self._diag_cache['mass_closure_error_kg']:.3e} kg")
        return True

    def diagnostics(self) -> Dict[str, object]:
        if self._diag_cache is None:
            return {"flow_accum_kgps": np.zeros(self.shape), "ocean_inflow_kgps": 0.0, "mass_closure_error_kg": 0.0,
                    "lake_volume_kg": (np.zeros(self.n_lakes) if self.n_lakes > 0 else None)}
        return dict(self._diag_cache)

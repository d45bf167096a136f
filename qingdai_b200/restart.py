"""Restart / topography / ocean-state files without netCDF4.

Same function names, arguments, variable names, dimensions and attributes as the reference's helpers, written as
NetCDF-3 through ``qingdai_b200.ncio`` so that they work on boxes without netCDF4/HDF5 and stay readable by the
reference (netCDF4 opens NetCDF-3 classic files):

  save_restart / load_restart            scripts/run_simulation.py:63-123, 161-184
  save_topography                        scripts/run_simulation.py:125-159
  load_topography_from_netcdf            pygcm/topography.py:428-575   (incl. the bilinear / nearest regrid)
  save_ocean / load_ocean                scripts/run_simulation.py:186-246

Additions: ``dtype="f8"`` on the writers (the reference stores float32; a float64 restart resumes bit-exactly) and
``save_checkpoint`` / ``load_checkpoint`` for a ``Simulation``: EVERY device field, mask, counter and clock, plus the
routing buffer / lake volumes and the sub-daily ecology state when those are coupled, so that
run(n) -> save -> load -> run(m) equals run(n + m) bit for bit (tests/test_hostcheck.py, tests/test_gpu.py).
"""
from __future__ import annotations

import os

import numpy as np

from .ncio import Dataset

_RESTART_FIELDS = ["u", "v", "h", "T_s", "cloud_cover", "q", "h_ice", "uo", "vo", "eta", "Ts", "W_land", "S_snow", "C_snow", "land_mask"]


def _mkdir_for(path):
    os.makedirs(os.path.dirname(path) or ".", exist_ok=True)


def _coords(ds, grid):
    ds.createDimension("lat", grid.n_lat)
    ds.createDimension("lon", grid.n_lon)
    vlat = ds.createVariable("lat", "f4", ("lat",))
    vlon = ds.createVariable("lon", "f4", ("lon",))
    vlat[:] = np.asarray(grid.lat, dtype=np.float32)
    vlon[:] = np.asarray(grid.lon, dtype=np.float32)


def save_restart(path, grid, gcm, ocean, land_mask, W_land=None, S_snow=None, C_snow=None, t_seconds=None, dtype="f4"):
    """run_simulation.py:63-123.  Minimal prognostic state; ``dtype="f8"`` keeps full precision."""
    _mkdir_for(path)
    np_t = np.float32 if dtype == "f4" else np.float64
    with Dataset(path, "w") as ds:
        _coords(ds, grid)

        def wvar(name, data):
            if data is None:
                return
            var = ds.createVariable(name, dtype, ("lat", "lon"))
            var[:] = np.asarray(data, dtype=np_t)

        for name in ("u", "v", "h", "T_s"):
            wvar(name, getattr(gcm, name))
        for name in ("cloud_cover", "q", "h_ice"):
            wvar(name, getattr(gcm, name, None))
        if ocean is not None:
            for name in ("uo", "vo", "eta", "Ts"):
                wvar(name, getattr(ocean, name, None))
        wvar("W_land", W_land)
        wvar("S_snow", S_snow)
        wvar("C_snow", C_snow)
        wvar("land_mask", land_mask)
        vts = ds.createVariable("t_seconds", "f8")
        vts[...] = float(t_seconds) if (t_seconds is not None) else 0.0
        ds.setncattr("title", "Qingdai GCM Restart")
        ds.setncattr("creator", "PyGCM for Qingdai")
        ds.setncattr("note", "Contains minimal prognostic fields for warm restart (incl. t_seconds).")
        ds.setncattr("format", "v1")


def load_restart(path):
    """run_simulation.py:161-184: dict of arrays, None for variables the file does not hold."""
    out = {}
    with Dataset(path, "r") as ds:
        def rvar(name):
            try:
                return ds.variables[name][:].data
            except Exception:
                return None
        out["lat"] = ds.variables["lat"][:].data
        out["lon"] = ds.variables["lon"][:].data
        for name in _RESTART_FIELDS:
            out[name] = rvar(name)
        try:
            out["t_seconds"] = float(ds.variables["t_seconds"][...])
        except Exception:
            out["t_seconds"] = None
    return out


def save_topography(path, grid, land_mask, base_albedo_map, friction_map, elevation=None, dtype="f4"):
    """run_simulation.py:125-159: lat, lon, land_mask (u1), base_albedo, friction, optional elevation."""
    _mkdir_for(path)
    np_t = np.float32 if dtype == "f4" else np.float64
    with Dataset(path, "w") as ds:
        _coords(ds, grid)
        vmask = ds.createVariable("land_mask", "u1", ("lat", "lon"))
        vmask[:] = np.asarray(land_mask, dtype=np.uint8)
        for name, data in (("base_albedo", base_albedo_map), ("friction", friction_map), ("elevation", elevation)):
            if data is None:
                continue
            v = ds.createVariable(name, dtype, ("lat", "lon"))
            v[:] = np.asarray(data, dtype=np_t)
        ds.setncattr("title", "Qingdai Topography")
        ds.setncattr("source", "scripts/run_simulation.py")
        ds.setncattr("format", "v1")


def _regrid(src_lat, src_lon, field, tgt_lat_mesh, tgt_lon_mesh, is_mask):
    """topography.py:484-520: bilinear (nearest for the mask) on a longitude-cyclic source, latitude clipped to the
    source range, non-finite results replaced by the nearest neighbour."""
    from scipy.interpolate import RegularGridInterpolator
    lon_ext = np.concatenate([src_lon - 360.0, src_lon, src_lon + 360.0])
    ext = np.concatenate([field, field, field], axis=1)
    pts = np.stack([np.clip(tgt_lat_mesh.ravel(), src_lat.min(), src_lat.max()), tgt_lon_mesh.ravel()], axis=-1)

    def run(method):
        f = RegularGridInterpolator((src_lat, lon_ext), ext, bounds_error=False, fill_value=None, method=method)
        return f(pts).reshape(tgt_lat_mesh.shape)

    if is_mask:
        return np.where(run("nearest") >= 0.5, 1, 0).astype(np.uint8)
    vals = run("linear")
    if np.any(~np.isfinite(vals)):
        vals = np.where(np.isfinite(vals), vals, run("nearest"))
    return vals


def load_topography_from_netcdf(path, grid, *, regrid="auto", verbose=False):
    """topography.py:428-575 -> (elevation, land_mask, base_albedo, friction) on ``grid``.

    Longitudes are normalised to [0, 360) and sorted, descending latitudes flipped, a duplicated 0/360 seam column
    dropped; fields whose shape and coordinates match the grid are taken as they are, anything else is regridded
    (``regrid="never"``: shape mismatch raises, matching shapes are taken without a coordinate check)."""
    with Dataset(path, "r") as ds:
        lat = np.asarray(ds["lat"][:], dtype=float)
        lon = np.asarray(ds["lon"][:], dtype=float)
        if np.nanmin(lon) < 0.0 or np.nanmax(lon) <= 180.0:
            lon = np.mod(lon, 360.0)
            lon[lon < 0] += 360.0
        lat_up = bool(np.all(np.diff(lat) > 0))
        if not lat_up:
            lat = lat[::-1]
        order = np.argsort(lon)
        lon = lon[order]

        def field(name):
            a = np.asarray(ds[name][:])
            if not lat_up:
                a = a[::-1, :]
            return a[:, order]
        elev, mask, base, fric = field("elevation"), field("land_mask"), field("base_albedo"), field("friction")
    if lon.size >= 2 and np.isclose(lon[0], 0.0) and np.isclose(lon[-1], 360.0):
        lon = lon[:-1]
        elev, mask, base, fric = elev[:, :-1], mask[:, :-1], base[:, :-1], fric[:, :-1]
    same_shape = elev.shape == (grid.n_lat, grid.n_lon)
    if same_shape and (regrid == "never" or (np.allclose(lat, grid.lat, atol=1e-6) and np.allclose(lon, grid.lon, atol=1e-6))):
        out = elev.astype(float), mask.astype(np.uint8), base.astype(float), fric.astype(float)
    else:
        if not same_shape and regrid == "never":
            raise ValueError(f"Topography grid mismatch: source {elev.shape} vs target {(grid.n_lat, grid.n_lon)} and regrid='never'.")
        la, lo = grid.lat_mesh, grid.lon_mesh
        out = (_regrid(lat, lon, elev, la, lo, False), _regrid(lat, lon, mask, la, lo, True),
               _regrid(lat, lon, base, la, lo, False), _regrid(lat, lon, fric, la, lo, False))
    if verbose:
        w = np.cos(np.deg2rad(grid.lat_mesh))
        print(f"[Topo] Loaded: {path}; land fraction {float((w * (out[1] == 1)).sum() / (w.sum() + 1e-15)):.3f}")
    return out


def save_ocean(path, grid, ocean, day_value=None, dtype="f4"):
    """run_simulation.py:186-218.  Returns True on success (the reference swallows errors the same way)."""
    try:
        _mkdir_for(path)
        np_t = np.float32 if dtype == "f4" else np.float64
        with Dataset(path, "w") as ds:
            _coords(ds, grid)
            for name in ("uo", "vo", "eta", "Ts"):
                data = getattr(ocean, name, None)
                if data is not None:
                    v = ds.createVariable(name, dtype, ("lat", "lon"))
                    v[:] = np.asarray(data, dtype=np_t)
            ds.setncattr("title", "Qingdai Ocean State")
            ds.setncattr("source", "scripts/run_simulation.py")
            if day_value is not None:
                ds.setncattr("day", float(day_value))
        return True
    except Exception as e:          # noqa: BLE001  (reference behaviour: report and carry on)
        print(f"[Ocean] Save failed: {e}")
        return False


def load_ocean(path):
    """run_simulation.py:220-246."""
    out = {"uo": None, "vo": None, "eta": None, "Ts": None, "day": None}
    try:
        with Dataset(path, "r") as ds:
            for name in ("uo", "vo", "eta", "Ts"):
                try:
                    out[name] = ds.variables[name][:].data
                except Exception:
                    out[name] = None
            try:
                out["day"] = float(ds.getncattr("day"))
            except Exception:
                out["day"] = None
    except Exception as e:          # noqa: BLE001
        print(f"[Ocean] Load failed '{path}': {e}")
    return out


# ------------------------------------------------------------------------------------------------ exact checkpoints
def save_checkpoint(path, sim):
    """Every float64 field and mask of every member of ``sim``'s engine, the step counters that drive the Shapiro /
    band-stop cadences and the simulation clock; with routing: the runoff buffer, the accumulation clock and the lake
    volumes (routing.py:232-331); with the sub-daily ecology: the LAI layers, the canopy-cache clock and the cadence
    counters (population.py:57-71,252-286, adapter.py:150-156)."""
    from .engine import F, M
    import ctypes as C
    e = sim.engine
    _mkdir_for(path)
    with Dataset(path, "w") as ds:
        ds.createDimension("member", e.batch)
        ds.createDimension("lat", e.nlat)
        ds.createDimension("lon", e.nlon)
        for name in sorted(F, key=F.get):
            v = ds.createVariable("f_" + name, "f8", ("member", "lat", "lon"))
            for b in range(e.batch):
                v[b] = e.get(name, b)
        for name in sorted(M, key=M.get):
            v = ds.createVariable("m_" + name, "u1", ("member", "lat", "lon"))
            for b in range(e.batch):
                v[b] = e.get_mask(name, b)
        atm, oc, ce = e.counters()
        clock = ds.createVariable("clock", "f8", ())
        clock[...] = float(sim.t)
        ds.setncattr("format", "qd-checkpoint-v1")
        ds.setncattr("step_index", int(sim.step_index))
        ds.setncattr("atm_counter", int(atm))
        ds.setncattr("oc_counter", int(oc))
        ds.setncattr("has_cloud_eff", int(ce))
        ds.setncattr("dt", float(sim.dt))
        ds.setncattr("with_routing", int(sim.routing is not None))
        ds.setncattr("with_eco", int(sim.eco is not None))
        if sim.routing is not None:
            rr = sim.routing
            ds.createDimension("cell", rr.n_cells)
            v = ds.createVariable("routing_buffer_kg", "f8", ("cell",))
            v[:] = rr.buffer_kg
            ds.setncattr("routing_t_accum", float(rr.t_accum))
            if rr.lake_volume_kg is not None and rr.lake_volume_kg.size:
                ds.createDimension("lake", int(rr.lake_volume_kg.size))
                v = ds.createVariable("routing_lake_volume_kg", "f8", ("lake",))
                v[:] = rr.lake_volume_kg
        if sim.eco is not None:
            steps, have = C.c_int(0), C.c_int(0)
            e._chk(e.lib.qd_eco_state(e.ctx, C.byref(steps), C.byref(have), 0), "qd_eco_state")
            ds.setncattr("eco_step_count", int(steps.value))
            ds.setncattr("eco_have_alpha", int(have.value))
            from .engine import S
            sc = e.scalars()[0]
            ds.setncattr("eco_hours", float(sc[S["eco_hours"]]))
            ds.setncattr("eco_next_hours", float(sc[S["eco_next"]]))
            ds.setncattr("eco_cached", int(sc[S["eco_cached"]] != 0.0))
            pop = sim.eco.pop
            if pop is not None:
                lay = pop.LAI_layers_SK
                ds.createDimension("species", lay.shape[0])
                ds.createDimension("cohort", lay.shape[1])
                v = ds.createVariable("eco_lai_layers", "f8", ("species", "cohort", "lat", "lon"))
                v[:] = lay


def load_checkpoint(path, sim):
    """Inverse of save_checkpoint into a Simulation built with the same grid, batch, parameters and topography."""
    from .engine import F, M
    import ctypes as C
    e = sim.engine
    with Dataset(path, "r") as ds:
        if ds.getncattr("format") != "qd-checkpoint-v1":
            raise ValueError(f"{path!r} is not a qingdai_b200 checkpoint")
        shape = (ds.dimensions["member"].size, ds.dimensions["lat"].size, ds.dimensions["lon"].size)
        if shape != (e.batch, e.nlat, e.nlon):
            raise ValueError(f"checkpoint holds {shape}, the simulation is {(e.batch, e.nlat, e.nlon)}")
        if float(ds.getncattr("dt")) != float(sim.dt):
            raise ValueError("checkpoint was written with a different dt")
        if bool(ds.getncattr("with_routing")) != (sim.routing is not None) or bool(ds.getncattr("with_eco")) != (sim.eco is not None):
            raise ValueError("checkpoint and simulation disagree about routing / ecology")
        # every check happens BEFORE the first write: a refused checkpoint leaves the simulation untouched
        land = ds.variables["m_land"][:].data
        for b in range(e.batch):
            if not np.array_equal(land[b], e.get_mask("land", b)):
                raise ValueError("checkpoint land mask differs from the simulation's topography")
            for static in ("friction", "base_albedo", "elevation"):
                if not np.array_equal(ds.variables["f_" + static][:].data[b], e.get(static, b)):
                    raise ValueError(f"checkpoint {static} map differs from the simulation's topography")
        missing = [n for n in F if "f_" + n not in ds.variables and not n.startswith("x")]
        if missing:
            raise ValueError(f"checkpoint lacks the fields {missing}")
        if sim.eco is not None:
            # first the population and its clocks (qd_eco_reset re-snapshots the LAI), then every field on top of it
            pop = sim.eco.pop
            if pop is not None:
                pop.set_lai_layers(ds.variables["eco_lai_layers"][:].data, reset_clock=False)
            e.eco_reset(float(ds.getncattr("eco_hours")), float(ds.getncattr("eco_next_hours")),
                        cached=bool(ds.getncattr("eco_cached")), step_count=int(ds.getncattr("eco_step_count")))
            steps, have = C.c_int(int(ds.getncattr("eco_step_count"))), C.c_int(int(ds.getncattr("eco_have_alpha")))
            e._chk(e.lib.qd_eco_state(e.ctx, C.byref(steps), C.byref(have), 1), "qd_eco_state")
        if sim.routing is not None:
            rr = sim.routing
            buf = np.ascontiguousarray(ds.variables["routing_buffer_kg"][:].data, dtype=np.float64)
            e._chk(e.lib.qd_route_buffer(e.ctx, rr.member, C.c_void_p(buf.ctypes.data), 1), "qd_route_buffer")
            rr.t_accum = float(ds.getncattr("routing_t_accum"))
            if rr.lake_volume_kg is not None and rr.lake_volume_kg.size:
                rr.lake_volume_kg[:] = ds.variables["routing_lake_volume_kg"][:].data
            rr._diag_cache = None
        for name in sorted(F, key=F.get):
            if "f_" + name not in ds.variables:
                continue                                   # a scratch slot added after the checkpoint was written
            data = ds.variables["f_" + name][:].data
            for b in range(e.batch):
                e.set(name, data[b], b)
        for name in sorted(M, key=M.get):
            if name == "land":
                continue                                   # static input of the Simulation; checked, not overwritten
            data = ds.variables["m_" + name][:].data
            for b in range(e.batch):
                e.set_mask(name, data[b], b)
        e.set_counters(int(ds.getncattr("atm_counter")), int(ds.getncattr("oc_counter")), int(ds.getncattr("has_cloud_eff")))
        sim.t = float(ds.variables["clock"][...])
        sim.step_index = int(ds.getncattr("step_index"))
